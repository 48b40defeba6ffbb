"""Device-resident ClickedItemsState against the host class, batch by batch."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu


def test_device_state_tracks_host_state():
    from chameleon_recsys_b200.clicked_items_state import ClickedItemsState
    from chameleon_recsys_b200.device_state import DeviceClickedItemsState
    from chameleon_recsys_b200.harness import make_problem
    pb = make_problem('tiny', profile='B')
    V = pb.plan.num_items
    host = ClickedItemsState(0.01, 300, 50, V)                      # 36 s window, 300 rows: cut-off and clip both bite
    dev = DeviceClickedItemsState(host)
    it = pb.input_fn()
    for step in range(25):
        f, l = it.get_next()
        all_items = np.concatenate([f['item_clicked'], l['label_last_item']], axis=1)
        host.update_from_batch(f['item_clicked'], f['event_timestamp'], l['label_last_item'])
        dev.update(torch.from_numpy(all_items).cuda(), torch.from_numpy(np.ascontiguousarray(f['event_timestamp'])).cuda(),
                   has_clicks=bool(all_items.any()))
        assert np.array_equal(dev.buffer_ids().cpu().numpy(), host.get_recent_clicks_buffer()), step
        assert np.array_equal(dev.articles_recent_pop_norm().cpu().numpy(),
                              host.get_articles_recent_pop_norm().astype(np.float32)), step
    back = dev.to_host(ClickedItemsState(0.01, 300, 50, V))
    assert np.array_equal(back.pop_recent_clicks_buffer, host.pop_recent_clicks_buffer)
    assert np.array_equal(back.get_articles_recent_pop_norm(), host.get_articles_recent_pop_norm())   # float64, bit-exact
    assert np.array_equal(back.get_articles_pop(), host.get_articles_pop())
    assert np.array_equal(back.get_articles_recent_pop(), host.get_articles_recent_pop())
