"""Device engine of the NAR hot path: owns the HBM-resident state (weights, Adam slots, ACR
table, metadata) and runs one training / evaluation step as a sequence of libnar_b200 kernels
on torch's current CUDA stream.  torch = allocator + streams + NCCL plumbing only.

Step order follows the reference graph (nar_module/nar/nar_model.py):
  sampler (:265-276) -> features (:314-370) -> CAR (:374-405) -> RNN (:408, :1308-1342) ->
  FC1/FC2 (:410-438) -> scorer (:444-517) -> loss (:639-704) -> Adam (:706-722)
with one structural difference: only the valid positions (mask == 1) are materialised.
Padded positions never reach the loss (:660-664), so skipping them changes no output.

HBM layout (per step; L = valid positions, n_cand = 1+K, R = L + L*n_cand rows):
  X   [R, Fp]   feature rows: inputs [0,L), then per position: positive, K negatives
  H1  [R, C]    leaky(X W1 + b1)         E [R, C]  tanh(H1 W2 + b2)  (CAR embeddings)
  GX  [L, 2Hp]  x Wx + b per RNN layer   HO/GT/CD [L, Hp] state / gate / candidate
  F1  [L, 512]  PR [L, C] predicted embedding
  PD  [Rc, C]   cand * pred   Z1 [Rc,128] Z2 [Rc,64] Z3 [Rc,32]   logits [L, n_cand]
Backward: with the auxiliary stream (default) the gradients dH1 / dX / dPD have their own buffers, because the weight
gradients that read H1 / X / PD run concurrently with the dgrad chain; the single-stream schedule (NAR_AUX_STREAM=0)
reuses the forward buffers in place (dH1 over H1, dX over X, dprod over PD).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import ops
from ._lib import ACT_LEAKY, ACT_NONE, ACT_TANH, FeaturePlanC, NarError
from .dp import shard_sessions
from .plan import (SEG_ACR, SEG_CTX_EMBED, SEG_ITEM_EMB, SEG_META_EMBED, FeaturePlan, ParamLayout, round_up)


class StepBatch:
    """Host-side description of one step (numpy views) + device staging."""
    pass


class NarEngine:
    def __init__(self, plan: FeaturePlan, layout: ParamLayout, content_article_embeddings_matrix: np.ndarray,
                 articles_metadata: Dict[str, np.ndarray], *, negative_samples: int, negative_sample_from_buffer: int,
                 softmax_temperature: float, reg_weight_decay: float, lr: float,
                 recent_clicks_buffer_max_size: int, recent_clicks_for_normalization: int,
                 elapsed_days_smooth_log_base: float = 1.3, popularity_smooth_log_base: float = 2.0,
                 ranking: str = 'mlp', rnn_cell: str = 'ugrnn', sampler_seed: int = 42, device: Optional[int] = None,
                 fwd_precision: int = 3, bwd_precision: int = 1, process_group=None, max_batch: int = 0):
        if not torch.cuda.is_available():
            raise NarError('NarEngine needs a CUDA (sm_100a) device; there is no CPU fallback')
        if rnn_cell != 'ugrnn':
            raise NotImplementedError("rnn_cell=%r: only the reference's UGRNNCell is implemented" % rnn_cell)
        if ranking not in ('mlp', 'cosine'):
            raise ValueError(ranking)
        self.dev = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        self.plan, self.layout = plan, layout
        self.K = int(negative_samples)
        self.n_from_buffer = int(negative_sample_from_buffer)
        self.tau = float(softmax_temperature)
        self.reg = float(reg_weight_decay)
        self.lr = float(lr)
        self.buf_len = int(recent_clicks_buffer_max_size)
        self.n_norm = int(recent_clicks_for_normalization)
        self.lb_rec, self.lb_nov = float(elapsed_days_smooth_log_base), float(popularity_smooth_log_base)
        self.ranking = ranking
        self.seed = int(sampler_seed)
        self.fwd_prec, self.bwd_prec = int(fwd_precision), int(bwd_precision)
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.rank = torch.distributed.get_rank(process_group) if process_group is not None else 0
        self.C, self.H, self.Hp, self.layers = layout.C, layout.H, layout.Hp, layout.layers
        self.V = plan.num_items
        d = self.dev
        # ---- resident tables
        acr = np.zeros((self.V, plan.acr_ld), dtype=np.float32)
        acr[:, :plan.acr_dim] = np.asarray(content_article_embeddings_matrix, dtype=np.float32)
        self.acr = torch.from_numpy(acr).to(d)
        self.created_at = torch.from_numpy(np.asarray(articles_metadata['created_at_ts'], dtype=np.int64)).to(d)
        self.meta = [torch.from_numpy(np.asarray(articles_metadata[n], dtype=np.int64)).to(d) for n in plan.meta_names]
        # ---- parameters (flat fp32 buffers; same offsets for grads / Adam slots)
        n = layout.total
        self.params = torch.zeros(n, device=d)
        self.grads = torch.zeros(n, device=d)
        self.adam_m = torch.zeros(n, device=d)
        self.adam_v = torch.zeros(n, device=d)
        self.params_lo = torch.zeros(n, device=d)      # w - tf32_trunc(w): B_lo plane of the 3xTF32 forward GEMMs
        self.global_step = 0
        self.WhT = [torch.zeros(2 * self.Hp, self.Hp, device=d) for _ in range(self.layers)]
        self.stats = torch.zeros(24, device=d)
        self.loss_dev = torch.zeros(4, device=d)          # [xe_sum, reg, -, -]
        self.loss_host = torch.zeros(4).pin_memory()
        self._loss_hosts = [self.loss_host, torch.zeros(4).pin_memory()]    # two in flight: submit(n+1) before result(n)
        self._loss_slot = 0
        self._bufs: Dict[str, torch.Tensor] = {}
        self._pinned: Dict[str, torch.Tensor] = {}
        self._sampler_ws = None
        self._planc_static = None
        self._side = None
        self._prep_flip = 0
        self.use_side_stream = os.environ.get('NAR_SIDE_STREAM', '1') == '1'
        self._slot_events: Dict[str, torch.cuda.Event] = {}       # prepare() slot -> end of the last step that read it
        self._pin_events: Dict[str, torch.cuda.Event] = {}        # staging slot -> its last H2D copy
        self._aux = None
        self.use_aux_stream = os.environ.get('NAR_AUX_STREAM', '1') == '1'
        self.split_fwd = os.environ.get('NAR_SPLIT_FWD', '1') == '1'     # session branch under the candidate CAR GEMMs
        self.split_bwd = os.environ.get('NAR_SPLIT_BWD', '0') == '1'     # measured slower (1.77 vs 1.63 ms): off
        self._views: dict = {}
        self.last: Dict[str, torch.Tensor] = {}
        self.ops = ops
        ops.context(self.dev.index)     # fail loudly here if the library / device is unusable

    # ------------------------------------------------------------------ parameters
    def view(self, key: str, buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        b = self.params if buf is None else buf
        ck = (key, b.data_ptr())
        v = self._views.get(ck)
        if v is None:
            t = self.layout.by_key[key]
            v = b[t.offset:t.offset + t.size].view(t.rows, t.ld)
            self._views[ck] = v
        return v

    def set_params(self, logical: Dict[str, np.ndarray]):
        flat = self.layout.to_internal(logical)
        self.params.copy_(torch.from_numpy(flat))
        self.adam_m.zero_(); self.adam_v.zero_(); self.grads.zero_()
        ops.tf32_lo(self.params, self.layout.total, self.params_lo)
        self.global_step = 0

    def get_params(self) -> Dict[str, np.ndarray]:
        return self.layout.to_logical(self.params.detach().cpu().numpy())

    def get_grads(self) -> Dict[str, np.ndarray]:
        return self.layout.to_logical(self.grads.detach().cpu().numpy())

    def state_dict(self) -> dict:
        return {'params': self.layout.to_logical(self.params.cpu().numpy()),
                'adam_m': self.layout.to_logical(self.adam_m.cpu().numpy()),
                'adam_v': self.layout.to_logical(self.adam_v.cpu().numpy()),
                'global_step': self.global_step}

    def load_logical_state(self, params, adam_m, adam_v, global_step: int):
        """Set weights + Adam slots from logical (TF-shaped) dicts, e.g. to start a parity step from a given state."""
        self.load_state_dict({'params': params, 'adam_m': adam_m, 'adam_v': adam_v, 'global_step': global_step})

    def load_state_dict(self, sd: dict):
        self.params.copy_(torch.from_numpy(self.layout.to_internal(sd['params'])))
        self.adam_m.copy_(torch.from_numpy(self.layout.to_internal(sd['adam_m'])))
        self.adam_v.copy_(torch.from_numpy(self.layout.to_internal(sd['adam_v'])))
        self.global_step = int(sd['global_step'])
        ops.tf32_lo(self.params, self.layout.total, self.params_lo)

    # ------------------------------------------------------------------ buffers
    def _buf(self, name: str, rows: int, cols: int, dtype=torch.float32, cap_rows: int = 0) -> torch.Tensor:
        """Named device buffer.  ``cap_rows`` = the most rows this buffer can ever need (every position of every session
        valid): allocated once at that size - a (re)allocation inside the training loop is a device-wide sync, and
        with 180 GB of HBM the worst case of the reference configurations (a few GB) is cheap."""
        need = max(1, rows) * cols
        t = self._bufs.get(name)
        if t is None or t.numel() < need or t.dtype != dtype:
            cap = int(need * 1.25) + 1024
            worst = max(1, cap_rows) * cols
            # worst-case sizing only while it is cheap: the stress configuration (8192 sessions x 500 negatives) would
            # ask for a terabyte; beyond 8 GiB per buffer the 1.25x growth policy applies instead
            if worst * torch.empty((), dtype=dtype).element_size() <= (8 << 30):
                cap = max(cap, worst)
            t = torch.empty(cap, device=self.dev, dtype=dtype)
            self._bufs[name] = t
        return t[:max(1, rows) * cols].view(max(1, rows), cols)

    def _pin(self, name: str, nbytes: int) -> torch.Tensor:
        t = self._pinned.get(name)
        if t is None or t.numel() < nbytes:
            t = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8).pin_memory()
            self._pinned[name] = t
        return t

    # ------------------------------------------------------------------ staging (host -> HBM, one copy)
    def stage(self, features: Dict[str, np.ndarray], labels: Dict[str, np.ndarray], buffer: np.ndarray,
              pop_norm: np.ndarray, slot: str = 'stage', stream: Optional[torch.cuda.Stream] = None) -> dict:
        """Pack the step inputs into one pinned buffer and issue one async H2D copy.
        ``features``/``labels`` hold the GLOBAL batch (all data-parallel ranks see the same arrays)."""
        item_clicked = np.ascontiguousarray(features['item_clicked'], dtype=np.int64)
        Bg, T = item_clicked.shape
        # this rank's sessions + compact valid positions (session-major; flat index into the GLOBAL [Bg*T] arrays)
        sh = shard_sessions(np.asarray(features['session_size']), T, self.world, self.rank)
        s0, per, lens, L, L_global = sh['s0'], sh['per'], sh['lens'], sh['L'], sh['L_global']
        sess_off, pos_idx = sh['sess_off'], sh['pos_idx']
        all_items = np.concatenate([item_clicked, np.asarray(labels['label_last_item'], dtype=np.int64).reshape(Bg, 1)], axis=1)
        ev = np.ascontiguousarray(features['event_timestamp'], dtype=np.int64)
        parts = [('all_items', all_items, np.int64), ('event_ts', ev, np.int64),
                 ('item_clicked', item_clicked, np.int64),
                 ('label_next', np.ascontiguousarray(labels['label_next_item'], dtype=np.int64), np.int64),
                 ('buffer', np.ascontiguousarray(buffer, dtype=np.int64), np.int64),
                 ('max_ts', np.asarray([ev.max() if ev.size else 0], dtype=np.int64), np.int64)]
        for name in self.plan.ctx_int_names:
            parts.append(('ci/' + name, np.ascontiguousarray(features[name], dtype=np.int64), np.int64))
        parts.append(('pop_norm', np.ascontiguousarray(pop_norm, dtype=np.float32), np.float32))
        for name in self.plan.ctx_float_names:
            parts.append(('cf/' + name, np.ascontiguousarray(features[name], dtype=np.float32), np.float32))
        parts.append(('pos_idx', pos_idx, np.int32))
        parts.append(('sess_off', sess_off, np.int32))
        offs, off = {}, 0
        for name, arr, dt in parts:
            off = round_up(off, 16)
            offs[name] = (off, arr.size, dt, arr.shape)
            off += arr.size * np.dtype(dt).itemsize
        total = round_up(off, 16)
        worst = total + 4 * (per * T - pos_idx.size) + 64       # pos_idx is the only part whose size varies step to step
        pin = self._pin(slot, worst)
        busy = self._pin_events.get(slot)
        if busy is not None:
            busy.synchronize()                # the previous H2D copy out of this pinned slot (issued >= 2 steps ago) is done
        pin_np = pin.numpy()
        for name, arr, dt in parts:
            o, nel, _, _ = offs[name]
            pin_np[o:o + nel * np.dtype(dt).itemsize].view(dt)[:] = arr.reshape(-1)
        dev = self._buf(slot, total, 1, torch.uint8, cap_rows=worst).view(-1)
        copy_stream = stream if stream is not None else torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            dev[:total].copy_(pin[:total], non_blocking=True)
        copied = torch.cuda.Event()
        copied.record(copy_stream)
        self._pin_events[slot] = copied
        tmap = {np.int64: torch.int64, np.float32: torch.float32, np.int32: torch.int32}
        tens = {}
        for name, (o, nel, dt, shp) in offs.items():
            tens[name] = dev[o:o + nel * np.dtype(dt).itemsize].view(tmap[dt]).view(*shp) if nel > 0 else \
                torch.zeros(shp, dtype=tmap[dt], device=self.dev)
        return {'t': tens, 'Bg': Bg, 'B': per, 'T': T, 'L': L, 'L_global': L_global, 's0': s0,
                'h2d_bytes': total, 'lens': lens, 'slot': slot}

    # ------------------------------------------------------------------ feature plan for this step
    def _plan_c(self, st: dict) -> FeaturePlanC:
        """Per-step plan = cached static part (tables, column map) + this step's staged input pointers."""
        if self._planc_static is None:
            self._planc_static = bytes(self._plan_c_static())
        p = FeaturePlanC.from_buffer_copy(self._planc_static)
        t = st['t']
        for i, n in enumerate(self.plan.ctx_int_names):
            p.ctx_int[i] = t['ci/' + n].data_ptr()
        for i, n in enumerate(self.plan.ctx_float_names):
            p.ctx_float[i] = t['cf/' + n].data_ptr()
        p.pop_norm = t['pop_norm'].data_ptr()
        p.stats = st['stats'].data_ptr()
        return p

    def _plan_c_static(self) -> FeaturePlanC:
        p = FeaturePlanC()
        pl = self.plan
        p.n_segments = len(pl.segments)
        p.row_ld = pl.Fp
        for i, s in enumerate(pl.segments):
            sg = p.seg[i]
            sg.kind, sg.col, sg.width, sg.card, sg.src = s.kind, s.int_col, s.width, s.card, s.src
            sg.ld, sg.table, sg.grad = 0, None, None
            if s.kind == SEG_ACR:
                sg.ld, sg.table = pl.acr_ld, self.acr.data_ptr()
            elif s.kind in (SEG_ITEM_EMB, SEG_CTX_EMBED, SEG_META_EMBED):
                pt = self.layout.by_key[s.param]
                sg.ld = pt.ld
                sg.table = self.params.data_ptr() + 4 * pt.offset
                sg.grad = self.grads.data_ptr() + 4 * pt.offset
        for i, m in enumerate(self.meta):
            p.meta[i] = m.data_ptr()
        p.created_at_ts = self.created_at.data_ptr()
        p.gamma = self.view('gamma').data_ptr()
        p.beta = self.view('beta').data_ptr()
        p.stats = self.stats.data_ptr()
        p.log_base_recency, p.log_base_novelty = self.lb_rec, self.lb_nov
        # column -> segment map and the column ranges outside the wide (vector-copied) segments
        if pl.Fp > len(p.col_seg):
            raise NarError('feature rows wider than NAR_MAX_COLS')
        cs = np.full(len(p.col_seg), 255, dtype=np.uint8)
        wide = np.zeros(pl.Fp, dtype=bool)
        for i, s in enumerate(pl.segments):
            cs[s.int_col:s.int_col + s.width] = i
            if s.kind in (SEG_ACR, SEG_ITEM_EMB):
                wide[s.int_col:s.int_col + s.width] = True
        C.memmove(p.col_seg, cs.ctypes.data, len(p.col_seg))
        edges = np.flatnonzero(np.diff(np.concatenate([[True], wide, [True]]).astype(np.int8)))
        ranges = list(zip(edges[0::2], edges[1::2]))
        if len(ranges) > 4:
            raise NarError('more than 4 narrow column ranges')
        p.n_narrow = len(ranges)
        for i, (a, b) in enumerate(ranges):
            p.narrow_begin[i], p.narrow_end[i] = int(a), int(b)
        return p

    # ------------------------------------------------------------------ GEMM helpers
    def _fwd(self, X, Wkey, bkey, Y, M, act, K=None, N=None):
        W = self.view(Wkey)
        K = W.shape[0] if K is None else K
        N = W.shape[1] if N is None else N
        ops.gemm(X, W, Y, M, N, K, a_kmajor=True, b_kmajor=False, bias=self.view(bkey).view(-1) if bkey else None,
                 act=act, precision=self.fwd_prec,
                 b_lo=self.view(Wkey, self.params_lo) if self.fwd_prec == 3 else None)

    def _dgrad(self, dY, Wkey, dX, M, dact=ACT_NONE, aux=None, N_out=None, K_in=None):
        """dX[M, in] = dY[M, out] * W^T  (W stored [in, out]) optionally times act'(aux)."""
        W = self.view(Wkey)
        n_in = W.shape[0] if N_out is None else N_out
        n_out = W.shape[1] if K_in is None else K_in
        ops.gemm(dY, W, dX, M, n_in, n_out, a_kmajor=True, b_kmajor=True, dact=dact, aux=aux, precision=self.bwd_prec)

    def _split(self, M, N, Kred) -> int:
        tiles = ((M + 127) // 128) * ((N + 127) // 128)
        kt = (Kred + 31) // 32
        s = max(1, min((2 * 148 + tiles - 1) // tiles, kt // 4 if kt >= 4 else 1))
        return s

    def _wgrad(self, X, dY, Wkey, rows, n_in=None, n_out=None):
        """dW[in, out] += X[rows, in]^T * dY[rows, out]  (split-K, atomics into the grad buffer)."""
        dW = self.view(Wkey, self.grads)
        n_in = dW.shape[0] if n_in is None else n_in
        n_out = dW.shape[1] if n_out is None else n_out
        ops.gemm(X, dY, dW, n_in, n_out, rows, a_kmajor=False, b_kmajor=False, accumulate=True,
                 split_k=0, precision=self.bwd_prec)      # 0 = library picks the split (about two waves of CTAs)

    def _on_aux(self, *calls, done_event: bool = False):
        """Weight / bias gradients are only needed by Adam, so they leave the critical path: each group is launched on
        an auxiliary stream behind an event recorded where its inputs became final, while the main stream continues
        with the dgrad chain.  Small wgrads (a handful of CTAs each) and the column sums then overlap the main
        stream's kernels instead of serialising with them.  Inputs of a deferred group are never overwritten later in
        the step (the in-place dgrads of the single-stream version write to their own buffers here)."""
        if not self.use_aux_stream:
            for fn in calls:
                fn()
            return None
        if self._aux is None:
            self._aux = torch.cuda.Stream(device=self.dev)
            self._aux_events = [torch.cuda.Event() for _ in range(32)]       # reused round-robin: no per-step creation
            self._aux_ev_i = 0
        ev = self._aux_events[self._aux_ev_i]
        self._aux_ev_i = (self._aux_ev_i + 1) % len(self._aux_events)
        ev.record()
        self._aux.wait_event(ev)
        with ops.on_stream(self._aux):                # only libnar launches inside: no torch op runs on the aux stream
            for fn in calls:
                fn()
        self._aux_dirty = True
        if done_event:                                # lets the main stream wait for THIS group, not the whole queue
            dev = self._aux_events[self._aux_ev_i]
            self._aux_ev_i = (self._aux_ev_i + 1) % len(self._aux_events)
            dev.record(self._aux)
            return dev
        return None

    def _join_aux(self):
        if self.use_aux_stream and self._aux is not None and getattr(self, '_aux_dirty', False):
            ev = self._aux_events[self._aux_ev_i]
            self._aux_ev_i = (self._aux_ev_i + 1) % len(self._aux_events)
            ev.record(self._aux)
            torch.cuda.current_stream().wait_event(ev)
            self._aux_dirty = False

    def _bgrad(self, dY, bkey, rows, cols):
        ops.colsum_add(dY, rows, cols, dY.stride(0), self.view(bkey, self.grads).view(-1))

    # ------------------------------------------------------------------ the step
    def prepare(self, st: dict, step_id: int, stream: Optional[torch.cuda.Stream] = None) -> dict:
        """Everything of a step that does not depend on the weights: negatives (nar_model.py:265-276), the row
        lists and the recency / novelty statistics.  May run one step AHEAD on a side stream (``stream``) while the
        previous step's GEMMs occupy the SMs - the reference's tf.data prefetch(1) gives the same look-ahead
        (datasets.py:142).  Results live in per-slot buffers and are handed over through a CUDA event."""
        t = st['t']
        B, Bg, T, L, s0 = st['B'], st['Bg'], st['T'], st['L'], st['s0']
        K = self.K
        n_cand = K + 1
        R = L + L * n_cand
        Rmax = B * T * (n_cand + 1)                       # every position of every local session valid
        # per-slot result buffers are only needed when this runs ahead of the step that is still executing
        if stream is not None:
            self._prep_flip ^= 1
        slot = ('/ahead%d' % self._prep_flip) if stream is not None else ''
        cur = torch.cuda.current_stream()
        run_on = stream if stream is not None else cur
        # two result slots alternate: the step that consumed this slot two prepare() calls ago may still be running (its
        # gather backward reads the row lists at the very end), so the side stream waits for that step's end first
        last_use = self._slot_events.get(slot) if stream is not None else None
        if last_use is not None:
            run_on.wait_event(last_use)
        with torch.cuda.stream(run_on):
            need = ops.sample_negatives_workspace(Bg, T + 1, self.buf_len, K)
            if self._sampler_ws is None or self._sampler_ws.numel() < need:
                self._sampler_ws = torch.zeros(need, dtype=torch.uint8, device=self.dev)
            neg = self._buf('neg' + slot, Bg * T, K, torch.int64)
            if B < Bg:
                neg.zero_()
            neg_local = neg.view(-1)[s0 * T * K:(s0 + B) * T * K].view(B, T, K)
            ops.sample_negatives(t['all_items'], s0, B, t['buffer'], K, self.n_from_buffer, self.seed, step_id, neg_local,
                                 self._sampler_ws)
            stats = self._buf('stats' + slot, 24, 1).view(-1)
            row_pos = self._buf('row_pos' + slot, R, 1, torch.int32, cap_rows=Rmax).view(-1)
            row_item = self._buf('row_item' + slot, R, 1, torch.int64, cap_rows=Rmax).view(-1)
            if L > 0:
                ops.build_rows(t['pos_idx'], L, t['item_clicked'], t['label_next'], neg, K, row_pos, row_item)
                ops.feature_stats(t['buffer'], self.n_norm, self.created_at, t['pop_norm'], t['max_ts'], self.lb_rec,
                                  self.lb_nov, row_pos, row_item, R, L, n_cand, t['event_ts'], stats)
            ev = None
            if stream is not None:
                ev = torch.cuda.Event()
                ev.record(run_on)
        st['prep'] = {'neg': neg, 'neg_local': neg_local, 'row_pos': row_pos, 'row_item': row_item, 'stats': stats,
                      'event': ev, 'step_id': step_id, 'slot': slot}
        return st

    def step(self, st: dict, train: bool = True, keep: bool = False) -> dict:
        """Run one step on staged inputs.  Returns device tensors (loss parts, logits, negatives)."""
        with ops.on_stream(torch.cuda.current_stream()):      # one stream lookup per step instead of one per launch
            out = self._step(st, train, keep)
        prep = st.get('prep')
        if prep is not None and prep.get('slot'):             # ran-ahead results: mark when this step is done with them
            ev = torch.cuda.Event()
            ev.record()
            self._slot_events[prep['slot']] = ev
        return out

    def _step(self, st: dict, train: bool, keep: bool) -> dict:
        t = st['t']
        B, Bg, T, L, s0 = st['B'], st['Bg'], st['T'], st['L'], st['s0']
        K, C_, Hp, Fp = self.K, self.C, self.Hp, self.plan.Fp
        n_cand = K + 1
        Rc = L * n_cand
        R = L + Rc
        Lmax = B * T                                      # worst case: every position of every local session valid
        Rcmax, Rmax = Lmax * n_cand, Lmax * (n_cand + 1)
        inv_count = 1.0 / max(1, st['L_global'])
        step_id = self.global_step + 1
        self.loss_dev.zero_()
        prep = st.get('prep')
        if prep is None or prep['step_id'] != step_id:
            prep = self.prepare(st, step_id)['prep']          # inline, on the current stream
        if prep['event'] is not None:
            torch.cuda.current_stream().wait_event(prep['event'])
        neg, neg_local, row_pos, row_item = prep['neg'], prep['neg_local'], prep['row_pos'], prep['row_item']
        st['stats'] = prep['stats']
        out = {'negatives': neg_local, 'L': L}
        if L == 0:
            out.update(loss=self.loss_dev, logits=None)
            return out
        planc = self._plan_c(st)
        X = self._buf('X', R, Fp, cap_rows=Rmax)
        ops.gather_features(planc, row_pos, row_item, R, L, n_cand, t['event_ts'], t['max_ts'], X)
        # ---- CAR (nar_model.py:374-405)
        H1 = self._buf('H1', R, C_, cap_rows=Rmax)
        E = self._buf('E', R, C_, cap_rows=Rmax)
        HO, GT, CD, GX = [], [], [], []
        F1 = self._buf('F1', L, 512, cap_rows=Lmax)
        PR = self._buf('PR', L, C_, cap_rows=Lmax)
        for i in range(self.layers):
            GX.append(self._buf('GX%d' % i, L, 2 * Hp, cap_rows=Lmax))
            HO.append(self._buf('HO%d' % i, L, Hp, cap_rows=Lmax)); GT.append(self._buf('GT%d' % i, L, Hp, cap_rows=Lmax))
            CD.append(self._buf('CD%d' % i, L, Hp, cap_rows=Lmax))

        def session_branch():
            # ---- RNN (nar_model.py:408, :1308-1342) and session representation (:410-438): rows [0, L) only
            rnn_in = E
            for i in range(self.layers):
                self._fwd(rnn_in, 'rnn%d/Wx' % i, 'rnn%d/b' % i, GX[i], L, ACT_NONE)
                ops.ugrnn_fwd(GX[i], self.view('rnn%d/Wh' % i), t['sess_off'], B, Hp, HO[i], GT[i], CD[i])
                rnn_in = HO[i]
            self._fwd(HO[-1], 'W3', 'b3', F1, L, ACT_LEAKY)
            self._fwd(F1, 'W4', 'b4', PR, L, ACT_TANH)

        if self.use_aux_stream and self.split_fwd and Rc > 0:
            # The session branch (a chain of small kernels on L rows) only needs the CAR embeddings of the INPUT rows:
            # those go first, then the branch runs on the auxiliary stream under the big CAR GEMMs of the candidate rows.
            self._fwd(X[:L], 'W1', 'b1', H1[:L], L, ACT_LEAKY)
            self._fwd(H1[:L], 'W2', 'b2', E[:L], L, ACT_TANH)
            self._on_aux(session_branch)
            self._fwd(X[L:], 'W1', 'b1', H1[L:], Rc, ACT_LEAKY)
            self._fwd(H1[L:], 'W2', 'b2', E[L:], Rc, ACT_TANH)
            self._join_aux()
        else:
            self._fwd(X, 'W1', 'b1', H1, R, ACT_LEAKY)
            self._fwd(H1, 'W2', 'b2', E, R, ACT_TANH)
            session_branch()
        # ---- scorer + loss (nar_model.py:444-517, :639-667)
        Ec = E[L:]
        logits = self._buf('logits', L, n_cand, cap_rows=Lmax)
        dE = self._buf('dE', R, C_, cap_rows=Rmax) if train else None
        dPR = self._buf('dPR', L, C_, cap_rows=Lmax) if train else None
        if self.ranking == 'mlp':
            PD = self._buf('PD', Rc, C_, cap_rows=Rcmax)
            Z1 = self._buf('Z1', Rc, 128, cap_rows=Rcmax); Z2 = self._buf('Z2', Rc, 64, cap_rows=Rcmax); Z3 = self._buf('Z3', Rc, 32, cap_rows=Rcmax)
            ops.mul_pred(Ec, PR, L, n_cand, C_, PD)
            self._fwd(PD, 'M1', 'c1', Z1, Rc, ACT_LEAKY)
            self._fwd(Z1, 'M2', 'c2', Z2, Rc, ACT_LEAKY)
            self._fwd(Z2, 'M3', 'c3', Z3, Rc, ACT_LEAKY)
            dZ3 = self._buf('dZ3', Rc, 32, cap_rows=Rcmax) if train else None
            m4 = self.view('M4'); c4 = self.view('c4')
            ops.score_softmax_ce(Z3, 32, 32, m4, m4.stride(0), c4, L, n_cand, 1.0 / self.tau, inv_count, logits,
                                 self.loss_dev[0:1], dZ3, self.view('M4', self.grads) if train else None,
                                 self.view('c4', self.grads) if train else None)
        else:
            ops.cosine_softmax_ce(Ec, PR, L, n_cand, C_, 1.0 / self.tau, inv_count, logits, self.loss_dev[0:1],
                                  dE[L:] if train else None, dPR)
        if self.reg > 0.0:
            # every rank holds the same weights: add the regulariser once (rank 0) so that the sum-allreduce is exact
            if self.rank == 0:
                ops.l2_loss_add(self.params, self.layout.reg_end, self.reg, self.loss_dev[1:2])
        out.update(loss=self.loss_dev, logits=logits)
        if keep:
            # X and H1 are overwritten in place by the backward pass: keep copies for the parity tests
            self.last = dict(X=X.clone(), H1=H1.clone(), E=E, HO=HO, F1=F1, PR=PR, logits=logits, row_pos=row_pos,
                             row_item=row_item, stats=st['stats'].clone(), neg=neg_local)
        if not train:
            return out
        # =================================================================== backward
        inplace = not self.use_aux_stream          # single-stream version reuses H1 / X / PD for their gradients
        if self.ranking == 'mlp':
            dZ2 = self._buf('dZ2', Rc, 64, cap_rows=Rcmax); dZ1 = self._buf('dZ1', Rc, 128, cap_rows=Rcmax)
            self._on_aux(lambda: self._wgrad(Z2, dZ3, 'M3', Rc), lambda: self._bgrad(dZ3, 'c3', Rc, 32))
            self._dgrad(dZ3, 'M3', dZ2, Rc, dact=ACT_LEAKY, aux=Z2)
            self._on_aux(lambda: self._wgrad(Z1, dZ2, 'M2', Rc), lambda: self._bgrad(dZ2, 'c2', Rc, 64))
            self._dgrad(dZ2, 'M2', dZ1, Rc, dact=ACT_LEAKY, aux=Z1)
            self._on_aux(lambda: self._wgrad(PD, dZ1, 'M1', Rc), lambda: self._bgrad(dZ1, 'c1', Rc, 128))
            dPD = PD if inplace else self._buf('dPD', Rc, C_, cap_rows=Rcmax)
            self._dgrad(dZ1, 'M1', dPD, Rc)                      # d(prod); over PD in place when the wgrad is stream-ordered
            ops.mul_pred_bwd(dPD, Ec, PR, L, n_cand, C_, dE[L:], dPR, cand_act=ACT_TANH)   # candidate rows: through the CAR tanh
        else:
            ops.act_bwd(dE[L:], Ec, Rc * C_, ACT_TANH, dE[L:])
        dH1 = H1 if inplace else self._buf('dH1', R, C_, cap_rows=Rmax)
        dX = X if inplace else self._buf('dX', R, Fp, cap_rows=Rmax)

        def session_backward(deferred):
            """FC2 / FC1 (nar_model.py:410-426) -> BPTT -> d(E) of the input rows.  ``deferred(*calls)`` runs the weight /
            bias gradients: on the auxiliary stream (single-branch schedule) or inline (when this whole branch already
            runs there)."""
            ops.act_bwd(dPR, PR, L * C_, ACT_TANH, dPR)
            deferred(lambda: self._wgrad(F1, dPR, 'W4', L), lambda: self._bgrad(dPR, 'b4', L, C_))
            dF1 = self._buf('dF1', L, 512, cap_rows=Lmax)
            self._dgrad(dPR, 'W4', dF1, L, dact=ACT_LEAKY, aux=F1)
            ho_last = HO[-1]
            deferred(lambda: self._wgrad(ho_last, dF1, 'W3', L), lambda: self._bgrad(dF1, 'b3', L, 512))
            dHO = self._buf('dHO', L, Hp, cap_rows=Lmax)
            self._dgrad(dF1, 'W3', dHO, L)
            for i in reversed(range(self.layers)):
                Wh = self.view('rnn%d/Wh' % i)
                ops.transpose(Wh, Hp, 2 * Hp, 2 * Hp, self.WhT[i], Hp)
                dGX = self._buf('dGX%d' % i, L, 2 * Hp, cap_rows=Lmax); HPV = self._buf('HPV%d' % i, L, Hp, cap_rows=Lmax)
                ops.ugrnn_bwd(dHO, HO[i], GT[i], CD[i], self.WhT[i], t['sess_off'], B, Hp, dGX, HPV)
                x_in = E if i == 0 else HO[i - 1]
                deferred(lambda x_in=x_in, dGX=dGX, i=i: self._wgrad(x_in, dGX, 'rnn%d/Wx' % i, L),
                         lambda HPV=HPV, dGX=dGX, i=i: self._wgrad(HPV, dGX, 'rnn%d/Wh' % i, L),
                         lambda dGX=dGX, i=i: self._bgrad(dGX, 'rnn%d/b' % i, L, 2 * Hp))
                if i == 0:
                    self._dgrad(dGX, 'rnn0/Wx', dE, L, dact=ACT_TANH, aux=E)      # input rows of dE (pre-tanh)
                else:
                    dprev = self._buf('dHO_b%d' % i, L, Hp, cap_rows=Lmax)
                    self._dgrad(dGX, 'rnn%d/Wx' % i, dprev, L)
                    dHO = dprev

        def car_backward(lo, n, deferred):
            """CAR backward (shared weights) for rows [lo, lo+n)."""
            h1, de, dh1, x, dx = H1[lo:lo + n], dE[lo:lo + n], dH1[lo:lo + n], X[lo:lo + n], dX[lo:lo + n]
            deferred(lambda: self._wgrad(h1, de, 'W2', n), lambda: self._bgrad(de, 'b2', n, C_))
            self._dgrad(de, 'W2', dh1, n, dact=ACT_LEAKY, aux=h1)      # dH1(pre); over H1 in place when stream-ordered
            deferred(lambda: self._wgrad(x, dh1, 'W1', n), lambda: self._bgrad(dh1, 'b1', n, C_))
            self._dgrad(dh1, 'W1', dx, n)                               # dX; over X in place when stream-ordered

        def inline(*calls):
            for fn in calls:
                fn()

        if self.use_aux_stream and self.split_bwd and Rc > 0:
            # two branches: the candidate rows' CAR backward (big GEMMs, main stream) does not depend on the session
            # branch (FC -> BPTT -> input rows' CAR backward: small kernels), which runs on the auxiliary stream
            ev_rows = self._on_aux(lambda: session_backward(inline), lambda: car_backward(0, L, inline), done_event=True)
            car_backward(L, Rc, self._on_aux)
            torch.cuda.current_stream().wait_event(ev_rows)             # dX[:L] is final
        else:
            session_backward(self._on_aux)
            car_backward(0, R, self._on_aux)
        ops.gather_features_bwd(planc, row_pos, row_item, R, L, n_cand, t['event_ts'], t['max_ts'], dX,
                                self.view('gamma', self.grads).view(-1), self.view('beta', self.grads).view(-1))
        self._join_aux()
        return out

    def apply_gradients(self):
        """NCCL sum-allreduce of the flat gradient buffer (data parallel), then TF-Adam."""
        if self.world > 1:
            torch.distributed.all_reduce(self.grads, group=self.pg)
        self.global_step += 1
        ops.adam_tf(self.params, self.grads, self.adam_m, self.adam_v, self.layout.total, self.layout.reg_end,
                    self.reg, self.lr, self.global_step, params_lo=self.params_lo)

    # ---- pipelined interface: submit step n, overlap staging + prepare of step n+1 (side stream), then result(n)
    def side_stream(self) -> torch.cuda.Stream:
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        return self._side

    def stage_ahead(self, features, labels, buffer, pop_norm, slot: str, after: Optional[torch.cuda.Event] = None) -> dict:
        """Stage the NEXT step while the current one runs: the host packs the batch into the slot's pinned buffer; the
        H2D copy and the weight-independent front (sampler, row lists, statistics) run on a side stream next to the
        current step's GEMMs (measured on B200: 2.26 vs 2.35 ms per step).  NAR_SIDE_STREAM=0 queues the copy behind
        the running step on the main stream instead and leaves the front inline."""
        if self.use_side_stream:
            side = self.side_stream()
            if after is not None:
                side.wait_event(after)        # the step that last read this slot's buffers (two slots alternate) is done
            st = self.stage(features, labels, buffer, pop_norm, slot=slot, stream=side)
            return self.prepare(st, self.global_step + 1, stream=side)
        return self.stage(features, labels, buffer, pop_norm, slot=slot)

    def submit(self, st: dict, keep: bool = False) -> dict:
        """Queue one training step; nothing here waits for the GPU.  ``result(out)`` later waits for THIS step only
        (event), so the caller may queue step n+1 before reading the loss of step n - no bubble between steps."""
        self.grads.zero_()
        out = self.step(st, train=True, keep=keep)
        if st['L'] > 0 or self.world > 1:
            self.apply_gradients()
        if self.world > 1:
            torch.distributed.all_reduce(self.loss_dev, group=self.pg)
        self._loss_slot ^= 1
        host = self._loss_hosts[self._loss_slot]
        host.copy_(self.loss_dev, non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        out['stage'], out['loss_host'], out['done'] = st, host, done
        return out

    def result(self, out: dict) -> dict:
        out['done'].synchronize()
        host = out['loss_host']
        out['xe_loss'] = float(host[0]); out['reg_loss'] = float(host[1])
        out['total_loss'] = out['xe_loss'] + out['reg_loss']
        return out

    # ---- evaluation (ModeKeys.EVAL): forward + ranking of the 1+K candidates + HR@n / MRR@n accumulators
    def share_params(self, other: 'NarEngine'):
        """Use ``other``'s weights (same ParamLayout) without a copy: what Estimator.evaluate does when it restores the
        training graph's variables into the evaluation graph (nar_trainer_gcom.py:523)."""
        if other.layout.total != self.layout.total:
            raise ValueError('parameter layouts differ')
        self.params, self.params_lo = other.params, other.params_lo
        self.global_step = other.global_step
        self._views = {}
        self._planc_static = None

    def eval_step(self, features, labels, buffer, pop_norm, top_n: int, metrics: Optional[torch.Tensor] = None,
                  step_id: Optional[int] = None, keep: bool = False) -> dict:
        """One evaluation batch: negatives with this engine's (eval) sampling hparams, forward, loss, then
        rank_items_by_predicted_prob (nar_model.py:777-795) and the streaming HR@n / MRR@n sums (:835-885).
        ``metrics`` [3] float64 device accumulator {hits, sum of reciprocal ranks, valid labels} (counts stay exact)."""
        st = self.stage(features, labels, buffer, pop_norm, slot='eval')
        if step_id is not None:
            self.prepare(st, step_id)
            st['prep']['step_id'] = self.global_step + 1          # step() checks the id it would use itself
        out = self.step(st, train=False, keep=keep)
        L, n_cand = st['L'], self.K + 1
        if metrics is None:
            metrics = torch.zeros(3, device=self.dev, dtype=torch.float64)
        assert metrics.dtype == torch.float64
        out['metrics'] = metrics
        if L > 0:
            prep = st['prep']
            pred_ids = self._buf('pred_ids', L, n_cand, torch.int64, cap_rows=st['B'] * st['T'])
            pred_probs = self._buf('pred_probs', L, n_cand, cap_rows=st['B'] * st['T'])
            ops.rank_candidates(out['logits'], prep['row_item'][L:], L, n_cand, int(top_n), pred_ids, pred_probs, metrics)
            out['predicted_item_ids'], out['predicted_item_probs'] = pred_ids, pred_probs
        if self.world > 1:
            torch.distributed.all_reduce(self.loss_dev, group=self.pg)
        self.loss_host.copy_(self.loss_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out['xe_loss'] = float(self.loss_host[0]); out['reg_loss'] = float(self.loss_host[1])
        out['total_loss'] = out['xe_loss'] + out['reg_loss']
        out['stage'] = st
        return out

    def train_step(self, features, labels, buffer, pop_norm, keep: bool = False, sync: bool = True) -> dict:
        st = self.stage(features, labels, buffer, pop_norm)
        self.grads.zero_()
        out = self.step(st, train=True, keep=keep)
        if st['L'] > 0 or self.world > 1:
            self.apply_gradients()
        if self.world > 1:
            torch.distributed.all_reduce(self.loss_dev, group=self.pg)
        self.loss_host.copy_(self.loss_dev, non_blocking=True)
        out['stage'] = st
        if sync:
            torch.cuda.current_stream().synchronize()
            out['xe_loss'] = float(self.loss_host[0]); out['reg_loss'] = float(self.loss_host[1])
            out['total_loss'] = out['xe_loss'] + out['reg_loss']
        return out
