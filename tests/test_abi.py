"""The C-ABI library builds, loads without a GPU and exports every symbol include/nar_b200.h declares;
the product path fails loudly (no CPU fallback) when there is no CUDA device."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'nar_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(nar_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from chameleon_recsys_b200 import _lib, build
    build.build_library()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert lib.nar_abi_version() == 2
    # the ctypes mirrors of the ABI structs have the C layout (a padding mismatch would corrupt every call)
    for which, cls in enumerate((_lib.FeaturePlanC, _lib.ModelCfg, _lib.StepIO, _lib.RowLayout, _lib.GemmEpilogue, _lib.Segment)):
        import ctypes
        assert lib.nar_abi_struct_size(which) == ctypes.sizeof(cls), (which, cls)
    assert lib.nar_status_string(-3)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from chameleon_recsys_b200 import _lib
    with pytest.raises(_lib.NarError):
        _lib.Context(0)
    from chameleon_recsys_b200.harness import make_problem
    from chameleon_recsys_b200.nar_model import NARModuleModel
    from chameleon_recsys_b200.hparams import ModeKeys
    pb = make_problem('tiny', profile='A')
    with pytest.raises(Exception):
        NARModuleModel(ModeKeys.TRAIN, None, None, pb.session_features_config, pb.articles_features_config, 64, 1e-3, 1.0,
                       10, 300, pb.content_article_embeddings_matrix, articles_metadata=pb.articles_metadata,
                       CAR_embedding_size=64, rnn_units=64, internal_features_config=pb.internal_features_config)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'chameleon_recsys_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert 'import oracle' not in src and 'from oracle' not in src, fn
