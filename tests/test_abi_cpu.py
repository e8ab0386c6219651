"""CPU-side checks of the C ABI: the library builds, loads, exports every symbol declared in
include/sfb200.h, and its argument validation works without a GPU (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from slotformer_b200.build import build_extension
    from slotformer_b200 import engine
    build_extension()
    return engine.load()


def _declared(debug):
    with open(os.path.join(ROOT, 'include', 'sfb200.h')) as f:
        text = f.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    if not debug:
        text = re.sub(r'#ifdef SFB_DEBUG.*?#endif', '', text, flags=re.S)
    return set(re.findall(r'\b(sfb_[a-z0-9_]+)\s*\(', text))


def test_exports_every_declared_symbol(lib):
    from slotformer_b200 import engine
    declared = _declared(debug=False)
    assert declared == set(engine.exported_symbols())
    for name in declared:
        assert getattr(lib, name) is not None
    # the product library carries no debug hooks and no process-wide knobs
    for name in ('sfb_debug_set_profile', 'sfb_debug_umma_gemm', 'sfb_sa_set_cta_limit'):
        assert not hasattr(lib, name)


def test_debug_library_exports_the_debug_hooks():
    from slotformer_b200.build import build_extension
    from slotformer_b200 import engine
    build_extension(debug=True)
    dbg = engine.load_debug()
    declared = _declared(debug=True)
    assert declared == set(engine.exported_symbols(debug=True))
    for name in declared:
        assert getattr(dbg, name) is not None


def test_version_and_strerror(lib):
    assert lib.sfb_version() == 210
    assert lib.sfb_strerror(0) == b'ok'
    assert b'shape' in lib.sfb_strerror(-1)
    assert b'aligned' in lib.sfb_strerror(-2)
    assert lib.sfb_launch_count() >= 0


def test_workspace_sizes(lib):
    # grows with the batch (q~, partials) and with the x^ ring of one frame chunk
    a = lib.sfb_sa_workspace_bytes(8, 4096, 128, 128, 256, 2, 0)
    b = lib.sfb_sa_workspace_bytes(16, 4096, 128, 128, 256, 2, 0)
    one_iter = lib.sfb_sa_workspace_bytes(8, 4096, 128, 128, 256, 1, 0)
    assert 0 < one_iter < a < b
    assert a - one_iter >= 8 * 4096 * 128 * 2          # fp16 x^ ring only when n_iter > 1
    assert lib.sfb_sa_workspace_bytes(0, 4096, 128, 128, 256, 2, 0) == 0
    # fp16 copies of in/out proj (hi and lo halves: three-term products) + per-layer qkv, out, ffn1, ffn2 (256-byte aligned), then one fp32 block of
    # biases + LayerNorm affine per layer (9 d + F floats)
    def up256(n):
        return (n + 255) // 256 * 256
    d, Ds, F, L = 128, 128, 512, 4
    want = up256((4 * d * Ds + L * (3 * d * d + d * d + 2 * F * d)) * 2) + L * (9 * d + F) * 4
    assert lib.sfb_rollout_workspace_bytes(Ds, d, F, L) == want
    # output-feature rows are padded to multiples of 128 (out_proj with Ds = 192 -> 256 rows)
    assert lib.sfb_rollout_workspace_bytes(192, 256, 1024, 8) == \
        up256((2 * (256 * 192 + 256 * 256) + 8 * (768 * 256 + 256 * 256 + 2 * 1024 * 256)) * 2) + 8 * (9 * 256 + 1024) * 4


def test_null_and_shape_validation_without_gpu(lib):
    from slotformer_b200.engine import _SAWeights, _ROWeights
    w = _SAWeights()
    assert lib.sfb_sa_forward(None, 0, 0, None, None, None, ctypes.byref(w), 1, 64, 128, 128,
                              256, 4, 1, 1e-6, 0, 0, 0, None, 0, None) == -5
    rw = _ROWeights()
    assert lib.sfb_rollout_forward(None, None, ctypes.byref(rw), 1, 1, 1, 128, 128, 512, 8, 1,
                                   0, 0, 0, None, 0, None) == -5
    assert lib.sfb_rollout_prepare(None, 128, 128, 512, None, 0, None) == -5


def test_engine_refuses_cpu_tensors():
    import torch
    from slotformer_b200.engine import SfbError, SlotAttentionEngine
    eng = SlotAttentionEngine()
    with pytest.raises(SfbError):
        eng.forward(torch.zeros(1, 64, 128), torch.zeros(1, 4, 128), {}, 1, 1e-6, 256)


def test_modules_keep_reference_state_dict_keys():
    import cases
    from helpers import golden, ro_module, sa_module
    c, w, _, _ = cases.sa_case('sa_tiny')
    m = sa_module(c, w, 'cpu')
    assert list(m.state_dict().keys()) == golden('sa_tiny')['keys'].tolist()
    for name in ('ro_tiny', 'ro_cfg5'):
        c, w, _ = cases.ro_case(name)
        g = golden(name)
        m = ro_module(c, w, 'cpu', enc_t_pe=g['enc_t_pe'])
        assert sorted(m.state_dict().keys()) == sorted(g['keys'].tolist())
        # the sinusoid table the module builds equals the reference's enc_t_pe
        fresh = ro_module(c, w, 'cpu')
        assert (fresh.enc_t_pe.numpy() == g['enc_t_pe']).all()


def test_invalidate_kernel_caches_resets_every_launcher():
    """in-place weight updates through ``.data`` do not change the (data_ptr, _version) cache key;
    ``slotformer_b200.invalidate_kernel_caches(model)`` is the explicit hook."""
    import slotformer_b200
    from slotformer_b200.engine import TransitionEngine
    from slotformer_b200.base_slots.models import StoSAVi
    import wrapper_cases as W
    m = W.build_savi(StoSAVi)
    m.slot_attention._engine._key = ('stale',)
    m.__dict__['_transition_engine'] = TransitionEngine()
    m._transition_engine._key = ('stale',)
    m.__dict__['_loop_graphs'] = {'k': object()}
    assert slotformer_b200.invalidate_kernel_caches(m) == 3
    assert m.slot_attention._engine._key is None and m._transition_engine._key is None
    assert '_loop_graphs' not in m.__dict__

