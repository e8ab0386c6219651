"""Encoder tail: + SoftPositionEmbed -> LayerNorm(64) -> Linear -> ReLU -> Linear -> Slot
Attention's LayerNorm statistics as ONE sm_100a kernel writing fp16 operand tiles (csrc/enc_tail.cu), against the
same chain in stock PyTorch fp32 (reference savi.py:367-377, :66) and through StoSAVi against the fp32 feature path.

Tolerances: t = LN(features) is stored as fp16 (unit variance, |t| of a few): 2e-3 of max|t|.  Slots extracted from
the tiles vs from the fp32 feature grid: 5e-4 relative (measured ~1e-4; fp16 GEMM operands in the tail)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import wrapper_cases as W
from helpers import rel_max
from slotformer_b200 import engine

DEV = 'cuda:0'


def test_enc_tail_abi_validates_without_gpu():
    import ctypes
    from slotformer_b200.build import build_extension
    build_extension()
    lib = engine.load()
    assert lib.sfb_enc_tail_workspace_bytes(128) == 16384 + 32768 + (256 + 256) * 4
    assert lib.sfb_enc_tail_workspace_bytes(192) == 0
    assert lib.sfb_enc_tail_tiles_bytes(3, 4096, 128) == 3 * 4096 * 128 * 2
    assert lib.sfb_enc_tail_tiles_bytes(1, 400, 128) == 512 * 128 * 2          # ragged: whole 128-pixel tiles of a chunk
    w = engine._EncTailWeights()
    assert lib.sfb_enc_tail_prepare(ctypes.byref(w), 128, None, 0, None) == -5
    assert lib.sfb_enc_tail_forward(None, 0, 1, 64, 64, 128, None, 0, None, 0, 0, 0, None) == -5


def _torch_tail(m, cnn):
    """the reference chain in stock PyTorch (fp32, no TF32): [F, 64, H, W] -> y [F, N, C] and t = LN(y) without affine"""
    x = cnn + m.encoder_pos_embedding.dense(m.encoder_pos_embedding.grid).permute(0, 3, 1, 2)
    y = m.encoder_out_layer(x.flatten(2, 3).permute(0, 2, 1).contiguous())
    return y, F.layer_norm(y, (y.shape[-1],))


@pytest.mark.gpu
def test_tiles_match_the_pytorch_chain():
    from slotformer_b200.base_slots.models import StoSAVi
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = W.build_savi(StoSAVi).to(DEV)
    gen = torch.Generator(device=DEV).manual_seed(4)
    cnn = torch.randn((5, 64, 64, 64), device=DEV, generator=gen) * 1.5 + 0.3
    with torch.no_grad():
        named = dict(m.named_parameters())
        tiles = engine.EncoderTailEngine().forward(cnn, {k: named[k] for k in engine.ENC_TAIL_KEYS}, 128)
        _, t_ref = _torch_tail(m, cnn)
    assert tiles.shape == (5, 4096, 128)
    got = tiles.to_dense()
    assert torch.isfinite(got).all()
    assert rel_max(got.cpu().numpy(), t_ref.cpu().numpy()) < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize('hw', [(64, 64), (20, 20), (9, 12)])
def test_channels_last_input_gives_the_same_tiles(hw):
    """SFB_ET_NHWC: the CNN output in torch.channels_last memory (what cuDNN's tensor-core convolutions write) is read
    as it is -- bit-identical tiles to the NCHW route (same per-pixel arithmetic, only the TMA boxes differ); also a
    channels-last slice with a frame stride (frames 1..3 of a larger buffer), ragged grids with zero tail rows."""
    from slotformer_b200.base_slots.models import StoSAVi
    from slotformer_b200.base_slots.models.utils import build_grid
    H, Wd = hw
    m = W.build_savi(StoSAVi).to(DEV)
    gen = torch.Generator(device=DEV).manual_seed(6)
    cnn = torch.randn((5, 64, H, Wd), device=DEV, generator=gen) * 1.3 - 0.2
    named = dict(m.named_parameters())
    wts = {k: named[k] for k in engine.ENC_TAIL_KEYS}
    eng = engine.EncoderTailEngine()
    with torch.no_grad():
        a = eng.forward(cnn, wts, 128)
        n0 = engine.launch_count()
        cl = cnn.contiguous(memory_format=torch.channels_last)
        assert not cl.is_contiguous()
        b = eng.forward(cl, wts, 128)
        c = eng.forward(cl[1:4], wts, 128)
        assert engine.launch_count() - n0 == 2                    # no layout copy in between
    assert torch.equal(a.data, b.data)
    assert torch.equal(a.data[1:4], c.data)


@pytest.mark.gpu
def test_tiles_ragged_grid_and_tail_tiles_are_zero():
    """A 20 x 20 grid (N = 400: three full tiles + a 16-pixel tail inside a 512-pixel chunk): pixels beyond N are zero
    rows, and Slot Attention on the tiles agrees with Slot Attention on the fp32 features."""
    from slotformer_b200.base_slots.models import StoSAVi
    from slotformer_b200.base_slots.models.utils import build_grid
    torch.backends.cuda.matmul.allow_tf32 = False
    m = W.build_savi(StoSAVi).to(DEV)
    H = Wd = 20
    gen = torch.Generator(device=DEV).manual_seed(5)
    cnn = torch.randn((3, 64, H, Wd), device=DEV, generator=gen)
    with torch.no_grad():
        named = dict(m.named_parameters())
        tiles = engine.EncoderTailEngine().forward(cnn, {k: named[k] for k in engine.ENC_TAIL_KEYS}, 128)
        grid = build_grid((H, Wd)).to(DEV)
        x = cnn + m.encoder_pos_embedding.dense(grid).permute(0, 3, 1, 2)
        y = m.encoder_out_layer(x.flatten(2, 3).permute(0, 2, 1).contiguous())
        t_ref = F.layer_norm(y, (128,))
        assert rel_max(tiles.to_dense().cpu().numpy(), t_ref.cpu().numpy()) < 2e-3
        raw = tiles.data.view(3, -1, 2, 128, 64)                  # frame, tile, panel, row, 64 halves
        assert raw.shape[1] == 4 and (raw[:, 3, :, 16:] == 0).all()      # rows 16.. of the last tile: pixels >= 400
        s0 = torch.randn((3, 5, 128), device=DEV, generator=gen)
        a = m.slot_attention(y, s0)
        b = m.slot_attention(tiles, s0)
    assert rel_max(b.cpu().numpy(), a.cpu().numpy()) < 5e-4


@pytest.mark.gpu
def test_stosavi_with_fused_tail_matches_the_feature_grid_path():
    """StoSAVi.encode in inference: fused tail (default) vs the stock layers + fp32 feature grid, eager and graphed."""
    from slotformer_b200.base_slots.models import StoSAVi
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = W.build_savi(StoSAVi).to(DEV)
    img = W.savi_input().to(DEV)
    with torch.no_grad():
        m.testing = True
        n0 = engine.launch_count()
        fused = m({'img': img})['post_slots']
        fused_launches = engine.launch_count() - n0
        m.use_cuda_graph = False
        fused_eager = m({'img': img})['post_slots']
        m.fuse_encoder_tail = False
        plain = m({'img': img})['post_slots']
    assert torch.equal(fused, fused_eager)
    assert rel_max(fused.cpu().numpy(), plain.cpu().numpy()) < 5e-4
    assert fused_launches > 0
