"""slotformer_b200.pipeline.ClipPipeline (host images -> StoSAVi.encode -> SlotRollouter -> host on three streams): every
batch gets the results of the plain calls, with and without the graph-captured encode stage."""
import numpy as np
import pytest
import torch

import wrapper_cases as W


@pytest.mark.gpu
@pytest.mark.parametrize('capture', [False, True])
def test_clip_pipeline_matches_plain_calls(capture):
    from slotformer_b200.base_slots.models import StoSAVi
    from slotformer_b200.pipeline import ClipPipeline
    from slotformer_b200.video_prediction.models import SlotRollouter
    dev = torch.device('cuda', 0)
    savi = W.build_savi(StoSAVi).to(dev)
    torch.manual_seed(5)
    ro = SlotRollouter(5, 128, 3, d_model=128, num_layers=2, num_heads=8, ffn_dim=256).to(dev).eval()
    rs = np.random.RandomState(21)
    batches = [torch.from_numpy(rs.uniform(-1, 1, size=(2, 3, 3, 64, 64)).astype(np.float32)).pin_memory() for _ in range(5)]
    with torch.no_grad():
        want = []
        for img in batches:
            savi._reset_rnn()
            _, slots, _ = savi.encode(img.to(dev))
            want.append((slots.cpu(), ro(slots, 4).cpu()))
        pipe = ClipPipeline(savi, ro, 4, dev, capture_encode=capture)
        outs = [(torch.empty((2, 3, 5, 128)).pin_memory(), torch.empty((2, 4, 5, 128)).pin_memory()) for _ in batches]
        for img, (hs, hp) in zip(batches, outs):
            pipe.submit(img, hs, hp)
        pipe.drain()
    for (ws, wp), (hs, hp) in zip(want, outs):
        assert torch.equal(ws, hs) and torch.equal(wp, hp)
