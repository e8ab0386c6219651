"""Clip-batch pipeline: host images -> StoSAVi.encode -> SlotRollouter -> host, software-pipelined on three streams.

The reference runs this chain batch after batch (base_slots/extract_slots.py:18-38 followed by
video_prediction/rollout_clevrer_slots.py:45-58).  Its three stages use the GPU very differently:

  copy    pinned host images -> device (PCIe)
  encode  cuDNN CNN + the fused encoder tail (sfb_enc_tail_forward): every SM, ~4 ms for 384 frames of 128 x 128
  loop    the serial frame loop (per frame: sfb_transition_forward + sfb_sa_forward, one CUDA-graph replay) and the
          rollout (sfb_rollout_forward), then the D2H copy of slots + predictions: latency-bound, most SMs idle

so consecutive batches are overlapped: the copy and the encode stage of batch i+1 run on their own streams under the
loop stage of batch i.  The encode stage of each of the two image buffers is captured into ONE CUDA graph (the host
issues ~20 launches per batch for it otherwise).  Per-batch results are those of ``savi.encode`` + ``rollouter``.
"""
import torch


class ClipPipeline:
    """``submit(h_img, h_slots, h_pred)`` enqueues one batch of clips (pinned host tensors: images [B, T, 3, H, W] in,
    slots [B, T, K, D] and predicted slots [B, pred_len, K, D] out) and returns at once; ``drain()`` waits for every
    batch submitted so far.  Inference only (call under ``torch.no_grad()``); the batch shape is fixed by the first
    ``submit`` (or ``warmup``)."""

    def __init__(self, savi, rollouter, pred_len, device=None, capture_encode=True):
        self.savi, self.ro, self.pred_len = savi, rollouter, int(pred_len)
        self.dev = torch.device(device) if device is not None else savi.device
        self.capture_encode = capture_encode
        self.s_copy = torch.cuda.Stream(self.dev)
        self.s_enc = torch.cuda.Stream(self.dev)
        self.s_loop = torch.cuda.Stream(self.dev, priority=-1)    # the latency-bound stage gets SMs as CNN CTAs retire
        self._d_img = None
        self._graphs = [None, None]
        self._n = 0
        self._pending = None          # (features, ready event, h_slots, h_pred) of the batch whose loop stage is due
        self._img_free = [None, None]     # encode stage that last read image buffer sl
        self._tiles_free = [None, None]   # loop stage that last read the features of buffer sl

    # ------------------------------------------------------------------ stages
    def _stage_copy(self, h_img, sl):
        with torch.cuda.stream(self.s_copy):
            if self._img_free[sl] is not None:
                self.s_copy.wait_event(self._img_free[sl])
            self._d_img[sl].copy_(h_img, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.s_copy)
        return ev

    def _stage_encode(self, sl, copied):
        with torch.cuda.stream(self.s_enc):
            self.s_enc.wait_event(copied)
            if self._graphs[sl] is not None:
                if self._tiles_free[sl] is not None:
                    self.s_enc.wait_event(self._tiles_free[sl])   # the loop stage two batches back has taken its features
                self._graphs[sl][0].replay()
                feats = self._graphs[sl][1]
            else:
                feats = self.savi.encode_features(self._d_img[sl])
                # allocated on this stream, read on the loop stream: the allocator must not hand the block out before then
                (feats if isinstance(feats, torch.Tensor) else feats.data).record_stream(self.s_loop)
            ev = torch.cuda.Event()
            ev.record(self.s_enc)
        self._img_free[sl] = ev
        return feats, ev

    def _stage_loop(self, feats, ready, h_slots, h_pred, sl):
        with torch.cuda.stream(self.s_loop):
            self.s_loop.wait_event(ready)
            self.savi._reset_rnn()
            _, slots, _ = self.savi.encode(None, feats=feats)
            pred = self.ro(slots, self.pred_len)
            h_slots.copy_(slots, non_blocking=True)
            h_pred.copy_(pred, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.s_loop)
        self._tiles_free[sl] = ev

    # ------------------------------------------------------------------ API
    def warmup(self, h_img):
        """Allocates the two device image buffers, runs both stages once per buffer and captures the encode stage."""
        cur = torch.cuda.current_stream(self.dev)
        self._d_img = [torch.empty(h_img.shape, dtype=torch.float32, device=self.dev) for _ in range(2)]
        for sl in range(2):
            self._d_img[sl].copy_(h_img, non_blocking=True)
            feats = self.savi.encode_features(self._d_img[sl])
            self.savi._reset_rnn()
            _, slots, _ = self.savi.encode(None, feats=feats)
            self.ro(slots, self.pred_len)
        torch.cuda.synchronize(self.dev)
        if self.capture_encode:
            for sl in range(2):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    ft = self.savi.encode_features(self._d_img[sl])
                self._graphs[sl] = (g, ft)
        for s in (self.s_copy, self.s_enc, self.s_loop):
            s.wait_stream(cur)

    def submit(self, h_img, h_slots, h_pred):
        if self._d_img is None:
            self.warmup(h_img)
        sl = self._n & 1
        self._n += 1
        copied = self._stage_copy(h_img, sl)
        feats, ready = self._stage_encode(sl, copied)
        if self._pending is not None:                 # the previous batch's loop stage runs under this batch's CNN
            self._stage_loop(*self._pending)
        self._pending = (feats, ready, h_slots, h_pred, sl)

    def drain(self):
        if self._pending is not None:
            self._stage_loop(*self._pending)
            self._pending = None
        torch.cuda.synchronize(self.dev)
