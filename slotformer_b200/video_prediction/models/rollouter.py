"""Autoregressive slot-Transformer rollouters backed by the persistent sm_100a kernel.

Drop-in for reference ``SlotRollouter`` (slotformer/video_prediction/models/slotformer.py:48-134)
and ``SingleStepSlotRollouter`` (.../single_step_slotformer.py:6-90): same constructor
arguments, same parameter names (``in_proj``, ``transformer_encoder.layers.<i>.*``,
``enc_t_pe``, ``out_proj``), same ``forward(x, pred_len)``.
"""
import torch
from torch import nn

from ...autograd import KernelForward, warn_once
from ...engine import RolloutEngine


def get_sin_pos_enc(seq_len, d_model):
    """Sinusoidal table [1, seq_len, d_model]; row 0 is the OLDEST frame (largest position),
    sin half then cos half (reference slotformer.py:10-16)."""
    freq = 1.0 / (10000 ** (torch.arange(0.0, d_model, 2.0) / d_model))
    pos = torch.arange(seq_len - 1, -1, -1).type_as(freq)
    ang = torch.outer(pos, freq)
    return torch.cat([ang.sin(), ang.cos()], dim=-1).unsqueeze(0)


def build_pos_enc(pos_enc, input_len, d_model):
    """'' -> None; 'learnable' -> zero-init trainable; 'sin*' -> frozen sinusoid."""
    if not pos_enc:
        return None
    if pos_enc == 'learnable':
        return nn.Parameter(torch.zeros(1, input_len, d_model))
    if 'sin' in pos_enc:
        return nn.Parameter(get_sin_pos_enc(input_len, d_model), requires_grad=False)
    raise NotImplementedError(f'unsupported pos enc {pos_enc}')


class Rollouter(nn.Module):
    """Interface of slot-dynamics predictors."""

    def forward(self, x):
        raise NotImplementedError

    def burnin(self, x):
        pass

    def reset(self):
        pass


class SlotRollouter(Rollouter):
    """Pre-LN Transformer encoder over a sliding window of history_len * num_slots tokens."""

    _mode = 'slide'

    def __init__(self, num_slots, slot_size, history_len, t_pe='sin', slots_pe='', d_model=128,
                 num_layers=4, num_heads=8, ffn_dim=512, norm_first=True):
        super().__init__()
        self.num_slots = num_slots
        self.history_len = history_len
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.norm_first = norm_first

        self.in_proj = nn.Linear(slot_size, d_model)
        layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=num_heads,
                                           dim_feedforward=ffn_dim, norm_first=norm_first,
                                           batch_first=True)
        self.transformer_encoder = nn.TransformerEncoder(encoder_layer=layer,
                                                         num_layers=num_layers,
                                                         enable_nested_tensor=False)
        self.enc_t_pe = build_pos_enc(t_pe, history_len, d_model)
        self.enc_slots_pe = build_pos_enc(slots_pe, num_slots, d_model)
        self.out_proj = nn.Linear(d_model, slot_size)
        self._engine = RolloutEngine()

    # -- helpers ---------------------------------------------------------- #
    def _token_pe(self):
        """[T*K, d] positional row per window token (time-major, slot-minor)."""
        srcs = [self.enc_t_pe] + ([self.enc_slots_pe] if self.enc_slots_pe is not None else [])
        key = tuple((t.data_ptr(), t._version, t.requires_grad and torch.is_grad_enabled()) for t in srcs)
        if getattr(self, '_pe_key', None) == key:
            return self._pe_cache
        pe = self._token_pe_uncached()
        if not any(k[2] for k in key):
            self._pe_key, self._pe_cache = key, pe
        return pe

    def _token_pe_uncached(self):
        T = self.enc_t_pe.shape[1]
        pe = self.enc_t_pe[0].unsqueeze(1).expand(T, self.num_slots, -1)
        if self.enc_slots_pe is not None:
            pe = pe + self.enc_slots_pe[0].unsqueeze(0)
        return pe.reshape(T * self.num_slots, -1).contiguous()

    def _weights(self):
        out = {}
        for name, p in self.named_parameters():
            if name.startswith(('in_proj.', 'out_proj.', 'transformer_encoder.')):
                out[name] = p.detach()
        return out

    def _needs_autograd(self, x):
        if not torch.is_grad_enabled():
            return False
        return x.requires_grad or any(p.requires_grad for p in self.parameters())

    def _window(self, in_x):
        return in_x, None

    def _autograd_forward(self, x, pred_len):
        """Differentiable restatement (GPU eager) used only when gradients are required."""
        K = self.num_slots
        in_x = x.flatten(1, 2)
        pe = self._token_pe().unsqueeze(0)
        preds = []
        for _ in range(pred_len):
            if self._mode == 'slide':
                win, win_pe = in_x, pe
            else:
                win = in_x[:, -self.num_cond_tokens:]
                win_pe = pe[:, -win.shape[1]:]
            h = self.transformer_encoder(self.in_proj(win) + win_pe)
            pred = self.out_proj(h[:, -K:])
            preds.append(pred)
            in_x = torch.cat([in_x[:, K:] if self._mode == 'slide' else in_x, pred], dim=1)
        return torch.stack(preds, dim=1)

    def forward(self, x, pred_len):
        """x [B, history_len, num_slots, slot_size] -> [B, pred_len, num_slots, slot_size]."""
        assert x.shape[1] == self.history_len, 'wrong burn-in steps'
        if not self.norm_first:
            raise NotImplementedError('the sm_100a engine implements the pre-LN encoder only')
        if self._needs_autograd(x):
            dropout = self.training and any(m.p > 0 for m in self.modules() if isinstance(m, nn.Dropout))
            if dropout or not x.is_cuda:
                # train() mode: nn.TransformerEncoderLayer's dropout (p = 0.1, slotformer.py:72-78) is part of the
                # reference's training forward and is not in the kernel -> the differentiable restatement runs
                if dropout and x.is_cuda:
                    warn_once('ro_dropout', 'SlotRollouter in train() mode with dropout > 0: the forward pass runs the '
                              'PyTorch restatement, not the sm_100a kernel (call .eval() or set dropout to 0 to train '
                              'on the kernel forward)')
                return self._autograd_forward(x, pred_len)
            # gradients required, no dropout: FORWARD on the kernel, backward through the restatement
            params = [p for n, p in self.named_parameters()
                      if n.startswith(('in_proj.', 'out_proj.', 'transformer_encoder.')) or n in ('enc_t_pe', 'enc_slots_pe')]
            return KernelForward.apply(lambda h: self._kernel(h, pred_len), lambda h: self._autograd_forward(h, pred_len),
                                       1, 1, x, *params)
        return self._kernel(x, pred_len)

    def _kernel(self, x, pred_len):
        return self._engine.forward(
            x.detach().float(), self._weights(), self._token_pe().detach().float(),
            self.num_layers, self.num_heads, pred_len, mode=self._mode,
            cond_len=getattr(self, 'cond_len', 0), flags=getattr(self, 'engine_flags', 0))

    @property
    def dtype(self):
        return self.in_proj.weight.dtype

    @property
    def device(self):
        return self.in_proj.weight.device


class SingleStepSlotRollouter(SlotRollouter):
    """Rollouter conditioned on the first frame only (PHYRE): the window grows by one frame
    per step until it holds ``cond_len`` frames, then slides."""

    _mode = 'grow'

    def __init__(self, num_slots, slot_size, history_len, cond_len, t_pe='sin', slots_pe='',
                 d_model=128, num_layers=4, num_heads=8, ffn_dim=512, norm_first=True):
        super().__init__(num_slots=num_slots, slot_size=slot_size, history_len=history_len,
                         t_pe=t_pe, slots_pe=slots_pe, d_model=d_model, num_layers=num_layers,
                         num_heads=num_heads, ffn_dim=ffn_dim, norm_first=norm_first)
        assert self.history_len == 1, \
            'SingleStepSlotRollouter performs rollout using only initial frame'
        self.cond_len = cond_len
        self.num_cond_tokens = cond_len * num_slots
        self.enc_t_pe = build_pos_enc(t_pe, cond_len, d_model)
