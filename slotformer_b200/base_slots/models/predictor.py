"""Slot transition functions t -> t+1 used inside SAVi (interface of reference
base_slots/models/predictor.py).  Tiny (B*K rows once per frame): stock PyTorch, not a kernel
target."""
import torch
from torch import nn


class Predictor(nn.Module):
    def forward(self, x):
        raise NotImplementedError

    def burnin(self, x):
        pass

    def reset(self):
        pass


class TransformerPredictor(Predictor):
    """Transformer encoder over the K slots of a frame."""

    def __init__(self, d_model=128, num_layers=1, num_heads=4, ffn_dim=256, norm_first=True):
        super().__init__()
        layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=num_heads,
                                           dim_feedforward=ffn_dim, norm_first=norm_first,
                                           batch_first=True)
        self.transformer_encoder = nn.TransformerEncoder(encoder_layer=layer,
                                                         num_layers=num_layers,
                                                         enable_nested_tensor=False)

    def forward(self, x):
        return self.transformer_encoder(x)


class ResidualMLPPredictor(Predictor):
    """LayerNorm + MLP with a residual connection (taken after the norm if ``norm_first``)."""

    def __init__(self, channels, norm_first=True):
        super().__init__()
        assert len(channels) >= 2
        self.ln = nn.LayerNorm(channels[0])
        layers = []
        for cin, cout in zip(channels[:-2], channels[1:-1]):
            layers += [nn.Linear(cin, cout), nn.ReLU()]
        layers.append(nn.Linear(channels[-2], channels[-1]))
        self.mlp = nn.Sequential(*layers)
        self.norm_first = norm_first

    def forward(self, x):
        normed = self.ln(x)
        skip = normed if self.norm_first else x
        return self.mlp(normed) + skip


class RNNPredictorWrapper(Predictor):
    """Runs ``base_predictor`` and then a recurrent cell over time (hidden state kept between
    calls; ``reset`` clears it)."""

    def __init__(self, base_predictor, input_size=128, hidden_size=256, num_layers=1,
                 rnn_cell='LSTM', sg_every=None):
        super().__init__()
        cells = {'LSTM': nn.LSTM, 'GRU': nn.GRU, 'RNN': nn.RNN}
        assert rnn_cell in cells
        self.base_predictor = base_predictor
        self.rnn = cells[rnn_cell](input_size=input_size, hidden_size=hidden_size,
                                   num_layers=num_layers)
        self.step = 0
        self.hidden_state = None
        self.out_projector = nn.Linear(hidden_size, input_size)
        self.sg_every = sg_every

    def _detach_state(self):
        if isinstance(self.hidden_state, torch.Tensor):
            self.hidden_state = self.hidden_state.detach()
        elif self.hidden_state is not None:
            self.hidden_state = tuple(h.detach() for h in self.hidden_state)

    def forward(self, x):
        if self.sg_every is not None and self.step > 0 and self.step % self.sg_every == 0:
            x = x.detach()
            self._detach_state()
        out = self.base_predictor(x)
        shape = out.shape
        if not (out.is_cuda and torch.cuda.is_current_stream_capturing()):
            # (inside a CUDA-graph capture it would move the weights into the graph's pool.)  flatten_parameters() allocates
            # a new flat buffer on every call: once per set of weight addresses is enough
            ptrs = tuple(w.data_ptr() for w in self.rnn._flat_weights)
            if self.__dict__.get('_flat_ptrs') != ptrs:
                self.rnn.flatten_parameters()
                self.__dict__['_flat_ptrs'] = tuple(w.data_ptr() for w in self.rnn._flat_weights)
        out, self.hidden_state = self.rnn(out.reshape(1, -1, shape[-1]), self.hidden_state)
        self.step += 1
        return self.out_projector(out[0]).view(shape)

    def burnin(self, x):
        """Warm the recurrent state up on [B, T, K, D] slots."""
        self.reset()
        B, T = x.shape[:2]
        out = self.base_predictor(x.flatten(0, 1)).unflatten(0, (B, T))
        out = out.transpose(1, 0).reshape(T, -1, x.shape[-1])
        _, self.hidden_state = self.rnn(out, self.hidden_state)
        self.step = T

    def reset(self):
        self.step = 0
        self.hidden_state = None
