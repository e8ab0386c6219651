"""Small helpers shared by the slot models (interface of reference base_slots/models/utils.py)."""
import torch
from torch import nn


def get_lr(optimizer):
    return optimizer.param_groups[0]['lr']


def torch_stack(tensors, dim):
    return torch.stack(tensors) if tensors[0].dim() < dim else torch.stack(tensors, dim=dim)


def torch_cat(tensors, dim):
    return torch.cat(tensors) if tensors[0].dim() <= dim else torch.cat(tensors, dim=dim)


def assert_shape(actual, expected, message=''):
    assert list(actual) == list(expected), \
        f'Expected shape: {expected} but passed shape: {actual}. {message}'


def to_rgb_from_tensor(x):
    """[-1, 1] image tensor -> [0, 1]."""
    return (x * 0.5 + 0.5).clamp(0, 1)


def build_grid(resolution):
    """[1, H, W, 4] grid of (y, x, 1-y, 1-x) coordinates in [0, 1]."""
    axes = [torch.linspace(0.0, 1.0, steps=r) for r in resolution]
    yx = torch.stack(torch.meshgrid(*axes, indexing='ij'), dim=-1)
    yx = yx.reshape(resolution[0], resolution[1], -1).unsqueeze(0)
    return torch.cat([yx, 1.0 - yx], dim=-1)


class SoftPositionEmbed(nn.Module):
    """Adds a learned linear embedding of the normalised pixel coordinates to a feature map."""

    def __init__(self, hidden_size, resolution):
        super().__init__()
        self.dense = nn.Linear(4, hidden_size)
        self.register_buffer('grid', build_grid(resolution))

    def forward(self, inputs):
        """inputs [B, C, H, W]."""
        return inputs + self.dense(self.grid).permute(0, 3, 1, 2).contiguous()
