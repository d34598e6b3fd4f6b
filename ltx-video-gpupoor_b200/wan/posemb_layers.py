"""3-axis RoPE tables for Wan — drop-in for wan/modules/posemb_layers.py:432-473 (get_rotary_pos_embed) and the
helpers it reaches (:299-430).  Integer grid positions are exact; trig is fp32 on the host exactly as the
reference computes it (tables are then uploaded once per video)."""
from typing import Sequence, Tuple

import torch


def get_1d_rotary_pos_embed(dim: int, pos: torch.Tensor, theta: float = 10000.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """posemb_layers.py:381-430, use_real=True, no rescale/interpolation"""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
    freqs = torch.outer(pos, freqs)
    return freqs.cos().repeat_interleave(2, dim=1), freqs.sin().repeat_interleave(2, dim=1)


def get_1d_rotary_pos_embed_riflex(dim: int, pos: torch.Tensor, theta: float, k: int, L_test: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """posemb_layers.py:8-62 (RIFLEx, use_real=True): the k-th frequency of the TIME axis is lowered to 0.9 * 2*pi / L_test so that the
    extrapolated length stays inside one period of that component."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
    freqs[k - 1] = 0.9 * 2 * torch.pi / L_test
    freqs = torch.outer(pos, freqs)
    return freqs.cos().repeat_interleave(2, dim=1).float(), freqs.sin().repeat_interleave(2, dim=1).float()


def get_rotary_pos_embed(latents_size: Sequence[int], enable_RIFLEx: bool = False):
    """posemb_layers.py:432-473: latents_size = (F, H, W) of the latent; patch (1,2,2); head_dim 128 split [44,42,42];
    enable_RIFLEx: time axis through get_1d_rotary_pos_embed_riflex with k = 6 (:308) and L_test = latent frames (:470)."""
    patch = [1, 2, 2]
    assert all(s % patch[i] == 0 for i, s in enumerate(latents_size))
    sizes = [s // patch[i] for i, s in enumerate(latents_size)]
    grids = [torch.linspace(0, n, n + 1, dtype=torch.float32)[:n] for n in sizes]          # :136-141
    grid = torch.stack(torch.meshgrid(*grids, indexing="ij"), dim=0)
    cos, sin = [], []
    for i, d in enumerate([44, 42, 42]):
        if i == 0 and enable_RIFLEx:
            c, s = get_1d_rotary_pos_embed_riflex(d, grid[i].reshape(-1), 10000.0, k=6, L_test=latents_size[0])
        else:
            c, s = get_1d_rotary_pos_embed(d, grid[i].reshape(-1), 10000.0)
        cos.append(c); sin.append(s)
    return torch.cat(cos, dim=1), torch.cat(sin, dim=1)
