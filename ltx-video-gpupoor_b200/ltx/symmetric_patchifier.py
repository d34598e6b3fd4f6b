"""SymmetricPatchifier — drop-in for ltx_video/models/transformers/symmetric_patchifier.py:54-84 and the
coordinate helpers of vae_encode.py:190-225.  Pure index/permutation work: bit-exact."""
from typing import Tuple

import torch


class SymmetricPatchifier:
    def __init__(self, patch_size: int = 1):
        self._patch_size = (1, patch_size, patch_size)

    @property
    def patch_size(self):
        return self._patch_size

    def get_latent_coords(self, latent_num_frames, latent_height, latent_width, batch_size, device):
        """symmetric_patchifier.py:33-51: int64 [b, 3, n] top-left (t, y, x) of every patch."""
        g = torch.meshgrid(torch.arange(0, latent_num_frames, self._patch_size[0], device=device),
                           torch.arange(0, latent_height, self._patch_size[1], device=device),
                           torch.arange(0, latent_width, self._patch_size[2], device=device), indexing="ij")
        c = torch.stack(g, dim=0).reshape(3, -1)
        return c.unsqueeze(0).repeat(batch_size, 1, 1)

    def patchify(self, latents: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """b c (f p1) (h p2) (w p3) -> b (f h w) (c p1 p2 p3)"""
        b, c, f, h, w = latents.shape
        p1, p2, p3 = self._patch_size
        coords = self.get_latent_coords(f, h, w, b, latents.device)
        x = latents.reshape(b, c, f // p1, p1, h // p2, p2, w // p3, p3).permute(0, 2, 4, 6, 1, 3, 5, 7)
        return x.reshape(b, (f // p1) * (h // p2) * (w // p3), c * p1 * p2 * p3), coords

    def unpatchify(self, latents: torch.Tensor, output_height: int, output_width: int, out_channels: int) -> torch.Tensor:
        """b (f h w) (c p q) -> b c f (h p) (w q)"""
        p, q = self._patch_size[1], self._patch_size[2]
        h, w = output_height // p, output_width // q
        b, n, _ = latents.shape
        f = n // (h * w)
        x = latents.reshape(b, f, h, w, out_channels, p, q).permute(0, 4, 1, 2, 5, 3, 6)
        return x.reshape(b, out_channels, f, h * p, w * q)


def latent_to_pixel_coords_from_factors(latent_coords: torch.Tensor, scale_factors, causal_fix: bool = False):
    """vae_encode.py:214-225"""
    px = latent_coords * torch.tensor(scale_factors, device=latent_coords.device)[None, :, None]
    if causal_fix:
        px[:, 0] = (px[:, 0] + 1 - scale_factors[0]).clamp(min=0)
    return px
