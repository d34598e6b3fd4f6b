"""Seeded random initialisation with the reference's key names for benchmarks (no checkpoints in this environment).
Mirrors the layout of WanVAE_ (wan/modules/vae.py:503-530): conv2.*, decoder.* ."""
import math
from typing import Dict

import torch


def random_wan_vae_decoder_state_dict(seed: int = 0, dim: int = 96, z_dim: int = 16, dim_mult=(1, 2, 4, 4), num_res_blocks: int = 2,
                                      temperal_upsample=(True, True, False)) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k):
        fan = cin * math.prod(k)
        sd[name + ".weight"] = (torch.rand(cout, cin, *k, generator=g) * 2 - 1) / math.sqrt(fan)
        sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) / math.sqrt(fan)

    def gamma(name, c):
        sd[name + ".gamma"] = 1.0 + 0.1 * torch.randn(c, generator=g)

    def res(p, cin, cout):
        gamma(p + "residual.0", cin); conv(p + "residual.2", cout, cin, (3, 3, 3))
        gamma(p + "residual.3", cout); conv(p + "residual.6", cout, cout, (3, 3, 3))
        if cin != cout:
            conv(p + "shortcut", cout, cin, (1, 1, 1))

    dm = list(dim_mult)
    dims = [dim * u for u in [dm[-1]] + dm[::-1]]
    d0 = dims[0]
    conv("conv2", z_dim, z_dim, (1, 1, 1)); conv("decoder.conv1", d0, z_dim, (3, 3, 3))
    res("decoder.middle.0.", d0, d0)
    gamma("decoder.middle.1.norm", d0); conv("decoder.middle.1.to_qkv", 3 * d0, d0, (1, 1)); conv("decoder.middle.1.proj", d0, d0, (1, 1))
    res("decoder.middle.2.", d0, d0)
    idx, c_last = 0, d0
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            cin = cin // 2
        for _ in range(num_res_blocks + 1):
            res(f"decoder.upsamples.{idx}.", cin, cout); cin = cout; idx += 1
        c_last = cout
        if i != len(dm) - 1:
            conv(f"decoder.upsamples.{idx}.resample.1", cout // 2, cout, (3, 3))
            if temperal_upsample[i]:
                conv(f"decoder.upsamples.{idx}.time_conv", 2 * cout, cout, (3, 1, 1))
            c_last = cout // 2; idx += 1
    gamma("decoder.head.0", c_last); conv("decoder.head.2", 3, c_last, (3, 3, 3))
    return sd
