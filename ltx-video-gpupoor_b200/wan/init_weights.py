"""Seeded random initialisation with the reference's key names for benchmarks (no checkpoints in this environment):
`random_wan_vae_decoder_state_dict` mirrors the layout of WanVAE_ (wan/modules/vae.py:503-530: conv2.*, decoder.*),
`seeded_wan_state_dict` that of WanModel (wan/modules/model.py:656-813)."""
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


def seeded_wan_state_dict(cfg: dict, seed: int = 0, num_layers=None) -> Dict[str, torch.Tensor]:
    """WanModel t2v weights (reference key names, fp32, CPU) from one seeded generator: uniform(+-1/sqrt(fan_in)) linears,
    norm weights 1 + 0.1 N(0,1), modulation tables N(0,1)/sqrt(dim).  The draw order is fixed — it is the recipe the golden
    fixtures under tests/golden/wan_*.pt were recorded with, which lets `bench.py` check the sequence-parallel forward
    against a fixture of the unmodified reference without test infrastructure (tests/test_cabi_and_host_cpu.py pins the
    equality).  head.head is NOT zero-initialised as the reference does (model.py:1160: the output would be identically 0)."""
    g = torch.Generator().manual_seed(seed)
    D, Fd = cfg["dim"], cfg["ffn_dim"]
    L = cfg["num_layers"] if num_layers is None else num_layers
    patch = tuple(cfg.get("patch_size", (1, 2, 2)))
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, o, i):
        b = 1.0 / math.sqrt(i)
        sd[name + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * b
        sd[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * b

    def norm_w(name):
        sd[name] = 1.0 + 0.1 * torch.randn(D, generator=g)

    pk = cfg.get("in_dim", 16) * math.prod(patch)
    sd["patch_embedding.weight"] = ((torch.rand(D, pk, generator=g) * 2 - 1) / math.sqrt(pk)).view(D, cfg.get("in_dim", 16), *patch)
    sd["patch_embedding.bias"] = (torch.rand(D, generator=g) * 2 - 1) / math.sqrt(pk)
    lin("text_embedding.0", D, cfg.get("text_dim", 4096)); lin("text_embedding.2", D, D)
    lin("time_embedding.0", D, cfg.get("freq_dim", 256)); lin("time_embedding.2", D, D)
    lin("time_projection.1", 6 * D, D)
    for i in range(L):
        p = f"blocks.{i}."
        for a in ("self_attn", "cross_attn"):
            for n in ("q", "k", "v", "o"):
                lin(p + a + "." + n, D, D)
            norm_w(p + a + ".norm_q.weight"); norm_w(p + a + ".norm_k.weight")
        norm_w(p + "norm3.weight")
        sd[p + "norm3.bias"] = 0.1 * torch.randn(D, generator=g)
        lin(p + "ffn.0", Fd, D); lin(p + "ffn.2", D, Fd)
        sd[p + "modulation"] = torch.randn(1, 6, D, generator=g) / D ** 0.5
    lin("head.head", cfg.get("out_dim", 16) * math.prod(patch), D)
    sd["head.modulation"] = torch.randn(1, 2, D, generator=g) / D ** 0.5
    return sd
