"""Per-kernel breakdown of the full-size LTX VAE decode (BASELINE configs[2]: latent (1,128,16,16,24) -> 121x512x768),
CUDA events around every launch (ops.PROFILER); with `once` as argv[1] a single decode and no timing (for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ltx_video_gpupoor_b200 import ops
from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import CausalVideoAutoencoder, vae_decode
from ltx_video_gpupoor_b200.ltx.init_weights import random_vae_decoder_state_dict

dev = torch.device("cuda")
vae = CausalVideoAutoencoder()
vae.load_state_dict(random_vae_decoder_state_dict(seed=1, device=dev), device=dev)
z = torch.randn(1, 128, 16, 16, 24, device=dev)
if len(sys.argv) > 1 and sys.argv[1] == "once":
    vae_decode(z, vae, True, vae_per_channel_normalize=True)
    torch.cuda.synchronize()
    sys.exit(0)
for _ in range(2):
    vae_decode(z, vae, True, vae_per_channel_normalize=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    vae_decode(z, vae, True, vae_per_channel_normalize=True)
b.record()
torch.cuda.synchronize()
print(f"decode: {a.elapsed_time(b) / 3:.2f} ms")
ops.PROFILER = []
vae_decode(z, vae, True, vae_per_channel_normalize=True)
torch.cuda.synchronize()
prof, ops.PROFILER = ops.PROFILER, None
agg = {}
for name, kind, amount, e0, e1 in prof:
    d = agg.setdefault((name, kind, amount), [0.0, 0])
    d[0] += e0.elapsed_time(e1); d[1] += 1
tot = sum(v[0] for v in agg.values())
print(f"instrumented total {tot:.2f} ms over {sum(v[1] for v in agg.values())} launches")
for (name, kind, amount), (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    rate = amount * n / (ms * 1e-3) / (1e12 if kind == "flop" else 1e9)
    print(f"{name:24s} {kind} {amount:12.4g} x{n:2d}  {ms:7.3f} ms  {ms / tot * 100:5.1f}%  {rate:8.1f} {'TF/s' if kind == 'flop' else 'GB/s'}")
