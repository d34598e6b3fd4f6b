"""Full-size Wan VAE decode (832x480x81 -> latent 16x21x60x104): time, peak memory, per-kernel breakdown."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ltx_video_gpupoor_b200 import ops
from ltx_video_gpupoor_b200.wan.vae import WanVAE
from ltx_video_gpupoor_b200.wan.init_weights import random_wan_vae_decoder_state_dict

dev = torch.device("cuda", 0)
vae = WanVAE(device=dev)
vae.load_state_dict(random_wan_vae_decoder_state_dict(seed=1), device=dev)
z = torch.randn(16, 21, 60, 104, device=dev)
vae.decode([z], 0)
torch.cuda.synchronize()
torch.cuda.reset_peak_memory_stats()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
y = vae.decode([z], 0)[0]
b.record()
torch.cuda.synchronize()
print(f"decode {tuple(y.shape)}: {a.elapsed_time(b):.1f} ms, peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
ops.PROFILER = []
vae.decode([z], 0)
torch.cuda.synchronize()
prof, ops.PROFILER = ops.PROFILER, None
agg = {}
for name, kind, amount, e0, e1 in prof:
    d = agg.setdefault(name, [kind, 0.0, 0.0, 0])
    d[1] += amount; d[2] += e0.elapsed_time(e1); d[3] += 1
for k, (kind, amt, ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f"  {k:22s} {n:4d} launches {ms:8.2f} ms  {amt / ms / (1e9 if kind == 'flop' else 1e6):9.1f} {'TFLOP/s' if kind == 'flop' else 'GB/s'}")
