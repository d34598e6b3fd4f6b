"""Timing of the two headline attention shapes for a diagnostic build selected with LTXB200_LIB (results of the ablation builds are wrong by
construction; only the time is read).  Usage: LTXB200_LIB=... python profiles/scripts/attn_ablation_probe.py"""
import sys; sys.path.insert(0, "/root/repo")
import torch
from ltx_video_gpupoor_b200 import ops
def t(name, fn, flops, reps=int(__import__("os").environ.get("REPS", "10"))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"{name}: {ms:.3f} ms {flops/ms/1e9:.1f} TFLOP/s", flush=True)
for (B, N, H, d) in [(1, 32760, 12, 128), (3, 6144, 32, 64)]:
    qkv = torch.randn(B, N, 3 * H * d, device="cuda").bfloat16()
    q, k, v = [qkv[:, :, i * H * d:(i + 1) * H * d].unflatten(-1, (H, d)) for i in range(3)]
    t(f"B{B} N{N} H{H} d{d}", lambda: ops.attention(q, k, v), 4.0 * B * H * N * N * d)
