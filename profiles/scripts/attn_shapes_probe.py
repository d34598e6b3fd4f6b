import sys; sys.path.insert(0, "/root/repo")
import torch
from ltx_video_gpupoor_b200 import ops
def t(name, fn, flops, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"{name}: {ms:.3f} ms {flops/ms/1e9:.1f} TFLOP/s", flush=True)
for (B, N, H, d) in [(3, 6144, 32, 128), (1, 6144, 32, 128), (3, 6144, 32, 64), (1, 32760, 12, 128), (2, 32760, 12, 128)]:
    qkv = torch.randn(B, N, 3 * H * d, device="cuda").bfloat16()
    q, k, v = [qkv[:, :, i * H * d:(i + 1) * H * d].unflatten(-1, (H, d)) for i in range(3)]
    t(f"fused-qkv layout B{B} N{N} H{H} d{d}", lambda: ops.attention(q, k, v), 4.0 * B * H * N * N * d)
    q2, k2, v2 = q.contiguous(), k.contiguous(), v.contiguous()
    t(f"contiguous      B{B} N{N} H{H} d{d}", lambda: ops.attention(q2, k2, v2), 4.0 * B * H * N * N * d)
