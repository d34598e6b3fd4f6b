"""Cross-attention launches of the two headline steps (LTX: d 64, 3 x 32 heads x 6144 queries x 256 prompt keys; Wan-1.3B: d 128, 2 x 12 x 32760 x 512)
for same-box A/B of builds selected with LTXB200_LIB."""
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
    print(f"{name}: {ms:.4f} ms {flops/ms/1e9:.1f} TFLOP/s", flush=True)
for (B, N, H, d, L) in [(3, 6144, 32, 64, 256), (2, 32760, 12, 128, 512)]:
    q = torch.randn(B, N, H, d, device="cuda").bfloat16()
    k, v = [torch.randn(B, L, H, d, device="cuda").bfloat16() for _ in range(2)]
    t(f"cross-attention d{d} B{B} N{N} H{H} L{L}", lambda: ops.attention(q, k, v), 4.0 * B * H * N * L * d)
    kl = torch.tensor([L] * B, dtype=torch.int32, device="cuda")
    t(f"cross-attention d{d} key_lens full", lambda: ops.attention(q, k, v, key_lens=kl), 4.0 * B * H * N * L * d)
