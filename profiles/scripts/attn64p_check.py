"""Parity of the d = 64 attention path on shapes that take attention64p_kernel (several items per CTA, partial last query tile,
strided q/k/v out of one fused QKV buffer), then the LTX self-attention timing.  LTXB200_ATTN64P=1 selects attention64p_kernel (default:
attention_fwd_kernel<64,false>)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ltx_video_gpupoor_b200 import ops

dev = "cuda"


def ref(q, k, v):
    o = torch.nn.functional.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2))
    return o.transpose(1, 2)


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


worst = 0.0
for (B, H, Lq, Lk, scale) in [(1, 2, 128, 256, 1.0), (2, 3, 256, 384, 1.0), (1, 4, 300, 1280, 1.0), (2, 40, 600, 512, 1.0), (3, 32, 1000, 2048, 3.0),
                              (1, 160, 515, 768, 0.3), (1, 1, 77, 4096, 5.0), (1, 2, 128, 96, 1.0), (2, 3, 300, 960, 2.0), (3, 32, 1000, 1920, 1.0),
                              (2, 40, 600, 192, 4.0)]:
    g = torch.Generator(device=dev).manual_seed(B * 1000 + Lq)
    qkv = torch.randn(B, max(Lq, Lk), 3 * H * 64, device=dev, generator=g) * scale
    qkv = qkv.bfloat16()
    q, k, v = [qkv[:, :, i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3)]
    q, k, v = q[:, :Lq], k[:, :Lk], v[:, :Lk]
    o = ops.attention(q, k, v)
    torch.cuda.synchronize()
    e = rel(o, ref(q, k, v))
    worst = max(worst, e)
    print(f"B{B} H{H} Lq{Lq} Lk{Lk} x{scale}: rel-L2 {e:.2e}", flush=True)
    assert e < 1e-2 and bool(torch.isfinite(o.float()).all())
    o2 = ops.attention(q, k, v)
    assert torch.equal(o, o2), "not deterministic"
print("worst", worst, "ATTN64P =", os.environ.get("LTXB200_ATTN64P", "0"))

B, N, H, d = 3, 6144, 32, 64
qkv = torch.randn(B, N, 3 * H * d, device=dev).bfloat16()
q, k, v = [qkv[:, :, i * H * d:(i + 1) * H * d].unflatten(-1, (H, d)) for i in range(3)]
for rep in range(3):
    for _ in range(3):
        ops.attention(q, k, v)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ops.attention(q, k, v)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"attention d64 B3 N6144 H32: {ms:.4f} ms  {4.0 * B * H * N * N * d / ms / 1e9:.1f} TFLOP/s", flush=True)
