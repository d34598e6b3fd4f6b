"""Parity of the d = 128 attention path with the CTA-pair kernel selected (LTXB200_ATTN128_2CTA=1 -> attention128p2_kernel): several items per
cluster, partial query groups (Lq % 512), partial key blocks, strided q/k/v out of one fused QKV buffer, key bias, per-batch key lengths,
large logits (lazy rescale), the accumulate epilogue; deterministic.  Default (0) runs the same shapes through attention_fwd_kernel<128>."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ltx_video_gpupoor_b200 import ops

dev = "cuda"


def ref(q, k, v, bias=None):
    m = None if bias is None else bias[:, None, None, :].float()
    o = torch.nn.functional.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2), attn_mask=m)
    return o.transpose(1, 2)


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


worst = 0.0
for (B, H, Lq, Lk, scale) in [(1, 2, 128, 256, 1.0), (2, 3, 512, 384, 1.0), (1, 4, 300, 1280, 1.0), (2, 40, 600, 512, 1.0), (3, 12, 1000, 2040, 3.0),
                              (1, 160, 515, 768, 0.3), (1, 1, 77, 4096, 5.0), (1, 2, 1300, 97, 1.0), (2, 3, 2049, 1000, 2.0), (1, 1, 64, 1, 1.0)]:
    g = torch.Generator(device=dev).manual_seed(B * 1000 + Lq)
    qkv = (torch.randn(B, max(Lq, Lk), 3 * H * 128, device=dev, generator=g) * scale).bfloat16()
    q, k, v = [qkv[:, :, i * H * 128:(i + 1) * H * 128].unflatten(-1, (H, 128)) for i in range(3)]
    q, k, v = q[:, :Lq], k[:, :Lk], v[:, :Lk]
    o = ops.attention(q, k, v)
    torch.cuda.synchronize()
    e = rel(o, ref(q, k, v))
    worst = max(worst, e)
    print(f"B{B} H{H} Lq{Lq} Lk{Lk} x{scale}: rel-L2 {e:.2e}", flush=True)
    assert e < 1e-2 and bool(torch.isfinite(o.float()).all())
    assert torch.equal(o, ops.attention(q, k, v)), "not deterministic"
# key bias, key lengths, accumulate
B, H, Lq, Lk = 3, 4, 700, 384
g = torch.Generator(device=dev).manual_seed(7)
q, k, v = [torch.randn(B, L, H, 128, device=dev, generator=g).bfloat16() for L in (Lq, Lk, Lk)]
lens = [384, 130, 1]
bias = torch.zeros(B, Lk, device=dev)
for b, n in enumerate(lens):
    bias[b, n:] = -10000.0
ob = ops.attention(q, k, v, key_bias=bias)
ol = ops.attention(q, k, v, key_lens=torch.tensor(lens, dtype=torch.int32, device=dev))
torch.cuda.synchronize()
e1, e2 = rel(ob, ref(q, k, v, bias)), rel(ol, ob)
print(f"key bias {e1:.2e}  key lengths vs bias {e2:.2e}", flush=True)
assert e1 < 1e-2 and e2 < 4e-3
worst = max(worst, e1)
print("worst", worst, "ATTN128_2CTA =", os.environ.get("LTXB200_ATTN128_2CTA", "0"))
