"""Launch the hot kernels at LTX-2B bench shapes a few times (for ncu captures and quick timing)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ltx_video_gpupoor_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = "cuda"
torch.manual_seed(0)


def timeit(name, fn, flops):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"{name}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)


if which in ("all", "attn"):
    B, N, H, d = 3, 6144, 32, 64
    qkv = torch.randn(B, N, 3 * H * d, device=dev).bfloat16()
    q, k, v = [qkv[:, :, i * H * d:(i + 1) * H * d].unflatten(-1, (H, d)) for i in range(3)]
    timeit("attention d64 B3 N6144 H32", lambda: ops.attention(q, k, v), 4.0 * B * H * N * N * d)
if which in ("all", "attn", "xattn"):
    B, N, H, d, L = 3, 6144, 32, 64, 256
    q = torch.randn(B, N, H, d, device=dev).bfloat16()
    kv = torch.randn(B, L, 2 * H * d, device=dev).bfloat16()
    k, v = kv[:, :, :H * d].unflatten(-1, (H, d)), kv[:, :, H * d:].unflatten(-1, (H, d))
    bias = torch.zeros(B, L, device=dev)
    timeit("cross-attention d64 B3 N6144 L256 bias", lambda: ops.attention(q, k, v, key_bias=bias), 4.0 * B * H * N * L * d)
    timeit("cross-attention d64 B3 N6144 L256 nobias", lambda: ops.attention(q, k, v), 4.0 * B * H * N * L * d)
    for lens in ([256, 256, 256], [200, 120, 200], [100, 30, 100]):
        kl = torch.tensor(lens, dtype=torch.int32, device=dev)
        timeit(f"cross-attention d64 B3 N6144 L256 key_lens {lens} (flops counted on 256 keys)", lambda: ops.attention(q, k, v, key_lens=kl), 4.0 * B * H * N * L * d)
if which in ("all", "attn128"):
    B, N, H, d = 1, 32760, 12, 128
    q, k, v = [torch.randn(B, N, H, d, device=dev).bfloat16() for _ in range(3)]
    timeit("attention d128 B1 N32760 H12 (Wan-1.3B)", lambda: ops.attention(q, k, v), 4.0 * B * H * N * N * d)
    B, N, H, d = 1, 8192, 12, 128
    q, k, v = [torch.randn(B, N, H, d, device=dev).bfloat16() for _ in range(3)]
    timeit("attention d128 B1 N8192 H12", lambda: ops.attention(q, k, v), 4.0 * B * H * N * N * d)
if which in ("all", "gemm"):
    M, N, K = 18432, 8192, 2048
    a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16() * 0.02; bias = torch.randn(N, device=dev).bfloat16()
    timeit("gemm ffn_up 18432x8192x2048 gelu", lambda: ops.gemm(a, w, bias, act=ops.ACT_GELU_TANH), 2.0 * M * N * K)
    M, N, K = 18432, 2048, 8192
    a2 = torch.randn(M, K, device=dev).bfloat16(); w2 = torch.randn(N, K, device=dev).bfloat16() * 0.02
    timeit("gemm ffn_down 18432x2048x8192", lambda: ops.gemm(a2, w2, None), 2.0 * M * N * K)
    M, N, K = 18432, 6144, 2048
    w3 = torch.randn(N, K, device=dev).bfloat16() * 0.02
    timeit("gemm qkv 18432x6144x2048", lambda: ops.gemm(a, w3, None), 2.0 * M * N * K)
if which in ("all", "ew"):
    M, D = 18432, 2048
    x = torch.randn(M, D, device=dev).bfloat16()
    ada = torch.randn(3, 6, D, device=dev).bfloat16() * 0.1
    def timeit_bw(name, fn, nbytes):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        print(f"{name}: {ms * 1e3:.1f} us  {nbytes / ms / 1e6:.0f} GB/s", flush=True)
    sc, sh = ada[:, 0], ada[:, 1]
    out = torch.empty_like(x)
    timeit_bw("norm_mod rms+adaln 18432x2048", lambda: ops.norm_mod(x, sc, sh, rows_per_group=6144, out=out), 4.0 * M * D)
    qkv = torch.randn(M, 3 * D, device=dev).bfloat16()
    w = torch.randn(D, device=dev).bfloat16()
    cos = torch.randn(6144, D, device=dev).bfloat16(); sin = torch.randn(6144, D, device=dev).bfloat16()
    timeit_bw("qk_norm_rope q,k 18432x2048 (in fused qkv)", lambda: ops.qk_norm_rope(qkv[:, :D], qkv[:, D:2 * D], w, w, cos, sin, tokens_per_batch=6144),
              8.0 * M * D + 4.0 * 6144 * D)
    q2 = torch.randn(M, D, device=dev).bfloat16()
    timeit_bw("qk_norm (cross-attn q only) 18432x2048", lambda: ops.qk_norm_rope(q2, None, w, None), 4.0 * M * D)
if which in ("wangemm",):
    M = 65520
    for (name, N, K, act) in (("wan qkv", 4608, 1536, None), ("wan o / q", 1536, 1536, None), ("wan ffn_up gelu", 8960, 1536, ops.ACT_GELU_TANH),
                              ("wan ffn_down", 1536, 8960, None)):
        a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16() * 0.02; bias = torch.randn(N, device=dev).bfloat16()
        if act is None:
            timeit(f"gemm {name} {M}x{N}x{K}", lambda: ops.gemm(a, w, bias), 2.0 * M * N * K)
        else:
            timeit(f"gemm {name} {M}x{N}x{K}", lambda: ops.gemm(a, w, bias, act=act), 2.0 * M * N * K)
        del a, w
if which in ("all", "wanew"):
    # Wan-1.3B self-attention q/k norm + RoPE on the fused QKV rows of both sequences (2 x 32760 tokens), fp32 [N, 128] tables
    M, D = 65520, 1536
    qkv = torch.randn(M, 3 * D, device=dev).bfloat16()
    w = torch.randn(D, device=dev).bfloat16()
    cos = torch.randn(32760, 128, device=dev); sin = torch.randn(32760, 128, device=dev)
    for _ in range(3):
        ops.qk_norm_rope_wan(qkv[:, :D], qkv[:, D:2 * D], w, w, cos, sin, head_dim=128, tokens_per_batch=32760)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ops.qk_norm_rope_wan(qkv[:, :D], qkv[:, D:2 * D], w, w, cos, sin, head_dim=128, tokens_per_batch=32760)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"qk_norm_rope_wan q,k 65520x1536: {ms * 1e3:.1f} us  {8.0 * M * D / ms / 1e6:.0f} GB/s", flush=True)
