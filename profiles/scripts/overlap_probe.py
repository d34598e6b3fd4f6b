"""Do the QKV GEMM (persistent, 1 CTA / SM, ~200 KB smem) and the q/k-norm + RoPE + scatter kernel (bounded grid, other stream) run
CONCURRENTLY on one GPU?  P = 1 'exchange' into a local buffer; times GEMM alone, scatter alone, both on two streams."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ltx_video_gpupoor_b200 import _lib, ops

lib = _lib.lib()
dev = torch.device("cuda")
M, D = 16380, 1536
x = torch.randn(M, D, device=dev).bfloat16()
w = (torch.randn(3 * D, D, device=dev) * 0.02).bfloat16()
b = torch.randn(3 * D, device=dev).bfloat16()
qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
qkv2 = torch.randn(M, 3 * D, device=dev).bfloat16()
wn = torch.ones(D, device=dev, dtype=torch.bfloat16)
cos = torch.randn(M, 128, device=dev); sin = torch.randn(M, 128, device=dev)
recv = torch.empty(M * 3 * D, device=dev, dtype=torch.bfloat16)
ctl = torch.zeros(1024, device=dev, dtype=torch.int32)
VP = ctypes.c_void_p * 1
side = torch.cuda.Stream()
epoch = [0]


def scatter(stream):
    epoch[0] += 1
    _lib.check(lib.ltxb200_qk_norm_rope_wan_scatter_rows_bf16(
        qkv2.data_ptr(), qkv2.stride(0), M, 0, M, D, wn.data_ptr(), wn.data_ptr(), cos.data_ptr(), sin.data_ptr(), 128, M, 0, 1e-6, 1, 1, 0,
        VP(recv.data_ptr()), VP(ctl.data_ptr()), epoch[0], ctl.data_ptr() + 256, lib.ltxb200_scatter_signal_ctas(M), 3, stream.cuda_stream), "scatter")


def gemm():
    ops.gemm(x, w, b, out=qkv)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return a.elapsed_time(e) / reps * 1e3


main = torch.cuda.current_stream()


def both(gemm_first=True):
    ev = torch.cuda.Event()
    ev.record(main)
    side.wait_event(ev)
    if gemm_first:
        gemm(); scatter(side)
    else:
        scatter(side); gemm()
    main.wait_stream(side)


t_g, t_s = timeit(gemm), timeit(lambda: scatter(main))
print(f"gemm alone {t_g:.1f} us, scatter alone {t_s:.1f} us, serial sum {t_g + t_s:.1f} us")
print(f"two streams, gemm enqueued first: {timeit(lambda: both(True)):.1f} us;  scatter enqueued first: {timeit(lambda: both(False)):.1f} us")
