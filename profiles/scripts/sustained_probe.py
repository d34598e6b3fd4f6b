"""Sustained (power-capped) throughput of the hot kernels: each one alone in a loop for ~3 s, TFLOP/s over the last 2 s, with the SM
clock and board power sampled through NVML meanwhile.  Under the 1 kW cap the time of the LTX step is its energy: a kernel's sustained
TFLOP/s is its energy efficiency, and cuBLAS on the same shape is the yardstick."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pynvml
from ltx_video_gpupoor_b200 import ops

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = "cuda"
torch.manual_seed(0)


def run(name, fn, flops, secs=3.0):
    samples = []
    stop = False

    def sampler():
        while not stop:
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
            time.sleep(0.05)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    th = threading.Thread(target=sampler); th.start()
    t0 = time.perf_counter()
    n = 0
    marks = []
    while time.perf_counter() - t0 < secs:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
        marks.append((time.perf_counter() - t0, n))
    stop = True; th.join()
    # last 2 seconds
    t_end, n_end = marks[-1]
    t_a, n_a = next((t, k) for t, k in marks if t >= t_end - 2.0)
    tf = flops * (n_end - n_a) / max(t_end - t_a, 1e-9) / 1e12
    tail = samples[len(samples) // 3:]
    clk = sorted(s[0] for s in tail)[len(tail) // 2]; pw = sorted(s[1] for s in tail)[len(tail) // 2]
    print(f"{name}: sustained {tf:7.1f} TFLOP/s   sm {clk} MHz   {pw:.0f} W", flush=True)


B, N, H, d = 3, 6144, 32, 64
qkv = torch.randn(B, N, 3 * H * d, device=dev).bfloat16()
q, k, v = [qkv[:, :, i * H * d:(i + 1) * H * d].unflatten(-1, (H, d)) for i in range(3)]
run("attention d64 B3 N6144 H32 (this repo)", lambda: ops.attention(q, k, v), 4.0 * B * H * N * N * d)
qt, kt, vt = [x.transpose(1, 2).contiguous() for x in (q, k, v)]
run("attention d64 torch SDPA (library)", lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt), 4.0 * B * H * N * N * d)
if "attn128" in sys.argv:
    B, N, H, d = 1, 32760, 12, 128
    q, k, v = [torch.randn(B, N, H, d, device=dev).bfloat16() for _ in range(3)]
    run("attention d128 B1 N32760 H12 (this repo)", lambda: ops.attention(q, k, v), 4.0 * B * H * N * N * d)
    qt, kt, vt = [x.transpose(1, 2).contiguous() for x in (q, k, v)]
    run("attention d128 torch SDPA (library)", lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt), 4.0 * B * H * N * N * d)
    sys.exit(0)
for (nm, M, Nn, K, act) in [("qkv", 18432, 6144, 2048, 0), ("ffn_up+gelu", 18432, 8192, 2048, ops.ACT_GELU_TANH), ("ffn_down", 18432, 2048, 8192, 0)]:
    a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(Nn, K, device=dev).bfloat16() * 0.02; bias = torch.randn(Nn, device=dev).bfloat16()
    run(f"gemm {nm} {M}x{Nn}x{K} (this repo)", lambda: ops.gemm(a, w, bias, act=act), 2.0 * M * Nn * K)
    if act:
        run(f"gemm {nm} torch linear + gelu (library)", lambda: torch.nn.functional.gelu(torch.nn.functional.linear(a, w, bias), approximate="tanh"), 2.0 * M * Nn * K)
    else:
        run(f"gemm {nm} torch linear (cuBLAS)", lambda: torch.nn.functional.linear(a, w, bias), 2.0 * M * Nn * K)
