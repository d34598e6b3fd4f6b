"""sustained_probe.py restricted to this repo's two attention shapes (for same-box A/B of builds selected with LTXB200_LIB)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pynvml
from ltx_video_gpupoor_b200 import ops
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
def run(name, fn, flops, secs=3.0):
    samples = []; stop = False
    def sampler():
        while not stop:
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
            time.sleep(0.05)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    th = threading.Thread(target=sampler); th.start()
    t0 = time.perf_counter(); n = 0; marks = []
    while time.perf_counter() - t0 < secs:
        for _ in range(20): fn()
        n += 20
        torch.cuda.synchronize()
        marks.append((time.perf_counter() - t0, n))
    stop = True; th.join()
    t_end, n_end = marks[-1]
    t_a, n_a = next((t, k) for t, k in marks if t >= t_end - 2.0)
    tf = flops * (n_end - n_a) / max(t_end - t_a, 1e-9) / 1e12
    tail = samples[len(samples) // 3:]
    clk = sorted(s[0] for s in tail)[len(tail) // 2]; pw = sorted(s[1] for s in tail)[len(tail) // 2]
    print(f"{name}: sustained {tf:7.1f} TFLOP/s   sm {clk} MHz   {pw:.0f} W", flush=True)
for (B, N, H, d) in ([(1, 32760, 12, 128)] if os.environ.get("ONLY128") else [(3, 6144, 32, 64), (1, 32760, 12, 128)]):
    qkv = torch.randn(B, N, 3 * H * d, device="cuda").bfloat16()
    q, k, v = [qkv[:, :, i * H * d:(i + 1) * H * d].unflatten(-1, (H, d)) for i in range(3)]
    run(f"attention d{d} B{B} N{N} H{H}", lambda: ops.attention(q, k, v), 4.0 * B * H * N * N * d)
