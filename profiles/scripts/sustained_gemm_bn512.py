import os, sys, time, threading
sys.path.insert(0, "/root/repo")
import torch, pynvml
from ltx_video_gpupoor_b200 import ops
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
M, Nn, K = 18432, 2048, 8192
a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(Nn, K, device="cuda") * 0.02).bfloat16(); bias = torch.randn(Nn, device="cuda").bfloat16()
fn = lambda: ops.gemm(a, w, bias)
for _ in range(3): fn()
torch.cuda.synchronize()
samples=[]; stop=False
def smp():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h)/1000.0)); time.sleep(0.05)
th=threading.Thread(target=smp); th.start()
t0=time.perf_counter(); n=0; marks=[]
while time.perf_counter()-t0 < 4.0:
    for _ in range(20): fn()
    n+=20; torch.cuda.synchronize(); marks.append((time.perf_counter()-t0, n))
stop=True; th.join()
te,ne=marks[-1]; ta,na=next((t,k) for t,k in marks if t>=te-2.5)
tail=samples[len(samples)//3:]
print(f"BN512={os.environ.get('LTXB200_GEMM_BN512','1')}: ffn_down sustained {2.0*M*Nn*K*(ne-na)/(te-ta)/1e12:.1f} TFLOP/s  sm {sorted(s[0] for s in tail)[len(tail)//2]} MHz  {sorted(s[1] for s in tail)[len(tail)//2]:.0f} W")
