"""Where does the time between kernels go in a Wan2.1-1.3B denoise step?  torch.profiler (CUPTI) trace of 2 steps on ONE GPU: total span,
sum of kernel time, the kernels that are not ours (ATen / NCCL), and the largest idle gaps with the kernels around them."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from ltx_video_gpupoor_b200 import ops
from ltx_video_gpupoor_b200.wan.fm_solvers_unipc import FlowUniPCMultistepScheduler
from ltx_video_gpupoor_b200.wan.posemb_layers import get_rotary_pos_embed

world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", lr)
torch.cuda.set_device(lr)
wl = bench.WAN_WORKLOADS["wan1.3b_832x480x81_sp"]
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
cfgp, group, sp_size = bench.wan_groups(wl, world)
model, cfg = bench.wan_build(wl, dev, cfgp, group)
shape = (16, 21, 60, 104)
g = torch.Generator().manual_seed(42)
ctx = torch.randn(128, 4096, generator=g).to(torch.bfloat16).to(dev)
ctx0 = torch.randn(128, 4096, generator=g).to(torch.bfloat16).to(dev)
lat = torch.randn(*shape, generator=g).to(dev)
freqs = get_rotary_pos_embed(shape[1:]); freqs = (freqs[0].to(dev), freqs[1].to(dev))
sch = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
sch.set_timesteps(50, device=dev, shift=5.0)
scratch = torch.empty(2 * 148, device=dev)


def step(lat, i):
    t = sch.timesteps_host[i]
    if cfgp is not None:
        c, u = cfgp.exchange(model([lat], t=torch.tensor([t], device=dev), context=[cfgp.select(ctx, ctx0)], freqs=freqs, x_id=cfgp.branch)[0])
    else:
        c, u = model([lat, lat], t=torch.tensor([t], device=dev), context=[ctx, ctx0], freqs=freqs)
    pred = ops.cfg_combine(c.contiguous(), u.contiguous(), 5.0, use_alpha=i > 5, scratch=scratch)
    return sch.step(pred.unsqueeze(0), t, lat.unsqueeze(0), return_dict=False)[0].squeeze(0)


for i in range(3):
    lat = step(lat, i)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3, 5):
        lat = step(lat, i)
    torch.cuda.synchronize()
if rank != 0:
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    sys.exit(0)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ev.sort(key=lambda e: e.time_range.start)
span = (ev[-1].time_range.end - ev[0].time_range.start) / 1e3
busy = sum(e.time_range.end - e.time_range.start for e in ev) / 1e3
print(f"2 steps: span {span:.2f} ms, sum of kernel/memcpy time {busy:.2f} ms, idle {span - busy:.2f} ms over {len(ev)} device activities")
agg = {}
for e in ev:
    d = agg.setdefault(e.name[:90], [0, 0.0]); d[0] += 1; d[1] += (e.time_range.end - e.time_range.start) / 1e3
print("-- device activities that are not b200:: kernels")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if "b200" not in k:
        print(f"   {ms:8.3f} ms  x{n:4d}  {k}")
print("-- b200 kernels")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if "b200" in k:
        print(f"   {ms:8.3f} ms  x{n:4d}  {k}")
gaps = []
for a, b in zip(ev, ev[1:]):
    gaps.append(((b.time_range.start - a.time_range.end) / 1e3, a.name[:60], b.name[:60]))
gaps.sort(reverse=True)
print("-- largest gaps (ms, after, before)")
for gp in gaps[:25]:
    print(f"   {gp[0]:7.3f}  {gp[1]}  ->  {gp[2]}")
hist = {}
for gp in gaps:
    b = "<2us" if gp[0] < 0.002 else "<5us" if gp[0] < 0.005 else "<10us" if gp[0] < 0.01 else "<50us" if gp[0] < 0.05 else ">=50us"
    d = hist.setdefault(b, [0, 0.0]); d[0] += 1; d[1] += gp[0]
print("-- gap histogram:", {k: (v[0], round(v[1], 3)) for k, v in hist.items()})

if world > 1:
    dist.barrier(); dist.destroy_process_group()
