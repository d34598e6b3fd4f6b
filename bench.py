#!/usr/bin/env python
"""bench.py — headline benchmark: LTX-Video 2B t2v 768x512x121, 30-step schedule, bf16, on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

One "step" = one denoise step of the reference loop (pipeline_ltx_video.py:1104-1256): all num_conds
transformer forwards (joint batch) + guidance + scheduler update.  `value` = denoise steps/s with all inputs
resident in HBM; `e2e` = the same metric through LTXVideoPipeline.__call__ with pinned HOST prompt embeddings
and a device->host read of the result.  LTX is single-GPU whole-model (replicas only): with N>1
every rank denoises its own video, no data-path collective, value = N*K / max-over-ranks time ("weak").
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1] with the 2B preset guidance (ltx_video/configs/ltxv-2b-0.9.6-dev.yaml:3-8)
    "ltx2b_768x512x121_cfg_stg": dict(height=512, width=768, num_frames=121, frame_rate=25.0, schedule_steps=30,
                                      guidance_scale=3.0, stg_scale=1.0, rescaling_scale=0.7, skip_block_list=[19],
                                      prompt_tokens=256, num_conds=3),
    "ltx2b_768x512x121_noguidance": dict(height=512, width=768, num_frames=121, frame_rate=25.0, schedule_steps=30,
                                         guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0, skip_block_list=None,
                                         prompt_tokens=256, num_conds=1),
}
LAYER_FLOPS = 1.048e12      # per layer per cond at N=6144
FWD_FLOPS = 29.36e12


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_sample(wl, layers=(1, 2)):
    """Bounded sample of one denoise step on the CPU: full-size (N=6144, L=256) transformer forward of the oracle
    with 1 and 2 layers; per-layer and fixed costs are separated and extrapolated to 28 layers x num_conds."""
    from oracle import ltx_oracle as O
    torch.manual_seed(0)
    f, h, w = wl["num_frames"] // 8 + 1, wl["height"] // 32, wl["width"] // 32
    N = f * h * w
    hidden = torch.randn(1, N, 128)
    enc = torch.randn(1, wl["prompt_tokens"], 4096)
    mask = torch.ones(1, wl["prompt_tokens"])
    coords = O.latent_to_pixel_coords(O.latent_coords(f, h, w, 1)).float()
    coords[:, 0] /= wl["frame_rate"]
    cs = O.precompute_freqs_cis(coords, 2048, 10000.0, (20, 2048, 2048))
    ts = torch.full((1, 1), 0.5)
    times = {}
    with torch.no_grad():
        for L in layers:
            sd = O.make_transformer_state_dict(O.LTX_2B, seed=0, num_layers=L)
            t0 = time.perf_counter()
            O.transformer_forward(sd, O.LTX_2B, hidden, cs, enc, ts, mask, None, None, (f, h, w))
            times[L] = time.perf_counter() - t0
            del sd
    t_layer = max(times[layers[1]] - times[layers[0]], 1e-9) / (layers[1] - layers[0])
    t_fixed = max(times[layers[0]] - layers[0] * t_layer, 0.0)
    step_s = wl["num_conds"] * (t_fixed + 28 * t_layer)
    return dict(step_s=step_s, t_layer=t_layer, t_fixed=t_fixed, raw=times, tokens=N)


def run_reference(args, wl_name, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = torch.get_num_threads()
    vals = []
    for i in range(args.warmup + args.steps):
        s = cpu_sample(wl, layers=(1, 2))
        if i >= args.warmup:
            vals.append(s["step_s"])
    step_s = sum(vals) / len(vals)
    v = 1.0 / step_s
    sample = (f"oracle port (torch fp32, {cores} threads) of Transformer3DModel.forward at full size N={s['tokens']}, "
              f"L={wl['prompt_tokens']}: 1- and 2-layer forwards timed, extrapolated to 28 layers x {wl['num_conds']} conds per step")
    line = {"impl": "reference", "metric": "denoise_steps_per_s", "value": v, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": wl_name, **{k: wl[k] for k in ("height", "width", "num_frames", "schedule_steps", "num_conds")}},
            "cpu_baseline": {"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ltx2b_768x512x121_cfg_stg", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--layers", type=int, default=28, help="debug only: fewer layers makes the number INVALID")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)
    wl = WORKLOADS[args.workload]

    if args.impl == "reference":
        return run_reference(args, args.workload, wl)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from ltx_video_gpupoor_b200 import _lib, ops
    from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import CausalVideoAutoencoder, vae_decode
    from ltx_video_gpupoor_b200.ltx.init_weights import random_transformer_state_dict, random_vae_decoder_state_dict
    from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import LTXVideoPipeline
    from ltx_video_gpupoor_b200.ltx.rf import RectifiedFlowScheduler
    from ltx_video_gpupoor_b200.ltx.skip_layer_strategy import SkipLayerStrategy
    from ltx_video_gpupoor_b200.ltx.symmetric_patchifier import SymmetricPatchifier
    from ltx_video_gpupoor_b200.ltx.transformer3d import LTX_2B_CONFIG, Transformer3DModel

    cfg = dict(LTX_2B_CONFIG, num_layers=args.layers)
    tr = Transformer3DModel(**cfg)
    tr.load_state_dict(random_transformer_state_dict(cfg, seed=0, device=dev), device=dev)
    vae = CausalVideoAutoencoder()
    vae.load_state_dict(random_vae_decoder_state_dict(seed=1, device=dev), device=dev)
    pipe = LTXVideoPipeline(vae=vae, transformer=tr, scheduler=RectifiedFlowScheduler(), patchifier=SymmetricPatchifier(1))

    # synthetic prompt embeddings of the named shape, in pinned host memory (e2e copies them every call)
    g = torch.Generator().manual_seed(42 + rank)
    Lp = wl["prompt_tokens"]
    pe_h = torch.randn(1, Lp, 4096, generator=g).to(torch.bfloat16).pin_memory()
    ne_h = torch.randn(1, Lp, 4096, generator=g).to(torch.bfloat16).pin_memory()
    pm_h = torch.ones(1, Lp).pin_memory()
    nm_h = torch.ones(1, Lp).pin_memory()
    call_kw = dict(height=wl["height"], width=wl["width"], num_frames=wl["num_frames"], frame_rate=wl["frame_rate"],
                   guidance_scale=wl["guidance_scale"], stg_scale=wl["stg_scale"], rescaling_scale=wl["rescaling_scale"],
                   skip_block_list=wl["skip_block_list"],
                   skip_layer_strategy=SkipLayerStrategy.AttentionValues if wl["skip_block_list"] else None,
                   is_video=True, vae_per_channel_normalize=True, return_dict=False)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timed loop ----------------
    st = pipe(prompt_embeds=pe_h.to(dev), prompt_attention_mask=pm_h.to(dev), negative_prompt_embeds=ne_h.to(dev),
              negative_prompt_attention_mask=nm_h.to(dev), num_inference_steps=wl["schedule_steps"],
              generator=torch.Generator(device=dev).manual_seed(42), output_type="latent", _prepare_only=True, **call_kw)
    S = wl["schedule_steps"]
    for i in range(args.warmup):
        pipe.denoise_step(st, i % S)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        pipe.denoise_step(st, (args.warmup + i) % S)
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = _lib.launch_count() - l0
    elapsed = e0.elapsed_time(e1) / 1e3
    if dist is not None:
        t = torch.tensor([elapsed], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t)
    steps_per_s = world * args.steps / elapsed

    # ---------------- roofline probe: one instrumented step (not part of the timed region) ----------------
    ops.PROFILER = []
    pipe.denoise_step(st, 0)
    torch.cuda.synchronize()
    prof, ops.PROFILER = ops.PROFILER, None
    agg = {}
    for name, kind, amount, a, b in prof:
        d = agg.setdefault(name, dict(kind=kind, amount=0.0, ms=0.0, n=0))
        d["amount"] += amount; d["ms"] += a.elapsed_time(b); d["n"] += 1
    total_ms = sum(d["ms"] for d in agg.values())
    top = max(agg, key=lambda k: agg[k]["ms"])
    pk = peaks()
    d = agg[top]
    if d["kind"] == "flop":
        achieved, peak, unit, bound = d["amount"] / (d["ms"] * 1e-3) / 1e12, pk["tf_sust"], "TFLOP/s", "tensor"
    else:
        achieved, peak, unit, bound = d["amount"] / (d["ms"] * 1e-3) / 1e9, pk["hbm"], "GB/s", "hbm"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(top)
        except Exception:
            traffic = None
    roofline = {"kernel": top, "bound": bound, "achieved": achieved, "peak": peak, "peak_source": pk["src"] + " (sustained: kernel timed inside a long step)",
                "unit": unit, "frac": achieved / peak, "traffic": traffic, "launches_per_step": d["n"],
                "avg_launch_ms": d["ms"] / d["n"], "share_of_step": d["ms"] / total_ms,
                "algorithmic_per_launch": d["amount"] / d["n"]}
    kernels = {k: {"ms_per_step": round(v["ms"], 3), "launches": v["n"], "share": round(v["ms"] / total_ms, 4),
                   ("tflops" if v["kind"] == "flop" else "gbs"): round(v["amount"] / (v["ms"] * 1e-3) / (1e12 if v["kind"] == "flop" else 1e9), 1)}
               for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}

    # ---------------- end-to-end through the public pipeline call, host buffers in the timed region ----------------
    K = max(args.steps, 2)
    pipe(prompt_embeds=pe_h, prompt_attention_mask=pm_h, negative_prompt_embeds=ne_h, negative_prompt_attention_mask=nm_h,
         num_inference_steps=2, generator=torch.Generator(device=dev).manual_seed(1), output_type="latent", **call_kw)
    barrier()
    t0 = time.perf_counter()
    lat = pipe(prompt_embeds=pe_h, prompt_attention_mask=pm_h, negative_prompt_embeds=ne_h,
               negative_prompt_attention_mask=nm_h, num_inference_steps=K,
               generator=torch.Generator(device=dev).manual_seed(2), output_type="latent", **call_kw)[0]
    lat_h = lat.to("cpu", non_blocking=False)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t)
    h2d = (pe_h.numel() + ne_h.numel()) * 2 + (pm_h.numel() + nm_h.numel()) * 4
    e2e = {"value": world * K / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": h2d / K,
           "d2h_bytes_per_step": lat_h.numel() * lat_h.element_size() / K, "steps_in_call": K,
           "api": "LTXVideoPipeline.__call__(prompt_embeds=<pinned host>, output_type='latent') + .cpu()"}

    # ---------------- VAE decode (s/video = 30 denoise steps + decode) ----------------
    decode_s = None
    if not args.no_decode:
        z = torch.randn(1, 128, wl["num_frames"] // 8 + 1, wl["height"] // 32, wl["width"] // 32, device=dev)
        vae_decode(z, vae, True, vae_per_channel_normalize=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        img = vae_decode(z, vae, True, vae_per_channel_normalize=True)
        b.record()
        torch.cuda.synchronize()
        decode_s = a.elapsed_time(b) / 1e3
        del img

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        s = cpu_sample(wl)
        cores = torch.get_num_threads()
        cpu_baseline = {"value": 1.0 / s["step_s"], "unit": "steps/s", "cores": cores, "kind": "port",
                        "sample": (f"oracle port (torch fp32, {cores} threads) of the transformer forward at full size "
                                   f"N={s['tokens']}: 1- and 2-layer forwards ({s['raw'][1]:.1f}s, {s['raw'][2]:.1f}s) "
                                   f"extrapolated to 28 layers x {wl['num_conds']} conds per denoise step")}

    ms_step = elapsed / args.steps * 1e3
    line = {
        "metric": "denoise_steps_per_s", "value": steps_per_s, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": args.workload, "model": "LTX-Video 2B (random-init, 28 layers)" if args.layers == 28 else f"INVALID: {args.layers} layers",
                   "height": wl["height"], "width": wl["width"], "num_frames": wl["num_frames"], "tokens": st.N,
                   "schedule_steps": S, "num_conds": wl["num_conds"], "prompt_tokens": Lp, "parallelism": f"replicas x{world}",
                   "l2_policy": "per-step working set (3.8 GB weights + activations) far exceeds the 126 MB L2; no flush needed"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "s_per_video": (S * ms_step / 1e3 + decode_s) if decode_s is not None else None, "vae_decode_s": decode_s,
        "model_tflops": wl["num_conds"] * FWD_FLOPS * (args.layers / 28) / (ms_step / 1e3) / 1e12,
        "kernels": kernels,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
