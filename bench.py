#!/usr/bin/env python
"""bench.py — headline benchmark: LTX-Video 2B t2v 768x512x121, 30-step schedule, bf16, on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

One "step" = one denoise step of the reference loop (pipeline_ltx_video.py:1104-1256): all num_conds
transformer forwards (joint batch) + guidance + scheduler update.  `value` = denoise steps/s with all inputs
resident in HBM; `e2e` = the same metric through LTXVideoPipeline.__call__ with pinned HOST prompt embeddings
and a device->host read of the result.  LTX is single-GPU whole-model (replicas only): with N>1
every rank denoises its own video, no data-path collective, value = N*K / max-over-ranks time ("weak").

The part of the path that SHARDS is Wan's Ulysses sequence parallelism (BASELINE.json configs[3]); every run of the default workload
therefore also carries a `wan_sp` block: ONE Wan2.1-1.3B 832x480x81 video split over all N ranks (strong scaling; N = 1 is the
single-GPU step, so the driver's 1/2/4/8 runs form a self-contained curve), preceded by an in-run parity check of the
sequence-parallel forward against the fixture recorded from the unmodified reference (tests/golden/wan_t2v_h4.pt).
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
    # BASELINE.json configs[2]: i2v — the conditioning image goes through the VAE encoder, per-token timesteps, full decode
    "ltx2b_768x512x121_i2v_cfg_stg": dict(height=512, width=768, num_frames=121, frame_rate=25.0, schedule_steps=30,
                                          guidance_scale=3.0, stg_scale=1.0, rescaling_scale=0.7, skip_block_list=[19],
                                          prompt_tokens=256, num_conds=3, i2v=True),
    # the model the reference APP actually loads (ltx_video/ltxv.py:171-194: LTX-Video 13B 0.9.7, 48 layers, 32 heads x 128): same step, 13B widths
    "ltx13b_768x512x121_cfg_stg": dict(height=512, width=768, num_frames=121, frame_rate=25.0, schedule_steps=30,
                                       guidance_scale=3.0, stg_scale=1.0, rescaling_scale=0.7, skip_block_list=[28],
                                       prompt_tokens=256, num_conds=3, arch="13b"),
    "ltx2b_768x512x121_noguidance": dict(height=512, width=768, num_frames=121, frame_rate=25.0, schedule_steps=30,
                                         guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0, skip_block_list=None,
                                         prompt_tokens=256, num_conds=1),
}
WAN_WORKLOADS = {
    # BASELINE.json configs[3]: Wan2.1 T2V-1.3B 832x480x81, 50 steps, CFG 5.0 (2 forwards/step), Ulysses SP over all ranks
    "wan1.3b_832x480x81_sp": dict(model="1.3B", width=832, height=480, frame_num=81, schedule_steps=50, shift=5.0,
                                  guide_scale=5.0, prompt_tokens=128, fwd_flops=283.0e12),
    # BASELINE.json configs[4]: Wan2.1 T2V-14B 1280x720x81
    "wan14b_1280x720x81_sp": dict(model="14B", width=1280, height=720, frame_num=81, schedule_steps=50, shift=5.0,
                                  guide_scale=5.0, prompt_tokens=128, fwd_flops=6523.0e12),
}
LAYER_FLOPS = 1.048e12      # per layer per cond at N=6144
FWD_FLOPS = 29.36e12
LTX_ARCH = {"2b": dict(layers=28, overrides={}, name="LTX-Video 2B"),
            "13b": dict(layers=48, overrides=dict(attention_head_dim=128, cross_attention_dim=4096), name="LTX-Video 13B")}


def ltx_fwd_flops(wl, layers, tokens):
    """28*N*D^2 (q,k,v,o + cross q,o + FFN 4x) + 4*N^2*D self-attention + 4*N*L*D cross-attention, per layer per cond"""
    D = 32 * (128 if wl.get("arch") == "13b" else 64)
    return layers * (28.0 * tokens * D * D + 4.0 * tokens * tokens * D + 4.0 * tokens * wl["prompt_tokens"] * D)


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
def cpu_sample(wl, layers=(1, 3)):
    """Bounded sample of one denoise step on the CPU: full-size (N=6144, L=256) transformer forward of the oracle
    with 1 and 3 layers (a two-layer difference per sample keeps host timing noise out of the slope); per-layer and
    fixed costs are separated and extrapolated to 28 layers x num_conds."""
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


def ltx_config(workload, wl, layers, world):
    """`config` of the JSON line: identical for the b200 arm and the reference arm."""
    tokens = (wl["num_frames"] // 8 + 1) * (wl["height"] // 32) * (wl["width"] // 32)
    arch = LTX_ARCH[wl.get("arch", "2b")]
    return {"workload": workload, "network": f"{arch['name']} (random-init, {layers} layers)" if layers == arch["layers"] else f"INVALID: {layers} layers",
            "height": wl["height"], "width": wl["width"], "num_frames": wl["num_frames"], "tokens": tokens,
            "schedule_steps": wl["schedule_steps"], "num_conds": wl["num_conds"], "prompt_tokens": wl["prompt_tokens"],
            "parallelism": f"replicas x{world}", "prompt_mask": "all ones (every prompt token valid): zero key bias",
            "l2_policy": "per-step working set (3.8 GB weights + activations) far exceeds the 126 MB L2; no flush needed"}


def split_attention(launches, tf_sust):
    """(flops, ms) per attention launch of one step -> the self-attention launches (N x N keys: the large ones) and the cross-attention
    launches (N x prompt tokens) apart: the step's average mixes a tensor-bound-sized problem with 2-key-block work items."""
    big = max(f for f, _ in launches)
    out = {}
    for key, sel in (("self_attention", [x for x in launches if x[0] > 0.5 * big]), ("cross_attention", [x for x in launches if x[0] <= 0.5 * big])):
        if not sel:
            continue
        fl, ms = sum(f for f, _ in sel), sum(m for _, m in sel)
        tf = fl / (ms * 1e-3) / 1e12
        out[key] = {"launches": len(sel), "ms_per_step": round(ms, 3), "tflops": round(tf, 1), "tensor_frac_of_sustained": round(tf / tf_sust, 3)}
    return out


def host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm uses every host core it can get."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except AttributeError:
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_reference_sample(wl):
    """One bounded sample of a full-size denoise step on the host cores.  Runs the UNMODIFIED reference (oracle/_ref, populated
    from /root/reference by oracle/build_ref.py) when it is there — kind "reference" — and the pinned oracle port otherwise."""
    cores = host_threads()
    from oracle import ref_runner
    if ref_runner.available():
        s = ref_runner.full_size_sample(wl)
        kind, what = "reference", "the unmodified reference Transformer3DModel.forward (oracle/_ref, fp32)"
    else:
        s = cpu_sample(wl)
        kind, what = "port", "oracle port of Transformer3DModel.forward (torch fp32)"
    sample = (f"{what} on {cores} host threads at full size N={s['tokens']}, L={wl['prompt_tokens']}: 1- and 3-layer forwards timed "
              f"({s['raw'][1]:.1f} s, {s['raw'][3]:.1f} s), EXTRAPOLATED to 28 layers x {wl['num_conds']} conds per denoise step")
    return s, {"value": 1.0 / s["step_s"], "unit": "steps/s", "cores": cores, "kind": kind, "sample": sample, "extrapolated": True,
               "sample_wall_s": s.get("sample_s", sum(s["raw"].values()))}


def run_reference(args, wl_name, wl):
    """The reference's own CPU implementation of the path on the box's host cores.  ONE bounded sample per run whatever K is:
    for the 2B workloads ONE WHOLE denoise step of the bench workload, for real — the unmodified reference Transformer3DModel.forward
    at full size, all 28 layers, all guidance conditions in one batch (30-60 s on the 16-32 host threads of a GPU box; nothing
    extrapolated) — plus BASELINE configs[0] — 256x256x9, 4 steps + VAE decode — through the reference's LTXVideoPipeline.
    LTXB200_REF_EXTRAPOLATE=1, the 13B workload or a missing oracle/_ref fall back to the 1- and 3-layer sample, marked extrapolated."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_start = time.perf_counter()
    from oracle import ref_runner
    real = ref_runner.available() and wl.get("arch", "2b") == "2b" and not os.environ.get("LTXB200_REF_EXTRAPOLATE")
    tr = None
    if real:
        cores = host_threads()
        tr = ref_runner.reference_2b_transformer()
        s = ref_runner.full_step_real(wl, tr)
        base = {"value": 1.0 / s["step_s"], "unit": "steps/s", "cores": cores, "kind": "reference", "extrapolated": False,
                "sample": (f"ONE whole denoise step, measured: the unmodified reference Transformer3DModel.forward (oracle/_ref, fp32) on {cores} host "
                           f"threads at full size N={s['tokens']}, L={wl['prompt_tokens']}, all 28 layers, {s['conds']} guidance conditions in one batch "
                           f"(STG skip-layer mask included): {s['step_s']:.1f} s"),
                "sample_wall_s": s["step_s"]}
    else:
        s, base = cpu_reference_sample(wl)
    v, step_s = base["value"], s["step_s"]
    config0 = None
    if not os.environ.get("LTXB200_BENCH_SKIP_CONFIG0"):
        try:
            if ref_runner.available():
                c0 = ref_runner.config0(tr=tr)
                config0 = {"workload": "BASELINE configs[0]: LTX-2B (28 layers) t2v 256x256x9, 4 steps + VAE decode, fp32, host CPU",
                           "measured": True, "extrapolated": False, "steps_per_s": c0["steps_per_s"], "denoise_loop_s": c0["loop_s"],
                           "s_per_video": c0["video_s"], "cores": base["cores"],
                           "api": "unmodified reference LTXVideoPipeline.__call__ (oracle/_ref)"}
        except Exception as exc:                      # never lose the line over the extra
            config0 = {"error": repr(exc)[:300]}
    line = {"impl": "reference", "metric": "denoise_steps_per_s", "value": v, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "extrapolated": bool(base.get("extrapolated")),
            "timed": {"samples": 1, "sample_wall_s": base["sample_wall_s"], "run_wall_s": None,
                      "note": "steps/warmup echo the request; ONE bounded sample is timed per run (see cpu_baseline.sample)"},
            "config": ltx_config(wl_name, wl, LTX_ARCH[wl.get("arch", "2b")]["layers"], args.gpus), "cpu_baseline": base, "config0_real": config0,
            "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["timed"]["run_wall_s"] = time.perf_counter() - t_start
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
# Wan2.1 T2V, Ulysses sequence parallel over all ranks (strong scaling: one video, N GPUs)
# ------------------------------------------------------------------------------------------------------------
def wan_cpu_sample(wl):
    """Bounded sample of one Wan denoise step on the CPU (oracle port, torch fp32): full-width WanModel.forward with 1 and 2
    layers at two reduced token counts; the per-layer cost is fitted as a*N + b*N^2 (projections/FFN vs self-attention)
    and extrapolated to the workload's token count, 30/40 layers and 2 forwards per step."""
    from oracle import wan_oracle as W
    cfg = dict(W.WAN_1_3B if wl["model"] == "1.3B" else W.WAN_14B)
    lat_full = (16, (wl["frame_num"] - 1) // 4 + 1, wl["height"] // 8, wl["width"] // 8)
    N_full = lat_full[1] * lat_full[2] * lat_full[3] // 4
    torch.manual_seed(0)
    t = torch.tensor([500])
    ctx = [torch.randn(wl["prompt_tokens"], 4096)]
    grids = [(4, 32, 32), (4, 32, 64)]            # latent (f, h, w) -> 1024 / 2048 tokens
    per_layer, fixed = {}, {}
    with torch.no_grad():
        for (f, h, w) in grids:
            n = f * h * w // 4
            x = [torch.randn(16, f, h, w)]
            cos, sin = W.rope_tables((f, h, w))
            tt = {}
            for L in (1, 2):
                sd = W.make_wan_state_dict(cfg, seed=0, num_layers=L)
                t0 = time.perf_counter()
                W.wan_forward(sd, cfg, x, t, ctx, cos, sin)
                tt[L] = time.perf_counter() - t0
                del sd
            per_layer[n] = max(tt[2] - tt[1], 1e-9)
            fixed[n] = max(tt[1] - per_layer[n], 0.0)
    (n1, t1), (n2, t2) = sorted(per_layer.items())
    b = max((t2 / n2 - t1 / n1) / (n2 - n1), 0.0)
    a = max(t1 / n1 - b * n1, 0.0)
    layer_full = a * N_full + b * N_full * N_full
    fixed_full = fixed[n2] * N_full / n2
    step_s = 2 * (fixed_full + cfg["num_layers"] * layer_full)
    return dict(step_s=step_s, tokens=N_full, fit=dict(a=a, b=b), sample_tokens=[n1, n2], per_layer_s=[t1, t2])


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def wan_groups(wl, world):
    """Ulysses over all ranks when the heads divide, otherwise (Wan-1.3B on 8 GPUs: 12 heads) cond / uncond on the two halves
    of the world with Ulysses inside each half (wan/distributed/cfg_parallel.py).  -> (cfgp | None, sp_group | None, sp_size)"""
    if world == 1:
        return None, None, 1
    import torch.distributed as dist
    use_cfgp = {"1": True, "0": False}.get(os.environ.get("LTXB200_CFG_PARALLEL", ""), "auto")
    heads = 12 if wl["model"] == "1.3B" else 40
    if use_cfgp is True or (use_cfgp == "auto" and heads % world != 0):
        from ltx_video_gpupoor_b200.wan.distributed.cfg_parallel import CfgParallel
        cfgp = CfgParallel()
        return cfgp, (cfgp.sp_group if world > 2 else None), world // 2
    return None, dist.group.WORLD, world


def wan_sp_parity(dev, world, cfgp, group, sp_size):
    """The 2-layer / 4-head forward of the fixture recorded from the UNMODIFIED reference (tests/golden/wan_t2v_h4.pt, fp64 reference,
    oracle/gen_golden_wan.py) on this run's sequence-parallel group — q/k-norm + RoPE + head scatter in token chunks on the side
    stream, attention with the return scatter, the peer-store head gather — against the single-GPU reference output.  Max over ranks."""
    from ltx_video_gpupoor_b200.wan.init_weights import seeded_wan_state_dict
    from ltx_video_gpupoor_b200.wan.model import WanModel
    from ltx_video_gpupoor_b200.wan.posemb_layers import get_rotary_pos_embed
    g = torch.load(os.path.join(ROOT, "tests", "golden", "wan_t2v_h4.pt"), weights_only=False)
    cfg = g["cfg"]
    if cfg["num_heads"] % sp_size:
        return {"skipped": f"fixture has {cfg['num_heads']} heads, group size {sp_size}"}
    m = WanModel(dim=cfg["dim"], ffn_dim=cfg["ffn_dim"], num_heads=cfg["num_heads"], num_layers=cfg["num_layers"], sp_group=group)
    m.load_state_dict(seeded_wan_state_dict(cfg, seed=g["seed_weights"]), device=dev)
    cos, sin = get_rotary_pos_embed(g["lat"].shape[1:])
    lat, ctx, ctx0 = g["lat"].to(dev), g["ctx"].to(dev), g["ctx0"].to(dev)
    if group is not None:
        ex = m._peer_exchange(2, (lat.shape[1] * lat.shape[2] * lat.shape[3] // 4) // sp_size)
        ex.min_chunk_rows = 8                      # the fixture has 96 tokens: chunk anyway, so the overlapped path is what is checked
    for _ in range(3):                             # several forwards: buffers, parities and epoch flags are reused across calls
        y = m([lat, lat], t=g["t"].to(dev), context=[ctx, ctx0], freqs=(cos, sin))
    torch.cuda.synchronize()
    err = max(rel_l2(a.cpu(), b) for a, b in zip(y, g["fwd"]))
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([err], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        err = float(t)
    m.close()
    return {"sp_parity_rel_l2": err, "tolerance": 2e-2, "ok": bool(err < 2e-2), "group_size": sp_size,
            "exchange": m.sp_exchange if group is not None else None,
            "fixture": "tests/golden/wan_t2v_h4.pt (unmodified reference WanModel.forward, fp64; 2 layers, 4 heads, 96 tokens, 2 sequences)"}


def wan_build(wl, dev, cfgp, group, layers=None):
    from ltx_video_gpupoor_b200.wan.model import WAN_T2V_1_3B, WAN_T2V_14B, WanModel
    cfg = dict(WAN_T2V_1_3B if wl["model"] == "1.3B" else WAN_T2V_14B)
    if layers:
        cfg["num_layers"] = layers
    gen = torch.Generator(device=dev).manual_seed(0)          # same seed on every rank: replicated weights
    D, Fd = cfg["dim"], cfg["ffn_dim"]

    def u(shape, fan_in):
        return ((torch.rand(shape, generator=gen, device=dev) * 2 - 1) / fan_in ** 0.5).to(torch.bfloat16)

    sd = {"patch_embedding.weight": u((D, 16, 1, 2, 2), 64), "patch_embedding.bias": u((D,), 64)}
    for n, o, i in (("text_embedding.0", D, 4096), ("text_embedding.2", D, D), ("time_embedding.0", D, 256),
                    ("time_embedding.2", D, D), ("time_projection.1", 6 * D, D), ("head.head", 64, D)):
        sd[n + ".weight"], sd[n + ".bias"] = u((o, i), i), u((o,), i)
    sd["head.modulation"] = (torch.randn(1, 2, D, generator=gen, device=dev) / D ** 0.5).to(torch.bfloat16)
    for li in range(cfg["num_layers"]):
        p = f"blocks.{li}."
        for a in ("self_attn", "cross_attn"):
            for n in ("q", "k", "v", "o"):
                sd[p + a + "." + n + ".weight"], sd[p + a + "." + n + ".bias"] = u((D, D), D), u((D,), D)
            sd[p + a + ".norm_q.weight"] = torch.ones(D, device=dev, dtype=torch.bfloat16)
            sd[p + a + ".norm_k.weight"] = torch.ones(D, device=dev, dtype=torch.bfloat16)
        sd[p + "norm3.weight"] = torch.ones(D, device=dev, dtype=torch.bfloat16)
        sd[p + "norm3.bias"] = torch.zeros(D, device=dev, dtype=torch.bfloat16)
        sd[p + "ffn.0.weight"], sd[p + "ffn.0.bias"] = u((Fd, D), D), u((Fd,), D)
        sd[p + "ffn.2.weight"], sd[p + "ffn.2.bias"] = u((D, Fd), Fd), u((D,), Fd)
        sd[p + "modulation"] = (torch.randn(1, 6, D, generator=gen, device=dev) / D ** 0.5).to(torch.bfloat16)
    model = WanModel(**{k: v for k, v in cfg.items() if k not in ("qk_norm", "cross_attn_norm")}, sp_group=group)
    model.load_state_dict(sd, device=dev)
    del sd
    return model, cfg


def wan_measure(wl, dev, world, local_rank, model, cfgp, steps, warmup, e2e=True):
    """Timed steps of ONE video split over all ranks (strong scaling), device-resident, CUDA events, max over ranks; then one
    instrumented step for the per-kernel table and (optionally) the end-to-end loop with host buffers."""
    from ltx_video_gpupoor_b200 import _lib, ops
    from ltx_video_gpupoor_b200.wan.fm_solvers_unipc import FlowUniPCMultistepScheduler
    from ltx_video_gpupoor_b200.wan.posemb_layers import get_rotary_pos_embed
    dist = None
    if world > 1:
        import torch.distributed as dist
    shape = (16, (wl["frame_num"] - 1) // 4 + 1, wl["height"] // 8, wl["width"] // 8)
    g = torch.Generator().manual_seed(42)
    ctx_h = torch.randn(wl["prompt_tokens"], 4096, generator=g).to(torch.bfloat16).pin_memory()
    ctx0_h = torch.randn(wl["prompt_tokens"], 4096, generator=g).to(torch.bfloat16).pin_memory()
    noise_h = torch.randn(*shape, generator=g).pin_memory()
    freqs = get_rotary_pos_embed(shape[1:])
    freqs = (freqs[0].to(dev), freqs[1].to(dev))
    S = wl["schedule_steps"]
    scratch = torch.empty(2 * 148, device=dev)

    def run_steps(lat, ctx, ctx0, sch, idx):
        for i in idx:
            t = sch.timesteps_host[i]
            if cfgp is not None:
                c, uu = cfgp.exchange(model([lat], t=torch.tensor([t], device=dev), context=[cfgp.select(ctx, ctx0)], freqs=freqs,
                                            x_id=cfgp.branch)[0])
            else:
                c, uu = model([lat, lat], t=torch.tensor([t], device=dev), context=[ctx, ctx0], freqs=freqs)
            pred = ops.cfg_combine(c.contiguous(), uu.contiguous(), wl["guide_scale"], use_alpha=i > 5, scratch=scratch)
            lat = sch.step(pred.unsqueeze(0), t, lat.unsqueeze(0), return_dict=False)[0].squeeze(0)
        return lat

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sch = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    sch.set_timesteps(S, device=dev, shift=wl["shift"])
    ctx, ctx0, lat = ctx_h.to(dev), ctx0_h.to(dev), noise_h.to(dev)
    lat = run_steps(lat, ctx, ctx0, sch, range(warmup))
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if os.environ.get("LTXB200_NCU_RANGE"):
        torch.cuda.profiler.start()
    e0.record()
    lat = run_steps(lat, ctx, ctx0, sch, range(warmup, warmup + steps))
    e1.record()
    if os.environ.get("LTXB200_NCU_RANGE"):
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    barrier()
    clocks = sampler.stop()
    launches = _lib.launch_count() - l0
    elapsed = e0.elapsed_time(e1) / 1e3
    if dist is not None:
        tt = torch.tensor([elapsed], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed = float(tt)

    ops.PROFILER = []                               # instrumented step: single stream (the chunk overlap is off), events per launch
    run_steps(lat, ctx, ctx0, sch, [warmup + steps])
    torch.cuda.synchronize()
    prof, ops.PROFILER = ops.PROFILER, None
    agg = {}
    attn_launches = []                              # (flops, ms) per attention launch: self- and cross-attention are reported apart below
    for name, kind, amount, a, b in prof:
        d = agg.setdefault(name, dict(kind=kind, amount=0.0, ms=0.0, n=0))
        ms_l = a.elapsed_time(b)
        d["amount"] += amount; d["ms"] += ms_l; d["n"] += 1
        if name == "attention_bf16":
            attn_launches.append((amount, ms_l))
    total_ms = sum(d["ms"] for d in agg.values())
    top = max(agg, key=lambda k: agg[k]["ms"])
    pk = peaks()
    d = agg[top]
    achieved = d["amount"] / (d["ms"] * 1e-3) / (1e12 if d["kind"] == "flop" else 1e9)
    peak = pk["tf_sust"] if d["kind"] == "flop" else pk["hbm"]
    roofline = {"kernel": top, "bound": "tensor" if d["kind"] == "flop" else "hbm", "achieved": achieved, "peak": peak,
                "peak_source": pk["src"], "unit": "TFLOP/s" if d["kind"] == "flop" else "GB/s", "frac": achieved / peak,
                "traffic": None, "launches_per_step": d["n"], "share_of_kernel_time": d["ms"] / total_ms}
    kernels = {k: {"ms_per_step": round(v["ms"], 3), "launches": v["n"], "share": round(v["ms"] / total_ms, 4),
                   ("tflops" if v["kind"] == "flop" else "gbs"): round(v["amount"] / (v["ms"] * 1e-3) / (1e12 if v["kind"] == "flop" else 1e9), 1)}
               for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
    try:
        if "attention_bf16" in kernels and attn_launches:
            kernels["attention_bf16"].update(split_attention(attn_launches, pk["tf_sust"]))
    except Exception as exc:                        # never lose the line over an extra
        kernels["attention_bf16"]["split_error"] = repr(exc)[:200]
    out = dict(elapsed=elapsed, ms_per_step=elapsed / steps * 1e3, steps_per_s=steps / elapsed, launches=int(launches), clocks=clocks,
               roofline=roofline, kernels=kernels, serialised_kernel_ms=total_ms, shape=shape)
    if e2e:
        # host noise + host prompt embeddings in, latents back to the host, K steps of a K-step schedule
        K = max(steps, 2)
        barrier()
        t0 = time.perf_counter()
        sch2 = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
        sch2.set_timesteps(K, device=dev, shift=wl["shift"])
        o = run_steps(noise_h.to(dev, non_blocking=True), ctx_h.to(dev, non_blocking=True), ctx0_h.to(dev, non_blocking=True), sch2, range(K))
        out_h = o.cpu()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([e2e_s], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_s = float(tt)
        out["e2e"] = {"value": K / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": (noise_h.numel() * 4 + 2 * ctx_h.numel() * 2) / K,
                      "d2h_bytes_per_step": out_h.numel() * 4 / K, "steps_in_call": K}
    return out


def wan_parallelism(world, cfgp):
    return f"cfg-parallel 2 x ulysses sp{world // 2}" if cfgp is not None else f"ulysses sp{world}"


def wan_sp_block(dev, world, rank, local_rank, steps=10, warmup=3):
    """The `wan_sp` block of the default (LTX) line: BASELINE configs[3] — Wan2.1 T2V-1.3B 832x480x81 — one video over all N ranks."""
    wl = WAN_WORKLOADS["wan1.3b_832x480x81_sp"]
    cfgp, group, sp_size = wan_groups(wl, world)
    parity = wan_sp_parity(dev, world, cfgp, group, sp_size)
    model, cfg = wan_build(wl, dev, cfgp, group)
    r = wan_measure(wl, dev, world, local_rank, model, cfgp, steps, warmup, e2e=False)
    model.close()
    del model
    torch.cuda.empty_cache()
    k = r["kernels"]
    pick = lambda n: k.get(n, {}).get("ms_per_step")
    return {"workload": "wan1.3b_832x480x81_sp (BASELINE configs[3]): ONE video over all ranks, strong scaling",
            "network": f"Wan2.1-T2V-1.3B (random-init, {cfg['num_layers']} layers)", "tokens": r["shape"][1] * r["shape"][2] * r["shape"][3] // 4,
            "n_gpus": world, "parallelism": wan_parallelism(world, cfgp), "steps": steps, "warmup": warmup,
            "ms_per_step": r["ms_per_step"], "steps_per_s": r["steps_per_s"], "s_per_50_step_video_denoise": 50 * r["ms_per_step"] / 1e3,
            "scaling": "strong", "forwards_per_step": 2,
            "exchange": (os.environ.get("LTXB200_SP_EXCHANGE", "p2p") + f" (peer-memory stores over NVLink fused into the producing kernels; "
                         f"token chunks per exchange: {os.environ.get('LTXB200_SP_CHUNKS', '1')})") if sp_size > 1 else None,
            "exchange_kernels_ms_per_step_serialised": {"qk_norm_rope_wan_scatter": pick("qk_norm_rope_wan_scatter_bf16"), "comm_wait": pick("comm_wait"),
                                                        "peer_allgather": pick("peer_allgather")},
            "serialised_kernel_ms": r["serialised_kernel_ms"], "gpu_launches": r["launches"], "clocks": r["clocks"],
            "kernels": k, "parity": parity, "sp_parity_rel_l2": parity.get("sp_parity_rel_l2"),
            "model_tflops_per_gpu": 2 * wl["fwd_flops"] / (r["ms_per_step"] / 1e3) / 1e12 / world}


def run_wan(args, wl):
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        cores = host_threads()
        smp = wan_cpu_sample(wl)
        step_s = smp["step_s"]
        v = 1.0 / step_s
        sample = (f"oracle port (torch fp32, {cores} threads) of WanModel.forward, full width, 1- and 2-layer forwards at "
                  f"{smp['sample_tokens']} tokens; per-layer cost fitted a*N + b*N^2 and EXTRAPOLATED to N={smp['tokens']}, all layers, 2 forwards/step")
        print(json.dumps({"impl": "reference", "metric": "denoise_steps_per_s", "value": v, "unit": "steps/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "extrapolated": True,
                          "config": {"workload": args.workload, "network": f"Wan2.1-T2V-{wl['model']}", "tokens": smp["tokens"]},
                          "cpu_baseline": {"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample, "extrapolated": True},
                          "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfgp, group, sp_size = wan_groups(wl, world)
    parity = wan_sp_parity(dev, world, cfgp, group, sp_size) if wl["model"] == "1.3B" or sp_size in (1, 2, 4) else None
    model, cfg = wan_build(wl, dev, cfgp, group, layers=args.layers if args.layers != 28 else None)
    r = wan_measure(wl, dev, world, local_rank, model, cfgp, args.steps, args.warmup)
    shape, S = r["shape"], wl["schedule_steps"]
    decode_s = None
    if rank == 0 and not args.no_decode and wl["model"] == "1.3B":
        # WanVAE.decode of the full latent video on one GPU (text2video.py:590): part of s/video
        from ltx_video_gpupoor_b200.wan.vae import WanVAE
        from ltx_video_gpupoor_b200.wan.init_weights import random_wan_vae_decoder_state_dict
        vae = WanVAE(device=dev)
        vae.load_state_dict(random_wan_vae_decoder_state_dict(seed=1), device=dev)
        z = torch.randn(*shape, device=dev)
        vae.decode([z], 0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        img = vae.decode([z], 0)[0]
        b.record()
        torch.cuda.synchronize()
        decode_s = a.elapsed_time(b) / 1e3
        del img, vae
    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            cores = host_threads()
            smp = wan_cpu_sample(wl)
            cpu_baseline = {"value": 1.0 / smp["step_s"], "unit": "steps/s", "cores": cores, "kind": "port", "extrapolated": True,
                            "sample": (f"oracle port (torch fp32, {cores} threads) of WanModel.forward, full width, 1- and 2-layer forwards at "
                                       f"{smp['sample_tokens']} tokens; per-layer cost fitted a*N + b*N^2, extrapolated to N={smp['tokens']}, all layers, 2 forwards/step")}
        ms_step = r["ms_per_step"]
        line = {"metric": "denoise_steps_per_s", "value": r["steps_per_s"], "unit": "steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": args.workload, "network": f"Wan2.1-T2V-{wl['model']} (random-init, {cfg['num_layers']} layers)",
                           "latent": list(shape), "tokens": shape[1] * shape[2] * shape[3] // 4, "schedule_steps": S,
                           "forwards_per_step": 2, "parallelism": wan_parallelism(world, cfgp),
                           "sp_exchange": (os.environ.get("LTXB200_SP_EXCHANGE", "p2p") + (" (fused peer-memory stores over NVLink)" if os.environ.get("LTXB200_SP_EXCHANGE", "p2p") == "p2p" else " (all_to_all_single)")) if world > 1 else None,
                           "sp_chunks": int(os.environ.get("LTXB200_SP_CHUNKS", "1")) if sp_size > 1 else None,
                           "l2_policy": "per-step working set (weights + activations) far exceeds the 126 MB L2"},
                "e2e": r["e2e"], "gpu_launches": r["launches"], "clocks": r["clocks"], "roofline": r["roofline"], "cpu_baseline": cpu_baseline,
                "s_per_video_denoise": S * ms_step / 1e3, "vae_decode_s": decode_s,
                "s_per_video": (S * ms_step / 1e3 + decode_s) if decode_s is not None else None,
                "model_tflops_per_gpu": 2 * wl["fwd_flops"] * (cfg["num_layers"] / (30 if wl["model"] == "1.3B" else 40)) / (ms_step / 1e3) / 1e12 / world,
                "parity": parity, "kernels": r["kernels"]}
        print(json.dumps(line))
    model.close()
    if dist is not None:
        dist.destroy_process_group()

def ltx_cond_parallel_block(pipe, wl, call_kw, dev, dist, rank, world, steps, warmup, one_gpu_ms):
    """ONE LTX video over min(world, num_conds) GPUs: the [uncond, text, perturbed] rows of every denoise step on different ranks
    (ltx/distributed/cond_parallel.py), predictions exchanged once per step by NCCL broadcast.  Strong scaling of the LTX path's latency —
    the headline `value` stays replicas (throughput).  Every rank takes part in creating the group; ranks beyond it sit the block out."""
    import torch
    from ltx_video_gpupoor_b200.ltx.distributed.cond_parallel import cond_partition
    P = min(world, wl["num_conds"])
    group = dist.new_group(list(range(P)))
    if rank >= P:
        return None
    g = torch.Generator().manual_seed(4242)                  # the SAME inputs on every rank of the group
    Lp = wl["prompt_tokens"]
    pe, ne = torch.randn(1, Lp, 4096, generator=g).to(torch.bfloat16).to(dev), torch.randn(1, Lp, 4096, generator=g).to(torch.bfloat16).to(dev)
    pm = torch.ones(1, Lp, device=dev)
    S = wl["schedule_steps"]

    def prepare(cp):
        return pipe(prompt_embeds=pe, prompt_attention_mask=pm, negative_prompt_embeds=ne, negative_prompt_attention_mask=pm,
                    num_inference_steps=S, generator=torch.Generator(device=dev).manual_seed(7), output_type="latent", _prepare_only=True,
                    **(dict(call_kw, cond_parallel_group=group) if cp else call_kw))

    # parity first: two steps on the group against two steps of the whole batch on rank 0, same seed -> the same bits
    st = prepare(True)
    for i in range(2):
        pipe.denoise_step(st, i)
    identical = None
    if rank == 0:
        st1 = prepare(False)
        for i in range(2):
            pipe.denoise_step(st1, i)
        torch.cuda.synchronize()
        identical = bool(torch.equal(st.lat32, st1.lat32))
        del st1
    for i in range(warmup):
        pipe.denoise_step(st, (2 + i) % S)
    dist.barrier(group=group)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        pipe.denoise_step(st, (2 + warmup + i) % S)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    ms = float(t)
    n_tok = st.N * st.C * 2
    return {"workload": "ONE video of the headline workload over the ranks of a group: guidance conditions split, strong scaling of latency",
            "gpus": P, "conds_per_rank": [hi - lo for lo, hi in cond_partition(wl["num_conds"], P)], "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "steps_per_s": 1e3 / ms, "s_per_video_denoise": S * ms / 1e3, "one_gpu_ms_per_step": one_gpu_ms,
            "speedup_vs_one_gpu": one_gpu_ms / ms, "strong_efficiency": one_gpu_ms / ms / P,
            "exchange": f"NCCL broadcast of every owner's prediction rows once per step ({n_tok} B of bf16 per condition)",
            "bit_identical_to_single_gpu": identical, "scaling": "strong"}


def gpu_library_baseline(wl, cfg, tr, dev, ours_ms):
    """Per-step time of the library path on this GPU: the pinned oracle restatement of the reference's PyTorch modules in bf16 with
    torch.nn.functional.scaled_dot_product_attention (what LTX-Video-GPUPoor executes with `_attention = "sdpa"`), same random
    weights, same shapes as the timed step.  A reported baseline, measured, not extrapolated."""
    import torch.nn.functional as F
    from ltx_video_gpupoor_b200.ltx.init_weights import random_transformer_state_dict
    from oracle import ltx_oracle as O
    BF = torch.bfloat16
    sd = random_transformer_state_dict(cfg, seed=0, device=dev)
    B, Lp = wl["num_conds"], wl["prompt_tokens"]
    f, h, w = wl["num_frames"] // 8 + 1, wl["height"] // 32, wl["width"] // 32
    N = f * h * w
    g = torch.Generator().manual_seed(3)
    hidden = torch.randn(1, N, 128, generator=g).expand(B, N, 128).contiguous().to(dev, BF)
    enc = torch.randn(B, Lp, 4096, generator=g).to(dev, BF)
    mask = torch.ones(B, Lp, device=dev)
    t = torch.full((B, 1), 0.7, device=dev)
    coords = O.latent_to_pixel_coords(O.latent_coords(f, h, w, 1)).float()
    coords[:, 0] *= 1.0 / wl["frame_rate"]
    cos, sin = O.precompute_freqs_cis(coords.to(dev), 2048, 10000.0, (20, 2048, 2048), out_dtype=BF)

    def sdpa(q, k, v, bias=None):                  # utils/attention.py:99-116 on the GPU: torch SDPA on [B, H, L, d]
        return F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                              attn_mask=None if bias is None else bias.to(q.dtype)).transpose(1, 2)

    saved, O.attention_core = O.attention_core, sdpa
    try:
        with torch.no_grad():
            fwd = lambda: O.transformer_forward(sd, O.LTX_2B, hidden, (cos, sin), enc, t, mask, latent_shape=(f, h, w))
            y_ref = fwd()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                y_ref = fwd()
            b.record()
            torch.cuda.synchronize()
            lib_ms = a.elapsed_time(b) / 3
            y = tr(hidden, freqs_cis=(cos, sin), encoder_hidden_states=enc, timestep=t, encoder_attention_mask=mask,
                   latent_shape=(f, h, w), return_dict=False)[0]
            torch.cuda.synchronize()
    finally:
        O.attention_core = saved
    return {"what": "torch eager bf16 + F.scaled_dot_product_attention (the reference's GPU path, `_attention='sdpa'`): oracle restatement of the "
                    "reference modules on this GPU, same weights, 28 layers x %d conds x %d tokens, transformer forward only" % (B, N),
            "ms_per_step": lib_ms, "steps_per_s": 1e3 / lib_ms, "this_repo_ms_per_step": ours_ms, "speedup_vs_library": lib_ms / ours_ms,
            "rel_l2_between_them_one_forward": rel_l2(y.float().cpu(), y_ref.float().cpu()), "extrapolated": False}


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ltx2b_768x512x121_cfg_stg", choices=sorted(WORKLOADS) + sorted(WAN_WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-wan-sp", action="store_true", help="skip the Wan2.1 sequence-parallel block of the default line")
    ap.add_argument("--no-cond-parallel", action="store_true", help="skip the LTX guidance-condition-parallel block (N > 1)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the torch-eager-on-GPU library baseline")
    ap.add_argument("--layers", type=int, default=28, help="debug only: fewer layers makes the number INVALID")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)
    if args.workload in WAN_WORKLOADS:
        return run_wan(args, WAN_WORKLOADS[args.workload])
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
    from ltx_video_gpupoor_b200.ltx.init_weights import (random_transformer_state_dict, random_vae_decoder_state_dict,
                                                          random_vae_encoder_state_dict)
    from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import ConditioningItem, LTXVideoPipeline
    from ltx_video_gpupoor_b200.ltx.rf import RectifiedFlowScheduler
    from ltx_video_gpupoor_b200.ltx.skip_layer_strategy import SkipLayerStrategy
    from ltx_video_gpupoor_b200.ltx.symmetric_patchifier import SymmetricPatchifier
    from ltx_video_gpupoor_b200.ltx.transformer3d import LTX_2B_CONFIG, Transformer3DModel

    arch = LTX_ARCH[wl.get("arch", "2b")]
    if args.layers == 28:                          # the default = the architecture's full depth
        args.layers = arch["layers"]
    if wl.get("arch", "2b") != "2b":               # the CPU / torch-eager baselines are wired for the headline 2B architecture
        args.no_cpu_baseline = args.no_gpu_baseline = args.no_wan_sp = args.no_cond_parallel = True
    cfg = dict(LTX_2B_CONFIG, num_layers=args.layers, **arch["overrides"])
    tr = Transformer3DModel(**cfg)
    tr.load_state_dict(random_transformer_state_dict(cfg, seed=0, device=dev), device=dev)
    vae = CausalVideoAutoencoder()
    vsd = random_vae_decoder_state_dict(seed=1, device=dev)
    if wl.get("i2v"):
        vsd.update(random_vae_encoder_state_dict(seed=2, device=dev))
    vae.load_state_dict(vsd, device=dev)
    del vsd
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
    img_h = None
    if wl.get("i2v"):
        # synthetic conditioning frame in pinned host memory: encoded by the VAE inside every pipeline call (e2e includes it)
        img_h = (torch.rand(1, 3, 1, wl["height"], wl["width"], generator=g) * 2 - 1).pin_memory()
        call_kw.update(conditioning_items=[ConditioningItem(media_item=img_h, media_frame_number=0, conditioning_strength=1.0)],
                       image_cond_noise_scale=0.15)

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
    if os.environ.get("LTXB200_NCU_RANGE"):          # `ncu --profile-from-start off`: capture the timed region only
        torch.cuda.profiler.start()
    e0.record()
    for i in range(args.steps):
        pipe.denoise_step(st, (args.warmup + i) % S)
    e1.record()
    if os.environ.get("LTXB200_NCU_RANGE"):
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    barrier()
    clocks = sampler.stop()
    launches = _lib.launch_count() - l0
    elapsed = e0.elapsed_time(e1) / 1e3
    if dist is not None:
        t = torch.tensor([elapsed], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t)
    steps_per_s = world * args.steps / elapsed

    # ---------------- extra (NOT the headline): the same steps with share_stg_prefix=True ----------------
    # the perturbed STG condition's rows are copied from the text condition's rows up to the first skipped block instead of being
    # recomputed (bit-identical latents: tests/test_ltx_model_gpu.py::test_pipeline_shared_stg_prefix_is_bit_identical)
    shared = None
    if wl["num_conds"] == 3 and wl["skip_block_list"]:
        st.shared_prefix = (1, 1)
        for i in range(2):
            pipe.denoise_step(st, i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(args.steps):
            pipe.denoise_step(st, (args.warmup + i) % S)
        b.record()
        torch.cuda.synchronize()
        sh_ms = a.elapsed_time(b) / args.steps
        shared = {"ms_per_step": sh_ms, "steps_per_s": 1e3 / sh_ms,
                  "note": (f"optional pipeline kwarg share_stg_prefix=True: blocks 0..{min(wl['skip_block_list']) - 1} run for 2 conds, "
                           "the rest for 3; identical latents; not used for value / e2e")}
        st.shared_prefix = None

    # ---------------- roofline probe: one instrumented step (not part of the timed region) ----------------
    ops.PROFILER = []
    pipe.denoise_step(st, 0)
    torch.cuda.synchronize()
    prof, ops.PROFILER = ops.PROFILER, None
    agg = {}
    attn_launches = []                              # (flops, ms) per attention launch: self- and cross-attention are reported apart below
    for name, kind, amount, a, b in prof:
        d = agg.setdefault(name, dict(kind=kind, amount=0.0, ms=0.0, n=0))
        ms_l = a.elapsed_time(b)
        d["amount"] += amount; d["ms"] += ms_l; d["n"] += 1
        if name == "attention_bf16":
            attn_launches.append((amount, ms_l))
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
                "unit": unit, "frac": achieved / peak, "traffic": traffic,
                "traffic_source": ("static: dram__bytes_read.sum + dram__bytes_write.sum of this kernel's largest launch (FFN-down, 18432x2048x8192) from the "
                                   "committed ncu --set full capture profiles/r01c_ncu_gemm_raw.csv via profiles/traffic.json; NOT measured in this run "
                                   "(ncu cannot run inside bench.py)") if traffic is not None else None,
                "launches_per_step": d["n"],
                "avg_launch_ms": d["ms"] / d["n"], "share_of_step": d["ms"] / total_ms,
                "algorithmic_per_launch": d["amount"] / d["n"]}
    kernels = {k: {"ms_per_step": round(v["ms"], 3), "launches": v["n"], "share": round(v["ms"] / total_ms, 4),
                   ("tflops" if v["kind"] == "flop" else "gbs"): round(v["amount"] / (v["ms"] * 1e-3) / (1e12 if v["kind"] == "flop" else 1e9), 1)}
               for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}

    if "attention_bf16" in kernels and wl.get("arch", "2b") == "2b" and clocks.get("sm_mhz"):
        # the pipe that bounds attention at head dim 64 is the XU pipe (16 ex2 per clock and SM), not the tensor pipe: every score costs one
        # exponential and 4 * 64 = 256 flop; 3 of 8 exponentials run as an FMA-pipe polynomial (csrc/attention.cuh, LTXB200_ATTN_POLY_D64)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        xu_peak = sms * 16 * clocks["sm_mhz"] * 1e6 * 256 / (1.0 - 3.0 / 8.0) / 1e12
        kernels["attention_bf16"].update({"tensor_frac_of_sustained": round(kernels["attention_bf16"]["tflops"] / pk["tf_sust"], 3),
                                          "xu_bound_tflops_at_step_clock": round(xu_peak, 1),
                                          "xu_frac": round(kernels["attention_bf16"]["tflops"] / xu_peak, 3),
                                          "note": "head dim 64: bounded by the XU (ex2) pipe; xu_bound = SMs x 16 ex2/clk x median SM clock of the step x 256 flop per score / (5/8 on MUFU)"})

    try:
        if "attention_bf16" in kernels and attn_launches:
            kernels["attention_bf16"].update(split_attention(attn_launches, pk["tf_sust"]))
    except Exception as exc:                        # never lose the line over an extra
        kernels["attention_bf16"]["split_error"] = repr(exc)[:200]

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
    h2d = (pe_h.numel() + ne_h.numel()) * 2 + (pm_h.numel() + nm_h.numel()) * 4 + (img_h.numel() * 4 if img_h is not None else 0)
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

    # ---------------- the library bar on the same box (NOT this repo's path): the reference's GPU path = plain PyTorch bf16 modules
    # with torch SDPA (`_attention = "sdpa"`), here the pinned restatement of those modules on this GPU with the SAME weights,
    # all 28 layers x 3 conds, one transformer forward per step (guidance + scheduler arithmetic not included: it favours the baseline)
    # ---------------- extra at N > 1: ONE video over the guidance conditions' ranks (latency scaling of the LTX path) ----------------
    cond_par = None
    if dist is not None and world > 1 and wl["num_conds"] > 1 and not wl.get("i2v") and not args.no_cond_parallel:
        try:
            cond_par = ltx_cond_parallel_block(pipe, wl, call_kw, dev, dist, rank, world, args.steps, args.warmup, elapsed / args.steps * 1e3)
        except Exception as exc:                       # an extra must never cost the headline line
            cond_par = {"error": repr(exc)[:300]}
    gpu_lib = None
    if rank == 0 and not args.no_gpu_baseline:
        try:
            gpu_lib = gpu_library_baseline(wl, cfg, tr, dev, elapsed / args.steps * 1e3)
        except Exception as exc:                       # an extra must never cost the headline line
            gpu_lib = {"error": repr(exc)[:300]}
    # ---------------- the part of the path that shards: Wan2.1-1.3B Ulysses sequence parallel over ALL ranks (strong scaling) ----------------
    wan_sp = None
    if not args.no_wan_sp:
        del pipe, tr, vae, st
        torch.cuda.empty_cache()
        try:
            wan_sp = wan_sp_block(dev, world, rank, local_rank)
        except Exception as exc:
            import traceback
            wan_sp = {"error": repr(exc)[:300], "trace": traceback.format_exc()[-600:]}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        _, cpu_baseline = cpu_reference_sample(wl)

    ms_step = elapsed / args.steps * 1e3
    line = {
        "metric": "denoise_steps_per_s", "value": steps_per_s, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": ltx_config(args.workload, wl, args.layers, world),
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "s_per_video": (S * ms_step / 1e3 + decode_s) if decode_s is not None else None, "vae_decode_s": decode_s,
        "model_tflops": wl["num_conds"] * ltx_fwd_flops(wl, args.layers, (wl["num_frames"] // 8 + 1) * (wl["height"] // 32) * (wl["width"] // 32)) / (ms_step / 1e3) / 1e12,
        "kernels": kernels,
    }
    if shared is not None:
        line["stg_prefix_sharing"] = shared
        if decode_s is not None:
            shared["s_per_video"] = S * shared["ms_per_step"] / 1e3 + decode_s
    if cond_par is not None:
        line["ltx_cond_parallel"] = cond_par
    if gpu_lib is not None:
        line["gpu_library_baseline"] = gpu_lib
    if wan_sp is not None:
        line["wan_sp"] = wan_sp
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
