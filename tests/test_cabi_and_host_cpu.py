"""CPU-only: the C-ABI library loads and exports every symbol include/ltx_b200.h declares (no compute
calls without a GPU); host-side index/schedule logic of the drop-ins is bit-exact against the fixtures
recorded from the reference; the product path refuses to run without CUDA instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ltx_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ltxb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ltx_video_gpupoor_b200 import _lib
    lib = _lib.lib()                      # loads libltx_b200.so (built by __graft_entry__.build())
    names = _header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ltx_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names, "ctypes signatures out of sync with the header"
    assert lib.ltxb200_abi_version() == 4
    assert lib.ltxb200_error_string(-2).decode().startswith("pointer")


def test_bad_arguments_return_error_codes_not_crashes():
    from ltx_video_gpupoor_b200 import _lib
    lib = _lib.lib()
    # shape validation happens before any CUDA call, so this is safe without a GPU
    assert lib.ltxb200_gemm_bf16(None, 0, None, 0, 0, 8, 8, None, 0, 0, None, 0, None, 0, None, 0, 1, None) == -1
    assert lib.ltxb200_attention_bf16(None, 0, 0, None, 0, 0, None, 0, 0, None, 0, 0, 1, 1, 128, 128, 96, 0.0, None, None) == -5
    assert lib.ltxb200_norm_mod_bf16(None, 0, None, 0, 4, 100, None, None, 0, 1, None, None, 1e-6, 0, None) == -1
    assert lib.ltxb200_comm_wait(None, 2, 1, None) == -1
    assert lib.ltxb200_comm_alloc(0, None, None) == -1


def test_no_cpu_fallback():
    from ltx_video_gpupoor_b200 import _lib, ops
    with pytest.raises(_lib.LtxB200Error):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))


def test_missing_library_fails_loudly(monkeypatch):
    from ltx_video_gpupoor_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libltx_b200.so")
    with pytest.raises(_lib.LtxB200Error, match="no CPU/PyTorch fallback"):
        _lib.lib()


def test_scheduler_timesteps_bit_exact(golden_dir):
    from ltx_video_gpupoor_b200.ltx.rf import RectifiedFlowScheduler
    g = torch.load(os.path.join(golden_dir, "rf_scheduler.pt"), weights_only=False)
    for case in g.values():
        s = RectifiedFlowScheduler()
        s.set_timesteps(case["steps"], samples_shape=case["shape"], device="cpu")
        assert torch.equal(s.timesteps, case["timesteps"])


def test_patchifier_bit_exact(golden_dir):
    from ltx_video_gpupoor_b200.ltx.symmetric_patchifier import SymmetricPatchifier, latent_to_pixel_coords_from_factors
    g = torch.load(os.path.join(golden_dir, "patchifier.pt"), weights_only=False)
    p = SymmetricPatchifier(1)
    x = torch.arange(2 * 5 * 3 * 4 * 6, dtype=torch.float32).reshape(2, 5, 3, 4, 6)
    tok, coords = p.patchify(x)
    assert torch.equal(coords, g["coords"])
    assert torch.equal(latent_to_pixel_coords_from_factors(coords, (8, 32, 32)), g["px"])
    assert torch.equal(latent_to_pixel_coords_from_factors(coords, (8, 32, 32), True), g["px_fix"])
    assert torch.equal(p.unpatchify(tok, 4, 6, 5), x)
    assert torch.equal(tok, x.permute(0, 2, 3, 4, 1).reshape(2, 72, 5))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ltx-video-gpupoor_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f"{f} touches oracle/"


# ------------------------------------------------------------------ checkpoint formats, host-side only
def test_checkpoint_formats_round_trip(tmp_path):
    import json
    import torch
    from safetensors.torch import save_file
    from ltx_video_gpupoor_b200.ltx import checkpoint_io as C
    from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import LTX_VAE_CONFIG
    from ltx_video_gpupoor_b200.ltx.transformer3d import LTX_2B_CONFIG
    g = torch.Generator().manual_seed(0)
    tsd = {"model.diffusion_model.patchify_proj.weight": torch.randn(8, 4, generator=g),
           "model.diffusion_model.transformer_blocks.0.attn1.q_norm.weight": torch.randn(8, generator=g),
           "vae.decoder.conv_in.conv.weight": torch.randn(4, 2, 3, 3, 3, generator=g),
           "vae.per_channel_statistics.std-of-means": torch.rand(4, generator=g)}
    # --- single file with the configs in the metadata (transformer3d.py:313-325, causal_video_autoencoder.py:104-114)
    cfg = {"transformer": dict(LTX_2B_CONFIG, num_layers=1), "vae": dict(LTX_VAE_CONFIG)}
    f = tmp_path / "ltxv.safetensors"
    save_file(tsd, str(f), metadata={"config": json.dumps(cfg)})
    c, sd = C.load_transformer_checkpoint(f)
    assert c == cfg["transformer"] and set(sd) == set(tsd) and torch.equal(sd["vae.decoder.conv_in.conv.weight"], tsd["vae.decoder.conv_in.conv.weight"])
    c, sd = C.load_vae_checkpoint(f)
    assert c == cfg["vae"] and c["_class_name"] == "CausalVideoAutoencoder"
    # --- diffusers directory: config looked up in the mapping, keys renamed (diffusers_config_mapping.py:133-174)
    d = tmp_path / "repo"
    (d / "transformer").mkdir(parents=True); (d / "vae").mkdir()
    (d / "transformer" / "config.json").write_text(json.dumps(C.DIFFUSERS_TRANSFORMER_CONFIG))
    (d / "vae" / "config.json").write_text(json.dumps(C.DIFFUSERS_VAE_CONFIG))
    dsd = {"proj_in.weight": torch.randn(8, 4, generator=g), "time_embed.linear.bias": torch.randn(8, generator=g),
           "transformer_blocks.3.attn1.norm_q.weight": torch.randn(8, generator=g), "transformer_blocks.3.attn2.norm_k.weight": torch.randn(8, generator=g)}
    save_file({k: v for k, v in list(dsd.items())[:2]}, str(d / "transformer" / "diffusion_pytorch_model-00001-of-00002.safetensors"))
    save_file({k: v for k, v in list(dsd.items())[2:]}, str(d / "transformer" / "diffusion_pytorch_model-00002-of-00002.safetensors"))
    c, sd = C.load_transformer_checkpoint(d)
    assert c["num_layers"] == 28 and c["qk_norm"] == "rms_norm" and c["_class_name"] == "Transformer3DModel"
    assert set(sd) == {"patchify_proj.weight", "adaln_single.linear.bias", "transformer_blocks.3.attn1.q_norm.weight",
                       "transformer_blocks.3.attn2.k_norm.weight"}
    assert torch.equal(sd["patchify_proj.weight"], dsd["proj_in.weight"])
    vsd = {"decoder.mid_block.resnets.2.conv1.conv.weight": torch.zeros(1), "decoder.up_blocks.0.resnets.0.conv2.conv.bias": torch.zeros(1),
           "decoder.up_blocks.1.upsamplers.0.conv.conv.weight": torch.zeros(1), "decoder.up_blocks.2.conv_in.norm3.weight": torch.zeros(1),
           "decoder.up_blocks.2.conv_in.conv_shortcut.conv.weight": torch.zeros(1), "decoder.up_blocks.3.resnets.3.conv1.conv.weight": torch.zeros(1),
           "encoder.down_blocks.0.downsamplers.0.conv.weight": torch.zeros(1), "encoder.mid_block.resnets.0.conv1.conv.weight": torch.zeros(1),
           "latents_mean": torch.zeros(128), "latents_std": torch.ones(128)}
    save_file(vsd, str(d / "vae" / "diffusion_pytorch_model.safetensors"))
    c, sd = C.load_vae_checkpoint(d)
    assert c["blocks"] == LTX_VAE_CONFIG["blocks"]
    assert set(sd) == {"decoder.up_blocks.0.res_blocks.2.conv1.conv.weight", "decoder.up_blocks.1.res_blocks.0.conv2.conv.bias",
                       "decoder.up_blocks.2.conv.conv.weight", "decoder.up_blocks.4.norm3.norm.weight",
                       "decoder.up_blocks.4.conv_shortcut.weight", "decoder.up_blocks.9.res_blocks.3.conv1.conv.weight",
                       "encoder.down_blocks.1.conv.weight", "encoder.down_blocks.9.res_blocks.0.conv1.conv.weight",
                       "per_channel_statistics.mean-of-means", "per_channel_statistics.std-of-means"}
    import pytest
    (d / "vae" / "config.json").write_text(json.dumps(dict(C.DIFFUSERS_VAE_CONFIG, latent_channels=64)))
    with pytest.raises(AssertionError):
        C.load_vae_checkpoint(d)
    # --- legacy VAE directory with per_channel_statistics.json (causal_video_autoencoder.py:42-70)
    lv = tmp_path / "legacy"; lv.mkdir()
    (lv / "config.json").write_text(json.dumps(dict(LTX_VAE_CONFIG)))
    torch.save({"decoder.conv_in.conv.weight": torch.zeros(1)}, lv / "autoencoder.pth")
    (lv / "per_channel_statistics.json").write_text(json.dumps({"columns": ["std-of-means", "mean-of-means"], "data": [[1.0, 0.5], [2.0, 0.25]]}))
    c, sd = C.load_vae_checkpoint(lv)
    assert torch.equal(sd["per_channel_statistics.std-of-means"], torch.tensor([1.0, 2.0]))
    assert torch.equal(sd["per_channel_statistics.mean-of-means"], torch.tensor([0.5, 0.25]))
    # --- upsampler
    uf = tmp_path / "up.safetensors"
    ucfg = {"_class_name": "LatentUpsampler", "in_channels": 128, "mid_channels": 512, "num_blocks_per_stage": 4, "dims": 3,
            "spatial_upsample": True, "temporal_upsample": False}
    save_file({"final_conv.bias": torch.zeros(128)}, str(uf), metadata={"config": json.dumps(ucfg)})
    c, sd = C.load_upsampler_checkpoint(uf)
    assert c == ucfg and "final_conv.bias" in sd
    # --- LoRA merge: W += mult * alpha/rank * up @ down
    W0 = torch.randn(6, 5, generator=g)
    sdw = {"transformer_blocks.0.attn1.to_q.weight": W0.clone()}
    down, up = torch.randn(2, 5, generator=g), torch.randn(6, 2, generator=g)
    n = C.merge_lora(sdw, {"diffusion_model.transformer_blocks.0.attn1.to_q.lora_down.weight": down,
                           "diffusion_model.transformer_blocks.0.attn1.to_q.lora_up.weight": up,
                           "diffusion_model.transformer_blocks.0.attn1.to_q.alpha": torch.tensor(4.0)}, multiplier=0.5)
    assert n == 1 and torch.allclose(sdw["transformer_blocks.0.attn1.to_q.weight"], W0 + 0.5 * (4.0 / 2) * up @ down, atol=1e-6)


def test_torch_custom_ops_registered_with_fake_kernels():
    """The torch.library layer over the C ABI (custom_ops.py): every op is in torch.ops.ltxb200, traces with shape-only fake
    tensors (no GPU, no library call), and refuses CPU tensors at run time like the rest of the package."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    from ltx_video_gpupoor_b200 import custom_ops
    for name in custom_ops.OPS:
        assert hasattr(torch.ops.ltxb200, name), name
    with FakeTensorMode():
        a, w, b = torch.empty(48, 64, dtype=torch.bfloat16), torch.empty(256, 64, dtype=torch.bfloat16), torch.empty(256, dtype=torch.bfloat16)
        assert torch.ops.ltxb200.gemm(a, w, b, 1).shape == (48, 256)
        q, k = torch.empty(2, 96, 4, 64, dtype=torch.bfloat16), torch.empty(2, 80, 4, 64, dtype=torch.bfloat16)
        assert torch.ops.ltxb200.attention(q, k, k, None, 0.0).shape == (2, 96, 4, 64)
        x = torch.empty(1, 3, 8, 8, 64, dtype=torch.bfloat16)
        assert torch.ops.ltxb200.conv3d(x, torch.empty(128, 27 * 64, dtype=torch.bfloat16), None, True).shape == (1, 3, 8, 8, 128)
        assert torch.ops.ltxb200.norm_mod(a, None, None, 0, 1e-6, False).shape == (48, 64)
    with pytest.raises(Exception):
        torch.ops.ltxb200.gemm(torch.zeros(8, 64, dtype=torch.bfloat16), torch.zeros(16, 64, dtype=torch.bfloat16), None, 0)


def test_prepare_conditioning_keyframes_bit_exact(golden_dir):
    """LTXVideoPipeline.prepare_conditioning (host-side torch code; pre-encoded conditioning latents, so no GPU is involved) against the
    outputs of the unmodified reference method for first-frame, keyframe, first+last, mid-sequence and prefix-only-sequence conditioning
    (oracle/gen_golden_conditioning.py): tokens, pixel coordinates, mask and extra-token count."""
    from types import SimpleNamespace
    from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import ConditioningItem, LTXVideoPipeline
    from ltx_video_gpupoor_b200.ltx.symmetric_patchifier import SymmetricPatchifier
    g = torch.load(os.path.join(golden_dir, "ltx_conditioning.pt"), weights_only=False)
    F_l, H_l, W_l, num_frames, height, width = g["geom"]
    pipe = LTXVideoPipeline.__new__(LTXVideoPipeline)
    pipe.vae = SimpleNamespace(spatial_downscale_factor=32, temporal_downscale_factor=8)
    pipe.patchifier = SymmetricPatchifier(1)
    pipe.transformer = SimpleNamespace(config=SimpleNamespace(causal_temporal_positioning=True))
    assert len(g["cases"]) == 5
    for name, c in g["cases"].items():
        init = torch.randn(1, 128, F_l, H_l, W_l, generator=torch.Generator().manual_seed(1))
        items = [ConditioningItem(latents=g["enc"][n].clone(), media_frame_number=f, conditioning_strength=s) for n, f, s in c["items"]]
        tok, px, cm, extra = pipe.prepare_conditioning(items, init, num_frames, height, width, vae_per_channel_normalize=True,
                                                       generator=torch.Generator().manual_seed(2))
        assert extra == c["extra"] and torch.equal(tok, c["tokens"]) and torch.equal(cm, c["mask"]), name
        assert torch.equal(px.to(c["coords"].dtype), c["coords"]), name


def test_guidance_and_timestep_tables_bit_exact(golden_dir):
    """LTXVideoPipeline.__call__'s host-side tables (retrieve_timesteps with skip_initial / skip_final / strength / explicit timesteps, the
    guidance_timesteps mapping, list-valued guidance / STG / rescaling scales, nested skip_block_list, skip-layer masks, cond-batch size, first
    timestep argument) against what the unmodified reference built for its own presets (ltx_video/configs/*.yaml) and two edge cases
    (oracle/gen_golden_schedule.py).  Runs the product pipeline up to the loop on a CPU stand-in transformer: no kernel is launched."""
    from oracle.schedule_tables import compare, product_tables
    g = torch.load(os.path.join(golden_dir, "ltx_schedule_tables.pt"), weights_only=False)
    assert len(g["cases"]) == 7
    for name, c in g["cases"].items():
        lat = g["init_latents"].clone() if c["needs_latents"] else None
        mine = product_tables(g["num_layers"], g["geom"], g["pe"], g["pm"], g["ne"], g["pm"], lat, dict(c["kwargs"]))
        compare(name, mine, c["ref"])


def test_checkpoint_format_tables_match_reference(golden_dir):
    """The diffusers <-> native config mapping, the configs themselves and the ordered key-rename tables against the ones dumped from the
    unmodified reference module (oracle/gen_golden_formats.py; ltx_video/utils/diffusers_config_mapping.py)."""
    import json
    from oracle.format_tables import check_format_tables
    with open(os.path.join(golden_dir, "ltx_format_tables.json")) as f:
        check_format_tables(json.load(f))


def test_rf_scheduler_samplers_and_shiftings_bit_exact(golden_dir, tmp_path):
    """RectifiedFlowScheduler.set_timesteps for every sampler x shifting combination of the reference class ("Uniform" / "LinearQuadratic" /
    "Constant" x none / "SD3" (+ terminal stretch) / "SimpleDiffusion"), 1..40 steps, three sample shapes: 105 tables recorded from the unmodified
    reference scheduler (oracle/gen_golden_rf_variants.py), bit for bit; from_pretrained on a JSON config and on a single-file checkpoint."""
    import json
    from safetensors.torch import save_file
    from ltx_video_gpupoor_b200.ltx.rf import RectifiedFlowScheduler
    g = torch.load(os.path.join(golden_dir, "rf_scheduler_variants.pt"), weights_only=False)
    assert len(g["tables"]) == 105
    for (name, steps, shape), ref in g["tables"].items():
        s = RectifiedFlowScheduler.from_config(dict(g["configs"][name], num_train_timesteps=1000))
        s.set_timesteps(steps, samples_shape=shape)
        assert s.timesteps.dtype == ref.dtype and torch.allclose(s.timesteps, ref, rtol=0, atol=0, equal_nan=True), (name, steps, shape)
    cfg = dict(g["configs"]["linear_quadratic_sd3_terminal"], num_train_timesteps=1000, _class_name="RectifiedFlowScheduler")
    (tmp_path / "scheduler_config.json").write_text(json.dumps(cfg))
    save_file({"x": torch.zeros(1)}, str(tmp_path / "ckpt.safetensors"), metadata={"config": json.dumps({"scheduler": cfg})})
    for path in (tmp_path / "scheduler_config.json", tmp_path / "ckpt.safetensors"):
        s = RectifiedFlowScheduler.from_pretrained(path)
        s.set_timesteps(30, samples_shape=(1, 128, 16, 16, 24))
        assert torch.equal(s.timesteps, g["tables"][("linear_quadratic_sd3_terminal", 30, (1, 128, 16, 16, 24))])
    with pytest.raises(ValueError):
        RectifiedFlowScheduler(sampler="Nope")


def test_every_compute_entry_point_rejects_null_arguments():
    """All-null / all-zero arguments must come back as an error code from every compute entry point of the C ABI (validation runs before
    any CUDA call or pointer use, so this is safe on a machine without a GPU) — never a crash, never 0."""
    import ctypes
    from ltx_video_gpupoor_b200 import _lib
    lib = _lib.lib()
    skipped = {"ltxb200_abi_version", "ltxb200_error_string", "ltxb200_launch_count", "ltxb200_comm_open", "ltxb200_comm_close",
               "ltxb200_comm_free", "ltxb200_scatter_signal_ctas"}                                     # queries, and calls whose only argument is a CUDA IPC handle / pointer
    checked = 0
    for name in _lib.EXPORTED_SYMBOLS:
        if name in skipped:
            continue
        fn = getattr(lib, name)
        args = [0.0 if t is ctypes.c_float else (None if (t is ctypes.c_void_p or hasattr(t, "contents")) else 0) for t in fn.argtypes]
        rc = fn(*args)
        assert rc < 0, f"{name} accepted null arguments (rc {rc})"
        assert lib.ltxb200_error_string(rc), name
        checked += 1
    assert checked >= 35


def test_trim_conditioning_sequence_matches_reference(golden_dir):
    """LTXVideoPipeline.trim_conditioning_sequence (pipeline_ltx_video.py:1689-1707) on the grid recorded from the reference method."""
    import json
    from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import LTXVideoPipeline
    pipe = LTXVideoPipeline.__new__(LTXVideoPipeline)
    pipe.video_scale_factor = 8
    with open(os.path.join(golden_dir, "ltx_trim_sequence.json")) as f:
        rows = json.load(f)
    assert len(rows) > 200
    for start, n, target, expected in rows:
        assert pipe.trim_conditioning_sequence(start, n, target) == expected


def test_model_objects_accept_the_reference_loaders_module_calls():
    """`model.eval().requires_grad_(False)` (wan/text2video.py:101), `vae.to(torch.bfloat16)`, `latent_upsampler.to("cpu").eval()`
    (ltx_video/ltxv.py:176,199-200) keep working on the drop-in classes; anything that would change what the packed bf16 weights are is refused."""
    from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import CausalVideoAutoencoder
    from ltx_video_gpupoor_b200.ltx.latent_upsampler import LatentUpsampler
    from ltx_video_gpupoor_b200.ltx.transformer3d import Transformer3DModel
    from ltx_video_gpupoor_b200.wan.model import WanModel
    from ltx_video_gpupoor_b200.wan.vae import WanVAE
    objs = [Transformer3DModel(num_layers=1), CausalVideoAutoencoder(), LatentUpsampler(in_channels=128, mid_channels=256, num_blocks_per_stage=1, dims=3),
            WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=1), WanVAE()]
    for m in objs:
        assert m.eval().requires_grad_(False) is m
        assert m.to(torch.bfloat16) is m and m.to("cpu").eval() is m and m.to(device="cuda", dtype=torch.bfloat16) is m
        m._model_dtype = torch.bfloat16                      # ltxv.py:177,193 tag the objects
        with pytest.raises(NotImplementedError):
            m.to(torch.float16)
        with pytest.raises(NotImplementedError):
            m.requires_grad_(True)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the arm the driver runs first): one JSON line with the contract's keys, the b200 arm's config, zero copy
    bytes; under a multi-rank launch only rank 0 works and prints."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], cwd=root, env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env["LTXB200_BENCH_SKIP_CONFIG0"] = "1"          # the real 28-layer configs[0] run takes about a minute: covered by the bench itself
    env["OMP_NUM_THREADS"] = "1"                     # what torchrun exports: the arm has to take the host cores back itself
    env["LTXB200_REF_EXTRAPOLATE"] = "1"             # the arm's default times ONE WHOLE full-size step for real (2.5 min on this container's 8 threads,
                                                     # 30-60 s on a GPU box): here the 1- and 3-layer sample keeps the CPU suite short
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "20", "--warmup", "5"], cwd=root, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "denoise_steps_per_s" and d["unit"] == "steps/s" and d["value"] > 0
    assert d["higher_is_better"] is True and d["steps"] == 20 and d["warmup"] == 5 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["extrapolated"] is True and d["timed"]["samples"] == 1 and d["timed"]["run_wall_s"] < 300     # ONE bounded sample whatever K is
    assert d["config"]["workload"] == "ltx2b_768x512x121_cfg_stg" and d["config"]["tokens"] == 6144 and d["config"]["num_conds"] == 3
    have_ref = os.path.isdir(os.path.join(root, "oracle", "_ref", "ltx_video")) or os.path.isdir("/root/reference/ltx_video")
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["cores"] == (len(os.sched_getaffinity(0)) or 1) and d["cpu_baseline"]["extrapolated"] is True
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_cond_partition():
    """ltx/distributed/cond_parallel.py: contiguous condition ranges per rank; ranks beyond the conditions take none."""
    from ltx_video_gpupoor_b200.ltx.distributed.cond_parallel import cond_partition
    assert cond_partition(3, 1) == [(0, 3)]
    assert cond_partition(3, 2) == [(0, 2), (2, 3)]
    assert cond_partition(3, 3) == [(0, 1), (1, 2), (2, 3)]
    assert cond_partition(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert cond_partition(2, 8)[:3] == [(0, 1), (1, 2), (2, 2)]
    assert cond_partition(1, 2) == [(0, 1), (1, 1)]
    for conds in (1, 2, 3):
        for ranks in range(1, 9):
            parts = cond_partition(conds, ranks)
            assert len(parts) == ranks and parts[0][0] == 0 and max(hi for _, hi in parts) == conds
            assert all(a[1] == b[0] or b == (conds, conds) for a, b in zip(parts, parts[1:]))


def test_bench_attention_split():
    """bench.py reports the self-attention (N x N) and cross-attention (N x prompt) launches of a step apart: the split is by flops per launch."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    out = bench.split_attention([(927.7e9, 1.25)] * 28 + [(38.7e9, 0.075)] * 28, 1382.5)
    assert out["self_attention"]["launches"] == 28 and out["cross_attention"]["launches"] == 28
    assert abs(out["self_attention"]["tflops"] - 742.2) < 0.2 and abs(out["cross_attention"]["tflops"] - 516.0) < 0.2
    assert abs(out["self_attention"]["tensor_frac_of_sustained"] - 742.2 / 1382.5) < 1e-3
    assert list(bench.split_attention([(5.0e9, 1.0)], 1000.0)) == ["self_attention"]      # a step with one kind of launch only

