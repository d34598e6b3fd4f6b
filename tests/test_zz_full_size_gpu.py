"""Full-size parity on the GPU (VERDICT r1 "no full-size loop parity"): BASELINE.json's own sizes, not the reduced fixtures.

  * configs[1]: LTX-2B (28 layers) t2v 768x512x121 -> 6144 tokens, the 2B preset guidance (CFG 3 + STG 1 on block 19 + rescale 0.7 =
    3 conds), ALL 30 steps: per-step latents of the CUDA drop-in against the pinned oracle in fp32 on the same GPU (the truth), with the
    reference's own bf16 path (the oracle in bf16 + torch SDPA = what LTX-Video-GPUPoor executes on a GPU) run beside it as the noise
    floor.  The drift over 30 steps is MEASURED and printed; contract: <= 2e-2, or no worse than 1.25x the reference's own bf16 drift.
  * configs[2]: i2v from pixels at full size (VAE encode of the conditioning frame inside the call) + the full CausalVideoAutoencoder
    decode to (1, 3, 121, 512, 768): decoded frames of the drop-in against the fp32 oracle decode of the SAME latents, PSNR >= 40 dB.
  * Wan2.1-14B geometry (dim 5120, 40 heads, ffn 13824) at reduced depth against the oracle in fp32 on the GPU.
Weights: the product's seeded random init (reference key names), shared by all paths.  ~2 min of GPU time.
    python -m pytest tests/test_zz_full_size_gpu.py -m gpu -x -q -s"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

from oracle import ltx_oracle as O  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16


def _sdpa_core(q, k, v, bias=None):
    """utils/attention.py:99-116 on the GPU: torch SDPA on [B, H, L, d] (fp32 in -> fp32 math, bf16 in -> flash)"""
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                       attn_mask=None if bias is None else bias.to(q.dtype))
    return o.transpose(1, 2)


@pytest.fixture()
def exact_fp32():
    """The fp32 truth must not run on TF32 tensor cores — and not on cuDNN: on this image (torch 2.11 / cuDNN 9.22, B200) cuDNN's fp32
    conv3d returns wrong values for the decoder's 512 -> 4096 depth-to-space convolution (rel. error 0.37 against the CPU; torch's native
    CUDA convolution and the CPU agree to 1e-6: profiles/scripts/diag_vae3.py, profiles/r02_diag_cudnn_conv3d.log), which made a
    GPU-side oracle decode look 19 dB away from everything else."""
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.enabled)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.enabled = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.enabled = saved


def _ltx_pipe(with_encoder=False):
    from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import CausalVideoAutoencoder
    from ltx_video_gpupoor_b200.ltx.init_weights import (random_transformer_state_dict, random_vae_decoder_state_dict,
                                                          random_vae_encoder_state_dict)
    from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import LTXVideoPipeline
    from ltx_video_gpupoor_b200.ltx.rf import RectifiedFlowScheduler
    from ltx_video_gpupoor_b200.ltx.symmetric_patchifier import SymmetricPatchifier
    from ltx_video_gpupoor_b200.ltx.transformer3d import LTX_2B_CONFIG, Transformer3DModel
    sd = random_transformer_state_dict(dict(LTX_2B_CONFIG), seed=0, device=DEV)
    tr = Transformer3DModel(**LTX_2B_CONFIG)
    tr.load_state_dict(sd, device=DEV)
    vsd = random_vae_decoder_state_dict(seed=1, device=DEV)
    if with_encoder:
        vsd.update(random_vae_encoder_state_dict(seed=2, device=DEV))
    vae = CausalVideoAutoencoder()
    vae.load_state_dict(vsd, device=DEV)
    pipe = LTXVideoPipeline(vae=vae, transformer=tr, scheduler=RectifiedFlowScheduler(), patchifier=SymmetricPatchifier(1))
    return pipe, sd, vsd


def test_ltx_config1_30_step_loop_drift(monkeypatch, exact_fp32):
    from ltx_video_gpupoor_b200.ltx.skip_layer_strategy import SkipLayerStrategy
    pipe, sd, _ = _ltx_pipe()
    H, W, FR, fps, steps = 512, 768, 121, 25.0, 30
    f, h, w = FR // 8 + 1, H // 32, W // 32
    N = f * h * w
    g = torch.Generator().manual_seed(42)
    pe, ne = torch.randn(1, 256, 4096, generator=g).to(BF).float(), torch.randn(1, 256, 4096, generator=g).to(BF).float()
    pm = torch.ones(1, 256)
    pm[:, 200:] = 0                                  # a padded prompt: the masked cross-attention path, as a real prompt takes it
    noise = torch.randn(1, N, 128, generator=torch.Generator().manual_seed(7))
    ours = []
    # prepare_latents (pipeline_ltx_video.py:696-699) draws randn((b, N, C)) in prompt_embeds' dtype from the generator: with fp32
    # embeddings and seed 7 that is exactly `noise`
    pipe(height=H, width=W, num_frames=FR, frame_rate=fps, prompt_embeds=pe, prompt_attention_mask=pm, negative_prompt_embeds=ne,
         negative_prompt_attention_mask=pm, num_inference_steps=steps, guidance_scale=3.0, stg_scale=1.0, rescaling_scale=0.7,
         skip_block_list=[19], skip_layer_strategy=SkipLayerStrategy.AttentionValues, generator=torch.Generator().manual_seed(7),
         output_type="latent", return_dict=False, is_video=True, vae_per_channel_normalize=True, _per_step_latents=ours)
    torch.cuda.synchronize()
    assert len(ours) == steps
    monkeypatch.setattr(O, "attention_core", _sdpa_core)
    kw = dict(num_frames_lat=f, lat_h=h, lat_w=w, frame_rate=fps, num_steps=steps, neg_enc=ne.to(DEV), neg_mask=pm.to(DEV),
              guidance_scale=3.0, stg_scale=1.0, rescaling_scale=0.7, skip_block_list=[19], strategy=O.SKIP_ATTENTION_VALUES)
    truth, floor = [], []
    with torch.no_grad():
        O.denoise_loop({k: v.float() for k, v in sd.items()}, O.LTX_2B, noise.to(DEV), pe.to(DEV), pm.to(DEV),
                       per_step=truth, model_dtype=torch.float32, **kw)
        # bf16 modules, fp32 latents between the steps: a LOWER floor than the reference's real bf16 run, whose latents are bf16 too (:1061)
        O.denoise_loop(sd, O.LTX_2B, noise.to(DEV), pe.to(DEV), pm.to(DEV), per_step=floor, model_dtype=BF, **kw)
    worst = 0.0
    for i in range(steps):
        e, ef = O.rel_l2(ours[i].float().cpu(), truth[i].float().cpu()), O.rel_l2(floor[i].float().cpu(), truth[i].float().cpu())
        worst = max(worst, e)
        if i % 3 == 2 or i == steps - 1:
            print(f"configs[1] step {i + 1:2d}/30: latents rel_l2 vs the fp32 oracle: CUDA drop-in {e:.3e}; the reference's own bf16 path {ef:.3e}")
        assert e < max(2e-2, 1.25 * ef), f"step {i}: {e} vs floor {ef}"
    print(f"configs[1] 30-step loop: worst per-step drift of the drop-in {worst:.3e} (contract 2e-2)")


def test_ltx_config2_i2v_full_size_decode_psnr(exact_fp32):
    from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import ConditioningItem
    from ltx_video_gpupoor_b200.ltx.skip_layer_strategy import SkipLayerStrategy
    pipe, sd, vsd = _ltx_pipe(with_encoder=True)
    H, W, FR, fps = 512, 768, 121, 25.0
    g = torch.Generator().manual_seed(11)
    pe, ne = torch.randn(1, 256, 4096, generator=g).to(BF), torch.randn(1, 256, 4096, generator=g).to(BF)
    pm = torch.ones(1, 256)
    img = torch.rand(1, 3, 1, H, W, generator=g) * 2 - 1
    kw = dict(height=H, width=W, num_frames=FR, frame_rate=fps, prompt_embeds=pe, prompt_attention_mask=pm, negative_prompt_embeds=ne,
              negative_prompt_attention_mask=pm, num_inference_steps=8, guidance_scale=3.0, stg_scale=1.0, rescaling_scale=0.7,
              skip_block_list=[19], skip_layer_strategy=SkipLayerStrategy.AttentionValues, return_dict=False, is_video=True,
              vae_per_channel_normalize=True, image_cond_noise_scale=0.15,
              conditioning_items=[ConditioningItem(media_item=img, media_frame_number=0, conditioning_strength=1.0,
                                                   encode_noise=torch.randn(1, 128, 1, H // 32, W // 32, generator=g))])
    lat = pipe(generator=torch.Generator().manual_seed(3), output_type="latent", **kw)[0]
    frames = pipe(generator=torch.Generator().manual_seed(3), output_type="pt", **kw)[0]
    torch.cuda.synchronize()
    assert tuple(lat.shape) == (1, 128, 16, 16, 24) and tuple(frames.shape) == (1, 3, 121, 512, 768)
    assert torch.isfinite(frames).all() and float(frames.min()) >= 0.0 and float(frames.max()) <= 1.0
    with torch.no_grad():
        ref = O.postprocess(O.vae_decode({k: v.float() for k, v in vsd.items() if k.startswith("decoder.") or k.endswith("_of_means")},
                                         lat.float()))
    ps = O.psnr(frames.float().cpu(), ref.float().cpu())
    print(f"configs[2] i2v 768x512x121: decoded frames (1,3,121,512,768) PSNR vs the fp32 oracle decode of the same latents = {ps:.1f} dB")
    assert ps >= 40.0


def test_wan_14b_geometry_reduced_depth(monkeypatch, exact_fp32):
    """Wan2.1-14B widths (dim 5120, 40 heads of 128, ffn 13824, wan/configs/wan_t2v_14B.py:19-29) at 2 layers and 1920 tokens: WanModel vs the
    oracle in fp32 on the GPU; the norm / rope / GEMM template instances of the 14B shapes (NV = 20, N = 13824) are not touched by any
    1.3B test."""
    from ltx_video_gpupoor_b200.wan.init_weights import seeded_wan_state_dict
    from ltx_video_gpupoor_b200.wan.model import WanModel
    from ltx_video_gpupoor_b200.wan.posemb_layers import get_rotary_pos_embed
    from oracle import wan_oracle as W
    cfg = dict(W.WAN_14B, num_layers=2)
    sd = seeded_wan_state_dict(cfg, seed=3)
    m = WanModel(dim=cfg["dim"], ffn_dim=cfg["ffn_dim"], num_heads=cfg["num_heads"], num_layers=2)
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(5)
    lat = torch.randn(16, 5, 32, 48, generator=g).to(DEV)
    ctx, ctx0 = torch.randn(77, 4096, generator=g).to(DEV), torch.randn(40, 4096, generator=g).to(DEV)
    t = torch.tensor([612.0], device=DEV)
    cos, sin = get_rotary_pos_embed(lat.shape[1:])
    y = m([lat, lat], t=t, context=[ctx, ctx0], freqs=(cos, sin))
    torch.cuda.synchronize()
    monkeypatch.setattr(W, "attention_core", _sdpa_core)
    with torch.no_grad():
        y_ref = W.wan_forward({k: v.to(DEV) for k, v in sd.items()}, cfg, [lat, lat], t, [ctx, ctx0], cos.to(DEV), sin.to(DEV))
    for a, b in zip(y, y_ref):
        e = W.rel_l2(a.float().cpu(), b.float().cpu())
        print(f"Wan-14B geometry (2 layers, 1920 tokens): forward rel_l2 vs the fp32 oracle = {e:.3e}")
        assert e < 2e-2
