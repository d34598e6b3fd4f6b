"""LTXVideoPipeline — B200-native drop-in for ltx_video/pipelines/pipeline_ltx_video.py:222-1707.

Keeps the reference call signature (`__call__(height, width, num_frames, frame_rate, prompt_embeds=…,
num_inference_steps, guidance_scale, stg_scale, rescaling_scale, skip_block_list, latents,
conditioning_items, output_type, …)`) and semantics of the denoise loop (:1103-1256): timestep table,
per-step guidance tables, cond batching [uncond, text, perturbed], per-token timesteps for conditioned
tokens, cfg-star / STG / std-rescale guidance, rectified-flow Euler step with the conditioning mask, and
the final VAE decode.  What differs is execution: latents live in fp32 on the device, each step is
the transformer forward + one fused guidance/step kernel group, and nothing in the loop synchronises with the
host (CUDA-graph replay of the step was measured and not kept: the loop is not launch-bound).  Text encoding, prompt enhancement, media loading and the
multi-scale wrapper are out of scope.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Callable, List, Optional, Tuple, Union

import torch

from .. import ops
from .causal_video_autoencoder import CausalVideoAutoencoder, get_vae_size_scale_factor, vae_decode, vae_encode
from .rf import RectifiedFlowScheduler
from .skip_layer_strategy import SkipLayerStrategy
from .symmetric_patchifier import SymmetricPatchifier, latent_to_pixel_coords_from_factors
from .transformer3d import Transformer3DModel

BF16 = torch.bfloat16


@dataclass
class ConditioningItem:
    """pipeline_ltx_video.py:195-219.  `media_item` is pixels [b,3,f,h,w] in [-1,1] (encoded here with the VAE encoder);
    `latents` (extension) carries already-encoded, normalised latents [b,128,f_l,h_l,w_l]; `encode_noise` (extension) fixes
    the noise of `latent_dist.sample()` (the reference draws it from the global RNG, vae_encode.py:77)."""
    media_item: Optional[torch.Tensor] = None
    media_frame_number: int = 0
    conditioning_strength: float = 1.0
    media_x: Optional[int] = None
    media_y: Optional[int] = None
    latents: Optional[torch.Tensor] = None
    encode_noise: Optional[torch.Tensor] = None


def retrieve_timesteps(scheduler, num_inference_steps=None, device=None, timesteps=None, max_timestep=1.0,
                       skip_initial_inference_steps=0, skip_final_inference_steps=0, **kwargs):
    """pipeline_ltx_video.py:125-200"""
    if timesteps is not None:
        scheduler.set_timesteps(timesteps=timesteps, device=device, **kwargs)
        timesteps = scheduler.timesteps
        num_inference_steps = len(timesteps)
    else:
        scheduler.set_timesteps(num_inference_steps, device=device, **kwargs)
        timesteps = scheduler.timesteps
    if (skip_initial_inference_steps < 0 or skip_final_inference_steps < 0
            or skip_initial_inference_steps + skip_final_inference_steps >= num_inference_steps):
        raise ValueError("invalid skip inference step values: must be non-negative and the sum of "
                         "skip_initial_inference_steps and skip_final_inference_steps must be less than the number of inference steps")
    timesteps = timesteps[skip_initial_inference_steps: len(timesteps) - skip_final_inference_steps]
    if max_timestep < 1.0:
        if max_timestep < timesteps.min():
            raise ValueError(f"max_timestep {max_timestep} is smaller than the minimum timestep {timesteps.min()}")
        timesteps = timesteps[timesteps <= max_timestep]
    num_inference_steps = len(timesteps)
    scheduler.set_timesteps(timesteps=timesteps, device=device, **kwargs)
    return timesteps, num_inference_steps


class LTXVideoPipeline:
    def __init__(self, tokenizer=None, text_encoder=None, vae: CausalVideoAutoencoder = None,
                 transformer: Transformer3DModel = None, scheduler: RectifiedFlowScheduler = None,
                 patchifier: SymmetricPatchifier = None, prompt_enhancer_image_caption_model=None,
                 prompt_enhancer_image_caption_processor=None, prompt_enhancer_llm_model=None,
                 prompt_enhancer_llm_tokenizer=None, allowed_inference_steps: Optional[List[float]] = None):
        self.tokenizer, self.text_encoder = tokenizer, text_encoder
        self.vae, self.transformer, self.scheduler = vae, transformer, scheduler
        self.patchifier = patchifier or SymmetricPatchifier(1)
        self.video_scale_factor, self.vae_scale_factor, _ = get_vae_size_scale_factor(vae) if vae is not None else (8, 32, 32)
        self.allowed_inference_steps = allowed_inference_steps

    @property
    def _execution_device(self):
        return self.transformer.device

    # ---------------------------------------------------------------------------------------------
    def prepare_latents(self, latents, media_items, timestep, latent_shape, dtype, device, generator,
                        vae_per_channel_normalize: bool = True):
        """pipeline_ltx_video.py:632-710: noise drawn in the patchified shape (b, f*h*w, c) on the generator's device; `media_items`
        (img2img / vid2vid, :682-687) are encoded with the VAE — `latent_dist.sample()` draws from the global RNG as in the
        reference — and noised to the first timestep like user-provided latents."""
        assert latents is None or media_items is None, "Cannot provide both latents and media_items. Please provide only one of the two."
        assert (latents is None and media_items is None) or timestep < 1.0, \
            "Input media_item or latents are provided, but they will be replaced with noise."
        if media_items is not None:
            latents = vae_encode(media_items.to(device), self.vae, vae_per_channel_normalize=vae_per_channel_normalize)
        b, c, f, h, w = latent_shape
        gdev = generator.device if isinstance(generator, torch.Generator) else device
        noise = torch.randn((b, f * h * w, c), generator=generator, device=gdev, dtype=dtype).to(device)
        noise = noise.reshape(b, f, h, w, c).permute(0, 4, 1, 2, 3) * self.scheduler.init_noise_sigma
        if latents is None:
            return noise
        assert tuple(latents.shape) == tuple(latent_shape)
        return timestep * noise + (1 - timestep) * latents.to(device=device, dtype=dtype)

    def trim_conditioning_sequence(self, start_frame: int, sequence_num_frames: int, target_num_frames: int) -> int:
        """pipeline_ltx_video.py:1689-1707 (called by ltxv.py:495-499 while it loads conditioning media): the longest prefix of a
        conditioning sequence that fits before the end of the video and holds 8k + 1 frames."""
        fits = min(sequence_num_frames, target_num_frames - start_frame)
        return (fits - 1) // self.video_scale_factor * self.video_scale_factor + 1

    @staticmethod
    def resize_tensor(media_items: torch.Tensor, height: int, width: int) -> torch.Tensor:
        """pipeline_ltx_video.py:748-760: per-frame F.interpolate(mode="bilinear", align_corners=False) of [b, c, n, h, w] pixels
        (the fp32 resize kernel; planes = b*c*n)."""
        if tuple(media_items.shape[-2:]) != (height, width):
            media_items = ops.bilinear_resize(media_items.to(torch.float32).contiguous(), height, width).to(media_items.dtype)
        return media_items

    @staticmethod
    def _handle_non_first_conditioning_sequence(init_latents, cmask, latents, media_frame_number: int, strength: float,
                                                num_prefix_latent_frames: int = 2):
        """pipeline_ltx_video.py:1614-1687 with the defaults the caller uses (prefix mode "concat"): the conditioning sequence minus
        its first two latent frames is blended into the noise at its position; the two prefix frames are returned to be appended as
        extra tokens."""
        f_l, f_p = latents.shape[2], num_prefix_latent_frames
        assert f_l >= f_p and media_frame_number % 8 == 0
        if f_l > f_p:
            f0 = media_frame_number // 8 + f_p
            f1 = f0 + f_l - f_p
            init_latents[:, :, f0:f1] = torch.lerp(init_latents[:, :, f0:f1], latents[:, :, f_p:], strength)
            cmask[:, f0:f1] = strength
        return init_latents, cmask, latents[:, :, :f_p]

    def prepare_conditioning(self, conditioning_items, init_latents, num_frames, height, width,
                             vae_per_channel_normalize: bool = False, generator=None):
        """pipeline_ltx_video.py:1344-1548: conditioning from pixels (encoded with the VAE, :1420-1424) or pre-encoded latents.
        media_frame_number == 0: the latents are blended into the start of the noise and marked in the mask (:1427-1448);
        otherwise (:1449-1503) a single frame — or the two-frame prefix of a sequence whose remainder is blended in place — is
        noised, patchified and PREPENDED as extra tokens whose pixel coordinates carry the target frame number.
        Returns (patchified latents, pixel coords, conditioning mask | None, number of extra tokens)."""
        cmask = None
        extra_lat, extra_px, extra_mask, extra_n = [], [], [], 0
        scale = get_vae_size_scale_factor(self.vae)
        causal_fix = self.transformer.config.causal_temporal_positioning
        if conditioning_items:
            cmask = torch.zeros(init_latents[:, 0].shape, dtype=torch.float32, device=init_latents.device)
            for item in conditioning_items:
                fno = int(item.media_frame_number)
                if item.latents is None:
                    m = item.media_item
                    assert m is not None and m.ndim == 5 and m.shape[2] % 8 == 1                  # :1409-1415
                    assert fno >= 0 and fno + m.shape[2] <= num_frames
                    if item.media_x or item.media_y:                                              # :1556-1559
                        raise ValueError("Provide media_item in the target size for spatial conditioning.")
                    # :1402-1404 every item is brought to the target size first (the multi-scale first pass relies on it), so the
                    # spatial offset / border stripping of :1566-1612 never sees a smaller item
                    m = self.resize_tensor(m.to(init_latents.device), height, width)
                    lat = vae_encode(m, self.vae, vae_per_channel_normalize=vae_per_channel_normalize,
                                     noise=item.encode_noise).to(device=init_latents.device, dtype=init_latents.dtype)
                else:
                    lat = item.latents.to(device=init_latents.device, dtype=init_latents.dtype)
                if lat.shape[0] == 1 and init_latents.shape[0] > 1:          # one conditioning item for every video of the call
                    lat = lat.expand(init_latents.shape[0], -1, -1, -1, -1)
                _, _, f_l, h_l, w_l = lat.shape
                s = item.conditioning_strength
                if fno == 0:
                    init_latents[:, :, :f_l, :h_l, :w_l] = torch.lerp(init_latents[:, :, :f_l, :h_l, :w_l], lat, s)   # :1436-1445
                    cmask[:, :f_l, :h_l, :w_l] = s
                    continue
                assert (h_l, w_l) == tuple(init_latents.shape[-2:]), "non-first conditioning items must have the target size (:1410-1412)"
                if f_l > 1:                                                                        # a sequence (n_frames > 1)
                    init_latents, cmask, lat = self._handle_non_first_conditioning_sequence(init_latents, cmask, lat, fno, s)
                gdev = generator.device if isinstance(generator, torch.Generator) else lat.device
                noise = torch.randn(lat.shape, generator=generator if isinstance(generator, torch.Generator) else None,
                                    device=gdev, dtype=lat.dtype).to(lat.device)                  # :1466-1471
                lat = torch.lerp(noise, lat, s)
                tok, lc = self.patchifier.patchify(lat)
                px = latent_to_pixel_coords_from_factors(lc, scale, causal_fix)
                px[:, 0] += fno                                                                    # :1487
                extra_n += tok.shape[1]
                extra_lat.append(tok)
                extra_px.append(px)
                extra_mask.append(torch.full(tok.shape[:2], s, dtype=torch.float32, device=init_latents.device))
        tokens, coords = self.patchifier.patchify(init_latents)
        px = latent_to_pixel_coords_from_factors(coords, scale, causal_fix)
        if cmask is None:
            return tokens, px, None, 0
        cm, _ = self.patchifier.patchify(cmask.unsqueeze(1))
        cm = cm.squeeze(-1)
        if extra_lat:                                                                              # :1519-1528
            tokens = torch.cat([*extra_lat, tokens], dim=1)
            px = torch.cat([*extra_px, px], dim=2)
            cm = torch.cat([*extra_mask, cm], dim=1)
        return tokens, px, cm, extra_n

    def _step_noise(self, st):
        """rf.py:372 `torch.randn_like(sample)` for the stochastic sampler, drawn from the call's generator when there is one."""
        gen = st.generator if isinstance(st.generator, torch.Generator) else None
        gdev = gen.device if gen is not None else self._execution_device
        return torch.randn(st.N * st.C, generator=gen, device=gdev, dtype=torch.float32).to(self._execution_device)

    def denoise_step(self, st, i: int):
        """One iteration of the loop at pipeline_ltx_video.py:1104-1256: cond batch, timestep tensor, transformer
        forward, guidance, scheduler step.  No host synchronisation.  The batch is laid out cond-major like the reference's
        torch.cat([negative, positive, positive]) (:1034-1051): row c * b + j is cond c of sample j."""
        t = st.ts_host[i]
        N, C, num_conds, bsz, device = st.N, st.C, st.num_conds, st.bsz, self._execution_device
        if st.cmask_dev is not None and st.image_cond_noise_scale > 0.0:
            # :606-629 add timestep-dependent noise to hard-conditioned tokens (host-side glue, i2v only)
            gen = st.generator
            # drawn in the dtype the reference's latents have (= prompt_embeds' dtype, :1061), so the same generator gives the same noise
            noise = torch.randn(st.tokens_shape, generator=gen, device=gen.device if isinstance(gen, torch.Generator) else device,
                                dtype=st.init_tokens.dtype).to(device)
            need = (st.cmask_dev.view(bsz, N) > 1.0 - 1e-6).unsqueeze(-1)
            noised = st.init_tokens.float() + st.image_cond_noise_scale * noise.float() * (t ** 2)
            st.lat32 = torch.where(need, noised, st.lat32.view(bsz, N, C)).contiguous().view(-1)
            st.lat16 = st.lat32.to(BF16)
        cp = getattr(st, "cond_parallel", None)
        nl = st.num_local_conds                 # the condition rows THIS rank runs (all of them without cond_parallel_group)
        if cp is not None and st.ltxv_model is not None and cp.any_flag(bool(getattr(st.ltxv_model, "_interrupt", False)), device):
            return None                         # collective decision at the step boundary: no rank is left waiting in the exchange
                                                # (one small all-reduce + host read per step, only when the caller passed an interrupt source)
        noise_pred = None
        if nl:
            st.x_in.view(nl, bsz, N, C).copy_(st.lat16.view(1, bsz, N, C).expand(nl, bsz, N, C))
            if st.cmask_dev is None:
                st.t_in.fill_(t)
            else:
                st.t_in.view(nl, bsz, N).copy_(torch.clamp(1.0 - st.cmask_dev, max=t).view(1, bsz, N).expand(nl, bsz, N))   # min(t, 1-mask) :1145-1150
            noise_pred = self.transformer(
                st.x_in, freqs_cis=st.freqs_cis, encoder_hidden_states=st.enc_b, encoder_attention_mask=st.mask_b,
                timestep=st.t_in, skip_layer_mask=st.skip_layer_masks[i] if st.skip_layer_masks is not None else None,
                skip_layer_strategy=st.skip_layer_strategy, latent_shape=st.latent_shape[2:], joint_pass=st.joint_pass,
                ltxv_model=st.ltxv_model if cp is None else None, return_dict=False, shared_prefix=getattr(st, "shared_prefix", None),
                mixed=getattr(st, "mixed", False), encoder_key_lens=getattr(st, "key_lens_b", None))[0]
            if noise_pred is None:
                return None
        n = N * C
        if cp is not None:                      # every owner's prediction rows -> every rank; guidance + step stay replicated
            lo, hi = st.cond_ranges[cp.rank]
            if nl:
                st.pred_full[lo * bsz:hi * bsz].copy_(noise_pred.view(nl * bsz, N, C))
            cp.exchange(st.pred_full, bsz, st.cond_ranges)
            noise_pred = st.pred_full
        pred = noise_pred.view(-1)
        for j in range(bsz):            # guidance statistics (cfg-star projection, std rescale) are per sample (:1183-1222)
            ops.guidance_step(pred[j * n:], st.lat32[j * n:(j + 1) * n], st.ts_dev, t, num_conds=num_conds,
                              has_cfg=st.do_cfg, has_stg=st.do_stg, do_rescale=st.do_rescaling,
                              guidance_scale=st.guidance_scale[i], stg_scale=st.stg_scale[i], rescale=st.rescaling_scale[i],
                              channels=C, cond_mask=None if st.cmask_dev is None else st.cmask_dev[j * N:(j + 1) * N], scratch=st.scratch,
                              latents_bf16=st.lat16[j * n:(j + 1) * n],
                              noise=self._step_noise(st) if st.stochastic_sampling else None, cond_stride=bsz * n)
        return st

    # ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, height: int, width: int, num_frames: int, frame_rate: float, prompt=None, negative_prompt=None,
                 num_inference_steps: int = 20, timesteps: List[int] = None,
                 guidance_scale: Union[float, List[float]] = 4.5, skip_layer_strategy: Optional[SkipLayerStrategy] = None,
                 skip_block_list=None, stg_scale: Union[float, List[float]] = 1.0,
                 rescaling_scale: Union[float, List[float]] = 0.7, guidance_timesteps: Optional[List[int]] = None,
                 num_images_per_prompt: Optional[int] = 1, eta: float = 0.0, generator=None,
                 latents: Optional[torch.Tensor] = None, prompt_embeds: Optional[torch.Tensor] = None,
                 prompt_attention_mask: Optional[torch.Tensor] = None, negative_prompt_embeds: Optional[torch.Tensor] = None,
                 negative_prompt_attention_mask: Optional[torch.Tensor] = None, output_type: Optional[str] = "pil",
                 return_dict: bool = True, callback_on_step_end: Optional[Callable] = None,
                 conditioning_items: Optional[List[ConditioningItem]] = None, decode_timestep=0.0,
                 decode_noise_scale=None, mixed_precision: bool = False, offload_to_cpu: bool = False,
                 enhance_prompt: bool = False, text_encoder_max_tokens: int = 256, stochastic_sampling: bool = False,
                 media_items: Optional[torch.Tensor] = None, strength: Optional[float] = 1.0,
                 skip_initial_inference_steps: int = 0, skip_final_inference_steps: int = 0, joint_pass: bool = False,
                 pass_no: int = -1, ltxv_model=None, callback=None, **kwargs):
        if prompt is not None or prompt_embeds is None:
            raise NotImplementedError("text encoding is out of scope: pass prompt_embeds / prompt_attention_mask")
        is_video = kwargs.get("is_video", False)
        vae_per_channel_normalize = kwargs.get("vae_per_channel_normalize", True)
        image_cond_noise_scale = kwargs.get("image_cond_noise_scale", 0.0)
        per_step = kwargs.get("_per_step_latents", None)          # test hook: list that receives each step's latents
        device = self._execution_device
        tr = self.transformer
        batch_size = prompt_embeds.shape[0]
        if num_images_per_prompt and num_images_per_prompt > 1:                   # :849-872 every prompt repeated per image
            rep = lambda t_: None if t_ is None else t_.repeat_interleave(num_images_per_prompt, dim=0)
            prompt_embeds, prompt_attention_mask = rep(prompt_embeds), rep(prompt_attention_mask)
            negative_prompt_embeds, negative_prompt_attention_mask = rep(negative_prompt_embeds), rep(negative_prompt_attention_mask)
        bsz = batch_size * (num_images_per_prompt or 1)

        video_scale = self.video_scale_factor if is_video else 1
        latent_height, latent_width = height // self.vae_scale_factor, width // self.vae_scale_factor
        latent_num_frames = num_frames // video_scale
        if is_video:
            latent_num_frames += 1
        latent_shape = (bsz, tr.config.in_channels, latent_num_frames, latent_height, latent_width)

        ts, num_inference_steps = retrieve_timesteps(self.scheduler, num_inference_steps, None, timesteps, max_timestep=strength,
                                                     skip_initial_inference_steps=skip_initial_inference_steps,
                                                     skip_final_inference_steps=skip_final_inference_steps,
                                                     samples_shape=latent_shape)
        ts_host = [float(x) for x in ts]
        ts_dev = ts.to(device=device, dtype=torch.float32).contiguous()
        n_steps = len(ts_host)
        if self.allowed_inference_steps is not None:
            for t_ in [round(x, 4) for x in ts_host]:
                assert t_ in self.allowed_inference_steps, f"Invalid inference timestep {t_}."

        # ---- per-step guidance tables (:958-1017)
        if guidance_timesteps:
            mapping = []
            for t_ in ts:          # 0-d fp32 tensors, like the reference: `val <= t_` is evaluated in fp32, so a threshold EQUAL to a timestep matches
                idx = [i for i, val in enumerate(guidance_timesteps) if val <= t_]
                mapping.append(idx[0] if len(idx) > 0 else (len(guidance_timesteps) - 1))
        per = lambda v: [v] * n_steps if not isinstance(v, list) else [v[mapping[i]] for i in range(n_steps)]
        guidance_scale = [x if x > 1.0 else 0.0 for x in per(guidance_scale)]
        stg_scale, rescaling_scale = per(stg_scale), per(rescaling_scale)
        do_cfg = any(x > 1.0 for x in guidance_scale)
        do_stg = any(x > 0.0 for x in stg_scale)
        do_rescaling = any(x != 1.0 for x in rescaling_scale)
        num_conds = 1 + int(do_cfg) + int(do_stg)
        if skip_block_list is not None:
            if len(skip_block_list) == 0 or not isinstance(skip_block_list[0], list):
                skip_block_list = [skip_block_list] * n_steps
            else:
                skip_block_list = [skip_block_list[mapping[i]] for i in range(n_steps)]
        skip_layer_masks = None
        if do_stg and skip_block_list is not None:
            skip_layer_masks = [tr.create_skip_layer_mask(bsz, num_conds, num_conds - 1, sb) for sb in skip_block_list]

        # ---- cond batch [uncond, text, perturbed] (:1034-1051)
        pe = prompt_embeds.to(device=device, dtype=BF16)
        pm = prompt_attention_mask.to(device)
        enc_b, mask_b = pe, pm
        if do_cfg:
            enc_b = torch.cat([negative_prompt_embeds.to(device=device, dtype=BF16), pe], dim=0)
            mask_b = torch.cat([negative_prompt_attention_mask.to(device), pm], dim=0)
        if do_stg:
            enc_b = torch.cat([enc_b, pe], dim=0)
            mask_b = torch.cat([mask_b, pm], dim=0)
        enc_b, mask_b = enc_b.contiguous(), mask_b.contiguous()
        # an all-ones prompt mask is a zero key bias ((1 - 1) * -10000, transformer3d.py:411-415): pass none at all, so that the 28
        # cross-attention launches per forward take the kernel's unmasked path (one host read per CALL, not per step)
        key_lens_b = None
        if bool((mask_b == 1).all()):
            mask_b = None
        else:
            # a RIGHT-PADDED mask (ones, then zeros: what the tokenizer's padding="max_length" gives) is a per-sample key length: the
            # kernel then skips the padded key blocks and the bias pass altogether (Transformer3DModel.forward, `encoder_key_lens`)
            mh = (mask_b > 0).to("cpu")
            lens = mh.sum(dim=1)
            if bool((lens >= 1).all()) and bool((mh == (torch.arange(mh.shape[1])[None, :] < lens[:, None])).all()):
                key_lens_b = lens.to(device=device, dtype=torch.int32).contiguous()

        # ---- latents (:1056-1088); drawn in prompt_embeds' dtype like the reference (:1061), kept as an fp32 master copy
        noise_dtype = prompt_embeds.dtype if prompt_embeds.dtype in (torch.float32, BF16) else torch.float32
        if mixed_precision:
            noise_dtype = torch.float32                                               # :1061
        init = self.prepare_latents(latents, media_items, ts_host[0], latent_shape, noise_dtype, device, generator,
                                    vae_per_channel_normalize=vae_per_channel_normalize)
        tokens, pixel_coords, conditioning_mask, num_cond_latents = self.prepare_conditioning(
            conditioning_items, init.clone(), num_frames, height, width, vae_per_channel_normalize=vae_per_channel_normalize,
            generator=generator)
        init_tokens = tokens.clone()
        frac = pixel_coords.to(torch.float32)
        frac[:, 0] = frac[:, 0] * (1.0 / frame_rate)
        freqs_cis = tr.precompute_freqs_cis(frac[:1])
        N, C = tokens.shape[1], tokens.shape[2]
        lat32 = tokens.to(torch.float32).contiguous().view(-1)            # [b*N*C] fp32 master copy
        lat16 = tokens.to(BF16).contiguous().view(-1)                      # bf16 model input
        cmask_dev = None if conditioning_mask is None else conditioning_mask.to(device=device, dtype=torch.float32).contiguous().view(-1)
        scratch = torch.empty(8 * 148, device=device, dtype=torch.float32)
        # ---- extension: guidance-condition parallelism (ltx/distributed/cond_parallel.py).  `cond_parallel_group=` is a
        # torch.distributed process group whose ranks all make this same call (same inputs, same generator seed): each runs the
        # transformer on its share of the [uncond, text, perturbed] rows, predictions are exchanged once per step.
        cond_parallel, cond_ranges, pred_full, num_local = None, None, None, num_conds
        if kwargs.get("cond_parallel_group") is not None:
            from .distributed.cond_parallel import CondParallel
            cond_parallel = CondParallel(kwargs["cond_parallel_group"])
            cond_ranges = cond_parallel.ranges(num_conds)
            lo, hi = cond_ranges[cond_parallel.rank]
            num_local = hi - lo
            rows = slice(lo * bsz, hi * bsz)
            enc_b = enc_b[rows].contiguous()
            mask_b = None if mask_b is None else mask_b[rows].contiguous()
            key_lens_b = None if key_lens_b is None else key_lens_b[rows].contiguous()
            if skip_layer_masks is not None:
                def cut(m):
                    if m is None:
                        return None
                    part = m[:, rows].contiguous()
                    if getattr(m, "_ltxb200_host", None) is not None:
                        part._ltxb200_host = m._ltxb200_host[:, rows].contiguous()    # keeps the forward free of a device sync
                    return part
                skip_layer_masks = [cut(m) for m in skip_layer_masks]
            pred_full = torch.empty(num_conds * bsz, N, C, device=device, dtype=BF16)
        x_in = torch.empty(max(num_local, 1) * bsz, N, C, device=device, dtype=BF16)
        t_in = torch.empty(max(num_local, 1) * bsz, N if cmask_dev is not None else 1, device=device, dtype=torch.float32)

        if callback is not None:
            callback(-1, None, True, override_num_inference_steps=num_inference_steps, pass_no=pass_no)

        st = SimpleNamespace(**dict(
            ts_host=ts_host, ts_dev=ts_dev, num_conds=num_conds, do_cfg=do_cfg, do_stg=do_stg, do_rescaling=do_rescaling,
            guidance_scale=guidance_scale, stg_scale=stg_scale, rescaling_scale=rescaling_scale,
            skip_layer_masks=skip_layer_masks, skip_layer_strategy=skip_layer_strategy, enc_b=enc_b, mask_b=mask_b,
            freqs_cis=freqs_cis, N=N, C=C, bsz=bsz, lat32=lat32, lat16=lat16, cmask_dev=cmask_dev, scratch=scratch, x_in=x_in,
            t_in=t_in, latent_shape=latent_shape, joint_pass=joint_pass, ltxv_model=ltxv_model, generator=generator,
            image_cond_noise_scale=image_cond_noise_scale, init_tokens=init_tokens, tokens_shape=tuple(tokens.shape),
            stochastic_sampling=bool(stochastic_sampling), mixed=bool(mixed_precision), key_lens_b=key_lens_b,
            cond_parallel=cond_parallel, cond_ranges=cond_ranges, pred_full=pred_full, num_local_conds=num_local,
            # extension (off by default): the perturbed STG condition repeats the text condition's inputs, so its rows are
            # copies of the text rows until the first skipped block (Transformer3DModel.forward, `shared_prefix`)
            shared_prefix=(bsz, int(do_cfg) * bsz) if (kwargs.get("share_stg_prefix", False) and do_stg and skip_layer_masks is not None and bsz == 1
                                                       and cond_parallel is None) else None))
        self._state = st
        if kwargs.get("_prepare_only", False):
            return st
        for i in range(n_steps):
            if self.denoise_step(st, i) is None:
                return None
            if per_step is not None:
                per_step.append(st.lat32.view(bsz, N, C).clone())
            if callback is not None:
                prev = st.lat32.view(bsz, N, C)[0, num_cond_latents:].transpose(0, 1).reshape(C, latent_num_frames, latent_height, latent_width)
                callback(i, prev, False, pass_no=pass_no)
            if callback_on_step_end is not None:
                callback_on_step_end(self, i, ts_host[i], {})
        lat32 = st.lat32

        lat = lat32.view(bsz, N, C)[:, num_cond_latents:]
        lat = self.patchifier.unpatchify(lat, latent_height, latent_width, tr.in_channels // math.prod(self.patchifier.patch_size))
        if output_type != "latent":
            if getattr(self.vae.decoder, "timestep_conditioning", False):          # :1271-1285 re-noise for the conditioned decoder
                dt = decode_timestep[0] if isinstance(decode_timestep, list) else decode_timestep
                dn = dt if decode_noise_scale is None else (decode_noise_scale[0] if isinstance(decode_noise_scale, list) else decode_noise_scale)
                noise = kwargs.get("_decode_noise")                                # test hook; the reference draws torch.randn_like
                if noise is None:
                    noise = torch.randn(lat.shape, device=lat.device, dtype=lat.dtype)
                lat = lat * (1 - dn) + noise.to(lat.device, lat.dtype) * dn
                decode_timestep = torch.tensor([dt], dtype=torch.float32)
            else:
                decode_timestep = None
            image = vae_decode(lat.contiguous(), self.vae, is_video, vae_per_channel_normalize=vae_per_channel_normalize,
                               timestep=decode_timestep)
            image = (image.float() / 2 + 0.5).clamp(0, 1)          # VaeImageProcessor.postprocess (:1298)
        else:
            image = lat
        if not return_dict:
            return (image,)
        return image


class LTXMultiScalePipeline:
    """pipeline_ltx_video.py:1741-1903: first pass at `downscale_factor` resolution -> LatentUpsampler (x2) -> AdaIN against the
    first-pass latents -> second pass from the partially re-noised upsampled latents -> decode -> bilinear resize to the requested size.
    Text encoding is out of scope here as everywhere: pass prompt_embeds / prompt_attention_mask (and the negative pair) in kwargs."""

    def __init__(self, video_pipeline: LTXVideoPipeline, latent_upsampler):
        self.video_pipeline = video_pipeline
        self.vae = video_pipeline.vae
        self.latent_upsampler = latent_upsampler

    def _upsample_latents(self, latest_upsampler, latents: torch.Tensor) -> torch.Tensor:
        """:1761-1772 un_normalize_latents -> upsampler -> normalize_latents, both affine maps fused into the layout kernels."""
        std, mean = self.vae.std_of_means.float().contiguous(), self.vae.mean_of_means.float().contiguous()
        z = latents.to(latest_upsampler.device)
        z = (z if z.dtype in (torch.float32, BF16) else z.float()).contiguous()
        return ops.latent_from_ndhwc(latest_upsampler.forward_ndhwc(ops.latent_to_ndhwc(z, std, mean)), std, mean)

    def __call__(self, downscale_factor: float, first_pass: dict, second_pass: dict, *args, **kwargs):
        from .latent_upsampler import adain_filter_latent
        video_pipeline = self.video_pipeline
        original_output_type = kwargs["output_type"]
        original_width, original_height = kwargs["width"], kwargs["height"]
        x_width = int(kwargs["width"] * downscale_factor)
        downscaled_width = x_width - (x_width % video_pipeline.vae_scale_factor)
        x_height = int(kwargs["height"] * downscale_factor)
        downscaled_height = x_height - (x_height % video_pipeline.vae_scale_factor)
        kwargs["output_type"] = "latent"
        kwargs["width"], kwargs["height"] = downscaled_width, downscaled_height
        # VAE_tile_size = (z_tile, hw_tile): the reference's callers always pass it (vae.py:92-115 answers (4, 0) on any GPU >= 24 GB) to
        # decode in overlapping temporal / spatial tiles that are blended back together (vae.py:223-263, 365-408) — a low-memory
        # approximation of the whole-video decode.  180 GB hold the whole video, so the value is accepted and the exact decode runs.
        kwargs.pop("VAE_tile_size", None)
        ltxv_model = kwargs.get("ltxv_model")
        for k in ("prompt", "negative_prompt", "device", "enhance_prompt"):
            kwargs.pop(k, None)                         # consumed by encode_prompt / the prompt enhancer in the reference (:1824-1850)
        if kwargs.get("prompt_embeds") is None:
            raise NotImplementedError("text encoding is out of scope: pass prompt_embeds / prompt_attention_mask")
        if ltxv_model is not None and getattr(ltxv_model, "_interrupt", False):
            return None
        steps1, steps2 = kwargs.pop("num_inference_steps1"), kwargs.pop("num_inference_steps2")
        kwargs["return_dict"] = True
        original_kwargs = kwargs.copy()

        kwargs["joint_pass"], kwargs["pass_no"] = True, 1
        kwargs.update(**first_pass)
        kwargs["num_inference_steps"] = steps1
        latents = video_pipeline(*args, **kwargs)
        if latents is None:
            return None
        upsampled_latents = self._upsample_latents(self.latent_upsampler, latents)
        upsampled_latents = adain_filter_latent(latents=upsampled_latents, reference_latents=latents)

        kwargs = original_kwargs
        kwargs["latents"] = upsampled_latents
        kwargs["output_type"] = original_output_type
        kwargs["width"], kwargs["height"] = downscaled_width * 2, downscaled_height * 2
        kwargs["joint_pass"], kwargs["pass_no"] = False, 2
        kwargs.update(**second_pass)
        kwargs["num_inference_steps"] = steps2
        result = video_pipeline(*args, **kwargs)
        if result is None:
            return None
        if original_output_type != "latent" and tuple(result.shape[-2:]) != (original_height, original_width):
            result = ops.bilinear_resize(result.float().contiguous(), original_height, original_width)     # :1890-1901
        return result
