"""Checkpoint formats of the reference, host-side only:

* single-file `.safetensors` whose metadata carries `{"config": json({"transformer": …, "vae": …})}` — the layout the reference
  ships (transformer3d.py:313-325, causal_video_autoencoder.py:104-114, latent_upsampler.py:183-199; key prefixes
  `model.diffusion_model.` / `vae.` are stripped by the models' load_state_dict);
* the diffusers repository layout (`transformer/config.json` + `diffusion_pytorch_model*.safetensors`, `vae/…`) with the key
  renames of ltx_video/utils/diffusers_config_mapping.py:140-174 (transformer3d.py:278-311, causal_video_autoencoder.py:72-102);
* the legacy VAE directory (`config.json`, `autoencoder.pth`, `per_channel_statistics.json`; causal_video_autoencoder.py:42-70).

Everything here returns `(config: dict, state_dict: dict[str, Tensor on CPU])`; the models' `from_pretrained` classmethods build the
module from the config and hand the tensors to `load_state_dict` (which moves and repacks them on the GPU).
"""
from __future__ import annotations

import glob
import json
import os
from pathlib import Path
from typing import Dict, Tuple

import torch

PER_CHANNEL_STATISTICS_PREFIX = "per_channel_statistics."


def make_hashable_key(d):
    """diffusers_config_mapping.py:1-10: nested dict/list -> sorted tuples (dictionary key)."""
    def conv(v):
        if isinstance(v, list):
            return tuple(v)
        if isinstance(v, dict):
            return tuple(sorted((k, conv(x)) for k, x in v.items()))
        return v
    return tuple(sorted((k, conv(v)) for k, v in d.items()))


# The diffusers-side configs of Lightricks/LTX-Video (diffusers_config_mapping.py:13-62) and what they map to (:65-130)
DIFFUSERS_SCHEDULER_CONFIG = {
    "_class_name": "FlowMatchEulerDiscreteScheduler", "_diffusers_version": "0.32.0.dev0", "base_image_seq_len": 1024,
    "base_shift": 0.95, "invert_sigmas": False, "max_image_seq_len": 4096, "max_shift": 2.05, "num_train_timesteps": 1000,
    "shift": 1.0, "shift_terminal": 0.1, "use_beta_sigmas": False, "use_dynamic_shifting": True,
    "use_exponential_sigmas": False, "use_karras_sigmas": False}
DIFFUSERS_TRANSFORMER_CONFIG = {
    "_class_name": "LTXVideoTransformer3DModel", "_diffusers_version": "0.32.0.dev0", "activation_fn": "gelu-approximate",
    "attention_bias": True, "attention_head_dim": 64, "attention_out_bias": True, "caption_channels": 4096,
    "cross_attention_dim": 2048, "in_channels": 128, "norm_elementwise_affine": False, "norm_eps": 1e-06,
    "num_attention_heads": 32, "num_layers": 28, "out_channels": 128, "patch_size": 1, "patch_size_t": 1,
    "qk_norm": "rms_norm_across_heads"}
DIFFUSERS_VAE_CONFIG = {
    "_class_name": "AutoencoderKLLTXVideo", "_diffusers_version": "0.32.0.dev0", "block_out_channels": [128, 256, 512, 512],
    "decoder_causal": False, "encoder_causal": True, "in_channels": 3, "latent_channels": 128,
    "layers_per_block": [4, 3, 3, 3, 4], "out_channels": 3, "patch_size": 4, "patch_size_t": 1, "resnet_norm_eps": 1e-06,
    "scaling_factor": 1.0, "spatio_temporal_scaling": [True, True, True, False]}
OURS_SCHEDULER_CONFIG = {
    "_class_name": "RectifiedFlowScheduler", "_diffusers_version": "0.25.1", "num_train_timesteps": 1000, "shifting": "SD3",
    "base_resolution": None, "target_shift_terminal": 0.1}


def _ours_transformer_config():
    from .transformer3d import LTX_2B_CONFIG
    return dict(LTX_2B_CONFIG, _class_name="Transformer3DModel")


def _ours_vae_config():
    from .causal_video_autoencoder import LTX_VAE_CONFIG
    return dict(LTX_VAE_CONFIG, _class_name="CausalVideoAutoencoder")


def diffusers_and_ours_config_mapping():
    return {make_hashable_key(DIFFUSERS_SCHEDULER_CONFIG): OURS_SCHEDULER_CONFIG,
            make_hashable_key(DIFFUSERS_TRANSFORMER_CONFIG): _ours_transformer_config(),
            make_hashable_key(DIFFUSERS_VAE_CONFIG): _ours_vae_config()}


# substring renames, applied in this order to every key (diffusers_config_mapping.py:140-174)
TRANSFORMER_KEYS_RENAME_DICT = {"proj_in": "patchify_proj", "time_embed": "adaln_single", "norm_q": "q_norm", "norm_k": "k_norm"}
VAE_KEYS_RENAME_DICT = {
    "decoder.up_blocks.3.conv_in": "decoder.up_blocks.7", "decoder.up_blocks.3.upsamplers.0": "decoder.up_blocks.8",
    "decoder.up_blocks.3": "decoder.up_blocks.9", "decoder.up_blocks.2.upsamplers.0": "decoder.up_blocks.5",
    "decoder.up_blocks.2.conv_in": "decoder.up_blocks.4", "decoder.up_blocks.2": "decoder.up_blocks.6",
    "decoder.up_blocks.1.upsamplers.0": "decoder.up_blocks.2", "decoder.up_blocks.1": "decoder.up_blocks.3",
    "decoder.up_blocks.0": "decoder.up_blocks.1", "decoder.mid_block": "decoder.up_blocks.0",
    "encoder.down_blocks.3": "encoder.down_blocks.8", "encoder.down_blocks.2.downsamplers.0": "encoder.down_blocks.7",
    "encoder.down_blocks.2": "encoder.down_blocks.6", "encoder.down_blocks.1.downsamplers.0": "encoder.down_blocks.4",
    "encoder.down_blocks.1.conv_out": "encoder.down_blocks.5", "encoder.down_blocks.1": "encoder.down_blocks.3",
    "encoder.down_blocks.0.conv_out": "encoder.down_blocks.2", "encoder.down_blocks.0.downsamplers.0": "encoder.down_blocks.1",
    "encoder.down_blocks.0": "encoder.down_blocks.0", "encoder.mid_block": "encoder.down_blocks.9",
    "conv_shortcut.conv": "conv_shortcut", "resnets": "res_blocks", "norm3": "norm3.norm",
    "latents_mean": "per_channel_statistics.mean-of-means", "latents_std": "per_channel_statistics.std-of-means"}


def rename_keys(state_dict: Dict[str, torch.Tensor], table: Dict[str, str]) -> Dict[str, torch.Tensor]:
    out = {}
    for key, value in state_dict.items():
        new_key = key
        for old, new in table.items():
            new_key = new_key.replace(old, new)
        out[new_key] = value
    return out


def read_safetensors(path) -> Tuple[Dict[str, torch.Tensor], dict]:
    from safetensors import safe_open
    sd = {}
    with safe_open(str(path), framework="pt", device="cpu") as f:
        meta = f.metadata() or {}
        for k in f.keys():
            sd[k] = f.get_tensor(k)
    return sd, meta


def load_transformer_checkpoint(path) -> Tuple[dict, Dict[str, torch.Tensor]]:
    """transformer3d.py:271-326"""
    path = Path(path)
    if path.is_dir():
        with open(path / "transformer" / "config.json") as f:
            key = make_hashable_key(json.load(f))
        mapping = diffusers_and_ours_config_mapping()
        assert key in mapping, ("Provided diffusers checkpoint config for transformer is not suppported. "
                                "We only support diffusers configs found in Lightricks/LTX-Video.")
        sd = {}
        for part in sorted(glob.glob(str(path / "transformer" / "diffusion_pytorch_model*.safetensors"))):
            sd.update(read_safetensors(part)[0])
        return mapping[key], rename_keys(sd, TRANSFORMER_KEYS_RENAME_DICT)
    if path.is_file() and str(path).endswith(".safetensors"):
        sd, meta = read_safetensors(path)
        return json.loads(meta["config"])["transformer"], sd
    raise FileNotFoundError(f"{path}: expected a diffusers directory or a .safetensors file")


def load_vae_checkpoint(path) -> Tuple[dict, Dict[str, torch.Tensor]]:
    """causal_video_autoencoder.py:34-120"""
    path = Path(path)
    if path.is_dir() and (path / "autoencoder.pth").exists():
        with open(path / "config.json") as f:
            config = json.load(f)
        sd = torch.load(path / "autoencoder.pth", map_location="cpu")
        stats = path / "per_channel_statistics.json"
        if stats.exists():
            with open(stats) as f:
                data = json.load(f)
            cols = {c: torch.tensor(v) for c, v in zip(data["columns"], zip(*data["data"]))}
            sd[PER_CHANNEL_STATISTICS_PREFIX + "std-of-means"] = cols["std-of-means"]
            sd[PER_CHANNEL_STATISTICS_PREFIX + "mean-of-means"] = cols.get("mean-of-means", torch.zeros_like(cols["std-of-means"]))
        return config, sd
    if path.is_dir():
        with open(path / "vae" / "config.json") as f:
            key = make_hashable_key(json.load(f))
        mapping = diffusers_and_ours_config_mapping()
        assert key in mapping, ("Provided diffusers checkpoint config for VAE is not suppported. "
                                "We only support diffusers configs found in Lightricks/LTX-Video.")
        sd, _ = read_safetensors(path / "vae" / "diffusion_pytorch_model.safetensors")
        return mapping[key], rename_keys(sd, VAE_KEYS_RENAME_DICT)
    if path.is_file() and str(path).endswith(".safetensors"):
        sd, meta = read_safetensors(path)
        return json.loads(meta["config"])["vae"], sd
    raise FileNotFoundError(f"{path}: expected a VAE directory or a .safetensors file")


def load_upsampler_checkpoint(path) -> Tuple[dict, Dict[str, torch.Tensor]]:
    """latent_upsampler.py:177-199"""
    sd, meta = read_safetensors(path)
    return json.loads(meta["config"]), sd


def merge_lora(state_dict: Dict[str, torch.Tensor], lora: Dict[str, torch.Tensor], multiplier: float = 1.0,
               prefix: str = "diffusion_model.") -> int:
    """W += multiplier * (alpha / rank) * up @ down for every `<prefix><module>.lora_{down,up}.weight` (or lora_A / lora_B) pair —
    the merge the reference delegates to mmgp's offload.load_loras_into_model (ltx_video/ltxv.py, wan/text2video.py
    `offload.set_step_no_for_lora`): done once on the host copy of the weights, so the hot path stays a plain GEMM."""
    merged = 0
    for k in list(lora.keys()):
        for dn, un in ((".lora_down.weight", ".lora_up.weight"), (".lora_A.weight", ".lora_B.weight")):
            if not k.endswith(dn):
                continue
            mod = k[: -len(dn)]
            tgt = (mod[len(prefix):] if mod.startswith(prefix) else mod) + ".weight"
            if tgt not in state_dict or mod + un not in lora:
                continue
            down, up = lora[k].float(), lora[mod + un].float()
            alpha = float(lora[mod + ".alpha"]) if mod + ".alpha" in lora else float(down.shape[0])
            w = state_dict[tgt]
            state_dict[tgt] = (w.float() + multiplier * (alpha / down.shape[0]) * (up @ down)).to(w.dtype)
            merged += 1
    return merged
