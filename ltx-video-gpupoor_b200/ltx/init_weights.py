"""Random-init weights of the named architectures in the reference's state_dict layout (there is no
network for checkpoints; BASELINE.json asks for random-init weights of the named architecture).
Distributions follow the reference modules' defaults: nn.Linear / nn.Conv3d kaiming-uniform
(= U(+-1/sqrt(fan_in))), scale_shift_table ~ randn/sqrt(D) (attention.py:183-185, transformer3d.py:141-143)."""
import math
from typing import Dict

import torch

from .causal_video_autoencoder import LTX_VAE_CONFIG
from .transformer3d import LTX_2B_CONFIG


def _u(shape, bound, gen, device, dtype):
    return ((torch.rand(shape, generator=gen, device=device, dtype=torch.float32) * 2 - 1) * bound).to(dtype)


def random_transformer_state_dict(config: dict = LTX_2B_CONFIG, seed: int = 0, device="cuda",
                                  dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    gen = torch.Generator(device=device).manual_seed(seed)
    D = config["num_attention_heads"] * config["attention_head_dim"]
    sd = {}

    def lin(name, o, i):
        sd[name + ".weight"] = _u((o, i), 1 / math.sqrt(i), gen, device, dtype)
        sd[name + ".bias"] = _u((o,), 1 / math.sqrt(i), gen, device, dtype)

    sd["scale_shift_table"] = (torch.randn(2, D, generator=gen, device=device) / D ** 0.5).to(dtype)
    lin("patchify_proj", D, config["in_channels"])
    for i in range(config["num_layers"]):
        p = f"transformer_blocks.{i}."
        sd[p + "scale_shift_table"] = (torch.randn(6, D, generator=gen, device=device) / D ** 0.5).to(dtype)
        for a, kv in (("attn1", D), ("attn2", config["cross_attention_dim"])):
            sd[p + a + ".q_norm.weight"] = torch.ones(D, device=device, dtype=dtype)
            sd[p + a + ".k_norm.weight"] = torch.ones(D, device=device, dtype=dtype)
            lin(p + a + ".to_q", D, D); lin(p + a + ".to_k", D, kv); lin(p + a + ".to_v", D, kv); lin(p + a + ".to_out.0", D, D)
        lin(p + "ff.net.0.proj", 4 * D, D); lin(p + "ff.net.2", D, 4 * D)
    lin("proj_out", config["out_channels"], D)
    lin("adaln_single.emb.timestep_embedder.linear_1", D, 256)
    lin("adaln_single.emb.timestep_embedder.linear_2", D, D)
    lin("adaln_single.linear", 6 * D, D)
    lin("caption_projection.linear_1", D, config["caption_channels"])
    lin("caption_projection.linear_2", D, D)
    return sd


def random_vae_decoder_state_dict(config: dict = LTX_VAE_CONFIG, seed: int = 1, device="cuda",
                                  dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    gen = torch.Generator(device=device).manual_seed(seed)
    sd = {}

    def conv(name, o, i, k=3):
        b = 1 / math.sqrt(i * k ** 3)
        sd[name + ".weight"] = _u((o, i, k, k, k), b, gen, device, dtype)
        sd[name + ".bias"] = _u((o,), b, gen, device, dtype)

    ch = 128 * 2 ** len([b for b in config["blocks"] if b[0] == "res_x_y"])
    conv("decoder.conv_in.conv", ch, config["latent_channels"])
    for idx, (name, n) in enumerate(reversed(config["blocks"])):
        p = f"decoder.up_blocks.{idx}."
        if name == "res_x":
            for j in range(int(n)):
                conv(p + f"res_blocks.{j}.conv1.conv", ch, ch); conv(p + f"res_blocks.{j}.conv2.conv", ch, ch)
        elif name == "res_x_y":
            conv(p + "conv1.conv", ch // 2, ch); conv(p + "conv2.conv", ch // 2, ch // 2)
            conv(p + "conv_shortcut", ch // 2, ch, k=1)
            sd[p + "norm3.norm.weight"] = torch.ones(ch, device=device, dtype=dtype)
            sd[p + "norm3.norm.bias"] = torch.zeros(ch, device=device, dtype=dtype)
            ch //= 2
        else:
            conv(p + "conv.conv", 8 * ch, ch)
    conv("decoder.conv_out.conv", config["out_channels"] * config["patch_size"] ** 2, ch)
    sd["std_of_means"] = 0.5 + torch.rand(config["latent_channels"], generator=gen, device=device)
    sd["mean_of_means"] = 0.1 * torch.randn(config["latent_channels"], generator=gen, device=device)
    return sd


def random_vae_encoder_state_dict(config: dict = LTX_VAE_CONFIG, seed: int = 2, device="cuda",
                                  dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """encoder.* keys of CausalVideoAutoencoder (causal_video_autoencoder.py:313-557), random init."""
    gen = torch.Generator(device=device).manual_seed(seed)
    sd = {}

    def conv(name, o, i, k=3):
        b = 1 / math.sqrt(i * k ** 3)
        sd[name + ".weight"] = _u((o, i, k, k, k), b, gen, device, dtype)
        sd[name + ".bias"] = _u((o,), b, gen, device, dtype)

    ch = 128
    conv("encoder.conv_in.conv", ch, config["in_channels"] * config["patch_size"] ** 2)
    for idx, (name, n) in enumerate(config["blocks"]):
        p = f"encoder.down_blocks.{idx}."
        if name == "res_x":
            for j in range(int(n)):
                conv(p + f"res_blocks.{j}.conv1.conv", ch, ch); conv(p + f"res_blocks.{j}.conv2.conv", ch, ch)
        elif name == "res_x_y":
            conv(p + "conv1.conv", 2 * ch, ch); conv(p + "conv2.conv", 2 * ch, 2 * ch)
            conv(p + "conv_shortcut", 2 * ch, ch, k=1)
            sd[p + "norm3.norm.weight"] = torch.ones(ch, device=device, dtype=dtype)
            sd[p + "norm3.norm.bias"] = torch.zeros(ch, device=device, dtype=dtype)
            ch *= 2
        else:
            conv(p + "conv", ch, ch)
    conv("encoder.conv_out.conv", config["latent_channels"] + 1, ch)
    return sd
