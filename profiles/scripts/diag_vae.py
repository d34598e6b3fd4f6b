"""Decode parity of the LTX VAE at growing latent sizes: CUDA drop-in vs the fp32 oracle on the GPU (TF32 off), with the halo-tiled
convolution / the fused PixelNorm epilogue switched on and off (LTXB200_CONV_HALO, LTXB200_VAE_FUSED_NORM are read per call)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import CausalVideoAutoencoder, vae_decode
from ltx_video_gpupoor_b200.ltx.init_weights import random_vae_decoder_state_dict
from oracle import ltx_oracle as O

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda")
vsd = random_vae_decoder_state_dict(seed=1, device=dev)
vae = CausalVideoAutoencoder()
vae.load_state_dict(vsd, device=dev)
sd32 = {k: v.float() for k, v in vsd.items()}
sizes = [(2, 4, 6), (3, 8, 12), (5, 16, 24), (16, 16, 24)] if len(sys.argv) < 2 else [tuple(int(x) for x in sys.argv[1].split(","))]
for (f, h, w) in sizes:
    z = torch.randn(1, 128, f, h, w, device=dev, generator=torch.Generator(device=dev).manual_seed(f))
    with torch.no_grad():
        ref = O.postprocess(O.vae_decode(sd32, z)).cpu()
    for halo, fused in (("1", "1"), ("0", "1"), ("1", "0"), ("0", "0")):
        os.environ["LTXB200_CONV_HALO"], os.environ["LTXB200_VAE_FUSED_NORM"] = halo, fused
        img = vae_decode(z, vae, True, vae_per_channel_normalize=True)
        torch.cuda.synchronize()
        out = O.postprocess(img.float()).cpu()
        print(f"latent {f}x{h}x{w}: halo={halo} fused_norm={fused}: PSNR vs fp32 oracle = {O.psnr(out, ref):.1f} dB", flush=True)
        del img, out
    del ref
