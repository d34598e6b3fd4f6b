import sys, os, torch
sys.path.insert(0, os.getcwd())
from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import CausalVideoAutoencoder, vae_decode
from ltx_video_gpupoor_b200.ltx.init_weights import random_vae_decoder_state_dict
dev = "cuda"
vae = CausalVideoAutoencoder(); vae.load_state_dict(random_vae_decoder_state_dict(seed=1, device=dev), device=dev)
z = torch.randn(1, 128, 16, 16, 24, device=dev)
for _ in range(2): vae_decode(z, vae, True, vae_per_channel_normalize=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): img = vae_decode(z, vae, True, vae_per_channel_normalize=True)
b.record(); torch.cuda.synchronize()
print(f"LTX VAE decode 768x512x121: {a.elapsed_time(b)/5:.2f} ms  ({41.51e12/(a.elapsed_time(b)/5*1e-3)/1e12:.0f} TF/s)", img.float().abs().mean().item())
