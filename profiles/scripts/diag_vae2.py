import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import CausalVideoAutoencoder, vae_decode
from ltx_video_gpupoor_b200.ltx.init_weights import random_vae_decoder_state_dict
from oracle import ltx_oracle as O
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda")
z = torch.randn(1, 128, 2, 4, 6, generator=torch.Generator().manual_seed(2))
def ours(sd):
    vae = CausalVideoAutoencoder(); vae.load_state_dict(sd, device=dev)
    out = vae_decode(z.to(dev), vae, True, vae_per_channel_normalize=True).float().cpu()
    torch.cuda.synchronize(); return out
def oracle(sd, where):
    with torch.no_grad():
        return O.vae_decode({k: v.float().to(where) for k, v in sd.items()}, z.to(where)).float().cpu()
prod = random_vae_decoder_state_dict(seed=1, device=dev)
orc = O.make_vae_decoder_state_dict(seed=1)
for name, sd in (("oracle-init", orc), ("product-init", prod)):
    a_cpu, a_gpu = oracle(sd, "cpu"), oracle(sd, dev)
    b = ours(sd)
    b2 = ours({k: v.float().cpu() for k, v in sd.items()})
    print(name, "oracle cpu vs gpu rel_l2 %.2e" % O.rel_l2(a_gpu, a_cpu), "| ours vs oracle(cpu) rel_l2 %.3e PSNR %.1f" % (O.rel_l2(b, a_cpu), O.psnr(O.postprocess(b), O.postprocess(a_cpu))),
          "| ours(loaded from fp32 cpu copy) vs oracle %.3e" % O.rel_l2(b2, a_cpu), "| ours std %.3f oracle std %.3f" % (b.std(), a_cpu.std()), flush=True)
