"""GPU parity on BASELINE.json configs[0] EXACTLY: LTX-Video 2B at full depth (28 layers), t2v 256x256x9, 4 denoise steps +
VAE decode.  The fixture (tests/golden/ltx_config0.pt, oracle/gen_golden_config0.py) holds what the unmodified reference
pipeline produced in fp32 on the CPU from the same seeded weights, prompt embeddings and generator; the CUDA path runs the
same call in bf16.  Tolerances from BASELINE.json north_star: per-step latents <= 2e-2 relative L2, frames PSNR >= 40 dB.
(This file sorts last: generating the 1.9 B seeded fp32 weights on the host takes ~20 s.)"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import CausalVideoAutoencoder  # noqa: E402
from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import LTXVideoPipeline  # noqa: E402
from ltx_video_gpupoor_b200.ltx.rf import RectifiedFlowScheduler  # noqa: E402
from ltx_video_gpupoor_b200.ltx.symmetric_patchifier import SymmetricPatchifier  # noqa: E402
from ltx_video_gpupoor_b200.ltx.transformer3d import Transformer3DModel  # noqa: E402
from oracle import ltx_oracle as O  # noqa: E402


def test_baseline_config0_vs_reference_fixture(golden_dir):
    g = torch.load(os.path.join(golden_dir, "ltx_config0.pt"), weights_only=False)
    m = g["meta"]
    tr = Transformer3DModel(num_layers=m["num_layers"])
    tr.load_state_dict(O.make_transformer_state_dict(O.LTX_2B, seed=m["seed_weights"], num_layers=m["num_layers"]))
    vae = CausalVideoAutoencoder()
    vae.load_state_dict(O.make_vae_decoder_state_dict(seed=m["seed_vae"]))
    pipe = LTXVideoPipeline(vae=vae, transformer=tr, scheduler=RectifiedFlowScheduler(), patchifier=SymmetricPatchifier(1))
    pe = torch.randn(1, 32, 4096, generator=torch.Generator().manual_seed(m["seed_prompt"]))
    kw = dict(height=m["H"], width=m["W"], num_frames=m["F"], frame_rate=m["fps"], prompt_embeds=pe,
              prompt_attention_mask=torch.ones(1, 32), num_inference_steps=m["steps"], return_dict=False, is_video=True,
              vae_per_channel_normalize=True, guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0)
    per_step = []
    img = pipe(generator=torch.Generator("cpu").manual_seed(m["seed_noise"]), output_type="pt", _per_step_latents=per_step, **kw)[0]
    torch.cuda.synchronize()
    assert len(per_step) == m["steps"] and tuple(img.shape) == (1, 3, m["F"], m["H"], m["W"])
    for i, (a, b) in enumerate(zip(per_step, g["per_step_oracle"])):
        err = O.rel_l2(a.float().cpu(), b)
        print(f"config0 step {i}: latents rel_l2 vs the fp32 oracle (== reference to 1.3e-7) = {err:.3e}")
        assert err < 2e-2
    f, h, w = m["latent"]
    err = O.rel_l2(O.unpatchify(per_step[-1].float().cpu(), f, h, w), g["latents"])
    print(f"config0 final latents rel_l2 vs the reference fixture = {err:.3e}")
    assert err < 2e-2
    ps = O.psnr(img.float().cpu(), g["frames"].float())
    print(f"config0 decoded frames PSNR vs the reference fixture = {ps:.1f} dB")
    assert ps >= 40.0
