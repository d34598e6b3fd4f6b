"""The oracle's Wan denoise loop against the reference's OWN `WanT2V.generate` method (wan/text2video.py:281-596), not a loop re-driven
by hand: the UNMODIFIED method is called unbound on a stand-in `self` that carries what the t2v path reads (device, strides, a text
encoder returning fixed embeddings, a VAE whose decode is the identity so that the final LATENTS come back, and the unmodified reference
WanModel in fp64 behind a dtype-casting wrapper — fp64 for the reason given in oracle/gen_golden_wan.py).  Covers: the noise drawn from
`seed`, UniPC and dpm++, CFG with and without the CFG-Zero* projection around `cfg_zero_step`, the two-call (x_id 0 / 1) and the joint
pass, guide_scale == 1.  Fixture: tests/golden/wan_generate.pt (TEST INFRASTRUCTURE ONLY).
Build container only (needs /root/reference):  python oracle/gen_golden_wan_generate.py"""
import os
import sys
import types
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "refshim"))
import load_reference  # noqa: E402

load_reference.install()
from oracle import wan_oracle as W  # noqa: E402
from oracle.gen_golden_wan import GOLD, TINY, build_ref  # noqa: E402
from oracle.ltx_oracle import rel_l2  # noqa: E402

torch.set_grad_enabled(False)

WIDTH, HEIGHT, FRAMES = 96, 64, 9               # latent (16, 3, 8, 12)
CASES = {
    "unipc_cfg_zero_star": dict(sample_solver="unipc", sampling_steps=4, guide_scale=5.0, cfg_star_switch=True, cfg_zero_step=1, joint_pass=False),
    "unipc_joint_plain_cfg": dict(sample_solver="unipc", sampling_steps=4, guide_scale=5.0, cfg_star_switch=False, cfg_zero_step=5, joint_pass=True),
    "dpmpp_cfg_zero_star": dict(sample_solver="dpm++", sampling_steps=5, guide_scale=3.0, cfg_star_switch=True, cfg_zero_step=0, joint_pass=True),
    "unipc_no_guidance": dict(sample_solver="unipc", sampling_steps=3, guide_scale=1, cfg_star_switch=True, cfg_zero_step=5, joint_pass=False),
}


class _Fp64Model:
    """The reference WanModel (fp64) behind a wrapper that casts the fp32 noise / latents the loop feeds it."""

    def __init__(self, ref):
        self.ref, self.enable_teacache = ref, False

    def __call__(self, x, **kw):
        kw["context"] = [c.double() for c in kw["context"]]
        return self.ref([u.double() for u in x], **kw)


class _Absent(types.ModuleType):
    """Stand-in for a third-party package the reference imports at module level but the t2v path never calls."""
    __path__ = []

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return type(k, (), {"__init__": lambda self, *a, **kw: None, "__call__": lambda self, *a, **kw: None})


def import_reference_text2video():
    """wan/text2video.py imports the T5 / VACE / media helpers at module level: `ftfy`, `imageio`, `decord`, `rembg` (not installed) get
    empty stand-ins, and t5.py:478 evaluates torch.cuda.current_device() as a default argument.  Nothing of the reference is modified."""
    if not torch.cuda.is_available():
        torch.cuda.current_device = lambda: 0
    for name in ("ftfy", "imageio", "decord", "rembg"):
        sys.modules.setdefault(name, _Absent(name))
    import wan.text2video as T
    return T


def main():
    T = import_reference_text2video()
    cfg = TINY
    sd = {k: v.double() for k, v in W.make_wan_state_dict(cfg, seed=0).items()}
    ref = build_ref(cfg, sd)
    g = torch.Generator().manual_seed(3)
    ctx, ctx0 = torch.randn(20, 4096, generator=g).double(), torch.randn(11, 4096, generator=g).double()
    me = SimpleNamespace(device=torch.device("cpu"), dtype=torch.float64, _interrupt=False, sample_neg_prompt="neg", vae_stride=(4, 8, 8),
                         patch_size=(1, 2, 2), num_train_timesteps=1000, model=_Fp64Model(ref),
                         text_encoder=lambda prompts, device: [ctx if prompts[0] == "pos" else ctx0],
                         vae=SimpleNamespace(model=SimpleNamespace(z_dim=16), decode=lambda x0, tile: x0))
    out = {}
    for name, kw in CASES.items():
        seed = 1234
        lat_ref = T.WanT2V.generate(me, "pos", width=WIDTH, height=HEIGHT, frame_num=FRAMES, shift=5.0, seed=seed, n_prompt="",
                                    model_filename="wan2.1_text2video_1.3B_bf16.safetensors", **kw)
        noise = torch.randn(16, 3, 8, 12, dtype=torch.float32, generator=torch.Generator().manual_seed(seed))      # :410
        mine = W.t2v_denoise(sd, cfg, noise.double(), ctx, ctx0, steps=kw["sampling_steps"], shift=5.0, guide_scale=kw["guide_scale"],
                             cfg_star_switch=kw["cfg_star_switch"], cfg_zero_step=kw["cfg_zero_step"], sample_solver=kw["sample_solver"])
        e = rel_l2(mine, lat_ref)
        print(f"  WanT2V.generate[{name}]: rel_l2(oracle loop, reference method) = {e:.3e}")
        assert e < 5e-5, name
        out[name] = dict(kw=kw, seed=seed, latents=lat_ref.float().clone())
    torch.save(dict(cfg=cfg, ctx=ctx.float(), ctx0=ctx0.float(), geom=(WIDTH, HEIGHT, FRAMES), cases=out), os.path.join(GOLD, "wan_generate.pt"))
    print("written tests/golden/wan_generate.pt")


if __name__ == "__main__":
    main()
