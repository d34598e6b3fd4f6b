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
    "unipc_joint_slg": dict(sample_solver="unipc", sampling_steps=4, guide_scale=5.0, cfg_star_switch=True, cfg_zero_step=0, joint_pass=True,
                            slg_layers=[1], slg_start=0.25, slg_end=0.75),
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
                             cfg_star_switch=kw["cfg_star_switch"], cfg_zero_step=kw["cfg_zero_step"], sample_solver=kw["sample_solver"],
                             slg_layers=kw.get("slg_layers"), slg_start=kw.get("slg_start", 0.0), slg_end=kw.get("slg_end", 1.0))
        e = rel_l2(mine, lat_ref)
        print(f"  WanT2V.generate[{name}]: rel_l2(oracle loop, reference method) = {e:.3e}")
        assert e < 5e-5, name
        out[name] = dict(kw=kw, seed=seed, latents=lat_ref.float().clone())
    torch.save(dict(cfg=cfg, ctx=ctx.float(), ctx0=ctx0.float(), geom=(WIDTH, HEIGHT, FRAMES), cases=out), os.path.join(GOLD, "wan_generate.pt"))
    print("written tests/golden/wan_generate.pt")


def _mask(frame_num, lat_h, lat_w, any_end, added):
    """The oracle's statement of the conditioning-frame mask (image2video.py:232-244) -> [4, latent frames, lat_h, lat_w]."""
    m = torch.zeros(frame_num, lat_h, lat_w)
    m[0] = 1
    if any_end:
        m[-1] = 1
    parts = [m[:1].repeat(4, 1, 1), m[1:-1] if (any_end and added) else m[1:]]
    if any_end and added:
        parts.append(m[-1:].repeat(4, 1, 1))
    m = torch.cat(parts)
    return m.view(m.shape[0] // 4, 4, lat_h, lat_w).transpose(0, 1)


def main_i2v():
    """WanI2V.generate (wan/image2video.py:124-428), the reference's own method on a stand-in self: CLIP and the VAE encoder answer with fixed
    tensors (they are inputs of this path), so what is pinned is the frame / latent-frame arithmetic incl. the added end frame, the mask and
    `y = [mask | latent]`, the noise from `seed`, and the loop with clip_fea / y on the i2v WanModel."""
    import_reference_text2video()
    import wan.image2video as I
    from mmgp import offload
    from PIL import Image
    offload.last_offload_obj = SimpleNamespace(unload_all=lambda: None)          # image2video.py:264 frees the text encoder / CLIP here
    cfg = dict(TINY, model_type="i2v", in_dim=36, clip_dim=1280)
    sd = {k: v.double() for k, v in W.make_wan_state_dict(cfg, seed=1).items()}
    ref = build_ref(cfg, sd)
    g = torch.Generator().manual_seed(5)
    clip = torch.randn(1, 257, 1280, generator=g).double()
    ctx, ctx0 = torch.randn(20, 4096, generator=g).double(), torch.randn(11, 4096, generator=g).double()
    lat_y = {3: torch.randn(16, 3, 8, 12, generator=g).double(), 4: torch.randn(16, 4, 8, 12, generator=g).double()}   # by latent frames
    seen = {}

    def encode(videos, tile, any_end_frame=False):
        f = videos[0].shape[1]
        seen.update(frames=f, any_end_frame=any_end_frame)
        return [lat_y[(f - 2) // 4 + 2 if any_end_frame else (f - 1) // 4 + 1]]
    me = SimpleNamespace(device=torch.device("cpu"), dtype=torch.float64, VAE_dtype=torch.float32, _interrupt=False, sample_neg_prompt="neg",
                         vae_stride=(4, 8, 8), patch_size=(1, 2, 2), num_train_timesteps=1000, model=_Fp64Model(ref),
                         text_encoder=lambda prompts, device: [ctx if prompts[0] == "pos" else ctx0],
                         clip=SimpleNamespace(model=SimpleNamespace(image_size=224), visual=lambda imgs: clip),
                         vae=SimpleNamespace(encode=encode, decode=lambda x0, tile, any_end_frame=False: x0))
    img = Image.new("RGB", (96, 64), (120, 30, 200))
    out = {}
    for name, kw in {"start_image": dict(image_end=None, sampling_steps=4, guide_scale=5.0, cfg_star_switch=True, cfg_zero_step=1, joint_pass=True),
                     "start_and_end_image": dict(image_end=img, sampling_steps=3, guide_scale=5.0, cfg_star_switch=False, cfg_zero_step=5,
                                                 joint_pass=False)}.items():
        seed = 77
        lat_ref = I.WanI2V.generate(me, "pos", img, height=64, width=96, frame_num=9, shift=5.0, seed=seed, n_prompt="",
                                    sample_solver="unipc", model_filename="wan2.1_image2video_480p_14B_bf16.safetensors", **kw)
        any_end = kw["image_end"] is not None
        frames, lat_frames = (10, 4) if any_end else (9, 3)                                  # :191-199
        assert seen["frames"] == frames and seen["any_end_frame"] == any_end
        noise = torch.randn(16, lat_frames, 8, 12, dtype=torch.float32, generator=torch.Generator().manual_seed(seed))
        y = torch.cat([_mask(frames, 8, 12, any_end, True).double(), lat_y[lat_frames]])
        mine = W.t2v_denoise(sd, cfg, noise.double(), ctx, ctx0, steps=kw["sampling_steps"], shift=5.0, guide_scale=kw["guide_scale"],
                             cfg_star_switch=kw["cfg_star_switch"], cfg_zero_step=kw["cfg_zero_step"], clip_fea=clip, y=y)
        if any_end:
            mine = mine[:, :-1]                                                                # :422-424 drops the added frame after decoding
        e = rel_l2(mine, lat_ref)
        print(f"  WanI2V.generate[{name}]: rel_l2(oracle loop, reference method) = {e:.3e}")
        assert e < 5e-5, name
        out[name] = dict(kw={k: v for k, v in kw.items() if k != "image_end"}, any_end=any_end, seed=seed, frames=frames,
                         lat_frames=lat_frames, y=y.float(), latents=lat_ref.float().clone())
    torch.save(dict(cfg=cfg, clip=clip.float(), ctx=ctx.float(), ctx0=ctx0.float(), cases=out), os.path.join(GOLD, "wan_i2v_generate.pt"))
    print("written tests/golden/wan_i2v_generate.pt")


if __name__ == "__main__":
    main()
    main_i2v()
