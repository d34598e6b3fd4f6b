"""RectifiedFlowScheduler timestep tables for every sampler x shifting combination the reference class offers (rf.py:176-257:
"Uniform" / "LinearQuadratic" / "Constant" x None / "SD3" (+ terminal stretch) / "SimpleDiffusion") — host-side float tables, BIT-EXACT.
Runs the UNMODIFIED reference scheduler and the product scheduler side by side on the CPU (set_timesteps is host code in both) and
stores the reference's tables in tests/golden/rf_scheduler_variants.pt (TEST INFRASTRUCTURE ONLY).
Build container only (needs /root/reference):  python oracle/gen_golden_rf_variants.py"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "refshim"))
import load_reference  # noqa: E402

load_reference.install()

CONFIGS = {
    "uniform_none": dict(sampler="Uniform", shifting=None),
    "uniform_sd3": dict(sampler="Uniform", shifting="SD3", target_shift_terminal=None),
    "uniform_sd3_terminal": dict(sampler="Uniform", shifting="SD3", target_shift_terminal=0.1),        # OURS_SCHEDULER_CONFIG
    "uniform_simple_diffusion": dict(sampler="Uniform", shifting="SimpleDiffusion", base_resolution=32 ** 2),
    "linear_quadratic_none": dict(sampler="LinearQuadratic", shifting=None),
    "linear_quadratic_sd3_terminal": dict(sampler="LinearQuadratic", shifting="SD3", target_shift_terminal=0.1),
    "constant_shift3": dict(sampler="Constant", shifting=None, shift=3.0),
}
SHAPES = [(1, 128, 2, 8, 8), (1, 128, 16, 16, 24), (1, 6144, 128)]
STEPS = [1, 2, 7, 30, 40]


def product_tables(cfg, steps, shape):
    from ltx_video_gpupoor_b200.ltx.rf import RectifiedFlowScheduler
    s = RectifiedFlowScheduler.from_config(dict(cfg, num_train_timesteps=1000))
    s.set_timesteps(steps, samples_shape=shape)
    return s.timesteps.clone()


def same(a, b):
    """bit-exact, NaN == NaN (one step with the terminal stretch is 0/0 in the reference and here)"""
    return a.dtype == b.dtype and a.shape == b.shape and torch.allclose(a, b, rtol=0, atol=0, equal_nan=True)


def main():
    from ltx_video.schedulers.rf import RectifiedFlowScheduler as Ref
    out = {}
    for name, cfg in CONFIGS.items():
        for steps in STEPS:
            for shape in SHAPES:
                r = Ref.from_config(dict(cfg, num_train_timesteps=1000))
                r.set_timesteps(steps, samples_shape=torch.Size(shape), device="cpu")
                mine = product_tables(cfg, steps, shape)
                assert same(mine, r.timesteps), (name, steps, shape, mine, r.timesteps)
                out[(name, steps, shape)] = r.timesteps.clone()
        print(f"  rf[{name}]: {len(STEPS) * len(SHAPES)} tables bit-exact")
    torch.save(dict(configs=CONFIGS, tables=out), os.path.join(ROOT, "tests", "golden", "rf_scheduler_variants.pt"))
    print("written tests/golden/rf_scheduler_variants.pt")


if __name__ == "__main__":
    main()
