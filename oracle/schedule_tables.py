"""TEST INFRASTRUCTURE ONLY: runs the PRODUCT pipeline's host-side table construction (LTXVideoPipeline.__call__ up to the loop,
through its `_prepare_only` hook) on the CPU with a stand-in transformer — no kernel is launched — and compares the tables with
the ones recorded from the unmodified reference (oracle/gen_golden_schedule.py; pipeline_ltx_video.py:943-1029, 1120-1149)."""
from types import SimpleNamespace

import torch


def product_tables(num_layers, geom, pe, pm, ne, nm, latents, kw):
    from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import LTXVideoPipeline
    from ltx_video_gpupoor_b200.ltx.rf import RectifiedFlowScheduler
    from ltx_video_gpupoor_b200.ltx.skip_layer_strategy import SkipLayerStrategy
    from ltx_video_gpupoor_b200.ltx.symmetric_patchifier import SymmetricPatchifier
    from ltx_video_gpupoor_b200.ltx.transformer3d import Transformer3DModel

    class _CpuTransformer:                       # what the table code touches of Transformer3DModel
        device, dtype, in_channels = torch.device("cpu"), torch.float32, 128
        config = SimpleNamespace(in_channels=128, causal_temporal_positioning=False)
        create_skip_layer_mask = Transformer3DModel.create_skip_layer_mask

        def __init__(self):
            self.num_layers, self._skip_host = num_layers, {}

        def precompute_freqs_cis(self, frac):
            return None
    vae = SimpleNamespace(spatial_downscale_factor=32, temporal_downscale_factor=8)
    pipe = LTXVideoPipeline(vae=vae, transformer=_CpuTransformer(), scheduler=RectifiedFlowScheduler(), patchifier=SymmetricPatchifier(1))
    st = pipe(**geom, prompt_embeds=pe, prompt_attention_mask=pm, negative_prompt_embeds=ne, negative_prompt_attention_mask=nm,
              generator=torch.Generator().manual_seed(5), output_type="latent", return_dict=False, joint_pass=True, is_video=True,
              vae_per_channel_normalize=True, latents=latents, skip_layer_strategy=SkipLayerStrategy.AttentionValues,
              _prepare_only=True, **kw)
    masks = st.skip_layer_masks
    # the timestep argument of the first forward (:1120-1143): t expanded over the cond batch, one column
    first_t = st.ts_dev[0].reshape(1, 1).expand(st.num_conds, 1).clone().float()
    return dict(timesteps=st.ts_dev.clone().float().cpu(), num_inference_steps=len(st.ts_host),
                guidance_scale=[float(x) for x in st.guidance_scale], stg_scale=[float(x) for x in st.stg_scale],
                rescaling_scale=[float(x) for x in st.rescaling_scale], num_conds=int(st.num_conds),
                skip_layer_masks=None if masks is None else [None if m is None else m.float().clone() for m in masks],
                first_timestep_arg=first_t, first_batch_rows=int(st.x_in.shape[0]))


def compare(name, mine, ref):
    """Bit-exact comparison of the product's tables with the reference's."""
    assert torch.equal(mine["timesteps"], ref["timesteps"]), (name, mine["timesteps"], ref["timesteps"])
    for k in ("num_inference_steps", "guidance_scale", "stg_scale", "rescaling_scale", "num_conds", "first_batch_rows"):
        assert mine[k] == ref[k], (name, k, mine[k], ref[k])
    assert torch.equal(mine["first_timestep_arg"], ref["first_timestep_arg"]), name
    a, b = mine["skip_layer_masks"], ref["skip_layer_masks"]
    assert (a is None) == (b is None), name
    if a is not None:
        assert len(a) == len(b), name
        for x, y in zip(a, b):
            assert (x is None) == (y is None) and (x is None or torch.equal(x, y)), name
