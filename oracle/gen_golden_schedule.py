"""Per-step guidance / timestep tables of LTXVideoPipeline.__call__ (pipeline_ltx_video.py:943-1029: retrieve_timesteps with
skip_initial / skip_final / strength / explicit timesteps, the guidance_timesteps -> step mapping, list-valued guidance_scale /
stg_scale / rescaling_scale, nested skip_block_list, create_skip_layer_mask) — integer / indexing work, so BIT-EXACT
(TEST INFRASTRUCTURE ONLY).

The UNMODIFIED reference `__call__` is run on the CPU up to its first transformer call; the transformer's forward is a
recorder that reads the tables out of the calling frame's locals and answers None (which makes the reference return, :1172).
The product pipeline is run with its `_prepare_only` hook on a CPU stand-in transformer and must hold the same tables.
Cases: the reference's own presets (ltx_video/configs/*.yaml) plus the edge where explicit timesteps EQUAL the guidance
thresholds (the reference compares python doubles with fp32 tensor elements, i.e. in fp32).

Build container only (needs /root/reference):  python oracle/gen_golden_schedule.py
"""
import os
import sys
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "refshim"))
import load_reference  # noqa: E402

load_reference.install()
torch.set_grad_enabled(False)

NUM_LAYERS = 48                                   # LTX 13B depth: the 13B presets skip blocks up to 42
GEOM = dict(height=128, width=192, num_frames=17, frame_rate=25.0)      # latent (1,128,3,4,6)

DEV_13B = dict(guidance_scale=[1, 1, 6, 8, 6, 1, 1], stg_scale=[0, 0, 4, 4, 4, 2, 1], rescaling_scale=[1, 1, 0.5, 0.5, 1, 1, 1],
               guidance_timesteps=[1.0, 0.996, 0.9933, 0.9850, 0.9767, 0.9008, 0.6180],
               skip_block_list=[[], [11, 25, 35, 39], [22, 35, 39], [28], [28], [28], [28]])

CASES = {
    # ltxv-13b-0.9.7-dev.yaml first_pass / second_pass
    "13b_dev_first_pass": dict(DEV_13B, num_inference_steps=30, skip_final_inference_steps=3),
    "13b_dev_second_pass": dict(guidance_scale=[1], stg_scale=[1], rescaling_scale=[1], guidance_timesteps=[1.0], skip_block_list=[27],
                                num_inference_steps=30, skip_initial_inference_steps=17),
    # ltxv-13b-0.9.7-distilled.yaml first_pass / second_pass (explicit timesteps, guidance off)
    "13b_distilled_first_pass": dict(timesteps=[1.0000, 0.9937, 0.9875, 0.9812, 0.9750, 0.9094, 0.7250], guidance_scale=1, stg_scale=0,
                                     rescaling_scale=1, skip_block_list=[42], num_inference_steps=None),
    "13b_distilled_second_pass": dict(timesteps=[0.9094, 0.7250, 0.4219], guidance_scale=1, stg_scale=0, rescaling_scale=1,
                                      skip_block_list=[42], num_inference_steps=None),
    # ltxv-2b-0.9.6-dev.yaml
    "2b_dev": dict(guidance_scale=3, stg_scale=1, rescaling_scale=0.7, skip_block_list=[19], num_inference_steps=40),
    # explicit timesteps equal to the thresholds: val <= timestep is evaluated in fp32 by the reference
    "timesteps_equal_thresholds": dict(DEV_13B, timesteps=list(DEV_13B["guidance_timesteps"]), num_inference_steps=None),
    # img2img-style strength cut (max_timestep) with a step skipped at each end
    "strength_0p7": dict(guidance_scale=[1, 4], stg_scale=[0, 1], rescaling_scale=[1, 0.7], guidance_timesteps=[1.0, 0.5],
                         skip_block_list=[[], [19]], num_inference_steps=20, strength=0.7, skip_initial_inference_steps=1,
                         skip_final_inference_steps=1, _needs_latents=True),
}


def main():
    import ltx_video.pipelines.pipeline_ltx_video as R
    from ltx_video.models.autoencoders.causal_video_autoencoder import CausalVideoAutoencoder as RefVAE
    from ltx_video.models.transformers.symmetric_patchifier import SymmetricPatchifier as RefPatchifier
    from ltx_video.models.transformers.transformer3d import Transformer3DModel as RefTransformer
    from ltx_video.schedulers.rf import RectifiedFlowScheduler as RefScheduler
    from ltx_video.utils.diffusers_config_mapping import OURS_SCHEDULER_CONFIG, OURS_TRANSFORMER_CONFIG, OURS_VAE_CONFIG
    from ltx_video.utils.skip_layer_strategy import SkipLayerStrategy as RefStrategy
    from oracle.gen_golden import _NoInterrupt, _cuda_to_cpu
    from oracle.schedule_tables import compare, product_tables

    cfg = dict(OURS_TRANSFORMER_CONFIG)
    cfg.update(num_layers=NUM_LAYERS, num_attention_heads=2, attention_head_dim=32, cross_attention_dim=64)   # width is irrelevant here
    tr = RefTransformer.from_config(cfg).eval()
    rec = {}

    def recorder(*a, **k):
        f = sys._getframe()
        while f is not None and not (f.f_code.co_name == "__call__" and "guidance_mapping" in f.f_code.co_varnames):
            f = f.f_back
        L = f.f_locals
        masks = L["skip_layer_masks"]
        rec.update(timesteps=L["timesteps"].clone().float(), num_inference_steps=int(L["num_inference_steps"]),
                   guidance_scale=[float(x) for x in L["guidance_scale"]], stg_scale=[float(x) for x in L["stg_scale"]],
                   rescaling_scale=[float(x) for x in L["rescaling_scale"]], num_conds=int(L["num_conds"]),
                   skip_layer_masks=None if masks is None else [None if m is None else m.float().clone() for m in masks],
                   first_timestep_arg=k["timestep"].clone().float(), first_batch_rows=int(a[0].shape[0]))
        return (None,)
    tr.forward = recorder
    vae = RefVAE.from_config(dict(OURS_VAE_CONFIG)).eval()          # only the scale factors and isinstance() are used before the loop
    pipe = R.LTXVideoPipeline(tokenizer=None, text_encoder=None, vae=vae, transformer=tr,
                              scheduler=RefScheduler.from_config(dict(OURS_SCHEDULER_CONFIG)), patchifier=RefPatchifier(patch_size=1),
                              prompt_enhancer_image_caption_model=None, prompt_enhancer_image_caption_processor=None,
                              prompt_enhancer_llm_model=None, prompt_enhancer_llm_tokenizer=None)
    g = torch.Generator().manual_seed(3)
    pe, ne = torch.randn(1, 8, 4096, generator=g), torch.randn(1, 8, 4096, generator=g)
    pm = torch.ones(1, 8)
    init_latents = torch.randn(1, 128, 3, 4, 6, generator=g)
    out = {}
    for name, case in CASES.items():
        kw = {k: v for k, v in case.items() if not k.startswith("_")}
        lat = init_latents.clone() if case.get("_needs_latents") else None
        rec.clear()
        with _cuda_to_cpu():
            r = pipe(**GEOM, prompt_embeds=pe, prompt_attention_mask=pm, negative_prompt_embeds=ne, negative_prompt_attention_mask=pm,
                     generator=torch.Generator().manual_seed(5), output_type="latent", return_dict=False, joint_pass=True,
                     ltxv_model=_NoInterrupt(), is_video=True, vae_per_channel_normalize=True, latents=lat,
                     skip_layer_strategy=RefStrategy.AttentionValues, **kw)
        assert r is None and rec, name
        ref = dict(rec)
        mine = product_tables(NUM_LAYERS, GEOM, pe, pm, ne, pm, lat, kw)
        compare(name, mine, ref)
        n_masks = 0 if ref["skip_layer_masks"] is None else sum(m is not None for m in ref["skip_layer_masks"])
        print(f"  schedule[{name}]: {len(ref['timesteps'])} steps, num_conds {ref['num_conds']}, {n_masks} skip masks: bit-exact")
        out[name] = dict(kwargs=kw, needs_latents=bool(case.get("_needs_latents")), ref=ref)
    torch.save(dict(cases=out, geom=GEOM, num_layers=NUM_LAYERS, pe=pe, ne=ne, pm=pm, init_latents=init_latents),
               os.path.join(ROOT, "tests", "golden", "ltx_schedule_tables.pt"))
    print("written tests/golden/ltx_schedule_tables.pt")


if __name__ == "__main__":
    main()
