"""Does replaying one LTX denoise step as a CUDA graph beat eager launches?  (launch-gap probe)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import CausalVideoAutoencoder
from ltx_video_gpupoor_b200.ltx.init_weights import random_transformer_state_dict, random_vae_decoder_state_dict
from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import LTXVideoPipeline
from ltx_video_gpupoor_b200.ltx.rf import RectifiedFlowScheduler
from ltx_video_gpupoor_b200.ltx.skip_layer_strategy import SkipLayerStrategy
from ltx_video_gpupoor_b200.ltx.symmetric_patchifier import SymmetricPatchifier
from ltx_video_gpupoor_b200.ltx.transformer3d import LTX_2B_CONFIG, Transformer3DModel

dev = torch.device("cuda", 0)
cfg = dict(LTX_2B_CONFIG)
tr = Transformer3DModel(**cfg)
tr.load_state_dict(random_transformer_state_dict(cfg, seed=0, device=dev), device=dev)
vae = CausalVideoAutoencoder()
vae.load_state_dict(random_vae_decoder_state_dict(seed=1, device=dev), device=dev)
pipe = LTXVideoPipeline(vae=vae, transformer=tr, scheduler=RectifiedFlowScheduler(), patchifier=SymmetricPatchifier(1))
g = torch.Generator().manual_seed(42)
pe = torch.randn(1, 256, 4096, generator=g).to(torch.bfloat16).to(dev)
ne = torch.randn(1, 256, 4096, generator=g).to(torch.bfloat16).to(dev)
pm = torch.ones(1, 256, device=dev)
st = pipe(prompt_embeds=pe, prompt_attention_mask=pm, negative_prompt_embeds=ne, negative_prompt_attention_mask=pm,
          num_inference_steps=30, generator=torch.Generator(device=dev).manual_seed(42), output_type="latent", _prepare_only=True,
          height=512, width=768, num_frames=121, frame_rate=25.0, guidance_scale=3.0, stg_scale=1.0, rescaling_scale=0.7,
          skip_block_list=[19], skip_layer_strategy=SkipLayerStrategy.AttentionValues, is_video=True,
          vae_per_channel_normalize=True, return_dict=False)


def timed(fn, n=10):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, host


eager, host = timed(lambda: pipe.denoise_step(st, 5))
print(f"eager: {eager:.2f} ms/step (host enqueue {host:.2f} ms/step)")
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    pipe.denoise_step(st, 5)
torch.cuda.current_stream().wait_stream(s)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    pipe.denoise_step(st, 5)
graph, host = timed(gr.replay)
print(f"graph: {graph:.2f} ms/step (host enqueue {host:.2f} ms/step)")
