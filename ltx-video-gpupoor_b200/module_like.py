"""The handful of torch.nn.Module conveniences the reference's loaders call on its model objects (`model.eval().requires_grad_(False)`,
wan/text2video.py:101; `vae.to(VAE_dtype)`, `latent_upsampler.to("cpu").eval()`, ltx_video/ltxv.py:176,199-200), kept so that those lines
keep working after the import swap.  The classes here are not nn.Modules: their weights are packed once, in bf16, into HBM by
`load_state_dict`, there is no autograd state and nothing is ever offloaded — so these calls have nothing to change and return `self`."""
import torch


class ModuleLike:
    def eval(self):
        return self

    def train(self, mode: bool = False):
        if mode:
            raise NotImplementedError("inference only: there is no training mode")
        return self

    def requires_grad_(self, requires_grad: bool = False):
        if requires_grad:
            raise NotImplementedError("inference only: the packed weights carry no gradients")
        return self

    def to(self, *args, **kwargs):
        """`.to(torch.bfloat16)` / `.to(<cuda device>)` describe what the object already is; `.to("cpu")` is the reference's offloading idiom
        and is ignored (the weights stay resident).  Any other dtype is refused: the kernels are bf16."""
        for a in list(args) + [kwargs.get("dtype")]:
            if isinstance(a, torch.dtype) and a != torch.bfloat16:
                raise NotImplementedError(f"the packed weights are bfloat16; .to({a}) is not supported")
        return self

    def cpu(self):
        return self

    def cuda(self, device=None):
        return self
