from typing import Any

PipelineImageInput = Any


class VaeImageProcessor:
    def __init__(self, vae_scale_factor=8, **kwargs):
        self.vae_scale_factor = vae_scale_factor

    @staticmethod
    def denormalize(images):
        return (images / 2 + 0.5).clamp(0, 1)

    def postprocess(self, image, output_type="pil", do_denormalize=None):
        if output_type == "latent":
            return image
        return self.denormalize(image)
