"""Shim of diffusers.image_processor.VaeImageProcessor: only the tensor output types are supported."""


class VaeImageProcessor:
    def __init__(self, vae_scale_factor=8, **kwargs):
        self.vae_scale_factor = vae_scale_factor

    def postprocess(self, image, output_type="pil", do_denormalize=None):
        if output_type in ("latent", "pt"):
            return image
        raise NotImplementedError("shim: use output_type='pt' (PIL conversion is outside the hot path)")


PipelineImageInput = object
