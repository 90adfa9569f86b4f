"""Shim of diffusers.loaders: empty mixins (LoRA / attn-proc loading is not on the path)."""


class UNet2DConditionLoadersMixin:
    pass


class LoraLoaderMixin:
    pass


class TextualInversionLoaderMixin:
    pass


class IPAdapterMixin:
    pass


class FromSingleFileMixin:
    pass
