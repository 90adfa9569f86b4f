from oracle.blocks import TimestepEmbedding, Timesteps  # noqa: F401


class _NotOnPath:
    def __init__(self, *a, **k):
        raise NotImplementedError("shim: this embedding type is not on the reference's stage-2 path")


class GaussianFourierProjection(_NotOnPath):
    pass


class TextImageProjection(_NotOnPath):
    pass


class TextImageTimeEmbedding(_NotOnPath):
    pass


class TextTimeEmbedding(_NotOnPath):
    pass
