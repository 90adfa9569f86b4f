from oracle.blocks import get_activation  # noqa: F401
