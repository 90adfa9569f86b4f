"""pcdms_b200 — B200-native (sm_100a) implementation of the PCDMs stage-2 inpainting denoising hot path."""
__version__ = "0.1.0"
