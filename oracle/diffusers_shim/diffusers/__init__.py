"""ORACLE SHIM (test infrastructure only): a minimal stand-in for the `diffusers==0.24.0` package, exposing just the
symbols that /root/reference/src/models/stage2_inpaint_unet_2d_condition.py:21-44 and
/root/reference/src/pipelines/stage2_inpaint_pipeline.py:9-34 import, implemented on top of oracle/blocks.py and
oracle/schedulers.py.  It exists so that the reference's OWN classes can be executed unmodified in the build
container (diffusers itself is not installable: no network) and compared against oracle/unet.py + oracle/pipeline.py.
Never imported by the product (pcdms_b200/)."""
__version__ = "0.24.0"

from .pipeline_utils import DiffusionPipeline  # noqa: F401
