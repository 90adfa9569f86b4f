"""Imported first by the tools that flip the pcdm_set_* experiment hooks: builds the experiment variant of the library
(-DPCDM_EXPERIMENT, pcdms_b200/libpcdm_b200_exp.so) and points pcdms_b200.lib at it.  The release library has no
such hooks (include/pcdm_b200_experiment.h)."""
import os
import sys

sys.path.insert(0, ".")
from pcdms_b200 import build as _build  # noqa: E402

os.environ.setdefault("PCDM_B200_LIB", str(_build.build(experiment=True)))
