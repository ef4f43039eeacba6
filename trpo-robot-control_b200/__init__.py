"""trpo-robot-control_b200 -- B200-native natural-gradient solve (FVP + CG + TRPO update) behind the reference's C API.

The product is ``libtrpo_b200.so`` (hand-written sm_100a CUDA + C host code, C-ABI in ``include/trpo_b200.h``).
This package is the thin Python mirror used by the tests and ``bench.py``: ctypes bindings (``api``), the reference's
text formats (``textio``) and the synthetic batch generator (``synth``). There is no CPU fallback anywhere in it.

The directory name is not a Python identifier; load it with ``__graft_entry__.load_package()``.
"""
from . import api, synth, textio  # noqa: F401
from .api import (CG_GPU, FVP_GPU, TRPO_Update_GPU, Context, TRPOparam, ValueFunction, build_library,  # noqa: F401
                  library_path)
