"""bboptpy_b200 -- B200-native population engine behind bboptpy's optimizer API.

`from bboptpy_b200 import ActiveCMAES, ...` mirrors `from bboptpy import ...` for the classes on
the accelerated path (CMAES, ActiveCMAES, SepCMAES, CholeskyCMAES, IPopCMAES, BiPopCMAES, JADE, SHADE).  The
compute path is the C-ABI library libbbgpu.so (hand-written sm_100a CUDA); there is no CPU
fallback -- constructing an optimizer without a usable CUDA device raises.
"""
import glob as _glob
import os as _os

_HERE = _os.path.dirname(_os.path.abspath(__file__))

from . import _cabi  # noqa: E402,F401

if _glob.glob(_os.path.join(_HERE, "_bboptpy*.so")):
    from ._bboptpy import *  # noqa: E402,F401,F403
    from . import _bboptpy  # noqa: E402,F401
