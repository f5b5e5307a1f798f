"""demf_b200: B200-native implementation of the DeMF (VoteNet) data-parallel hot path.

Importing the package registers every model class under the reference's registry names
(the same import-side-effect mechanism as `import demf`, reference demf/__init__.py:1-5).
The CUDA library (libdemf_b200.so) is loaded lazily by the first op call and there is no CPU
fallback: see demf_b200/_lib.py.
"""
from .mm import bricks, losses, ms_deform_attn, pointnet_modules  # noqa: F401  (registration)
from . import modeling  # noqa: F401  (registration)

__version__ = "0.1.0"
