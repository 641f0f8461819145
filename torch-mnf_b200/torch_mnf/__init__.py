"""torch_mnf -- B200-native drop-in for the hot path of janosh/torch-mnf.

Same import paths, class names, constructor signatures and parameter names as the
reference (``torch_mnf.flows``, ``torch_mnf.layers``, ``torch_mnf.models``), but every
forward / inverse / kl_div runs in hand-written sm_100a CUDA kernels reached through the
C ABI of ``libmnf_b200.so`` (include/mnf_b200.h).  There is no CPU path: tensors must be
CUDA fp32 and the library must be built (``python torch-mnf_b200/build.py``).
"""

__version__ = "0.1.0"

from . import data, flows, layers, models  # noqa: F401,E402
