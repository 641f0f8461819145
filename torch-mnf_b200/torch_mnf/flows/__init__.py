"""Normalizing flows with the reference's public names (flows/__init__.py:17-23).

Every flow is an ``nn.Module`` with ``forward(z) -> (x, log_det)`` and (except ``RNVP``)
``inverse(x) -> (z, log_det)``; containers return ``(list_of_intermediates, log_det[B])``.
"""

from .affine_constant_flow import ActNormFlow, AffineConstantFlow
from .affine_half_flow import AffineHalfFlow
from .core import NormalizingFlow, NormalizingFlowModel
from .glow import Glow
from .maf import IAF, MAF
from .rnvp import RNVP
from .spline_flow import NSF_AR, NSF_CL

__all__ = [
    "ActNormFlow", "AffineConstantFlow", "AffineHalfFlow", "NormalizingFlow", "NormalizingFlowModel",
    "Glow", "IAF", "MAF", "RNVP", "NSF_AR", "NSF_CL",
]
