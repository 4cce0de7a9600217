"""massiv_b200 — massiv's stencil / windowed-load hot path on NVIDIA B200 (sm_100a).

Python host mirror of ``Data.Massiv.Array.Stencil`` over the C ABI of ``libmassiv_b200.so``
(include/massiv_b200.h).  See DESIGN.md and INTEGRATION.md.
"""
from ._abi import (IndexZeroException, MassivB200Error, SizeMismatchException)  # noqa: F401
from .stencil import (  # noqa: F401
    Border, Continue, DW, Device, DeviceArray, Edge, Fill, Padding, Reflect, Stencil, Sz, Wrap,
    applyStencil, avgStencil, compute, computeInto, computeSeparable, computeWithStride,
    getStencilCenter, getStencilSize, handleBorderIndex, idStencil, iterateStencil, iterateUntil, lifeStencil,
    makeConvolutionStencil, makeConvolutionStencilFromKernel, makeCorrelationStencil,
    makeCorrelationStencilFromKernel, makeUnsafeConvolutionStencil, makeUnsafeCorrelationStencil,
    mapStencil, maxStencil, maxStencils, minStencil, minStencils, noPadding, productStencil, samePadding, sqrt, sumStencil,
)

from . import integral  # noqa: F401,E402  (Data.Massiv.Array.Numeric.Integral)

__version__ = "0.1.0"
