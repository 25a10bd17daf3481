"""
temporalgps.jl_b200 — B200-native (sm_100a) LGSSM inference hot path of TemporalGPs.jl.

Layout: csrc/ (CUDA kernels + the C ABI of libtgpb200.so), _lib.py (ctypes binding, the stand-in
for the Julia ccall glue), lgssm.py (mirror of the reference's L2 entry points: logpdf, _filter,
posterior, marginals), gp.py (mirror of the GP layer that calls them: to_sde, GP, posterior...).
The directory name contains a dot, so load it with importlib (see __graft_entry__.load_package()).
"""
from . import _lib, gp, lgssm
from ._lib import (DimensionMismatch, Handle, PosDefException, TGPError, default_handle, TGP_ALGO_AUTO, TGP_ALGO_SCAN)
from .build import build
from .gp import (GP, ApproxPeriodicKernel, ArrayStorage, B200Storage, ConstantKernel, FiniteLTISDE, Matern12Kernel,
                 Matern32Kernel, Matern52Kernel, RectilinearGrid, RegularSpacing, SArrayStorage, SEKernel, Separable, to_sde,
                 with_lengthscale)
from .lgssm import (LGSSM, BottleneckEmissions, Fill, Forward, Gaussian, GaussMarkovModel, LargeOutputEmissions, Reverse,
                    ScalarEmissions, SmallOutputEmissions)

__all__ = [n for n in dir() if not n.startswith("_")]
