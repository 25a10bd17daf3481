"""The reference's integration-test kernel list (test/gp/lti_sde.jl:90-139), as (name, oracle kernel,
package-kernel factory) triples, plus its input/noise/mean grids (:148-162)."""
import numpy as np

from oracle import tgp_oracle as O


def _pk(pkg):
    M12, M32, M52, C = pkg.Matern12Kernel, pkg.Matern32Kernel, pkg.Matern52Kernel, pkg.ConstantKernel
    TK = pkg.gp.TransformedKernel
    return M12, M32, M52, C, TK


KERNELS = [
    ("base-Matern12", lambda: O.Matern12(), lambda p: p.Matern12Kernel()),
    ("base-Matern32", lambda: O.Matern32(), lambda p: p.Matern32Kernel()),
    ("base-Matern52", lambda: O.Matern52(), lambda p: p.Matern52Kernel()),
]
for s2 in (1e-1, 1.0, 10.0, 100.0):
    KERNELS.append((f"scaled-{s2}", (lambda s2=s2: s2 * O.Matern32()), (lambda p, s2=s2: s2 * p.Matern32Kernel())))
for lam in (1e-2, 0.1, 1.0, 10.0, 100.0):
    KERNELS.append((f"stretched-{lam}", (lambda lam=lam: O.Matern32().stretch(lam)),
                    (lambda p, lam=lam: p.gp.TransformedKernel(p.Matern32Kernel(), lam))))
for n in (7, 11):
    KERNELS.append((f"approx-periodic-{n}", (lambda n=n: O.ApproxPeriodic(n, 1.0)), (lambda p, n=n: p.ApproxPeriodicKernel(n, 1.0))))
KERNELS += [
    ("prod-M52-M32", lambda: (1.5 * (O.Matern52() * O.Matern32())).stretch(0.01),
     lambda p: p.gp.TransformedKernel(1.5 * (p.Matern52Kernel() * p.Matern32Kernel()), 0.01)),
    ("prod-M32-M52-Const", lambda: 3.0 * (O.Matern32() * O.Matern52() * O.Constant(1.0)),
     lambda p: 3.0 * (p.Matern32Kernel() * p.Matern52Kernel() * p.ConstantKernel(1.0))),
    ("sum-M12-M32", lambda: 1.5 * O.Matern12().stretch(0.1) + 0.3 * O.Matern32().stretch(1.1),
     lambda p: 1.5 * p.gp.TransformedKernel(p.Matern12Kernel(), 0.1) + 0.3 * p.gp.TransformedKernel(p.Matern32Kernel(), 1.1)),
    ("sum-M32-M52-Const", lambda: 2.0 * O.Matern32() + 0.5 * O.Matern52() + 1.0 * O.Constant(1.0),
     lambda p: 2.0 * p.Matern32Kernel() + 0.5 * p.Matern52Kernel() + 1.0 * p.ConstantKernel(1.0)),
]
KERNEL_IDS = [k[0] for k in KERNELS]

N = 13
MEANS = [("zero", None), ("const", 3.0), ("custom", lambda x: 2.0 * x)]


def inputs(regular, mod):
    rs = mod.RegularSpacing(0.0, 0.3, N)
    return rs if regular else rs.collect()


def noises(rng):
    return [("homoscedastic", 0.1), ("heteroscedastic", rng.uniform(size=N) + 0.1)]


def state_dim(ko):
    return O.lgssm_components(ko, O.RegularSpacing(0.0, 0.3, 2))[5][0].shape[0]
