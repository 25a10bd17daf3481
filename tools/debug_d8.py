import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
from oracle import tgp_oracle as O
from tests.util import random_lgssm, sample_y, to_pkg_model
pkg = g.load_package(); h = pkg.default_handle(0)
for D in (6, 8):
    for T in (1, 2, 16, 17, 33, 512, 513, 700, 5000):
        for chunk in (0, 1, 4):
            rng = np.random.default_rng(D)
            m = random_lgssm(rng, T, D, "forward", True); y = sample_y(rng, m)
            ms, Ps, l = O.filter_(m, y)
            h.set_chunk(chunk)
            lml, steps = pkg.lgssm.logpdf(to_pkg_model(pkg, m), y, h, per_step=True)
            mf, Pf = pkg.lgssm._filter(to_pkg_model(pkg, m), y, h)
            bad = np.nonzero(np.abs(steps - l) > 1e-6)[0]
            print(f"D={D} T={T} chunk={chunk} lml_err={abs(lml-l.sum()):.2e} max_step_err={np.abs(steps-l).max():.2e} m_err={np.abs(mf-ms).max():.2e} first_bad={bad[:5]}")
