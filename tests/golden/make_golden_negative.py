"""What the UNMODIFIED reference does when a particle starts with a negative susceptible count:

    python tests/golden/make_golden_negative.py        # writes tests/golden/negative_s0_reference.npz

pmcmc.py:156-169 draws I0 ~ Poisson(mu) and sets S0 = n_population - I0 without a clamp.  With n_population = 6 and mu = 4
about one particle in nine starts below zero; scipy returns nan for binom.pmf(k, n < 0, p) (and for norm.pdf with a negative
scale), np.random.choice refuses the weights and particle_filter returns (None, None, None) (pmcmc.py:187-192).  The
simulators themselves raise on a negative propensity (np.random.choice inside gillespie_algo.py:63).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

ref = ref_harness.load_reference()
pm, ga = ref["pmcmc"], ref["gillespie_algo"]
from scipy.stats import binom, norm  # noqa: E402

Y = np.array([[2, 4, 0], [1, 4, 1], [1, 3, 2], [0, 3, 3]], dtype=float)
np.random.seed(1)
state = np.random.get_state()
z, H, A = pm.particle_filter(Y, pm.ModelType.SIR, np.array([1.2, .4]), False, .4, 1000, 6, 4, jobs=1)
np.random.set_state(state)
I0 = np.random.poisson(4, 1000)                                           # the draw pmcmc.py:157 made
try:
    ga.sir_simulate([-1, 5, 0], np.array([1.2, .4]), 1.0, True)
    sim_err = "returned"
except Exception as e:                                                    # noqa: BLE001
    sim_err = f"{type(e).__name__}: {e}"
out = dict(Y=Y, theta=np.array([1.2, .4]), probs=.4, n_particles=1000, n_population=6, mu=4, numpy_seed=1,
           returned_none=z is None and H is None and A is None, n_negative_s0=int((6 - I0 < 0).sum()),
           binom_pmf_negative_n=np.array([binom.pmf(2, -1, .4), binom.pmf(0, -1, .4), binom.pmf(0, -3, .4)]),
           norm_pdf_negative_scale=np.array([norm.pdf(2.0, -1, .4 * -1 + .0001)]),
           simulator_error=sim_err)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "negative_s0_reference.npz"), **out)
for k, v in out.items():
    print(k, v if np.ndim(v) < 2 else np.asarray(v).shape)
