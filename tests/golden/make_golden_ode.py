"""Golden vectors of the ODE data synthesiser (pmcmc.py:16-113), made by running the UNMODIFIED reference here
(scipy odeint + the pandas daily sub-sampling):

    python tests/golden/make_golden_ode.py        # writes tests/golden/ode_synth.npz

Cases are the reference's own uses: tests/test_noise.py:16-20 (SIR), tests/test_pmcmc_seir.py (SEIR),
tests/test_pmcmc_sir_subgrps.py:15-21 (2 subgroups), plus a 3-group case and a coarse grid.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

ref = ref_harness.load_reference()
pm = ref["pmcmc"]
out = {}
t200 = np.linspace(0, 14, num=200)
out["t200"] = t200
out["sir"] = pm.sir_simulate_discrete((4800, 20, 0), t200, 2, 1).to_numpy(dtype=float)                       # time, S, I, R
out["sir_slow"] = pm.sir_simulate_discrete((9980, 20, 0), np.linspace(0, 100, num=1500), .4, .2).to_numpy(dtype=float)
out["t1500"] = np.linspace(0, 100, num=1500)
out["seir"] = pm.seir_simulate_discrete((4800, 0, 20, 0), t200, 2.0, 1.0, 1.0).to_numpy(dtype=float)
pop2 = np.array([[2000, 30, 0], [3000, 40, 0]]); beta2 = np.array([[5, 2], [1, 3]])
out["sub2"] = pm.sir_subgroups_simulate_discrete(pop2, t200, beta2, .5).to_numpy(dtype=float)                # S0,I0,R0,S1,I1,R1,time
pop3 = np.array([[1000, 10, 0], [2000, 5, 0], [1500, 0, 0]]); beta3 = np.array([[1.5, .2, .1], [.3, 1.0, .4], [.2, .5, 2.0]])
out["sub3"] = pm.sir_subgroups_simulate_discrete(pop3, t200, beta3, .4).to_numpy(dtype=float)
out["pop2"], out["beta2"], out["pop3"], out["beta3"] = pop2, beta2, pop3, beta3
tc = np.linspace(0, 10, num=21)                                           # half-day grid: day d = the sample AT d
out["tc"] = tc
out["sir_coarse"] = pm.sir_simulate_discrete((990, 10, 0), tc, 1.2, .4).to_numpy(dtype=float)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ode_synth.npz"), **out)
for k, v in out.items():
    print(k, np.asarray(v).shape)
