"""B200-native particle-filter / SSA / ABC engine behind the reference's Python entry points.

Drop-in modules (same names and signatures as GeorgeEfstathiadis/Stochastic-Epidemic-Modelling):
    gillespie_algo : sir_simulate, seir_simulate, sir_subgroups_simulate   (gillespie_algo.py)
    pmcmc          : ModelType, particle_filter, particle_path_sampler, particle_mcmc   (pmcmc.py:116-408)
    abc_algo       : abc_algo, distance_function   (abc_algo.py)
The directory name contains a hyphen, so import it through the repo-root shim:  `import sem_b200`.
`sem_b200.install_dropin()` additionally registers the three modules under their reference names so an
unmodified reference script (`from pmcmc import *`) picks them up.
"""
import sys

from . import _lib, engine  # noqa: F401
from . import gillespie_algo, pmcmc, abc_algo, helpers, results_io, sharded  # noqa: F401
from .pmcmc import ModelType, particle_filter, particle_path_sampler, particle_mcmc  # noqa: F401
from .gillespie_algo import sir_simulate, seir_simulate, sir_subgroups_simulate  # noqa: F401
from .abc_algo import distance_function  # noqa: F401

__version__ = "0.1.0"


def install_dropin():
    """Make `import gillespie_algo / pmcmc / abc_algo` resolve to this package's modules."""
    sys.modules["gillespie_algo"] = gillespie_algo
    sys.modules["pmcmc"] = pmcmc
    sys.modules["abc_algo"] = abc_algo
