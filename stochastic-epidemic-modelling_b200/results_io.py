"""The reference's de-facto checkpoint format: thetas.csv, likelihoods.csv and one
sampled_trajs_<compartment>.csv per state column written with np.savetxt(delimiter=","), and the warm-start
convention of the experiment scripts -- restart from thetas[-1] with sigma = cov(unique thinned thetas)
(tests/experiments/noise/noise_.1.py:18-26, 52-56)."""
import os

import numpy as np

COMPARTMENTS = {3: ["susceptible", "infected", "recovered"], 4: ["susceptible", "exposed", "infected", "recovered"]}


def _names(n_cols):
    if n_cols in COMPARTMENTS:
        return COMPARTMENTS[n_cols]
    return [f"{c}{g}" for g in range(n_cols // 3) for c in ("susceptible", "infected", "recovered")]


def save_results(directory, thetas, likelihoods, sampled_trajs):
    """np.savetxt layout of noise_.1.py:45-56 (directory is created if missing)."""
    os.makedirs(directory, exist_ok=True)
    np.savetxt(os.path.join(directory, "thetas.csv"), thetas, delimiter=",")
    np.savetxt(os.path.join(directory, "likelihoods.csv"), likelihoods, delimiter=",")
    for c, name in enumerate(_names(sampled_trajs.shape[2])):
        np.savetxt(os.path.join(directory, f"sampled_trajs_{name}.csv"), sampled_trajs[:, :, c], delimiter=",")


def load_results(directory):
    thetas = np.loadtxt(os.path.join(directory, "thetas.csv"), delimiter=",", ndmin=2)
    lik = np.loadtxt(os.path.join(directory, "likelihoods.csv"), delimiter=",")
    files = sorted(f for f in os.listdir(directory) if f.startswith("sampled_trajs_"))
    by_name = {f[len("sampled_trajs_"):-4]: np.loadtxt(os.path.join(directory, f), delimiter=",", ndmin=2) for f in files}
    names = _names(len(by_name))
    trajs = np.stack([by_name[n] for n in names], axis=2) if by_name else None
    return thetas, lik, trajs


def warm_start(directory, burn_in=100, thinning=20):
    """(theta_proposal, sigma) for the next run: last draw, covariance of the unique thinned post-burn-in draws
    (noise_.1.py:18-26)."""
    thetas = np.loadtxt(os.path.join(directory, "thetas.csv"), delimiter=",", ndmin=2)
    uniq = np.unique(thetas[burn_in:][::thinning], axis=0)
    return thetas[-1].tolist(), np.cov(uniq.T, ddof=0)
