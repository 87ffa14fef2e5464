"""Manufacture the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The reference has no golden vectors / known-answer tests of its own, so these
pins are made from its code run under fixed seeds.  Versions of the third-party arithmetic the
reference leans on (numpy legacy RandomState, scipy.stats -> Boost.Math) are recorded in
manifest.json because the reference pins none.

Every file holds inputs, the MT19937 state at the start of the replayed stream and the reference's
outputs.  The uniform stream itself is NOT stored; tests regenerate it with numpy from the state.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"  wrote {name}.npz  ({os.path.getsize(path)/1024:.1f} KiB)")


# ------------------------------------------------------------------ SSA known-answer cases
def make_ssa():
    cases = [
        # name, model, population, theta, max_time, seed
        ("ssa_sir_basic", "sir", [480, 20, 0], [2.0, 1.0], 3.0, 5),
        ("ssa_sir_unit_interval", "sir", [980, 20, 0], [2.0, 1.0], 1.0, 11),
        ("ssa_sir_extinct", "sir", [200, 3, 0], [0.5, 3.0], 10.0, 2),          # dies out mid-interval
        ("ssa_sir_i0_zero", "sir", [500, 0, 7], [2.0, 1.0], 1.0, 3),           # loop never entered
        ("ssa_sir_overshoot", "sir", [100, 1, 0], [1.0, 1.0], 1e-4, 4),        # first event overshoots
        ("ssa_sir_full_epidemic", "sir", [300, 5, 0], [3.0, 1.0], 40.0, 8),    # runs to extinction
        ("ssa_seir_basic", "seir", [480, 0, 20, 0], [4.0, 1.0, 1.0], 3.0, 6),
        ("ssa_seir_extinct", "seir", [100, 2, 1, 0], [0.3, 2.0, 3.0], 20.0, 7),
        ("ssa_seir_e_only", "seir", [100, 5, 0, 0], [2.0, 1.0, 1.0], 2.0, 9),  # I=0 but E>0 keeps going
        ("ssa_sub2_basic", "subgroups", [[200, 15, 0], [300, 20, 0]], ([[5, 2], [1, 3]], 0.5), 2.0, 10),
        ("ssa_sub2_asym", "subgroups", [[100, 0, 0], [150, 10, 0]], ([[15, 5], [1, 8]], 0.5), 1.0, 12),
        ("ssa_sub3_basic", "subgroups", [[80, 5, 0], [120, 0, 3], [60, 2, 0]],
         ([[3, 1, .5], [.2, 2, 1], [1, .7, 4]], 0.8), 1.5, 13),
        ("ssa_sub2_extinct", "subgroups", [[50, 1, 0], [50, 1, 0]], ([[.2, .1], [.1, .2]], 4.0), 30.0, 14),
    ]
    for name, model, pop, theta, tmax, seed in cases:
        out = rh.run_ssa(model, pop, theta, tmax, seed)
        if model == "subgroups":
            th = np.concatenate([np.array(theta[0], float).ravel(), [theta[1]]])
        else:
            th = np.array(theta, float)
        save(name, model=model, population=np.array(pop, dtype=np.int64), theta=th, max_time=float(tmax),
             seed=seed, **out)


# ------------------------------------------------------------------ particle-filter cases
def ode_truth(model, y0, t_end, theta):
    pm = rh.load_reference()["pmcmc"]
    t = np.linspace(0, t_end, num=200)
    if model == "SIR":
        df = pm.sir_simulate_discrete(tuple(y0), t, *theta)
    elif model == "SEIR":
        df = pm.seir_simulate_discrete(tuple(y0), t, *theta)
    else:
        df = pm.sir_subgroups_simulate_discrete(np.array(y0), t, np.array(theta[0], float), theta[1])
    cols = [c for c in df.columns if c != "time"]
    return df[cols].to_numpy(dtype=float)


def observe(truth, kind, p, rng):
    if kind == "binom":        # tests/test_particles.py:31-38 recipe
        return rng.binomial(truth.astype(int), p).astype(float)
    # tests/test_noise.py:23-32 recipe: X + N(0, ratio*X) cast to int
    return (truth + rng.normal(0, p * truth + 1e-12)).astype(int).astype(float)


def make_pf():
    sub_pop = [[200, 15, 0], [300, 20, 0]]
    sub_theta = ([[5, 2], [1, 3]], 0.5)
    cases = [
        # name, model, y0, t_end, theta, obs kind, probs, N, n_population, mu, seed
        ("pf_sir_binom_n100", "SIR", (480, 20, 0), 7, [2.0, 1.0], "binom", .1, 100, 500, 20, 101),
        ("pf_sir_binom_n8", "SIR", (480, 20, 0), 5, [2.0, 1.0], "binom", .1, 8, 500, 20, 102),
        ("pf_sir_binom_n1000", "SIR", (180, 20, 0), 5, [2.0, 1.0], "binom", .1, 1000, 200, 20, 103),
        ("pf_sir_normal_n100", "SIR", (480, 20, 0), 7, [2.0, 1.0], "normal", .1, 100, 500, 20, 104),
        ("pf_sir_binom_p05_n64", "SIR", (480, 20, 0), 6, [1.5, .8], "binom", .5, 64, 500, 20, 105),
        ("pf_seir_binom_n64", "SEIR", (480, 0, 20, 0), 6, [4.0, 1.0, 1.0], "binom", .1, 64, 500, 20, 106),
        ("pf_seir_normal_n64", "SEIR", (480, 0, 20, 0), 6, [4.0, 1.0, 1.0], "normal", .2, 64, 500, 20, 107),
        ("pf_sub_binom_n48", "SIR_SUBGROUPS", sub_pop, 5, sub_theta, "binom", .1, 48, [215, 320], [15, 20], 108),
        ("pf_sub2_binom_n48", "SIR_SUBGROUPS2", sub_pop, 5, sub_theta, "binom", .1, 48, [215, 320], [15, 20], 109),
        ("pf_sub2_normal_n48", "SIR_SUBGROUPS2", sub_pop, 5, sub_theta, "normal", .1, 48, [215, 320], [15, 20], 110),
    ]
    for name, model, y0, t_end, theta, kind, probs, N, npop, mu, seed in cases:
        rng = np.random.RandomState(seed + 1000)
        truth = ode_truth(model, y0, t_end, theta)
        Yfull = observe(truth, kind, probs, rng)
        if model == "SIR_SUBGROUPS2":      # tests/test_particles_subroups2.py:42: group-summed observation
            Y = Yfull[:, :3] + Yfull[:, 3:]
        else:
            Y = Yfull
        if model in ("SIR", "SEIR"):
            th_in = np.array(theta, float)
            th_flat = th_in
        else:
            th_in = (np.array(theta[0], float), float(theta[1]))
            th_flat = np.concatenate([th_in[0].ravel(), [th_in[1]]])
        out = rh.run_particle_filter(Y, model, th_in, kind == "normal", probs, N, npop, mu, seed)
        assert not out["collapsed"], name
        save(name, model=model, Y=Y, theta=th_flat, observations=(kind == "normal"), probs=float(probs),
             n_particles=N, n_population=np.array(npop, dtype=np.int64).reshape(-1),
             mu=np.array(mu, dtype=float).reshape(-1), seed=seed, **out)

    # collapse: observation larger than any particle's compartment -> all binomial weights 0
    Y = np.array([[9999., 9999., 9999.]] * 4)
    out = rh.run_particle_filter(Y, "SIR", np.array([2., 1.]), False, .1, 16, 500, 20, 120)
    assert out["collapsed"]
    save("pf_sir_collapse", model="SIR", Y=Y, theta=np.array([2., 1.]), observations=False, probs=.1,
         n_particles=16, n_population=np.array([500]), mu=np.array([20.]), seed=120,
         mt_key=out["mt_key"], mt_pos=out["mt_pos"], collapsed=True)

    # path sampler on the n100 case (pmcmc.py:236-248, incl. its off-by-one)
    g = np.load(os.path.join(HERE, "pf_sir_binom_n100.npz"))
    for seed in (0, 1, 2):
        ps = rh.run_path_sampler(g["hidden_process"], g["ancestry_matrix"], seed)
        save(f"path_sir_n100_s{seed}", source="pf_sir_binom_n100", seed=seed, **ps)


# ------------------------------------------------------------------ observation-weight known answers
def make_weights():
    """scipy.stats values at pmcmc.py:179,181 incl. the edge cases listed in SURVEY 8(a) a6."""
    from scipy.stats import binom, norm
    rng = np.random.RandomState(77)
    n = np.concatenate([rng.randint(0, 50, 200), rng.randint(50, 20000, 400), rng.randint(10**5, 10**6 + 1, 200),
                        [0, 0, 1, 5, 10, 10, 10**6, 10**6]]).astype(float)
    p_choices = np.array([.005, .05, .1, .5, .9, 1.0, 0.0])
    p = p_choices[rng.randint(0, len(p_choices), n.size)]
    k = np.floor(n * p + rng.normal(0, 1, n.size) * np.sqrt(n * p * (1 - p) + 1) * 2)
    k[-8:] = [0, 1, 2, 5, 0, 10, 10**5, 10**6 + 1]
    k[::17] = k[::17] + .5          # non-integer k -> pmf 0
    k[::23] = -1                    # below support
    pmf = binom.pmf(k, n, p)
    x = np.concatenate([rng.randint(0, 30, 100), rng.randint(30, 10**4, 300), [0, 0, 1]]).astype(float)
    pr = np.array([.05, .1, .2, .3, .5])[rng.randint(0, 5, x.size)]
    y = np.floor(x + rng.normal(0, 1, x.size) * (pr * x + 1) * 1.5)
    pdf = norm.pdf(y, x, pr * x + .0001)
    save("weights_known_answers", binom_k=k, binom_n=n, binom_p=p, binom_pmf=pmf,
         norm_y=y, norm_x=x, norm_probs=pr, norm_pdf=pdf)


# ------------------------------------------------------------------ ABC
def make_abc():
    truth = ode_truth("SIR", (480, 20, 0), 9, [2.0, 1.0])
    rng = np.random.RandomState(7)
    obs = truth + rng.normal(0, .1 * truth + 1e-12)      # tests/test_abc_sir.py:22-30 recipe (not cast)
    out = rh.run_abc(obs, 3, 45.0, {"beta": [0, 5], "gamma": [0, 5]}, 21)
    tr = out["trials"]
    print(f"  abc: {len(tr)} trials, {len(out['posterior']['beta'])} accepted")
    save("abc_sir_small", observed=obs, threshold=45.0, priors=np.array([[0., 5.], [0., 5.]]),
         no_of_samples=3, seed=21,
         n_start=np.array([t["n_start"] for t in tr]), theta=np.array([t["theta"] for t in tr]),
         mt_key=np.array([t["mt_key"] for t in tr]), mt_pos=np.array([t["mt_pos"] for t in tr]),
         I_sim=np.array([t["I_sim"] for t in tr]), R_sim=np.array([t["R_sim"] for t in tr]),
         distance=np.array([t["distance"] for t in tr]),
         post_beta=np.array(out["posterior"]["beta"]), post_gamma=np.array(out["posterior"]["gamma"]),
         trajectories=out["trajectories"])


def main():
    import scipy
    print("SSA cases"); make_ssa()
    print("PF cases"); make_pf()
    print("weights"); make_weights()
    print("ABC"); make_abc()
    manifest = dict(
        generated_by="tests/golden/make_golden.py",
        reference="/root/reference (GeorgeEfstathiadis/Stochastic-Epidemic-Modelling, unmodified; matplotlib stubbed)",
        numpy=np.__version__, scipy=scipy.__version__, python=sys.version.split()[0],
        note="parity pinned to these versions; the reference itself pins none (no requirements file).",
    )
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
