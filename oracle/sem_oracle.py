"""CPU oracle (pure-Python / numpy restatement) of the reference's particle-filter / SSA / ABC path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg, never by the product package.  Each function cites the reference file:line it restates
(paths relative to GeorgeEfstathiadis/Stochastic-Epidemic-Modelling).

Parity status: PINNED against the goldens in tests/golden/ that were produced by running the
unmodified reference here (tests/golden/make_golden.py, versions in tests/golden/manifest.json).
The reference itself ships no golden vectors and pins no numpy/scipy versions.

Third-party arithmetic restated here (not under /root/reference):
  * numpy legacy RandomState (numpy 2.3.5): exponential(scale) = -log(1.0-u)*scale with u one
    random_sample double; choice(R, p=p) = searchsorted(cumsum(p)/cumsum(p)[-1], u, side='right').
  * scipy.stats.binom.pmf / norm.pdf (scipy 1.18.1; binom -> Boost.Math binomial pdf): used directly
    here (scipy is the same dependency the reference calls); the C oracle and the CUDA kernels
    restate them (Loader's saddle-point form) and are pinned to scipy's values.

This file is the replay-mode oracle: it consumes a flat stream of uniforms in exactly the order the
reference consumes numpy's global MT19937 stream (jobs=1).  As a by-product it reports how many
uniforms each (step, particle) consumed, which is what the per-particle replay buffers of the C
oracle and of the CUDA kernels are cut from.
"""
import math

import numpy as np

MODEL_SIR, MODEL_SEIR, MODEL_SUBGROUPS, MODEL_SUBGROUPS2 = 0, 1, 2, 3
MODEL_IDS = {"SIR": 0, "SEIR": 1, "SIR_SUBGROUPS": 2, "SIR_SUBGROUPS2": 3}


class MTStream:
    """Sequential stream of numpy legacy random_sample doubles from a saved MT19937 state."""

    def __init__(self, key=None, pos=None, seed=None, chunk=1 << 14):
        self.rs = np.random.RandomState()
        if seed is not None:
            self.rs.seed(seed)
        else:
            self.rs.set_state(("MT19937", np.asarray(key, dtype=np.uint32), int(pos), 0, 0.0))
        self.chunk = chunk
        self.buf = np.empty(0)
        self.i = 0
        self.consumed = 0
        self.log = []          # every double handed out, in order

    def _fill(self):
        self.buf = self.rs.random_sample(self.chunk)
        self.i = 0

    def next(self):
        if self.i >= self.buf.size:
            self._fill()
        v = float(self.buf[self.i])
        self.i += 1
        self.consumed += 1
        self.log.append(v)
        return v

    def take(self, n):
        out = np.empty(n)
        for k in range(n):
            out[k] = self.next()
        return out


class ArrayStream:
    """Stream over an explicit array (per-particle replay buffers)."""

    def __init__(self, arr):
        self.arr = np.asarray(arr, dtype=float)
        self.i = 0

    def next(self):
        v = float(self.arr[self.i])
        self.i += 1
        return v


def _choose(rates, a0, u):
    """np.random.choice(R, p=rates/a0) for one draw: legacy choice builds cdf = cumsum(p), divides by
    its last element and returns searchsorted(cdf, u, side='right')  (gillespie_algo.py:63)."""
    p = np.asarray(rates, dtype=float) / a0
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return int(cdf.searchsorted(u, side="right"))


def sir_rates(S, I, R, beta, gamma, N):
    # gillespie_algo.py:37-40 -- left-to-right fp64: ((beta*S)*I)/N , gamma*I
    return [beta * S * I / N, gamma * I]


def ssa_sir(population, theta, max_time, stream, record=False):
    """gillespie_algo.py:10-75.  Returns (final_state, n_events[, times, states])."""
    beta, gamma = float(theta[0]), float(theta[1])
    S, I, R = [float(x) for x in population]
    N = S + I + R                                   # :35
    t = 0.0
    times, states = [0.0], [[S, I, R]]
    n_draw_pairs = 0
    while I > 0:                                    # :48
        rates = sir_rates(S, I, R, beta, gamma, N)  # :50
        a0 = 0 + rates[0] + rates[1]                # builtin sum(), :62
        tau = -math.log(1.0 - stream.next()) * (1 / a0)   # :62
        j = _choose(rates, a0, stream.next())       # :63  (both draws precede the break test)
        n_draw_pairs += 1
        if t + tau > max_time:                      # :65
            break
        t = t + tau                                 # :68
        if j == 0:
            S -= 1; I += 1                          # :44
        else:
            I -= 1; R += 1                          # :45
        if record:
            times.append(t); states.append([S, I, R])
    if record:
        return [S, I, R], n_draw_pairs, np.array(times), np.array(states)
    return [S, I, R], n_draw_pairs


def ssa_seir(population, theta, max_time, stream, record=False):
    """gillespie_algo.py:78-146; theta order (beta, alpha, gamma) (:92)."""
    beta, alpha, gamma = [float(x) for x in theta]
    S, E, I, R = [float(x) for x in population]
    N = S + E + I + R                               # :104
    t = 0.0
    times, states = [0.0], [[S, E, I, R]]
    n = 0
    while E > 0 or I > 0:                           # :119
        rates = [beta * S * I / N, alpha * E, gamma * I]   # :106-110
        a0 = 0 + rates[0] + rates[1] + rates[2]
        tau = -math.log(1.0 - stream.next()) * (1 / a0)    # :133
        j = _choose(rates, a0, stream.next())              # :134
        n += 1
        if t + tau > max_time:                      # :136
            break
        t = t + tau
        if j == 0:
            S -= 1; E += 1                          # :114
        elif j == 1:
            E -= 1; I += 1                          # :115
        else:
            I -= 1; R += 1                          # :116
        if record:
            times.append(t); states.append([S, E, I, R])
    if record:
        return [S, E, I, R], n, np.array(times), np.array(states)
    return [S, E, I, R], n


def ssa_subgroups(population, betas, gamma, max_time, stream, record=False):
    """gillespie_algo.py:148-233.  population (G,3); betas (G,G) with beta[a,b] = infector group a ->
    susceptible group b (:182-183).  Reaction order (dict insertion, :180-185): for a in groups:
    [infect b=0..G-1 by a], recover a."""
    pop = np.array(population, dtype=float)
    G = pop.shape[0]
    betas = np.asarray(betas, dtype=float)
    gamma = float(gamma)
    Ng = [0 + pop[g, 0] + pop[g, 1] + pop[g, 2] for g in range(G)]    # :176 builtin sum per group
    Ntot = 0
    for v in Ng:                                                      # sum(N) (:182)
        Ntot = Ntot + v
    x = pop.copy()
    t = 0.0
    times, states = [0.0], [x.ravel().copy()]
    n = 0
    infected = sum(x[g, 1] for g in range(G))
    while infected > 0:                             # :193
        rates = []
        for a in range(G):
            for b in range(G):
                rates.append(betas[a, b] * x[b, 0] * x[a, 1] / Ntot)   # :182
            rates.append(gamma * x[a, 1])                              # :184
        a0 = 0
        for r in rates:                             # builtin sum over list (:208)
            a0 = a0 + r
        tau = -math.log(1.0 - stream.next()) * (1 / a0)
        j = _choose(rates, a0, stream.next())       # :209-212
        n += 1
        if t + tau > max_time:                      # :215
            break
        t = t + tau
        a, k = divmod(j, G + 1)
        if k < G:
            x[k, 0] -= 1; x[k, 1] += 1              # :183 susceptible of group b=k infected
        else:
            x[a, 1] -= 1; x[a, 2] += 1              # :185
        if record:
            times.append(t); states.append(x.ravel().copy())
        infected = sum(x[g, 1] for g in range(G))   # :222
    if record:
        return x.ravel().tolist(), n, np.array(times), np.array(states)
    return x.ravel().tolist(), n


def observation_weights(Yrow, Xobs, observations, probs):
    """pmcmc.py:178-181: per-particle weight = MIN over observed columns of the pmf/pdf."""
    from scipy.stats import binom, norm
    if not observations:
        cols = [binom.pmf(Yrow[i], Xobs[:, i], probs) for i in range(len(Yrow))]
    else:
        cols = [norm.pdf(Yrow[i], Xobs[:, i], probs * Xobs[:, i] + .0001) for i in range(len(Yrow))]
    return np.min(np.array(cols), axis=0)


def observed_state(model, X, G):
    """pmcmc.py:172-175,228-231: SUBGROUPS2 observes the group sum of each compartment."""
    if model == MODEL_SUBGROUPS2:
        out = 0
        for g in range(G):
            out = out + X[:, 3 * g:3 * g + 3]
        return out
    return X


def particle_filter(Y, model, theta, observations, probs, X0, stream, dt=1):
    """pmcmc.py:123-233 from the start of the time loop (:177); X0 = hidden_process[0] is an input
    because the Poisson initialisation (:156-169) consumes a variable number of MT words.

    Returns dict(zetas, hidden_process, ancestry_matrix, weights, resample_u[T-1,N],
                 ssa_counts[T-1,N] (uniform PAIRS consumed per particle), ssa_u (flat list), collapsed).
    """
    Y = np.asarray(Y, dtype=float)
    T = Y.shape[0]
    X0 = np.asarray(X0, dtype=float)
    N, C = X0.shape
    G = C // 3 if model >= MODEL_SUBGROUPS else 1
    if model >= MODEL_SUBGROUPS:
        betas = np.asarray(theta[:G * G], dtype=float).reshape(G, G)
        gamma = float(theta[G * G])
    zetas = np.zeros(T); zetas[0] = 1                         # :146,154
    hidden = np.zeros((T, N, C)); hidden[0] = X0
    anc = np.zeros((T, N))
    weights = np.zeros((T, N))
    res_u = np.zeros((T - 1, N))
    counts = np.zeros((T - 1, N), dtype=np.int64)
    ssa_u = []
    for p in range(1, T):                                     # :177
        Xobs = observed_state(model, hidden[p - 1], G)
        w = observation_weights(Y[p - 1], Xobs, observations, probs)    # :178-181 (uses Y[p-1])
        weights[p] = w
        zetas[p] = zetas[p - 1] * np.mean(w)                  # :183
        tot = 0
        for v in w:                                           # builtin sum (:185), sequential fp64
            tot = tot + v
        with np.errstate(invalid="ignore", divide="ignore"):
            nw = w / tot
        # np.random.choice validation (legacy mtrand choice): NaN / negative / sum != 1 -> ValueError
        if np.isnan(nw.sum()) or (nw < 0).any() or abs(math.fsum(nw) - 1.) > math.sqrt(np.finfo(float).eps):
            return dict(collapsed=True, step=p)               # :191-192
        cdf = nw.cumsum()
        cdf /= cdf[-1]
        u = stream.take(N)                                    # :188-190  N doubles
        res_u[p - 1] = u
        a = cdf.searchsorted(u, side="right")
        anc[p] = a                                            # :193
        for j in range(N):                                    # :195-220 (jobs=1 => in order)
            st0 = hidden[p - 1, a[j]]
            start = len(stream.log) if hasattr(stream, "log") else None
            if model == MODEL_SIR:
                st, n = ssa_sir(st0, theta, dt, stream)
            elif model == MODEL_SEIR:
                st, n = ssa_seir(st0, theta, dt, stream)
            else:
                st, n = ssa_subgroups(st0.reshape(G, 3), betas, gamma, dt, stream)
            counts[p - 1, j] = n
            if start is not None:
                ssa_u.extend(stream.log[start:])
            hidden[p, j] = st                                 # :222-225
    return dict(collapsed=False, zetas=zetas, hidden_process=hidden, ancestry_matrix=anc, weights=weights,
                resample_u=res_u, ssa_counts=counts, ssa_u=np.array(ssa_u))


def particle_path_sampler(hidden_process, ancestry_matrix, chosen_path):
    """pmcmc.py:236-248 with the uniform final pick given as input.  Bug-compatible: indexes
    ancestry_matrix[p] (parents at p-1) to choose the particle at p."""
    T = hidden_process.shape[0]
    traj = np.zeros((T, hidden_process.shape[2]))
    traj[-1] = hidden_process[-1, chosen_path]
    for p in range(T - 2, -1, -1):
        chosen_path = int(ancestry_matrix[p, chosen_path])
        traj[p] = hidden_process[p, chosen_path]
    return traj


def abc_trial(observed_data, beta, gamma, n_start, stream):
    """One trial of abc_algo.py:33-99 after the prior / Poisson draws.

    Row k of the discretised trajectory = state at integer time k (last event with ceil(t)==k,
    forward filled), k = 0..T-1 (abc_algo.py:58-93).  Returns (traj[T,4], distance)."""
    obs = np.asarray(observed_data, dtype=float)
    T = obs.shape[0]
    max_time = T                                              # :23
    _, _, times, states = ssa_sir(n_start, [beta, gamma], max_time, stream, record=True)   # :41-46
    day = np.ceil(times)                                      # :60
    traj = np.zeros((T, 4))
    cur = states[0]
    e = 0
    ne = len(times)
    for k in range(T):
        # advance over every event whose ceil(time) <= k : the last one wins (:68-82); no event => ffill (:85-91)
        while e < ne and day[e] <= k:
            cur = states[e]
            e += 1
        traj[k] = [k, cur[0], cur[1], cur[2]]
    d = (np.mean(np.abs(traj[:, 2] - obs[:, 1])) + np.mean(np.abs(traj[:, 3] - obs[:, 2]))) / 2    # :10-13
    return traj, float(d)


# ------------------------------------------------------------------ ODE data synthesiser (pmcmc.py:16-113)
def ode_rhs(model, G, th, y):
    """differential_sir / differential_seir / differential_sir_subroups (pmcmc.py:16-52), N = sum(y) per evaluation."""
    N = 0.0
    for v in y:
        N += v
    if model == 0:
        return np.array([-th[0] * y[0] * y[1] / N, ((th[0] * y[0] / N) - th[1]) * y[1], th[1] * y[1]])
    if model == 1:
        return np.array([-th[0] * y[0] * y[2] / N, th[0] * y[0] * y[2] / N - th[1] * y[1], th[1] * y[1] - th[2] * y[2], th[2] * y[2]])
    d = np.zeros(3 * G)
    gamma = th[G * G]
    for i in range(G):
        force = 0.0
        for j in range(G):
            force += th[i * G + j] * y[3 * j + 1]
        inf = y[3 * i] * force / N
        d[3 * i] = -inf; d[3 * i + 1] = inf - gamma * y[3 * i + 1]; d[3 * i + 2] = gamma * y[3 * i + 1]
    return d


def ode_daily(model, G, y0, th, t, substeps=8):
    """Restatement of sem_ode_daily's scheme: classical RK4, `substeps` per grid interval, then the reference's daily
    sub-sampling (last grid point with ceil(t) == day, pmcmc.py:68-74).  Returns (days+1, C)."""
    t = np.asarray(t, dtype=np.float64)
    y = np.array(y0, dtype=np.float64).reshape(-1)
    th = np.asarray(th, dtype=np.float64).reshape(-1)
    days = np.ceil(t).astype(int)
    rows = np.zeros((days[-1] + 1, y.size))
    last = {d: np.nonzero(days == d)[0][-1] for d in range(days[-1] + 1)}
    keep = {k: d for d, k in last.items()}
    for k in range(t.size):
        if k > 0:
            h = (t[k] - t[k - 1]) / float(substeps)
            for _ in range(substeps):
                k1 = ode_rhs(model, G, th, y)
                k2 = ode_rhs(model, G, th, y + 0.5 * h * k1)
                k3 = ode_rhs(model, G, th, y + 0.5 * h * k2)
                k4 = ode_rhs(model, G, th, y + h * k3)
                y = y + h / 6.0 * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
        if k in keep:
            rows[keep[k]] = y
    return rows
