"""ctypes binding of libsem_b200.so (the C ABI declared in include/sem_b200.h).

There is NO CPU fallback: if the CUDA library is missing or fails to load, every entry point raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SEM_LIB_PATH") or os.path.join(HERE, "libsem_b200.so")   # SEM_LIB_PATH: debug builds

SEM_MAX_GROUPS = 4
MODEL_SIR, MODEL_SEIR, MODEL_SIR_SUBGROUPS, MODEL_SIR_SUBGROUPS2 = 0, 1, 2, 3
OBS_BINOMIAL, OBS_NORMAL = 0, 1
RESAMPLE_MULTINOMIAL, RESAMPLE_SYSTEMATIC = 0, 1
ARITH_REFERENCE, ARITH_FAST = 0, 1
ERR_REPLAY = -3
ERR_PEER = -5
SEM_MAX_RANKS = 8

EXPORTS = [
    "sem_abi_version", "sem_last_error", "sem_device_info", "sem_host_workspace_release",
    "sem_pf_workspace_bytes", "sem_pf_hist_elems", "sem_pf_ancestry_elems", "sem_pf_launch_count",
    "sem_pf_run", "sem_pf_iteration", "sem_pf_run_host", "sem_path_sample", "sem_hist_to_f64",
    "sem_ssa_simulate", "sem_ode_daily", "sem_abc_run", "sem_shard_init", "sem_shard_offspring", "sem_shard_propagate",
    "sem_xchg_bytes", "sem_xchg_alloc", "sem_xchg_open", "sem_xchg_close", "sem_xchg_free", "sem_peer_enable",
    "sem_xchg_reset", "sem_xchg_iteration_result", "sem_pf_sharded_supported", "sem_pf_run_sharded", "sem_pf_iteration_sharded",
    "sem_test_philox", "sem_test_binom_logpmf", "sem_test_norm_logpdf", "sem_test_poisson", "sem_test_fast_math",
]


class PfConfig(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("obs_kind", C.c_int32), ("resampler", C.c_int32), ("arith", C.c_int32),
        ("n_particles", C.c_int32), ("n_obs", C.c_int32), ("n_groups", C.c_int32), ("n_obs_cols", C.c_int32),
        ("n_filters", C.c_int32), ("block_particles", C.c_int32), ("store_history", C.c_int32), ("reserved", C.c_int32),
        ("probs", C.c_double), ("dt", C.c_double), ("seed", C.c_uint64), ("filter_id0", C.c_uint32),
        ("path_exact", C.c_uint32), ("mu", C.c_double * SEM_MAX_GROUPS), ("n_population", C.c_double * SEM_MAX_GROUPS),
    ]


class PfBuffers(C.Structure):
    _fields_ = [
        ("Y", C.c_void_p), ("theta", C.c_void_p), ("X0", C.c_void_p),
        ("replay_resample_u", C.c_void_p), ("replay_ssa_u", C.c_void_p), ("replay_ssa_off", C.c_void_p),
        ("X_hist", C.c_void_p), ("ancestry", C.c_void_p), ("log_zetas", C.c_void_p), ("status", C.c_void_p),
        ("n_events", C.c_void_p), ("workspace", C.c_void_p), ("iteration_result", C.c_void_p),
        ("probs_per_filter", C.c_void_p),
    ]


class OdeConfig(C.Structure):
    _fields_ = [("model", C.c_int32), ("n_groups", C.c_int32), ("n_sets", C.c_int32), ("n_grid", C.c_int32),
                ("n_rows", C.c_int32), ("substeps", C.c_int32), ("shared_y0", C.c_int32), ("shared_theta", C.c_int32)]


class ShardStep(C.Structure):
    _fields_ = [
        ("step", C.c_int32), ("particle_offset", C.c_int32), ("n_global", C.c_int64), ("u0", C.c_double),
        ("total", C.c_double), ("total_local", C.c_double), ("G", C.c_double), ("G_next", C.c_double), ("s", C.c_double),
        ("slot0", C.c_int64),
    ]


class XchgDesc(C.Structure):
    _fields_ = [
        ("world", C.c_int32), ("rank", C.c_int32), ("generation", C.c_uint32), ("launch_tag", C.c_uint32),
        ("timeout_s", C.c_double), ("arena", C.c_void_p * SEM_MAX_RANKS),
    ]


class SimConfig(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("n_groups", C.c_int32), ("arith", C.c_int32), ("n_sims", C.c_int32),
        ("shared_theta", C.c_int32), ("shared_x0", C.c_int32), ("record_capacity", C.c_int64),
        ("max_time", C.c_double), ("seed", C.c_uint64), ("sim_index0", C.c_uint32), ("daily_rows", C.c_int32),
    ]


class AbcConfig(C.Structure):
    _fields_ = [
        ("n_days", C.c_int32), ("arith", C.c_int32), ("early_reject", C.c_int32), ("reserved", C.c_int32),
        ("n_trials", C.c_int64), ("trial0", C.c_uint64), ("threshold", C.c_double), ("prior", C.c_double * 4),
        ("seed", C.c_uint64),
    ]


class SemError(RuntimeError):
    pass


_lib = None


def load():
    """Load the CUDA library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SemError(
            f"{LIB_PATH} not found: build it with `python stochastic-epidemic-modelling_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.sem_last_error.restype = C.c_char_p
    L.sem_abi_version.restype = C.c_int
    for name in ("sem_pf_workspace_bytes", "sem_pf_hist_elems", "sem_pf_ancestry_elems"):
        getattr(L, name).restype = C.c_size_t
        getattr(L, name).argtypes = [C.POINTER(PfConfig)]
    L.sem_pf_launch_count.restype = C.c_int
    L.sem_pf_launch_count.argtypes = [C.POINTER(PfConfig)]
    L.sem_pf_run.restype = C.c_int
    L.sem_pf_run.argtypes = [C.POINTER(PfConfig), C.POINTER(PfBuffers), C.c_void_p]
    L.sem_pf_iteration.restype = C.c_int
    L.sem_pf_iteration.argtypes = [C.POINTER(PfConfig), C.POINTER(PfBuffers), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sem_ode_daily.restype = C.c_int
    L.sem_ode_daily.argtypes = [C.POINTER(OdeConfig)] + [C.c_void_p] * 6
    L.sem_pf_run_host.restype = C.c_int
    L.sem_pf_run_host.argtypes = [C.POINTER(PfConfig)] + [C.c_void_p] * 8
    L.sem_shard_init.restype = C.c_int
    L.sem_shard_init.argtypes = [C.POINTER(PfConfig), C.POINTER(PfBuffers), C.c_int32, C.c_void_p, C.c_void_p]
    L.sem_shard_offspring.restype = C.c_int
    L.sem_shard_offspring.argtypes = [C.POINTER(PfConfig), C.POINTER(PfBuffers), C.POINTER(ShardStep), C.c_void_p, C.c_void_p]
    L.sem_shard_propagate.restype = C.c_int
    L.sem_shard_propagate.argtypes = [C.POINTER(PfConfig), C.POINTER(PfBuffers), C.POINTER(ShardStep), C.c_void_p, C.c_void_p,
                                      C.c_void_p]
    L.sem_xchg_bytes.restype = C.c_size_t
    L.sem_xchg_bytes.argtypes = [C.POINTER(PfConfig), C.c_int32]
    L.sem_xchg_alloc.restype = C.c_int
    L.sem_xchg_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]
    L.sem_xchg_open.restype = C.c_int
    L.sem_xchg_open.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.sem_xchg_close.restype = C.c_int
    L.sem_xchg_close.argtypes = [C.c_void_p]
    L.sem_xchg_free.restype = C.c_int
    L.sem_xchg_free.argtypes = [C.c_void_p]
    L.sem_peer_enable.restype = C.c_int
    L.sem_peer_enable.argtypes = [C.c_int32, C.c_int32]
    L.sem_xchg_reset.restype = C.c_int
    L.sem_xchg_reset.argtypes = [C.POINTER(PfConfig), C.c_int32, C.c_void_p, C.c_void_p]
    L.sem_xchg_iteration_result.restype = C.c_void_p
    L.sem_xchg_iteration_result.argtypes = [C.POINTER(PfConfig), C.c_int32, C.c_void_p]
    L.sem_pf_sharded_supported.restype = C.c_int
    L.sem_pf_sharded_supported.argtypes = [C.POINTER(PfConfig), C.c_int32]
    L.sem_pf_run_sharded.restype = C.c_int
    L.sem_pf_run_sharded.argtypes = [C.POINTER(PfConfig), C.POINTER(PfBuffers), C.POINTER(XchgDesc), C.c_void_p]
    L.sem_pf_iteration_sharded.restype = C.c_int
    L.sem_pf_iteration_sharded.argtypes = [C.POINTER(PfConfig), C.POINTER(PfBuffers), C.POINTER(XchgDesc), C.c_void_p, C.c_void_p, C.c_void_p]
    L.sem_path_sample.restype = C.c_int
    L.sem_path_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
    L.sem_hist_to_f64.restype = C.c_int
    L.sem_hist_to_f64.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.sem_ssa_simulate.restype = C.c_int
    L.sem_ssa_simulate.argtypes = [C.POINTER(SimConfig)] + [C.c_void_p] * 9
    L.sem_abc_run.restype = C.c_int
    L.sem_abc_run.argtypes = [C.POINTER(AbcConfig)] + [C.c_void_p] * 12
    L.sem_device_info.restype = C.c_int
    L.sem_device_info.argtypes = [C.POINTER(C.c_int)] * 3
    L.sem_test_philox.restype = C.c_int
    L.sem_test_binom_logpmf.restype = C.c_int
    L.sem_test_norm_logpdf.restype = C.c_int
    L.sem_test_poisson.restype = C.c_int
    L.sem_test_fast_math.restype = C.c_int
    L.sem_test_fast_math.argtypes = [C.c_void_p] * 4 + [C.c_int64]
    L.sem_test_poisson.argtypes = [C.c_double, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int64]
    if L.sem_abi_version() != 2:
        raise SemError("libsem_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc, what):
    if rc < 0:
        raise SemError(f"{what} failed ({rc}): {load().sem_last_error().decode()}")
    return rc
