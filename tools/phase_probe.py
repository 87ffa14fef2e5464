"""Debug: per-step phase timestamps of CTA 0, thread 0 in the cooperative kernel (run with SEM_NO_SPLIT=1 for a clean
reading of the SSA phase: in the balanced layout warp 0 is not on the critical path).  Builds a -DSEM_PHASES copy of the library into
/tmp and loads it instead of the in-tree one."""
import ctypes as C, os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = os.path.join(ROOT, "stochastic-epidemic-modelling_b200")
dbg = os.environ.get("SEM_PHASES_LIB") or os.path.join(ROOT, "tools", "micro", "libsem_b200_phases.so")      # git-ignored, travels with the snapshot
src = [os.path.join(pkg, "csrc", f) for f in ("sem_pf.cu", "sem_pf_xchg.cu", "sem_sim_abc.cu")]
if "--build" in sys.argv or not os.path.exists(dbg):                       # (build it in the container: python tools/phase_probe.py --build)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--fmad=false", "-DSEM_PHASES",
                           "-DSEM_ONLY_SIR", "-Xcompiler", "-fPIC", "-shared", "-ccbin", "/usr/bin/g++", "-o", dbg] + src)
    if "--build" in sys.argv:
        sys.exit(0)
import numpy as np, torch
import sem_b200, workloads
from sem_b200 import engine, _lib
_lib.LIB_PATH = dbg
_lib._lib = None
L = _lib.load()
Y = workloads.headline_Y()
N = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 100000
cfg = engine.make_pf_config(0, N, 101, probs=.1, seed=1, mu=[20], n_population=[10000], grid_barrier=True)   # pf_persistent
out = engine.alloc_pf_outputs(cfg)
for _ in range(3):
    engine.run_pf(cfg, Y, np.array([.4, .2]), out=out)
torch.cuda.synchronize()
buf = np.zeros(24 * 256, dtype=np.uint64)
assert L.sem_debug_phases(buf.ctypes.data_as(C.c_void_p)) == 0
ph = buf.reshape(256, 24)[1:100].astype(np.int64)
# slots: 0 before grid.sync, 1 after, 2 after combine, 6 after ancestor search, 7 after gather+setup, 3 after SSA+store (own warp),
# 4 after the CTA barrier, 8 after the thread's log-weight, 9 after the CTA max, 10 after exp + CTA scan, 5 after the stores.  SM cycle counter of CTA 0's SM (1.965 GHz).
order = [0, 1, 2, 6, 7, 12, 13, 14, 15, 11, 3, 4, 8, 9, 10, 5]
names = ["grid.sync", "combine", "ancestor search", "gather+setup", "batch setup (K draw)", "sort: range", "sort: bins", "sort: scan + write", "sort: read", "SSA+store (warp 0)", "wait CTA", "weights (warp 0)", "CTA max", "exp + CTA scan",
         "store L, partial"]
t = ph[:, order] / 1965.0
d = np.diff(t, axis=1)
print("per-step mean us (CTA 0, thread 0):")
for k, nm in enumerate(names):
    print(f"  {nm:24s} mean {d[:, k].mean():8.2f}  min {d[:, k].min():8.2f}  max {d[:, k].max():8.2f}")
step = (t[1:, 0] - t[:-1, 0])
print("  step-to-step            mean %.2f  total %.2f ms" % (step.mean(), step.sum() / 1e3))
for p in (1, 10, 28, 60, 95):
    print("  step", p + 1, np.round(d[p], 2))

# per-warp view of the SSA phase (CTA 0): when each warp left the loop (us after the phase started) and its candidate count
we = np.zeros(256 * 32, dtype=np.uint64); ww = np.zeros(256 * 32, dtype=np.uint32)
if hasattr(L, "sem_debug_warps") and L.sem_debug_warps(we.ctypes.data_as(C.c_void_p), ww.ctypes.data_as(C.c_void_p)) == 0:
    we = we.reshape(256, 32).astype(np.int64); ww = ww.reshape(256, 32)
    buf2 = buf.reshape(256, 24).astype(np.int64)
    for p in (11, 29, 45, 61):
        t0 = buf2[p, 11]
        nw = int((we[p] > 0).sum())
        rel = (we[p, :nw] - t0) / 1965.0
        print(f"  step {p}: SSA phase of warp w (us after the exchange) / its largest K;  CTA barrier released at {(buf2[p, 4] - t0) / 1965.0:.2f}")
        for s4 in range(4):
            print("     sched", s4, " ".join(f"{rel[w]:7.2f}/{ww[p, w]:<5d}" for w in range(s4, nw, 4)))


# ---- the sharded filter's kernel with one rank (push-form resampling through the record buffers)
from sem_b200 import sharded
pf = sharded.PeerFilter(0, 1, 0, Y, N, theta=np.array([.4, .2]), probs=.1, seed=1, mu=[20], n_population=[10000], want_handle=False)
for i in range(3):
    pf.run(np.array([.4, .2]), filter_id=i)
torch.cuda.synchronize()
assert pf.status == 0
L.sem_debug_phases_x.argtypes = [C.c_void_p]
assert L.sem_debug_phases_x(buf.ctypes.data_as(C.c_void_p)) == 0
ph = buf.reshape(256, 24)[1:100].astype(np.int64)
order = [0, 16, 17, 2, 18, 6, 7, 12, 13, 14, 15, 11, 3, 4, 8, 9, 10, 19, 5]
names = ["wait for the partials", "CTA max", "scan + s_pfx + barrier", "offspring (J, records)", "take record", "setup", "batch setup (K draw)", "sort: range",
         "sort: bins", "sort: scan + write", "sort: read", "SSA+store (warp 0)", "wait CTA", "weights (warp 0)", "CTA max", "exp + CTA scan", "(keep)", "fence + publish"]
t = ph[:, order] / 1965.0
d = np.diff(t, axis=1)
print("pf_persistent_x, one rank; per-step mean us (CTA 0, thread 0):")
for k, nm in enumerate(names):
    print(f"  {nm:24s} mean {d[:, k].mean():8.2f}  min {d[:, k].min():8.2f}  max {d[:, k].max():8.2f}")
step = (t[1:, 0] - t[:-1, 0])
print("  step-to-step            mean %.2f  total %.2f ms" % (step.mean(), step.sum() / 1e3))
pf.close()

# ---- per-CTA view: when each CTA left its SSA phase / published its partial (globaltimer, ns), steps 1..100 of the last launch
if hasattr(L, "sem_debug_cta_times_x"):
    L.sem_debug_cta_times_x.argtypes = [C.c_void_p]
    ct = np.zeros(2 * 128 * 160, dtype=np.uint64)
    assert L.sem_debug_cta_times_x(ct.ctypes.data_as(C.c_void_p)) == 0
    ct = ct.reshape(2, 128, 160).astype(np.int64)
    nbk = int((ct[0, 5] > 0).sum())
    e = ct[0, 1:100, :nbk] / 1e3; pub = ct[1, 1:99, :nbk] / 1e3
    dur = e[1:] - pub[:-1].max(axis=1, keepdims=True) if False else None
    print(f"per-CTA end of the SSA phase over {nbk} CTAs (us, per step): max - median mean {np.mean(e.max(1) - np.median(e, 1)):.2f}, "
          f"max - min mean {np.mean(e.max(1) - e.min(1)):.2f}, std mean {np.mean(e.std(1)):.2f}")
    last = np.bincount(e.argmax(1), minlength=nbk)
    print("  CTAs most often last:", [(int(i), int(last[i])) for i in np.argsort(-last)[:8]])
    rel = e - np.median(e, 1, keepdims=True)
    m = rel.mean(0)
    print("  mean lateness vs the step's median, by CTA: min %.2f  max %.2f  (CTAs %s)" % (m.min(), m.max(), np.argsort(-m)[:6].tolist()))
    for p in (10, 28, 44, 60):
        r = np.sort(rel[p])
        print(f"  step {p + 1}: lateness quantiles 0/25/50/75/90/99/100 %: " + " ".join(f"{np.quantile(r, q):6.2f}" for q in (0, .25, .5, .75, .9, .99, 1)))
