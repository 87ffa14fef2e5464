"""Debug: build a copy of the library with extra -D flags into /tmp, load it instead of the in-tree one and time the
headline filter pass (device events).  Usage: python tools/variant_bench.py "-DFLAG1 -DFLAG2" [N] [arith]"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
flags = sys.argv[1].split() if len(sys.argv) > 1 else []
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
arith = sys.argv[3] if len(sys.argv) > 3 else "fast32"
pkg = os.path.join(ROOT, "stochastic-epidemic-modelling_b200")
dbg = "/tmp/libsem_b200_variant_%d.so" % (abs(hash(tuple(flags))) % 100000)
src = [os.path.join(pkg, "csrc", f) for f in ("sem_pf.cu", "sem_sim_abc.cu")]
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--fmad=false"] + flags +
                      ["-Xcompiler", "-fPIC", "-shared", "-ccbin", "/usr/bin/g++", "-o", dbg] + src)
import numpy as np, torch
import sem_b200, workloads
from sem_b200 import engine, _lib
_lib.LIB_PATH = dbg
_lib._lib = None
_lib.load()
Y = workloads.headline_Y()
cfg = engine.make_pf_config(0, N, 101, probs=.1, seed=1, mu=[20], n_population=[10000], arith=arith)
out = engine.alloc_pf_outputs(cfg)
Yd = torch.from_numpy(Y).cuda(); th = torch.tensor([.4, .2], dtype=torch.float64).cuda()
for _ in range(3):
    res = engine.run_pf(cfg, Yd, th, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 10
for i in range(K):
    cfg.filter_id0 = i
    res = engine.run_pf(cfg, Yd, th, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"flags={' '.join(flags) or '-'} N={N} arith={arith}: {ms:.3f} ms/pass  {N*100/ms/1e3:.1f} M particle-steps/s  "
      f"events {int(res.n_events[0])}  logZ {float(res.log_zetas[0,-1]):.4f} launches {res.launches}")
