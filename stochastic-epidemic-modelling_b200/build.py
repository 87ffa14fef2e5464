"""Build libsem_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python stochastic-epidemic-modelling_b200/build.py [--force] [--verbose]

The library is built next to this file so that it travels with the source snapshot to the GPU box.
nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsem_b200.so")
SOURCES = ["sem_pf.cu", "sem_pf_xchg.cu", "sem_sim_abc.cu"]
HEADERS = ["sem_common.cuh", "sem_host.h", "sem_pf_dev.cuh", "sem_pf_host.h", "sem_logtab.inc",
           os.path.join("..", "..", "include", "sem_b200.h")]
OBJ_DIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--fmad=false",          # integer outcomes must not depend on FMA contraction (parity with the CPU oracle)
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


PF_KERNEL_FILES = ["sem_pf.cu", "sem_pf_xchg.cu", "sem_pf_dev.cuh", "sem_pf_host.h", "sem_common.cuh", "sem_host.h", "sem_logtab.inc"]


def source_hash(files=None):
    """sha256 over the sources the particle-filter kernels are compiled from (everything but the simulation / ABC unit and
    the public header) and the compiler flags: identifies WHAT was built, whatever the binary's own hash (kept beside
    measurements so that a profile can be recognised as belonging to the library in use)."""
    import hashlib
    h = hashlib.sha256()
    for name in sorted(files or PF_KERNEL_FILES):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode()); h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """One nvcc process per translation unit (in parallel), then one link step.  SEM_BUILD_DEFINES="-DSEM_ONLY_SIR" etc.
    passes extra defines (development builds)."""
    if not force and not stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = nvcc_path()
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []   # the image exports CC=/opt/gcc/bin/gcc
    extra = os.environ.get("SEM_BUILD_DEFINES", "").split()

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + ccbin + ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return obj, " ".join(cmd) + "\n" + res.stdout + res.stderr, res.returncode

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        done = list(ex.map(compile_one, SOURCES))
    log = "\n".join(d[1] for d in done)
    rc = max(d[2] for d in done)
    if rc == 0:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + ccbin + ["-o", LIB] + [d[0] for d in done]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += "\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr
        rc = res.returncode
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(log)
    if rc != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libsem_b200.so")
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
