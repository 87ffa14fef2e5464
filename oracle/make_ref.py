"""Copy the reference's three hot-path modules into oracle/_ref/ so that the UNMODIFIED Python reference can be timed on
the GPU box next to the GPU number (bench.py: cpu_baseline kind "reference", and --impl reference).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (the copies never enter the history) but travels with
the snapshot to the GPU box, where /root/reference does not exist.  Run in the build container:

    python oracle/make_ref.py          (also called by __graft_entry__.build() when /root/reference is present)

Files: gillespie_algo.py (SSA simulators), pmcmc.py (particle_filter / particle_mcmc), abc_algo.py -- byte-for-byte copies;
MANIFEST.json records their sha256.  oracle/ref_run.py imports them from there (matplotlib stubbed: pmcmc.py:5 imports it,
the hot path never uses it).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["gillespie_algo.py", "pmcmc.py", "abc_algo.py"]


def make(force=False):
    if not os.path.isdir(SRC):
        return False
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for f in FILES:
        src, dst = os.path.join(SRC, f), os.path.join(DST, f)
        if force or not os.path.exists(dst) or open(src, "rb").read() != open(dst, "rb").read():
            shutil.copyfile(src, dst)
        manifest[f] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "sha256": manifest, "note": "unmodified copies; test/measurement infrastructure, git-ignored"}, fh, indent=1)
    return True


if __name__ == "__main__":
    print("oracle/_ref:", "written" if make(force="--force" in sys.argv) else "reference not present, nothing done")
