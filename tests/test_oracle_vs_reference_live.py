"""Live pin of the C oracle: a short run of tests/fuzz_oracle_vs_reference.py -- the UNMODIFIED reference imported from
/root/reference against oracle/sem_oracle.c on the same MT19937 stream, random small filters and simulations.  The
reference only exists in the build container: skipped elsewhere (the GPU box has the committed goldens instead)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference is only present in the build container")
def test_random_filters_and_simulations_match_the_live_reference():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_oracle_vs_reference.py"), "80", "2024", "pf", "ssa"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert ": 0 bad" in r.stdout
