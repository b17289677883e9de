"""The reference's OWN tests, run against the drop-in (build container only: needs /root/reference).

Every ``model.make_likelihood_function(tree, ...)`` in the reference's test files below builds the
drop-in likelihood function instead (tests/refrun_plugin.py); the known answers, accessor checks,
optimiser runs, ancestral reconstructions, bootstraps ... of the reference's suite must all still
pass (373 tests).  The engine is the CPU oracle port here; the same arithmetic runs on the GPU behind the same
Defn (tests/test_gpu_parity.py, tools/validate_dropin_gpu.py)."""

import json
import os
import re
import subprocess
import sys

import pytest
from refutil import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="needs /root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# every reference test file that builds likelihood functions: the whole test_evolve directory
# (likelihood_function, newq, ns_substitution_model, parameter_controller, substitution_model,
# bootstrap, distance, models, simulation ...), the app layer's evo tests, ML tree search
FILES = ["test_evolve", "test_app/test_evo.py", "test_phylo.py"]
NEEDS_NETWORK = "test_get_app_tree_is_url"  # fails in the plain reference too: no network here


def test_reference_tests_pass_with_the_dropin(tmp_path):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(
        [os.path.join(ROOT, "tests"), ROOT, os.path.join(ROOT, "tests", "_shims"), "/root/reference/src"]
    )
    env.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    stats_dir = str(tmp_path)
    env["C3_REFRUN_STATS"] = stats_dir
    res = subprocess.run(
        [sys.executable, "-m", "pytest", "-p", "no:cacheprovider", "-p", "refrun_plugin", "-q", "--no-header",
         "-n", "4", "-k", f"not {NEEDS_NETWORK}", *FILES],
        cwd="/root/reference/tests", env=env, capture_output=True, text=True, timeout=2400,
    )  # fmt: skip
    tail = res.stdout[-2500:]
    assert res.returncode == 0, tail
    m = re.search(r"\b(\d+) passed", res.stdout)
    assert m and int(m.group(1)) >= 370 and " failed" not in res.stdout, tail
    # every (xdist worker) process reports how many likelihood functions went through the drop-in
    built = [json.load(open(os.path.join(stats_dir, f))) for f in os.listdir(stats_dir)]
    n_dropin, n_ref = sum(b["dropin"] for b in built), sum(b["reference"] for b in built)
    assert n_dropin > 200 and n_dropin > 5 * n_ref, (n_dropin, n_ref)
