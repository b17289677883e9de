"""The reference's OWN tests, run against the drop-in (build container only: needs /root/reference).

Every ``model.make_likelihood_function(tree, ...)`` in the reference's test files below builds the
drop-in likelihood function instead (tests/refrun_plugin.py); the known answers, accessor checks,
optimiser runs, ancestral reconstructions, bootstraps ... of the reference's suite must all still
pass.  The engine is the CPU oracle port here; the same arithmetic runs on the GPU behind the same
Defn (tests/test_gpu_parity.py, tools/validate_dropin_gpu.py)."""

import os
import re
import subprocess
import sys

import pytest
from refutil import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="needs /root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FILES = [
    ["test_evolve/test_likelihood_function.py"],
    ["test_evolve/test_newq.py", "test_evolve/test_ns_substitution_model.py",
     "test_evolve/test_parameter_controller.py", "test_evolve/test_substitution_model.py",
     "test_evolve/test_bootstrap.py", "test_evolve/test_scale_rules.py", "test_evolve/test_motifchange.py"],
]  # fmt: skip


@pytest.mark.parametrize("files", FILES, ids=["likelihood_function", "models_controller_bootstrap"])
def test_reference_tests_pass_with_the_dropin(files):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(
        [os.path.join(ROOT, "tests"), ROOT, os.path.join(ROOT, "tests", "_shims"), "/root/reference/src"]
    )
    env.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    res = subprocess.run(
        [sys.executable, "-m", "pytest", "-p", "no:cacheprovider", "-p", "refrun_plugin", "-q", "--no-header", *files],
        cwd="/root/reference/tests", env=env, capture_output=True, text=True, timeout=1500,
    )  # fmt: skip
    tail = res.stdout[-1500:]
    assert res.returncode == 0, tail
    m = re.search(r"make_likelihood_function calls: \{'dropin': (\d+), 'reference': (\d+)\}", res.stdout)
    assert m and int(m.group(1)) > 50 and int(m.group(1)) > 10 * int(m.group(2)), tail
    assert re.search(r"\b(\d+) passed", res.stdout) and " failed" not in res.stdout, tail
