"""The reference's OWN tests, run against the drop-in.

``cogent3_b200.install()`` (tests/refrun_plugin.py) makes every ``model.make_likelihood_function(tree, ...)``
in the reference's test files below build the drop-in likelihood function; the known answers, accessor
checks, optimiser runs, ancestral reconstructions, bootstraps ... of the reference's suite must all still
pass (373 tests).

* build container (no GPU): the engine behind the drop-in is the CPU oracle port -- host logic only;
* GPU box (``-m gpu``): the engine is libc3cuda.  The reference package and its tests directory come from
  ``baseline/_ref`` (an unmodified copy); its absence FAILS the test -- this is the
  driver-run evidence that real cogent3 + drop-in + CUDA engine reproduce the reference's known answers
  (``test_calclikelihood`` -250.686745262, ``test_codon_rate_het`` -6755.451985437475, ``test_loci``
  -9168.333116203343, ``test_nucleotide`` -148.6455087258624, ``test_codon`` -103.05742415448259, ...)."""

import json
import os
import re
import subprocess
import sys

import pytest
from refutil import reference_root, reference_tests_dir

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# every reference test file that builds likelihood functions: the whole test_evolve directory
# (likelihood_function, newq, ns_substitution_model, parameter_controller, substitution_model,
# bootstrap, distance, models, simulation ...), the app layer's evo tests, ML tree search
FILES = ["test_evolve", "test_app/test_evo.py", "test_phylo.py"]
NEEDS_NETWORK = "test_get_app_tree_is_url"  # fails in the plain reference too: no network here


def run_suite(tmp_path, cuda, workers):
    ref, ref_tests = reference_root(), reference_tests_dir()
    assert ref and ref_tests, "the reference package / tests are missing (baseline/_ref, DESIGN.md §8)"
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests"), ROOT, os.path.join(ROOT, "tests", "_shims"), ref])
    env["COGENT3_REFERENCE_ROOT"] = ref
    env.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    env.pop("C3_REFRUN_CUDA", None)
    if cuda:
        env["C3_REFRUN_CUDA"] = "1"
    stats_dir = str(tmp_path)
    env["C3_REFRUN_STATS"] = stats_dir
    cmd = [sys.executable, "-m", "pytest", "-p", "no:cacheprovider", "-p", "refrun_plugin", "-q", "--no-header",
           "-k", f"not {NEEDS_NETWORK}", *FILES]  # fmt: skip
    if workers > 1:
        cmd[cmd.index("-q"):cmd.index("-q")] = ["-n", str(workers)]
    res = subprocess.run(cmd, cwd=ref_tests, env=env, capture_output=True, text=True, timeout=2400)
    tail = res.stdout[-2500:] + res.stderr[-1500:]
    assert res.returncode == 0, tail
    m = re.search(r"\b(\d+) passed", res.stdout)
    assert m and int(m.group(1)) >= 370 and " failed" not in res.stdout, tail
    # every (xdist worker) process reports how many likelihood functions went through the drop-in
    built = [json.load(open(os.path.join(stats_dir, f))) for f in os.listdir(stats_dir)]
    n_dropin, n_ref = sum(b["dropin"] for b in built), sum(b["reference"] for b in built)
    assert n_dropin > 200 and n_dropin > 5 * n_ref, (n_dropin, n_ref)
    return int(m.group(1)), n_dropin, n_ref


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="build container only (needs /root/reference)")
def test_reference_tests_pass_with_the_dropin(tmp_path):
    run_suite(tmp_path, cuda=False, workers=4)


@pytest.mark.gpu
def test_reference_tests_pass_with_the_dropin_on_the_gpu(tmp_path):
    """373 of the reference's own tests through sm.make_likelihood_function -> CUDA engine"""
    passed, n_dropin, n_ref = run_suite(tmp_path, cuda=True, workers=1)
    print(f"reference suite on the CUDA engine: {passed} passed; {n_dropin} likelihood functions through "
          f"the drop-in, {n_ref} left with the reference (unaligned / phylo-HMM)")
