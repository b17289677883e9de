"""pytest plugin: run the REFERENCE's own test files with the drop-in class swapped in.

``cogent3_b200.install()`` (the product hook) replaces ``_SubstitutionModel.make_likelihood_function``
(evolve/substitution_model.py:313-391); the engine behind it is the CPU oracle port
(build container, no GPU) or, with C3_REFRUN_CUDA=1, the CUDA engine.  Unaligned / phylo-HMM
likelihood functions (out of scope) stay with the reference.  Used by
tests/test_reference_suite_dropin.py:

    cd /root/reference/tests && PYTHONPATH=<repo>/tests:<repo>:<repo>/tests/_shims:/root/reference/src \
        python -m pytest -p no:cacheprovider -p refrun_plugin test_evolve/test_likelihood_function.py
"""
import functools
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF = os.environ.get("COGENT3_REFERENCE_ROOT") or (
    "/root/reference/src" if os.path.isdir("/root/reference/src") else os.path.join(_ROOT, "baseline", "_ref"))
sys.path[:0] = [_ROOT, os.path.join(_ROOT, "tests", "_shims"), _REF]
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
# import order decides the reference's global numpy error state: maths/util.py:14 sets divide="raise",
# evolve/likelihood_tree.py:10 sets all="ignore", the last one imported wins.  Keep the order of a plain
# run of the reference's tests (likelihood_tree last).
import cogent3.maths.util  # noqa: E402, F401
from cogent3.evolve import substitution_model  # noqa: E402, F401

import cogent3_b200  # noqa: E402
from cogent3_b200 import dropin  # noqa: E402

if os.environ.get("C3_REFRUN_CUDA"):
    FACTORY = None
else:
    from oracle.engine import OracleEngine  # noqa: E402  (test double: build container, no GPU)

    FACTORY = functools.partial(OracleEngine, mode="port")

# the PRODUCT hook: sm.make_likelihood_function(tree) itself returns the CUDA class; unaligned /
# phylo-HMM likelihood functions stay with the reference (cogent3_b200/dropin.py: install)
cogent3_b200.install(engine_factory=FACTORY)
STATS = dropin.INSTALL_STATS


def pytest_sessionfinish(session, exitstatus):
    print("\nmake_likelihood_function calls:", STATS)
    out = os.environ.get("C3_REFRUN_STATS")  # a directory: one file per (xdist worker) process
    if out:
        import json

        with open(os.path.join(out, f"{os.getpid()}.json"), "w") as f:
            json.dump({"dropin": STATS["cuda"], "reference": STATS["reference"]}, f)
