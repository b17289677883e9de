"""pytest plugin: run the REFERENCE's own test files with the drop-in class swapped in.

``_SubstitutionModel.make_likelihood_function`` (evolve/substitution_model.py:313-391) is replaced
by ``cogent3_b200.dropin.make_likelihood_function``; the engine behind it is the CPU oracle port
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
_REF = "/root/reference/src" if os.path.isdir("/root/reference/src") else os.path.join(_ROOT, "baseline", "_ref")
sys.path[:0] = [_ROOT, os.path.join(_ROOT, "tests", "_shims"), _REF]
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
# import order decides the reference's global numpy error state: maths/util.py:14 sets divide="raise",
# evolve/likelihood_tree.py:10 sets all="ignore", the last one imported wins.  Keep the order of a plain
# run of the reference's tests (likelihood_tree last).
import cogent3.maths.util  # noqa: E402, F401
from cogent3.evolve import substitution_model  # noqa: E402

from cogent3_b200 import dropin  # noqa: E402
from oracle.engine import OracleEngine  # noqa: E402

FACTORY = None if os.environ.get("C3_REFRUN_CUDA") else functools.partial(OracleEngine, mode="port")
_orig = substitution_model._SubstitutionModel.make_likelihood_function
STATS = {"dropin": 0, "reference": 0}


def patched(self, tree, motif_probs_from_align=None, optimise_motif_probs=None, aligned=True, expm=None,
            digits=None, space=None, **kw):
    if not aligned or kw.get("sites_independent", True) is False:
        STATS["reference"] += 1
        return _orig(self, tree, motif_probs_from_align=motif_probs_from_align,
                     optimise_motif_probs=optimise_motif_probs, aligned=aligned, expm=expm, digits=digits,
                     space=space, **kw)
    STATS["dropin"] += 1
    return dropin.make_likelihood_function(self, tree, motif_probs_from_align=motif_probs_from_align,
                                           optimise_motif_probs=optimise_motif_probs, expm=expm, digits=digits,
                                           space=space, engine_factory=FACTORY, **kw)


substitution_model._SubstitutionModel.make_likelihood_function = patched


def pytest_sessionfinish(session, exitstatus):
    print("\nmake_likelihood_function calls:", STATS)
    out = os.environ.get("C3_REFRUN_STATS")  # a directory: one file per (xdist worker) process
    if out:
        import json

        with open(os.path.join(out, f"{os.getpid()}.json"), "w") as f:
            json.dump(STATS, f)
