"""make the reference importable for the tests.

Build container: /root/reference/src (read-only).  GPU box: /root/reference does not exist there, but
``baseline/_ref`` (a plain copy of the unmodified package and of its tests directory, git-ignored,
DESIGN.md §8) travels with the repo snapshot.  Three import-time dependencies of cogent3 that are not
in the image (stevedore / loky / scitrack, none on the arithmetic path) are the shims of tests/_shims."""

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SHIMS = os.path.join(HERE, "_shims")
_CANDIDATES = ("/root/reference/src", os.path.join(ROOT, "baseline", "_ref"))


def reference_root():
    env = os.environ.get("COGENT3_REFERENCE_ROOT")
    for p in ((env,) if env else ()) + _CANDIDATES:
        if p and os.path.isdir(os.path.join(p, "cogent3")):
            return p
    return None


REF_SRC = reference_root()


def reference_available():
    return REF_SRC is not None


def reference_tests_dir():
    """the reference's own tests directory (with its data/), or None"""
    for p in ("/root/reference/tests", os.path.join(ROOT, "baseline", "_ref", "tests")):
        if os.path.isdir(os.path.join(p, "test_evolve")):
            return p
    return None


def import_reference():
    if not reference_available():
        return None
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    os.environ.setdefault("COGENT3_REFERENCE_ROOT", REF_SRC)  # where the stevedore shim finds pyproject.toml
    for p in (SHIMS, REF_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    # import order decides the reference's global numpy error state (maths/util.py:14 vs
    # evolve/likelihood_tree.py:10): keep the order of a plain run of the reference's tests
    import cogent3
    import cogent3.maths.util  # noqa: F401
    from cogent3.evolve import substitution_model  # noqa: F401

    return cogent3
