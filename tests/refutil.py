"""make the live reference importable in the build container (CPU tests only).

/root/reference does not exist on the GPU box; tests that need it are skipped there."""

import os
import sys

REF_SRC = "/root/reference/src"
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shims")


def reference_available():
    return os.path.isdir(os.path.join(REF_SRC, "cogent3"))


def import_reference():
    if not reference_available():
        return None
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    for p in (SHIMS, REF_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    import cogent3

    return cogent3
