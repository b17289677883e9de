#!/bin/bash
# The reference's own test files against the drop-in with the CUDA engine (GPU box).  Needs an
# unmodified copy of the reference package AND of its tests directory under baseline/_ref/
# (git-ignored; `cp -r /root/reference/src/cogent3 /root/reference/tests baseline/_ref/`).
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
cd "$ROOT/baseline/_ref/tests" || exit 1
export PYTHONPATH="$ROOT/tests:$ROOT:$ROOT/tests/_shims:$ROOT/baseline/_ref"
export NUMBA_CACHE_DIR=/tmp/numba_cache C3_REFRUN_CUDA=1 COGENT3_REFERENCE_ROOT="$ROOT/baseline/_ref"
python -m pytest -p no:cacheprovider -p refrun_plugin -q --no-header -k "not test_get_app_tree_is_url" \
  test_evolve test_app/test_evo.py test_phylo.py "$@"
