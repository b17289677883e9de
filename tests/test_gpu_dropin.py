"""Driver-run evidence for the PRODUCT: unmodified cogent3 (``baseline/_ref``) + the drop-in + the CUDA
engine, on a B200.

Every test of tests/test_dropin.py is re-exported here under the ``gpu`` marker.  On a GPU box that module
takes cogent3 from ``baseline/_ref`` and leaves ``engine_factory`` at the product default (libc3cuda, no
CPU fallback), so the same bodies that check the host logic against the oracle engine in the build
container check here: the reference's own known answers through ``sm.make_likelihood_function(tree)``
(``cogent3_b200.install()``), lnL / per-pattern ``lh`` / per-column likelihoods / posteriors / ancestral
reconstructions / psubs against the stock class in the same process, ``lf.optimise()`` on brca1, undo,
pickling, multi-locus sums, the vectorised ``set_alignment``.  A missing reference FAILS (it does not skip)."""

import types

import pytest
import test_dropin as cases
from refutil import reference_available, reference_tests_dir


@pytest.mark.gpu
def test_the_reference_is_on_the_gpu_box_and_the_engine_is_cuda():
    assert reference_available() and reference_tests_dir(), (
        "baseline/_ref (unmodified copy of the reference package + its tests directory) is missing"
    )
    assert cases.ON_GPU and cases.FACTORY is None
    from cogent3_b200 import _lib

    assert _lib.device_count() > 0


def _clone(fn):
    new = types.FunctionType(fn.__code__, fn.__globals__, fn.__name__, fn.__defaults__, fn.__closure__)
    new.__dict__.update({k: (list(v) if k == "pytestmark" else v) for k, v in fn.__dict__.items()})
    new.__doc__ = fn.__doc__
    new.__kwdefaults__ = fn.__kwdefaults__
    return pytest.mark.gpu(new)


for _name in sorted(dir(cases)):
    if _name.startswith("test_") and isinstance(getattr(cases, _name), types.FunctionType):
        globals()[_name] = _clone(getattr(cases, _name))
del _name
