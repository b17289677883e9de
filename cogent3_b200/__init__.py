"""cogent3_b200: a B200-native likelihood engine behind cogent3's API.

Only the substitution-model hot path lives here: batched
P(t)=expm(Qt), Felsenstein pruning over per-node compressed site patterns, and the
site-weighted log-likelihood reduction, as hand-written sm_100a CUDA kernels in
``libc3cuda.so`` (C ABI: include/c3cuda.h).  The Python in this package is the
host-side mirror of the reference interface plus a ctypes binding.
"""

__version__ = "0.1.0"


def install(*args, **kw):
    """route cogent3's own ``sm.make_likelihood_function(tree)`` to the CUDA engine
    (see :func:`cogent3_b200.dropin.install`)"""
    from .dropin import install as _install

    return _install(*args, **kw)


def uninstall():
    from .dropin import uninstall as _uninstall

    return _uninstall()
