"""ctypes binding of libc3cuda (include/c3cuda.h).

There is NO CPU fallback: if the shared library is missing, cannot be loaded, or
no CUDA device is present, the product path raises.  (Loading the library itself
works without a GPU -- that is what the CPU test-suite checks -- but creating an
engine does not.)
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_uint8, c_void_p

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libc3cuda.so")

C3_OK, C3_ERR_INVALID, C3_ERR_CUDA, C3_ERR_MEMORY, C3_ERR_NUMERIC = 0, 1, 2, 3, 4


class C3Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libc3cuda error {code}: {message}")
        self.code = code


class C3Stats(ctypes.Structure):
    _fields_ = [
        ("evaluations", c_int64),
        ("kernel_launches", c_int64),
        ("last_launches", c_int64),
        ("last_expm_jobs", c_int64),
        ("last_nodes", c_int64),
        ("last_node_rows", c_int64),
        ("last_alg_bytes", c_int64),
        ("last_flops", c_int64),
        ("last_h2d_bytes", c_int64),
        ("device_bytes", c_int64),
        ("partial_bytes", c_int64),
        ("graph_replays", c_int64),
        ("n_levels", c_int32),
        ("padded_states", c_int32),
        ("wave_launches", c_int64),
        ("wave_entries", c_int64),
        ("wave_noops", c_int64),
        ("lean_evaluations", c_int64),
        ("chain_launches", c_int64),
        ("walk_launches", c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_pd = POINTER(c_double)
_pi32 = POINTER(c_int32)
_pi64 = POINTER(c_int64)

# name -> (restype, argtypes); every symbol include/c3cuda.h declares
SIGNATURES = {
    "c3_abi_version": (c_int, []),
    "c3_last_error": (c_char_p, []),
    "c3_device_count": (c_int, [POINTER(c_int)]),
    "c3_create": (c_int, [c_int, c_int, c_int, c_int, _pi32, _pi32, POINTER(c_void_p)]),
    "c3_destroy": (c_int, [c_void_p]),
    "c3_set_stream": (c_int, [c_void_p, c_void_p]),
    "c3_set_tip": (c_int, [c_void_p, c_int, _pd, c_int64]),
    "c3_set_node_patterns": (c_int, [c_void_p, c_int, _pi64, c_int64]),
    "c3_set_root_counts": (c_int, [c_void_p, _pd, c_int64]),
    "c3_finalize": (c_int, [c_void_p]),
    "c3_compress_alignment": (c_int, [c_void_p, POINTER(c_uint8), c_int64, _pd, c_int, c_int]),
    "c3_get_node_rows": (c_int, [c_void_p, _pi64]),
    "c3_get_node_indexes": (c_int, [c_void_p, c_int, _pi64]),
    "c3_get_tip_table": (c_int, [c_void_p, c_int, _pd]),
    "c3_get_root_counts": (c_int, [c_void_p, _pd]),
    "c3_get_column_index": (c_int, [c_void_p, c_int, _pi64]),
    "c3_set_column_index": (c_int, [c_void_p, _pi64, c_int64]),
    "c3_get_site_lh": (c_int, [c_void_p, c_int, _pd, c_int64]),
    "c3_set_eigen": (c_int, [c_void_p, c_int, _pd, _pd, _pd, c_int]),
    "c3_set_Q": (c_int, [c_void_p, c_int, _pd]),
    "c3_set_tn93": (c_int, [c_void_p, c_int, _pd, c_double, c_double]),
    "c3_set_edge_qslots": (c_int, [c_void_p, _pi32]),
    "c3_set_psub": (c_int, [c_void_p, c_int, c_int, _pd]),
    "c3_set_distances": (c_int, [c_void_p, _pd]),
    "c3_set_root_mprobs": (c_int, [c_void_p, _pd]),
    "c3_set_bin_probs": (c_int, [c_void_p, _pd]),
    "c3_set_fixed_motifs": (c_int, [c_void_p, _pi32]),
    "c3_invalidate": (c_int, [c_void_p]),
    "c3_set_rescaling": (c_int, [c_void_p, c_int, c_int]),
    "c3_get_rescaling": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int)]),
    "c3_get_root_exponents": (c_int, [c_void_p, _pi32, c_int64]),
    "c3_get_scaled_lh": (c_int, [c_void_p, c_int, _pd, c_int64]),
    "c3_evaluate": (c_int, [c_void_p, _pd]),
    "c3_evaluate_device": (c_int, [c_void_p, c_void_p]),
    "c3_update": (c_int, [c_void_p, _pd, _pd]),
    "c3_evaluate_many": (c_int, [POINTER(c_void_p), c_int, _pd]),
    "c3_update_many": (c_int, [POINTER(c_void_p), c_int, _pd, c_int64, _pd]),
    "c3_comm_mailbox": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_uint8)]),
    "c3_comm_attach": (c_int, [c_void_p, c_int, c_int, POINTER(c_uint8), POINTER(c_void_p)]),
    "c3_comm_detach": (c_int, [c_void_p]),
    "c3_hmm_reduce": (c_int, [c_void_p, c_int, _pi32, _pd, _pd, _pd, _pd]),
    "c3_get_lh": (c_int, [c_void_p, c_int, _pd, c_int64]),
    "c3_get_psub": (c_int, [c_void_p, c_int, c_int, _pd]),
    "c3_get_partials": (c_int, [c_void_p, c_int, c_int, _pd, c_int64]),
    "c3_get_stats": (c_int, [c_void_p, POINTER(C3Stats)]),
    "c3_set_profiling": (c_int, [c_void_p, c_int]),
    "c3_get_launch_profile": (c_int, [c_void_p, c_int, _pi32, _pd, _pi64, _pi64, POINTER(c_int)]),
    "c3_measure_dmma_peak": (c_int, [c_int, _pd]),
}

_lib = None


def load(path: str | None = None):
    """load libc3cuda.so and bind every entry point; raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("C3CUDA_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise C3Error(
            -1,
            f"{path} not found: build it with `python -m cogent3_b200.build` "
            "(there is no CPU fallback for the likelihood engine)",
        )
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.c3_abi_version() != 1:
        raise C3Error(-1, f"ABI version mismatch: {lib.c3_abi_version()}")
    _lib = lib
    return lib


def check(rc: int):
    if rc != C3_OK:
        msg = load().c3_last_error()
        msg = msg.decode("utf8", "replace") if msg else ""
        if rc == C3_ERR_NUMERIC:
            raise ArithmeticError(msg)
        if rc == C3_ERR_MEMORY:
            raise MemoryError(f"libc3cuda: {msg}")
        raise C3Error(rc, msg)


def device_count() -> int:
    n = c_int(0)
    rc = load().c3_device_count(byref(n))
    return n.value if rc == C3_OK else 0
