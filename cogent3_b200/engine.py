"""LikelihoodEngine: Python handle on one ``c3_engine`` (include/c3cuda.h).

The engine owns all device buffers (pattern tables, resident partial likelihoods,
P matrices); Python keeps only small parameter arrays.  Mirrors what the
reference's Defn cells hold for one locus (a10-a15) but with the
whole post-order sweep behind one call.
"""

from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_int, c_void_p

import numpy as np

from . import _lib
from .patterns import PatternTree


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(c_double))


def _i32ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _i64ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


class LikelihoodEngine:
    """GPU-resident Felsenstein pruning for one alignment on one device."""

    def __init__(self, patterns: PatternTree, n_bins: int = 1, device: int = 0, stream=None):
        self._lib = _lib.load()
        self._h = c_void_p()
        self.patterns = patterns
        topo = patterns.topo
        self.topo = topo
        self.n_states = int(patterns.n_states)
        self.n_bins = int(n_bins)
        self.n_nodes = topo.n_nodes
        ptr, idx = topo.child_csr()
        _lib.check(
            self._lib.c3_create(
                device, topo.n_nodes, self.n_states, self.n_bins, _i32ptr(ptr), _i32ptr(idx),
                byref(self._h),
            )
        )  # fmt: skip
        try:
            if stream is not None:
                self.set_stream(stream)
            for n, nd in enumerate(patterns.nodes):
                if nd.is_tip:
                    tab = np.ascontiguousarray(nd.table, dtype=np.float64)
                    _lib.check(self._lib.c3_set_tip(self._h, n, _dptr(tab), tab.shape[0]))
                else:
                    ix = np.ascontiguousarray(nd.indexes, dtype=np.int64)
                    _lib.check(
                        self._lib.c3_set_node_patterns(self._h, n, _i64ptr(ix), ix.shape[1])
                    )
            counts = np.ascontiguousarray(patterns.root.counts, dtype=np.float64)
            _lib.check(self._lib.c3_set_root_counts(self._h, _dptr(counts), counts.shape[0]))
            _lib.check(self._lib.c3_finalize(self._h))
            self.n_columns = int(len(patterns.root.index))
            self.set_column_index(patterns.root.index)
        except Exception:
            self.close()
            raise
        self.root_rows = int(patterns.root.n_rows)
        self._dist = np.zeros((self.n_nodes, self.n_bins), np.float64)

    @classmethod
    def from_alignment(cls, topo, codes, symbol_rows, n_bins=1, device=0, stream=None, keep_index=False):
        """engine whose pattern tables are built ON THE DEVICE from the raw alignment
        (c3_compress_alignment): ``codes`` uint8 [n_tips, L], tips in node order
        (``topo.tips``), ``symbol_rows`` [S, M] the 0/1 row of every code."""
        from .patterns import DevicePatternTree

        self = cls.__new__(cls)
        self._lib = _lib.load()
        self._h = c_void_p()
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        symbol_rows = np.ascontiguousarray(symbol_rows, dtype=np.float64)
        assert codes.ndim == 2 and codes.shape[0] == len(topo.tips)
        self.topo = topo
        self.n_states = int(symbol_rows.shape[1])
        self.n_bins = int(n_bins)
        self.n_nodes = topo.n_nodes
        self.n_columns = int(codes.shape[1])
        ptr, idx = topo.child_csr()
        _lib.check(
            self._lib.c3_create(
                device, topo.n_nodes, self.n_states, self.n_bins, _i32ptr(ptr), _i32ptr(idx),
                byref(self._h),
            )
        )  # fmt: skip
        try:
            if stream is not None:
                self.set_stream(stream)
            _lib.check(
                self._lib.c3_compress_alignment(
                    self._h, codes.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), self.n_columns,
                    _dptr(symbol_rows), symbol_rows.shape[0], int(bool(keep_index)),
                )
            )  # fmt: skip
            _lib.check(self._lib.c3_finalize(self._h))
        except Exception:
            self.close()
            raise
        self.patterns = DevicePatternTree(topo, self, self.n_states, self.n_columns)
        self.root_rows = int(self.patterns.rows()[topo.root])
        self._dist = np.zeros((self.n_nodes, self.n_bins), np.float64)
        return self

    # -- pattern tables held by the engine ----------------------------------------
    def fetch_node_rows(self):
        out = np.zeros(self.n_nodes, np.int64)
        _lib.check(self._lib.c3_get_node_rows(self._h, _i64ptr(out)))
        return out

    def fetch_node_indexes(self, node):
        rows = int(self.fetch_node_rows()[node])
        out = np.zeros((len(self.topo.children[node]), rows), np.int64)
        _lib.check(self._lib.c3_get_node_indexes(self._h, int(node), _i64ptr(out)))
        return out

    def fetch_tip_table(self, node):
        rows = int(self.fetch_node_rows()[node])
        out = np.zeros((rows, self.n_states), np.float64)
        _lib.check(self._lib.c3_get_tip_table(self._h, int(node), _dptr(out)))
        return out

    def fetch_root_counts(self):
        out = np.zeros(int(self.fetch_node_rows()[self.topo.root]), np.float64)
        _lib.check(self._lib.c3_get_root_counts(self._h, _dptr(out)))
        return out

    def fetch_column_index(self, node):
        out = np.zeros(self.n_columns, np.int64)
        _lib.check(self._lib.c3_get_column_index(self._h, int(node), _i64ptr(out)))
        return out

    def set_column_index(self, index):
        """the root's ``index`` (column -> pattern) when the tables came from the host;
        enables :meth:`get_site_lh`"""
        ix = np.ascontiguousarray(index, dtype=np.int64)
        _lib.check(self._lib.c3_set_column_index(self._h, _i64ptr(ix), ix.shape[0]))
        self.n_columns = int(ix.shape[0])

    def get_site_lh(self, bin_=0):
        """per-column likelihoods lh[root.index], gathered on the device"""
        out = np.empty(self.n_columns, np.float64)
        _lib.check(self._lib.c3_get_site_lh(self._h, int(bin_), _dptr(out), out.shape[0]))
        return out

    # -- sharded runs: the all-reduce fused into the reduction kernel ------------------------
    def comm_mailbox(self, nranks):
        """allocate this rank's mailbox; returns (device pointer, 64-byte CUDA IPC handle)"""
        ptr = c_void_p()
        handle = (ctypes.c_uint8 * 64)()
        _lib.check(self._lib.c3_comm_mailbox(self._h, int(nranks), byref(ptr), handle))
        return ptr.value, bytes(handle)

    def comm_attach(self, nranks, rank, ipc_handles=None, device_ptrs=None):
        """peers' mailboxes: ``ipc_handles`` = list of 64-byte handles (one per rank, other
        processes) or ``device_ptrs`` = list of device pointers valid in this process.  From
        now on ``evaluate`` returns the TOTAL over all ranks."""
        hbuf = pbuf = None
        if ipc_handles is not None:
            raw = b"".join(ipc_handles)
            assert len(raw) == 64 * nranks
            hbuf = (ctypes.c_uint8 * len(raw)).from_buffer_copy(raw)
        if device_ptrs is not None:
            pbuf = (c_void_p * nranks)(*[int(p) for p in device_ptrs])
        _lib.check(self._lib.c3_comm_attach(self._h, int(nranks), int(rank), hbuf, pbuf))

    def comm_detach(self):
        _lib.check(self._lib.c3_comm_detach(self._h))

    # -- lifetime ------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.c3_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        """``stream``: a ``torch.cuda.Stream`` / raw cudaStream_t int / None"""
        raw = getattr(stream, "cuda_stream", stream)
        if stream is not None and not raw:
            # torch's default stream is handle 0, which the C ABI reads as "create your own":
            # name the legacy default stream explicitly (cudaStreamLegacy) so that work issued by
            # torch on its current stream (an NCCL all-reduce of the result) is ordered after ours
            raw = 1
        _lib.check(self._lib.c3_set_stream(self._h, c_void_p(raw) if raw else None))

    # -- model state ---------------------------------------------------------
    def set_eigen(self, slot, roots, evT, evI):
        """eigen-system of one distinct Q (EigenExponentiator state)."""
        is_complex = any(np.iscomplexobj(a) for a in (roots, evT, evI))
        dt = np.complex128 if is_complex else np.float64
        roots = np.ascontiguousarray(roots, dtype=dt)
        evT = np.ascontiguousarray(evT, dtype=dt)
        evI = np.ascontiguousarray(evI, dtype=dt)
        M = self.n_states
        assert roots.shape == (M,) and evT.shape == (M, M) and evI.shape == (M, M)
        f = lambda a: a.view(np.float64).ctypes.data_as(ctypes.POINTER(c_double))  # noqa: E731
        _lib.check(
            self._lib.c3_set_eigen(self._h, slot, f(roots), f(evT), f(evI), int(is_complex))
        )

    def set_Q(self, slot, Q):
        """rate matrix for the Pade path."""
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        assert Q.shape == (self.n_states, self.n_states)
        _lib.check(self._lib.c3_set_Q(self._h, slot, _dptr(Q)))

    def set_tn93(self, slot, pi, alpha1, alpha2):
        """closed-form slot of the F81 / HKY85 / TN93 family (calc_TN93_P); pi in T, C, A, G order"""
        pi = np.ascontiguousarray(pi, dtype=np.float64)
        assert pi.shape == (4,) and self.n_states == 4
        _lib.check(self._lib.c3_set_tn93(self._h, slot, _dptr(pi), float(alpha1), float(alpha2)))

    def set_edge_qslots(self, qslots):
        q = np.ascontiguousarray(np.broadcast_to(qslots, (self.n_nodes, self.n_bins)), np.int32)
        _lib.check(self._lib.c3_set_edge_qslots(self._h, _i32ptr(q)))

    def set_psub(self, node, bin_, P):
        P = np.ascontiguousarray(P, dtype=np.float64)
        _lib.check(self._lib.c3_set_psub(self._h, int(node), int(bin_), _dptr(P)))

    def set_distances(self, dist):
        """``dist[n_nodes, n_bins]`` = length[edge] * rate[bin] (root row ignored)."""
        d = dist
        if not (isinstance(d, np.ndarray) and d.dtype == np.float64 and d.shape == (self.n_nodes, self.n_bins)
                and d.flags.c_contiguous):
            d = np.ascontiguousarray(np.broadcast_to(dist, (self.n_nodes, self.n_bins)), np.float64)
        self._dist = d
        _lib.check(self._lib.c3_set_distances(self._h, _dptr(d)))

    def set_root_mprobs(self, pi):
        pi = np.ascontiguousarray(pi, dtype=np.float64)
        assert pi.shape == (self.n_states,)
        _lib.check(self._lib.c3_set_root_mprobs(self._h, _dptr(pi)))

    def set_bin_probs(self, bprobs):
        bp = np.ascontiguousarray(bprobs, dtype=np.float64)
        assert bp.shape == (self.n_bins,)
        _lib.check(self._lib.c3_set_bin_probs(self._h, _dptr(bp)))

    def set_fixed_motifs(self, fixed):
        fm = np.full(self.n_nodes, -1, np.int32)
        if fixed is not None:
            fm[:] = np.asarray(fixed, np.int32)
        _lib.check(self._lib.c3_set_fixed_motifs(self._h, _i32ptr(fm)))

    def invalidate(self):
        _lib.check(self._lib.c3_invalidate(self._h))

    # -- per-pattern underflow rescaling ------------------------------------------
    RESCALE_OFF, RESCALE_ON, RESCALE_AUTO = 0, 1, 2

    def set_rescaling(self, mode="auto", tip_threshold=0):
        """``"off"``: the reference's plain FP64 products (-inf once a pattern underflows,
        evolve/likelihood_tree.py:10); ``"on"``: per-pattern power-of-two exponents carried up
        the tree; ``"auto"`` (default): switch on when a blocking evaluation is not finite."""
        m = {"off": 0, "on": 1, "auto": 2, False: 0, True: 1}.get(mode, mode)
        _lib.check(self._lib.c3_set_rescaling(self._h, int(m), int(tip_threshold)))

    def rescaling(self):
        """(active, number of scaled nodes)"""
        a, n = c_int(0), c_int(0)
        _lib.check(self._lib.c3_get_rescaling(self._h, byref(a), byref(n)))
        return bool(a.value), n.value

    def get_root_exponents(self):
        """E[p]: the true per-pattern likelihood is lh[p] * 2**E[p] (zeros when rescaling is off)"""
        out = np.empty(self.root_rows, np.int32)
        _lib.check(self._lib.c3_get_root_exponents(self._h, _i32ptr(out), out.shape[0]))
        return out

    def get_scaled_lh(self):
        """[K, root rows] root likelihoods as stored: the true value is ``out[b, p] * 2**E[p]``"""
        out = np.empty((self.n_bins, self.root_rows), np.float64)
        for b in range(self.n_bins):
            _lib.check(self._lib.c3_get_scaled_lh(self._h, b, _dptr(out[b]), out.shape[1]))
        return out

    # -- evaluation ----------------------------------------------------------
    def evaluate(self) -> float:
        out = c_double(0.0)
        _lib.check(self._lib.c3_evaluate(self._h, byref(out)))
        return out.value

    def update(self, dist) -> float:
        d = dist
        if not (isinstance(d, np.ndarray) and d.dtype == np.float64 and d.shape == (self.n_nodes, self.n_bins)
                and d.flags.c_contiguous):
            d = np.ascontiguousarray(np.broadcast_to(dist, (self.n_nodes, self.n_bins)), np.float64)
        self._dist = d
        out = c_double(0.0)
        _lib.check(self._lib.c3_update(self._h, _dptr(d), byref(out)))
        return out.value

    def evaluate_device(self, device_ptr: int):
        """asynchronous; writes lnL to ``device_ptr`` (a double in device memory)"""
        _lib.check(self._lib.c3_evaluate_device(self._h, c_void_p(device_ptr)))

    def hmm_reduce(self, patch_of_bin, weight_of_bin, switch_matrix, patch_probs) -> float:
        """phylo-HMM reduction over the sites of the last evaluation's per-bin root likelihoods
        (c3_hmm_reduce; SiteHmm.__call__ / log_dot_reduce of the reference)"""
        pb = np.ascontiguousarray(patch_of_bin, np.int32)
        wb = np.ascontiguousarray(weight_of_bin, np.float64)
        sw = np.ascontiguousarray(switch_matrix, np.float64)
        pp = np.ascontiguousarray(patch_probs, np.float64)
        assert pb.shape == (self.n_bins,) and wb.shape == (self.n_bins,) and sw.shape == (2, 2) and pp.shape == (2,)
        out = c_double(0.0)
        _lib.check(self._lib.c3_hmm_reduce(self._h, 2, _i32ptr(pb), _dptr(wb), _dptr(sw), _dptr(pp), byref(out)))
        return out.value

    # -- accessors -------------------------------------------------------------
    def get_lh(self, bin_=0):
        out = np.empty(self.root_rows, np.float64)
        _lib.check(self._lib.c3_get_lh(self._h, int(bin_), _dptr(out), out.shape[0]))
        return out

    def get_psub(self, node, bin_=0):
        M = self.n_states
        out = np.empty((M, M), np.float64)
        _lib.check(self._lib.c3_get_psub(self._h, int(node), int(bin_), _dptr(out)))
        return out

    def get_partials(self, node, bin_=0):
        rows = int(self.patterns.nodes[node].n_rows)
        out = np.empty((rows, self.n_states), np.float64)
        _lib.check(self._lib.c3_get_partials(self._h, int(node), int(bin_), _dptr(out), rows))
        return out

    def set_profiling(self, on=True):
        _lib.check(self._lib.c3_set_profiling(self._h, int(bool(on))))

    KIND_NAMES = ("expm_eigen", "expm_pade", "prune", "prune_root", "reduce", "prune_small")

    def launch_profile(self):
        """per-launch (kind, ms, algorithmic bytes, executed flops) of the last evaluation"""
        n_max = 4 * self.n_nodes + 8
        kind = np.zeros(n_max, np.int32)
        ms = np.zeros(n_max, np.float64)
        nbytes = np.zeros(n_max, np.int64)
        flops = np.zeros(n_max, np.int64)
        n = ctypes.c_int(0)
        _lib.check(
            self._lib.c3_get_launch_profile(
                self._h, n_max, _i32ptr(kind), _dptr(ms), _i64ptr(nbytes), _i64ptr(flops), byref(n)
            )
        )
        return [
            {"kind": self.KIND_NAMES[kind[i]], "ms": float(ms[i]), "bytes": int(nbytes[i]),
             "flops": int(flops[i])}
            for i in range(n.value)
        ]  # fmt: skip

    def stats(self):
        st = _lib.C3Stats()
        _lib.check(self._lib.c3_get_stats(self._h, byref(st)))
        return st.as_dict()


def evaluate_many(engines) -> np.ndarray:
    """evaluate several independent engines concurrently (c3_evaluate_many): every engine's
    kernels are launched on its own stream before any result is waited for"""
    n = len(engines)
    out = np.zeros(n, np.float64)
    if n == 0:
        return out
    handles = (c_void_p * n)(*[e._h.value for e in engines])
    lib = engines[0]._lib
    _lib.check(lib.c3_evaluate_many(handles, n, _dptr(out)))
    return out


def update_many(engines, distances) -> np.ndarray:
    """new distances for every engine, then ``evaluate_many`` -- one FFI crossing per batch
    (c3_update_many).  ``distances``: one table [n_nodes, n_bins] shared by all engines, or an array
    [n_engines, n_nodes, n_bins]."""
    n = len(engines)
    out = np.zeros(n, np.float64)
    if n == 0:
        return out
    d = np.ascontiguousarray(distances, np.float64)
    per = engines[0].n_nodes * engines[0].n_bins
    stride = 0 if d.size == per else per
    assert d.size in (per, n * per) and all(e.n_nodes * e.n_bins == per for e in engines)
    handles = (c_void_p * n)(*[e._h.value for e in engines])
    _lib.check(engines[0]._lib.c3_update_many(handles, n, _dptr(d), stride, _dptr(out)))
    return out


def measure_dmma_peak(device=0) -> float:
    """sustained FP64 tensor-core TFLOP/s of this GPU (DMMA m8n8k4 issue loop)."""
    out = c_double(0.0)
    _lib.check(_lib.load().c3_measure_dmma_peak(device, byref(out)))
    return out.value
