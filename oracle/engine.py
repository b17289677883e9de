"""CPU evaluator with the same interface as cogent3_b200.engine.LikelihoodEngine.

TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py): used by tests/ as
the checker and test double, by __graft_entry__.smoke() as the checker, and by
bench.py's cpu_baseline / --impl reference legs as the timed CPU implementation.

modes
  "numpy" : oracle.py, pure numpy restatement (the pinned oracle)
  "port"  : the reference's execution structure -- per bin and edge numpy.inner
            (BLAS, evolve/likelihood_calculation.py:86), then the C restatement of
            the numba kernels (oracle/pruning.c) single-threaded, exactly as numba
            runs them
  "fused" : one fused, OpenMP-parallel C loop per node over all bins -- stronger
            than the reference; lets the GPU be compared with ALL host cores
"""

from __future__ import annotations

import ctypes
import os

import numpy as np

from . import oracle

_LIB = None


def _clib():
    global _LIB
    if _LIB is None:
        from .build import build

        lib = ctypes.CDLL(build())
        pd = ctypes.POINTER(ctypes.c_double)
        lib.c3o_sum_input_likelihoods.argtypes = [
            ctypes.POINTER(ctypes.c_int64), pd, ctypes.POINTER(pd), ctypes.c_int, ctypes.c_int64,
            ctypes.c_int,
        ]  # fmt: skip
        lib.c3o_log_sum_across_sites.restype = ctypes.c_double
        lib.c3o_log_sum_across_sites.argtypes = [pd, pd, ctypes.c_int64]
        lib.c3o_fused_node.argtypes = [
            pd, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(pd),
            ctypes.POINTER(ctypes.POINTER(ctypes.c_int32)), ctypes.POINTER(pd), ctypes.c_int,
        ]  # fmt: skip
        lib.c3o_mix_log_sum.restype = ctypes.c_double
        lib.c3o_mix_log_sum.argtypes = [pd, pd, pd, pd, ctypes.c_int64, ctypes.c_int, ctypes.c_int]
        lib.c3o_max_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def _pd(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class _Lht:
    def __init__(self, nd, M):
        self.shape = [nd.n_rows, M]
        self.input_likelihoods = nd.table
        self.indexes = nd.indexes
        self.counts = nd.counts
        self.index = nd.index


class OracleEngine:
    def __init__(self, patterns, n_bins=1, device=0, stream=None, mode="numpy", n_threads=0):
        self.patterns = patterns
        self.topo = patterns.topo
        self.n_states = patterns.n_states
        self.n_bins = n_bins
        self.n_nodes = self.topo.n_nodes
        self.mode = mode
        self.n_threads = n_threads
        self.lht = [_Lht(nd, self.n_states) for nd in patterns.nodes]
        self.slots = {}
        self.qslots = np.zeros((self.n_nodes, n_bins), int)
        self.fixed_psub = {}
        self.dist = None
        self.pi = None
        self.bprobs = np.full(n_bins, 1.0 / n_bins)
        self.fixed = None
        self.root_rows = patterns.root.n_rows
        self.n_evaluations = 0
        if mode == "fused":
            self._init_fused()

    def close(self):
        pass

    # -- same setters as the CUDA engine ---------------------------------------
    def set_eigen(self, slot, roots, evT, evI):
        self.slots[slot] = oracle.EigenExponentiator(None, roots, evT.T, evT, evI)

    def set_Q(self, slot, Q):
        self.slots[slot] = oracle.PadeExponentiator(np.array(Q))

    def set_tn93(self, slot, pi, alpha1, alpha2):
        self.slots[slot] = oracle.TN93Exponentiator(pi, alpha1, alpha2)

    def set_edge_qslots(self, q):
        q = np.broadcast_to(q, (self.n_nodes, self.n_bins))
        for n, b in zip(*np.nonzero(q >= 0), strict=True):  # negative: leave as is
            self.qslots[n, b] = q[n, b]
            self.fixed_psub.pop((int(n), int(b)), None)

    def set_psub(self, node, bin_, P):
        self.fixed_psub[(node, bin_)] = np.array(P)

    def set_distances(self, d):
        self.dist = np.broadcast_to(d, (self.n_nodes, self.n_bins)).copy()

    def set_root_mprobs(self, pi):
        self.pi = np.array(pi, float)

    def set_bin_probs(self, bp):
        self.bprobs = np.array(bp, float)

    def set_fixed_motifs(self, fixed):
        self.fixed = None if fixed is None else list(fixed)

    def invalidate(self):
        pass

    def _psubs(self):
        out = []
        for b in range(self.n_bins):
            row = [None] * self.n_nodes
            for n in range(self.n_nodes - 1):
                if (n, b) in self.fixed_psub:
                    row[n] = self.fixed_psub[(n, b)]
                else:
                    row[n] = self.slots[int(self.qslots[n, b])](self.dist[n, b])
            out.append(row)
        return out

    # -- evaluation ----------------------------------------------------------------
    def evaluate(self):
        self.n_evaluations += 1
        self._ps = self._psubs()
        if self.mode == "numpy":
            lnL, self._lhs, self._plhs = oracle.log_likelihood(
                self.topo.children, self.lht, self._ps, self.pi, self.bprobs,
                fixed_motif=self.fixed, return_parts=True,
            )  # fmt: skip
            return lnL
        if self.mode == "port":
            return self._evaluate_port()
        if self.mode == "fused":
            return self._evaluate_fused()
        raise ValueError(self.mode)

    def update(self, dist):
        self.set_distances(dist)
        return self.evaluate()

    def _evaluate_port(self):
        lib = _clib()
        kids_of = self.topo.children
        M = self.n_states
        lhs = []
        if not hasattr(self, "_port_buf"):
            self._port_buf = [
                [np.ones((self.lht[n].shape[0], M)) if kids_of[n] else None for n in range(self.n_nodes)]
                for _ in range(self.n_bins)
            ]
        PD = ctypes.POINTER(ctypes.c_double)
        for b in range(self.n_bins):
            plh = [None] * self.n_nodes
            for n, kids in enumerate(kids_of):
                if not kids:
                    plh[n] = self.lht[n].input_likelihoods
                    continue
                inners = [np.inner(plh[k], self._ps[b][k]) for k in kids]  # BLAS, like numpy.inner
                out = self._port_buf[b][n]  # recycled, like PartialLikelihoodProductDefn
                ptrs = (PD * len(kids))(*[_pd(a) for a in inners])
                ix = self.lht[n].indexes
                lib.c3o_sum_input_likelihoods(
                    ix.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _pd(out), ptrs, len(kids),
                    out.shape[0], M,
                )  # fmt: skip
                self._apply_fixed(n, out)
                plh[n] = out
            lhs.append(np.inner(plh[-1], self.pi))
        self._lhs = lhs
        mix = oracle.weighted_sum_lh(self.bprobs, lhs) if self.n_bins > 1 else lhs[0]
        counts = self.lht[-1].counts
        return float(lib.c3o_log_sum_across_sites(_pd(np.ascontiguousarray(mix)), _pd(counts), len(counts)))

    def _apply_fixed(self, n, out):
        """fixed-motif mask of PartialLikelihoodProductDefnFixedMotif
        (evolve/likelihood_calculation.py:50-59)"""
        if self.fixed is None or self.fixed[n] in (None, -1):
            return
        keep = int(self.fixed[n])
        for m in range(out.shape[-1]):
            if m != keep:
                out[..., m] = 0.0

    def _init_fused(self):
        K, M = self.n_bins, self.n_states
        self._rows = []
        self._idx32 = []
        for n, nd in enumerate(self.patterns.nodes):
            self._rows.append(np.zeros((nd.n_rows, K, M)))
            self._idx32.append(None if nd.is_tip else np.ascontiguousarray(nd.indexes, np.int32))

    def _evaluate_fused(self):
        lib = _clib()
        K, M = self.n_bins, self.n_states
        PD = ctypes.POINTER(ctypes.c_double)
        PI = ctypes.POINTER(ctypes.c_int32)
        Pall = [None] * self.n_nodes
        for n in range(self.n_nodes - 1):
            Pn = np.ascontiguousarray(np.stack([self._ps[b][n] for b in range(K)]))
            if self.topo.is_tip[n]:
                # pre-multiplied tip rows, like the GPU engine's tip_inner tables
                tab = self.lht[n].input_likelihoods
                self._rows[n][:] = np.stack([np.inner(tab, Pn[b]) for b in range(K)], axis=1)
            else:
                Pall[n] = Pn
        for n, kids in enumerate(self.topo.children):
            if not kids:
                continue
            C = len(kids)
            src = (PD * C)(*[_pd(self._rows[k]) for k in kids])
            idx = (PI * C)(*[self._idx32[n][c].ctypes.data_as(PI) for c in range(C)])
            Ps = (PD * C)(*[_pd(Pall[k]) if Pall[k] is not None else None for k in kids])
            lib.c3o_fused_node(_pd(self._rows[n]), self._rows[n].shape[0], K, M, C, src, idx, Ps,
                               self.n_threads)  # fmt: skip
            self._apply_fixed(n, self._rows[n])
        root = self._rows[-1]
        self._lhs = [root[:, b, :] @ self.pi for b in range(K)]
        counts = self.lht[-1].counts
        return float(lib.c3o_mix_log_sum(_pd(root), _pd(self.pi), _pd(self.bprobs), _pd(counts),
                                         root.shape[0], K, M))  # fmt: skip

    # -- accessors -------------------------------------------------------------------
    def get_lh(self, bin_=0):
        return np.asarray(self._lhs[bin_])

    def get_psub(self, node, bin_=0):
        return self._ps[bin_][node]

    def get_partials(self, node, bin_=0):
        if self.mode == "fused":
            return self._rows[node][:, bin_, :]
        if self.topo.is_tip[node]:
            return np.inner(self._plhs[bin_][node], self._ps[bin_][node])
        return self._plhs[bin_][node]

    def stats(self):
        return {"evaluations": self.n_evaluations}


def host_threads():
    return int(_clib().c3o_max_threads()) if os.cpu_count() else 1
