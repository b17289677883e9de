"""TEST DOUBLE of cogent3_b200.engine.LikelihoodEngine backed by the CPU oracle.

Lets the host-side logic (likelihood.py, sharded.py, dropin.py) be exercised on a
machine without a GPU.  Lives under tests/ on purpose: the product never imports it.
"""

from __future__ import annotations

import numpy as np

from oracle import oracle


class _Lht:
    def __init__(self, nd):
        self.shape = [nd.n_rows, nd.table.shape[1] if nd.is_tip else None]
        self.input_likelihoods = nd.table
        self.indexes = nd.indexes
        self.counts = nd.counts
        self.index = nd.index


class OracleEngine:
    def __init__(self, patterns, n_bins=1, device=0, stream=None):
        self.patterns = patterns
        self.topo = patterns.topo
        self.n_states = patterns.n_states
        self.n_bins = n_bins
        self.n_nodes = self.topo.n_nodes
        self.lht = [_Lht(nd) for nd in patterns.nodes]
        for x in self.lht:
            x.shape[1] = self.n_states
        self.slots = {}
        self.qslots = np.zeros((self.n_nodes, n_bins), int)
        self.fixed_psub = {}
        self.dist = None
        self.pi = None
        self.bprobs = np.full(n_bins, 1.0 / n_bins)
        self.fixed = None
        self.root_rows = patterns.root.n_rows
        self.n_evaluations = 0

    def close(self):
        pass

    def set_eigen(self, slot, roots, evT, evI):
        self.slots[slot] = oracle.EigenExponentiator(None, roots, evT.T, evT, evI)

    def set_Q(self, slot, Q):
        self.slots[slot] = oracle.PadeExponentiator(np.array(Q))

    def set_edge_qslots(self, q):
        self.qslots = np.broadcast_to(q, (self.n_nodes, self.n_bins)).copy()

    def set_psub(self, node, bin_, P):
        self.fixed_psub[(node, bin_)] = np.array(P)

    def set_distances(self, d):
        self.dist = np.broadcast_to(d, (self.n_nodes, self.n_bins)).copy()

    def set_root_mprobs(self, pi):
        self.pi = np.array(pi)

    def set_bin_probs(self, bp):
        self.bprobs = np.array(bp)

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

    def evaluate(self):
        self.n_evaluations += 1
        self._ps = self._psubs()
        lnL, self._lhs, self._plhs = oracle.log_likelihood(
            self.topo.children, self.lht, self._ps, self.pi, self.bprobs,
            fixed_motif=self.fixed, return_parts=True,
        )  # fmt: skip
        return lnL

    def update(self, dist):
        self.set_distances(dist)
        return self.evaluate()

    def get_lh(self, bin_=0):
        return self._lhs[bin_]

    def get_psub(self, node, bin_=0):
        return self._ps[bin_][node]

    def get_partials(self, node, bin_=0):
        if self.topo.is_tip[node]:
            return np.inner(self._plhs[bin_][node], self._ps[bin_][node])
        return self._plhs[bin_][node]

    def stats(self):
        return {"evaluations": self.n_evaluations}
