"""MultiDeviceEngine: one likelihood function's site patterns sharded over several GPUs of ONE
process -- what a cogent3 user gets from ``make_likelihood_function(tree, devices=[0, ..., 7])``
without torchrun (the one-process-per-GPU variant is cogent3_b200/sharded.py).

The GLOBAL per-node compression stays what cogent3 built (``lht``, bit-exact, so every accessor keeps
its meaning): the root's pattern rows are split into contiguous blocks, one per device, and every
block gets the sub-tables it induces -- for each node the rows some row of its parent block refers
to, renumbered in ascending order, plus the all-gap row.  Counts are split with the root rows, so
``sum over shards`` is exactly the reference's ``sum over patterns`` regrouped.

Evaluation: parameters are replicated (every engine receives the same q slots, distances, pi, bin
probabilities), the N engines are launched together by ONE ``c3_evaluate_many`` call (each on its
own device and stream), and the sum over shards is FUSED into the reduction kernels through
peer-memory mailboxes (``c3_comm_attach(device_ptrs=...)``): every engine returns the total.  When the
devices cannot map each other's memory the shard values are added on the host in shard order.
"""

from __future__ import annotations

import numpy as np

from .engine import LikelihoodEngine, evaluate_many
from .patterns import NodePatterns, PatternTree


def split_root_rows(n_patterns: int, n_shards: int):
    """contiguous [start, stop) blocks of the root's pattern rows (the gap row excluded)"""
    base, extra = divmod(n_patterns, n_shards)
    out, start = [], 0
    for s in range(n_shards):
        stop = start + base + (1 if s < extra else 0)
        out.append((start, stop))
        start = stop
    return out


def induced_shard(patterns: PatternTree, start: int, stop: int) -> PatternTree:
    """the pattern tables of root rows [start, stop) (+ the gap row): per node, the rows its parent's
    kept rows refer to, in ascending (= first-seen) order, renumbered; the gap row stays last"""
    topo = patterns.topo
    N = topo.n_nodes
    keep = [None] * N  # sorted global row numbers kept at every node (the gap row always last)
    root = topo.root
    gap_root = patterns.nodes[root].n_rows - 1
    keep[root] = np.concatenate([np.arange(start, stop, dtype=np.int64), [gap_root]])
    nodes = [None] * N
    for n in range(N - 1, -1, -1):  # parents before children (post-order numbering)
        nd = patterns.nodes[n]
        rows = keep[n]
        if nd.is_tip:
            nodes[n] = NodePatterns(nd.name, True, None, np.zeros(len(rows)), None,
                                    table=np.ascontiguousarray(nd.table[rows]))  # fmt: skip
            continue
        sub = nd.indexes[:, rows]  # [C, kept rows] global child rows
        new_ix = np.empty(sub.shape, np.int64)  # C-contiguous [C, rows], like LikelihoodTreeEdge.indexes
        for c, child in enumerate(topo.children[n]):
            gap_c = patterns.nodes[child].n_rows - 1
            used = np.unique(np.concatenate([sub[c], [gap_c]]))  # ascending; gap_c is the largest
            if keep[child] is not None:  # (a DAG would merge here; a tree visits each child once)
                used = np.union1d(keep[child], used)
            keep[child] = used
            new_ix[c] = np.searchsorted(used, sub[c])
        counts = nd.counts[rows] if n == root else np.zeros(len(rows))
        nodes[n] = NodePatterns(nd.name, False, None, np.asarray(counts, np.float64), None, indexes=new_ix)
    shard = PatternTree(topo, nodes, patterns.n_states, int(nodes[root].counts.sum()))
    # the shard's own column numbering is never used: per-column values are gathered from the global
    # ``index`` over the concatenated per-pattern values (MultiDeviceEngine.get_site_lh)
    nodes[root].index = np.zeros(0, np.int64)
    return shard


class MultiDeviceEngine:
    """``LikelihoodEngine`` interface over N engines on N devices (see the module docstring)"""

    def __init__(self, patterns: PatternTree, n_bins: int = 1, devices=(0,), engine_cls=LikelihoodEngine,
                 fuse=True):  # fmt: skip
        devices = [int(d) for d in devices]
        assert len(devices) >= 1
        self.patterns = patterns
        self.topo = patterns.topo
        self.n_states = int(patterns.n_states)
        self.n_bins = int(n_bins)
        self.n_nodes = self.topo.n_nodes
        self.devices = devices
        self.root_rows = int(patterns.root.n_rows)
        self.n_columns = int(len(patterns.root.index)) if patterns.root.index is not None else 0
        n_pat = self.root_rows - 1
        n = max(1, min(len(devices), max(n_pat, 1)))
        self.blocks = split_root_rows(n_pat, n)
        self.engines = []
        try:
            for (a, b), dev in zip(self.blocks, devices, strict=False):
                self.engines.append(engine_cls(induced_shard(patterns, a, b), n_bins=n_bins, device=dev))
        except Exception:
            self.close()
            raise
        self.fused = False
        if fuse and len(self.engines) > 1 and hasattr(self.engines[0], "comm_mailbox"):
            try:
                n = len(self.engines)
                ptrs = [e.comm_mailbox(n)[0] for e in self.engines]
                for r, e in enumerate(self.engines):
                    e.comm_attach(n, r, device_ptrs=ptrs)
                self.fused = True
            except Exception:  # noqa: BLE001 - no peer mapping between these devices: add on the host
                for e in self.engines:
                    try:
                        e.comm_detach()
                    except Exception:  # noqa: BLE001
                        pass
                self.fused = False

    # -- lifetime ------------------------------------------------------------
    def close(self):
        for e in getattr(self, "engines", []):
            e.close()
        self.engines = []

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # -- replicated model state --------------------------------------------------
    def _all(self, name, *args):
        for e in self.engines:
            getattr(e, name)(*args)

    def set_eigen(self, slot, roots, evT, evI):
        self._all("set_eigen", slot, roots, evT, evI)

    def set_Q(self, slot, Q):
        self._all("set_Q", slot, Q)

    def set_edge_qslots(self, qslots):
        self._all("set_edge_qslots", qslots)

    def set_psub(self, node, bin_, P):
        self._all("set_psub", node, bin_, P)

    def set_distances(self, dist):
        self._all("set_distances", dist)

    def set_root_mprobs(self, pi):
        self._all("set_root_mprobs", pi)

    def set_bin_probs(self, bprobs):
        self._all("set_bin_probs", bprobs)

    def set_fixed_motifs(self, fixed):
        self._all("set_fixed_motifs", fixed)

    def invalidate(self):
        self._all("invalidate")

    def set_rescaling(self, mode="auto", tip_threshold=0):
        self._all("set_rescaling", mode, tip_threshold)

    # -- evaluation ----------------------------------------------------------
    def evaluate(self) -> float:
        if len(self.engines) == 1:
            return self.engines[0].evaluate()
        vals = self._evaluate_many()
        if self.fused:
            return float(vals[0])  # every engine holds the total (added in shard order on the device)
        total = 0.0
        for v in vals:  # shard order: deterministic
            total += float(v)
        return total

    def _evaluate_many(self):
        if hasattr(self.engines[0], "_h"):
            return evaluate_many(self.engines)  # one FFI crossing; all launched before any is waited for
        return [e.evaluate() for e in self.engines]

    def update(self, dist) -> float:
        self.set_distances(dist)
        return self.evaluate()

    # -- accessors (global numbering) -------------------------------------------
    def get_lh(self, bin_=0):
        """per-pattern root likelihoods in the GLOBAL pattern numbering (gap row last)"""
        parts = [e.get_lh(bin_) for e in self.engines]
        return np.concatenate([p[:-1] for p in parts] + [parts[0][-1:]])

    def get_site_lh(self, bin_=0):
        return self.get_lh(bin_)[np.asarray(self.patterns.root.index)]

    def get_psub(self, node, bin_=0):
        return self.engines[0].get_psub(node, bin_)

    def rescaling(self):
        return self.engines[0].rescaling()

    def stats(self):
        out = None
        for e in self.engines:
            st = e.stats()
            if out is None:
                out = dict(st)
            else:
                for k, v in st.items():
                    if k not in ("n_levels", "padded_states", "evaluations"):
                        out[k] += v
        out["devices"] = len(self.engines)
        out["fused_allreduce"] = bool(self.fused)
        return out
