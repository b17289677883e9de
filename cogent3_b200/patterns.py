"""Per-node site-pattern compression, vectorised and bit-exact.

cogent3 compresses the alignment *at every tree node*: a tip keeps the unique
motifs it shows, an internal node keeps the unique tuples of its children's
pattern indices, both in first-seen column order, each followed by one extra
"all gap" row with count 0 (cogent3 evolve/likelihood_tree.py:13-70 ``
_LikelihoodTreeEdge.__init__``, :186-203 ``_indexed``, :223-278
``make_likelihood_tree_leaf``).  The reference does this with a Python dict per
column (6.8-8.2 s per 64 taxa x 100k columns); here the same
tables are produced with sorts (numpy) so 1M-10M column alignments are
tractable.  The results (``uniq``, ``indexes``, ``counts``, ``index``) are
integer-identical to the reference's; tests/test_patterns.py checks that
against the oracle's dict-loop restatement and against golden dumps of the live
reference.
"""

from __future__ import annotations

import numpy as np

from .tree import Topology

_I64_MAX = np.iinfo(np.int64).max

try:  # optional: only its hash-based ranking is used
    from pandas import factorize as _factorize
except Exception:  # noqa: BLE001
    _factorize = None


def first_seen_unique(keys: np.ndarray):
    """``_indexed`` for a 1-D integer key array.

    Returns (first[U], counts[U], index[L]) where ``first`` are the column
    numbers at which each distinct key first appears (ascending), ``counts`` how
    often each occurs and ``index[c]`` the first-seen rank of column c's key."""
    keys = np.asarray(keys)
    L = keys.shape[0]
    if L == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.float64), np.zeros(0, np.int64)
    kmin = keys.min()
    span = int(keys.max()) - int(kmin) + 1
    if span <= max(4 * L, 1 << 16) and span <= (1 << 27):
        # small key range: O(L) scatter instead of an O(L log L) sort
        k = (keys - kmin).astype(np.int64, copy=False)
        # first column of every key: assign the column numbers in REVERSE order -- for a repeated index numpy
        # keeps the last value assigned, i.e. the smallest column (ufunc.at does the same 10x slower)
        first_of_key = np.full(span, L, np.int64)
        first_of_key[k[::-1]] = np.arange(L - 1, -1, -1, dtype=np.int64)
        present = np.flatnonzero(first_of_key < L)
        first = first_of_key[present]
        order = np.argsort(first, kind="stable")
        first = first[order]
        rank_of_key = np.empty(span, np.int64)
        rank_of_key[present[order]] = np.arange(len(present), dtype=np.int64)
        index = rank_of_key[k]
    elif _factorize is not None:
        # wide key range: a hash table (pandas.factorize numbers the keys in first-seen order, O(L));
        # the sort-based numpy path below is the fallback where pandas is not installed
        index, _uniques = _factorize(np.ascontiguousarray(keys, np.int64))
        index = index.astype(np.int64, copy=False)
        n_unique = len(_uniques)
        first = np.full(n_unique, L, np.int64)
        first[index[::-1]] = np.arange(L - 1, -1, -1, dtype=np.int64)  # (first column of every key, as above)
    else:
        _, first_s, inverse = np.unique(keys, return_index=True, return_inverse=True)
        order = np.argsort(first_s, kind="stable")
        rank = np.empty(len(order), np.int64)
        rank[order] = np.arange(len(order), dtype=np.int64)
        index = rank[inverse.reshape(-1)]
        first = first_s[order].astype(np.int64)
    counts = np.bincount(index, minlength=len(first)).astype(np.float64)
    return first, counts, index


def _combine_keys(assignments, sizes):
    """mixed-radix key of the per-child pattern indices; folds pairwise through a
    dense re-ranking whenever the radix product would overflow int64."""
    key = assignments[0].astype(np.int64)
    span = int(sizes[0])
    for a, s in zip(assignments[1:], sizes[1:], strict=True):
        s = int(s)
        if span > _I64_MAX // max(s, 1):
            _, _, key = first_seen_unique(key)
            span = int(key.max()) + 1 if len(key) else 1
            if span > _I64_MAX // max(s, 1):  # pragma: no cover - >9e18 patterns
                raise OverflowError("pattern key overflow")
        key = key * s + a
        span = span * s
    return key


class NodePatterns:
    """pattern table of one node (mirrors the attributes of cogent3's
    ``LikelihoodTreeLeaf`` / ``LikelihoodTreeEdge`` that the hot path reads)."""

    __slots__ = ("ambig", "counts", "index", "indexes", "is_tip", "name", "table", "uniq")

    def __init__(self, name, is_tip, uniq, counts, index, table=None, indexes=None, ambig=None):
        self.name = name
        self.is_tip = is_tip
        self.uniq = uniq
        self.counts = counts
        self.index = index
        self.table = table  # tips: [U+1, M] indicator rows (float64)
        self.indexes = indexes  # internal: [C, P+1] int64 child rows
        self.ambig = ambig

    @property
    def n_rows(self):
        """P+1: patterns plus the trailing gap row"""
        return len(self.counts)


class PatternTree:
    """the ``lht`` of the reference, flattened: one NodePatterns per node."""

    def __init__(self, topo: Topology, nodes, n_states, n_columns):
        self.topo = topo
        self.nodes = nodes
        self.n_states = n_states
        self.n_columns = n_columns

    @property
    def root(self):
        return self.nodes[self.topo.root]

    def rows(self):
        return np.array([nd.n_rows for nd in self.nodes], np.int64)

    def work_units(self, n_bins=1, nodes=None):
        """(U_mv, U_gp) of SURVEY §8(d): K*sum_edges rows(child), K*sum_edges
        rows(parent) over edges whose parent is in ``nodes`` (default all)."""
        u_mv = u_gp = 0
        for n, kids in enumerate(self.topo.children):
            if not kids or (nodes is not None and n not in nodes):
                continue
            for k in kids:
                u_mv += self.nodes[k].n_rows
                u_gp += self.nodes[n].n_rows
        return n_bins * u_mv, n_bins * u_gp


def compress_tip(codes, symbol_rows, name):
    """tip table from integer symbol codes (one per column).

    ``symbol_rows[S, M]`` holds the 0/1 indicator row of every symbol the caller's
    encoding can produce (unambiguous states first, then ambiguity codes; what
    ``get_matched_array`` returns, likelihood_tree.py:206-220)."""
    codes = np.asarray(codes)
    first, counts, index = first_seen_unique(codes)
    uniq = codes[first].astype(np.int64)
    M = symbol_rows.shape[1]
    table = np.concatenate([symbol_rows[uniq].reshape(len(uniq), M), np.ones((1, M))]).astype(
        np.float64
    )
    uniq = np.concatenate([uniq, [-1]])
    counts = np.concatenate([counts, [0.0]])
    return NodePatterns(name, True, uniq, counts, index, table=table, ambig=table.sum(axis=1))


def compress_node(children, name):
    """internal-node table from its children's (already compressed) tables."""
    assignments = [c.index for c in children]
    sizes = [c.n_rows for c in children]
    key = _combine_keys(assignments, sizes)
    first, counts, index = first_seen_unique(key)
    uniq = np.stack([a[first] for a in assignments], axis=1).astype(np.int64)
    gap = np.array([[s - 1 for s in sizes]], np.int64)
    uniq = np.concatenate([uniq.reshape(len(first), len(children)), gap])
    counts = np.concatenate([counts, [0.0]])
    indexes = np.ascontiguousarray(uniq.T)
    ambig = np.prod([c.ambig[ix] for ix, c in zip(indexes, children, strict=True)], axis=0)
    return NodePatterns(name, False, uniq, counts, index, indexes=indexes, ambig=ambig)


def compress_alignment(topo: Topology, codes_by_tip, symbol_rows) -> PatternTree:
    """build the whole per-node pattern tree.

    ``codes_by_tip``: mapping tip name -> integer code array [L], or a 2-D array
    [n_tips, L] in ``topo.tip_names`` order."""
    if not isinstance(codes_by_tip, dict):
        arr = np.asarray(codes_by_tip)
        codes_by_tip = dict(zip(topo.tip_names, arr, strict=True))
    symbol_rows = np.asarray(symbol_rows, np.float64)
    nodes = [None] * topo.n_nodes
    L = None
    for n, kids in enumerate(topo.children):
        if not kids:
            codes = codes_by_tip[topo.names[n]]
            L = len(codes) if L is None else L
            if len(codes) != L:
                raise ValueError("sequences differ in length")
            nodes[n] = compress_tip(codes, symbol_rows, topo.names[n])
        else:
            nodes[n] = compress_node([nodes[k] for k in kids], topo.names[n])
    return PatternTree(topo, nodes, symbol_rows.shape[1], L)


def from_reference_lht(lht_root):
    """flatten a cogent3 ``LikelihoodTreeEdge`` hierarchy (the value of the "lht"
    Defn, evolve/likelihood_calculation.py:115-122) into (Topology, PatternTree).
    Used by the drop-in glue: the reference's own compression stays the source of
    truth there, only the arithmetic moves to the GPU."""
    names, children, nodes = [], [], []

    def rec(edge):
        kids = []
        if hasattr(edge, "_indexed_children"):
            for _ix, child in edge._indexed_children:
                kids.append(rec(child))
        nid = len(names)
        names.append(edge.edge_name)
        children.append(kids)
        if kids:
            nodes.append(
                NodePatterns(
                    edge.edge_name,
                    False,
                    np.asarray(edge.uniq, np.int64),
                    np.asarray(edge.counts, np.float64),
                    np.asarray(edge.index, np.int64),
                    indexes=np.ascontiguousarray(edge.indexes, dtype=np.int64),
                    ambig=np.asarray(edge.ambig, np.float64),
                )
            )
        else:
            nodes.append(
                NodePatterns(
                    edge.edge_name,
                    True,
                    list(edge.uniq),
                    np.asarray(edge.counts, np.float64),
                    np.asarray(edge.index, np.int64),
                    table=np.ascontiguousarray(edge.input_likelihoods, dtype=np.float64),
                    ambig=np.asarray(edge.ambig, np.float64),
                )
            )
        return nid

    rec(lht_root)
    topo = Topology(names, children)
    M = nodes[0].table.shape[1] if nodes[0].is_tip else lht_root.shape[1]
    return topo, PatternTree(topo, nodes, int(lht_root.shape[1]), len(lht_root.index))


class _DeviceNode:
    """lazy view of one node's table held by a CUDA engine (fetched on first use)"""

    def __init__(self, tree, n):
        self._t = tree
        self._n = n
        self.name = tree.topo.names[n]
        self.is_tip = not tree.topo.children[n]
        self._cache = {}

    @property
    def n_rows(self):
        return int(self._t.rows()[self._n])

    def _get(self, key, fn):
        if key not in self._cache:
            self._cache[key] = fn()
        return self._cache[key]

    @property
    def indexes(self):
        return self._get("indexes", lambda: self._t.engine.fetch_node_indexes(self._n))

    @property
    def uniq(self):
        if self.is_tip:
            raise AttributeError("tip symbols stay on the host side of the caller")
        return np.ascontiguousarray(self.indexes.T)

    @property
    def table(self):
        return self._get("table", lambda: self._t.engine.fetch_tip_table(self._n))

    @property
    def ambig(self):
        if self.is_tip:
            return self.table.sum(axis=1)
        raise AttributeError("ambig of internal nodes is not kept on the device")

    @property
    def index(self):
        return self._get("index", lambda: self._t.engine.fetch_column_index(self._n))

    @property
    def counts(self):
        if self._n == self._t.topo.root:
            return self._get("counts", self._t.engine.fetch_root_counts)
        idx = self.index  # raises unless the numbering was kept (keep_index)
        c = np.bincount(idx, minlength=self.n_rows).astype(np.float64)
        return c


class DevicePatternTree(PatternTree):
    """the pattern tree of an alignment compressed ON THE DEVICE (c3_compress_alignment):
    same attributes as PatternTree, tables fetched from the engine only when read."""

    def __init__(self, topo: Topology, engine, n_states, n_columns):
        self.topo = topo
        self.engine = engine
        self.n_states = n_states
        self.n_columns = n_columns
        self._rows = None
        self.nodes = [_DeviceNode(self, n) for n in range(topo.n_nodes)]

    def rows(self):
        if self._rows is None:
            self._rows = self.engine.fetch_node_rows()
        return self._rows
