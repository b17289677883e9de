"""Flat tree topology for the likelihood engine.

Node ids are in post-order (every child id < its parent id); the root is the
last node.  The edge above node ``n`` is identified by ``n`` (the root has
none).  Children are kept in tree order because the reference's per-node
pattern tables are keyed on the *ordered* tuple of child pattern indices
(cogent3 evolve/likelihood_tree.py:24,47 -- ``assignments = [c.index for c in
children]``) and its Defn graph walks ``edge.children`` in the same order
(evolve/likelihood_calculation.py:83-87).
"""

from __future__ import annotations

import numpy as np


class Topology:
    def __init__(self, names, children):
        self.names = list(names)
        self.children = [list(c) for c in children]
        n = len(self.names)
        assert len(self.children) == n
        self.parent = np.full(n, -1, np.int32)
        for p, kids in enumerate(self.children):
            for k in kids:
                if not k < p:
                    raise ValueError("node ids must be in post-order")
                self.parent[k] = p
        if n and (self.parent[:-1] < 0).any():
            raise ValueError("only the last node may be the root")
        # level = height above the tips (tips 0); nodes of one level are independent
        self.level = np.zeros(n, np.int32)
        for p, kids in enumerate(self.children):
            if kids:
                self.level[p] = 1 + max(self.level[k] for k in kids)
        self.is_tip = np.array([len(k) == 0 for k in self.children], bool)

    # -- basic properties -------------------------------------------------
    @property
    def n_nodes(self):
        return len(self.names)

    @property
    def root(self):
        return self.n_nodes - 1

    @property
    def tips(self):
        return [n for n in range(self.n_nodes) if self.is_tip[n]]

    @property
    def tip_names(self):
        return [self.names[n] for n in self.tips]

    @property
    def edges(self):
        """ids of all nodes with an edge above them (everything but the root)"""
        return list(range(self.n_nodes - 1))

    def index(self, name):
        return self.names.index(name)

    def ancestors(self, n):
        out = []
        n = int(self.parent[n])
        while n >= 0:
            out.append(n)
            n = int(self.parent[n])
        return out

    def child_csr(self):
        ptr = np.zeros(self.n_nodes + 1, np.int32)
        for n, kids in enumerate(self.children):
            ptr[n + 1] = ptr[n] + len(kids)
        idx = np.array([k for kids in self.children for k in kids], np.int32)
        return ptr, idx

    # -- construction -----------------------------------------------------
    @classmethod
    def from_newick(cls, text):
        """minimal newick reader (names, branch lengths).  Unnamed internal
        nodes get cogent3's names ``edge.0, edge.1, ...`` in the order their
        closing bracket is met; the root is ``root``
        (cogent3 core/tree.py name_unnamed_nodes semantics)."""
        topo, lengths = _parse_newick(text)
        return topo, lengths

    @classmethod
    def from_parent_arrays(cls, names, children):
        return cls(names, children)

    def to_newick(self, lengths=None):
        def rec(n):
            kids = self.children[n]
            s = "(" + ",".join(rec(k) for k in kids) + ")" if kids else ""
            if n != self.root:
                s += self.names[n]
                if lengths is not None:
                    s += f":{float(lengths[n])!r}"
            return s

        return rec(self.root) + ";"


def _parse_newick(text):
    text = text.strip()
    if text.endswith(";"):
        text = text[:-1]
    names, children, lengths = [], [], []
    pos = 0
    unnamed = []

    def read_label():
        nonlocal pos
        start = pos
        while pos < len(text) and text[pos] not in ",():;":
            pos += 1
        return text[start:pos].strip().strip("'\"")

    def node():
        nonlocal pos
        kids = []
        if pos < len(text) and text[pos] == "(":
            pos += 1
            while True:
                kids.append(node())
                if text[pos] == ",":
                    pos += 1
                    continue
                if text[pos] == ")":
                    pos += 1
                    break
                raise ValueError(f"bad newick at {pos}")
        name = read_label()
        length = None
        if pos < len(text) and text[pos] == ":":
            pos += 1
            length = float(read_label())
        nid = len(names)
        names.append(name)
        children.append(kids)
        lengths.append(length)
        if not name:
            unnamed.append(nid)
        return nid

    root = node()
    assert root == len(names) - 1
    k = 0
    for nid in unnamed:
        if nid == root:
            continue
        names[nid] = f"edge.{k}"
        k += 1
    names[root] = names[root] or "root"
    lens = np.array([np.nan if x is None else x for x in lengths], float)
    return Topology(names, children), lens


def random_join_tree(n_tips, rng, low=0.02, high=0.3):
    """SURVEY §8(d) synthetic tree: random-join (Yule-like) unrooted binary tree on
    tips ``t0..t{n-1}``: repeatedly join two uniformly chosen clades until three
    remain, join those at the root; branch lengths U(low, high).

    Returns (Topology, lengths[n_nodes]) with node ids in post-order."""
    assert n_tips >= 3
    # build with temporary ids, then renumber in post-order
    kids = {i: [] for i in range(n_tips)}
    label = {i: f"t{i}" for i in range(n_tips)}
    clades = list(range(n_tips))
    nxt = n_tips
    e = 0
    while len(clades) > 3:
        i, j = sorted(rng.choice(len(clades), size=2, replace=False))
        a, b = clades[i], clades[j]
        kids[nxt] = [a, b]
        label[nxt] = f"edge.{e}"
        e += 1
        clades = [c for k, c in enumerate(clades) if k not in (i, j)] + [nxt]
        nxt += 1
    kids[nxt] = list(clades)
    label[nxt] = "root"
    root = nxt
    order = []

    def rec(n):
        for k in kids[n]:
            rec(k)
        order.append(n)

    import sys

    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 4 * n_tips + 100))
    try:
        rec(root)
    finally:
        sys.setrecursionlimit(old)
    new_id = {old_id: i for i, old_id in enumerate(order)}
    names = [label[o] for o in order]
    children = [[new_id[k] for k in kids[o]] for o in order]
    topo = Topology(names, children)
    lengths = rng.uniform(low, high, size=topo.n_nodes)
    lengths[topo.root] = np.nan
    return topo, lengths
