"""Seeded synthetic workloads of SURVEY §8(d) (numpy only).

  * tree: :func:`cogent3_b200.tree.random_join_tree`
  * "evolved" alignment: root states i.i.d. from pi, every edge resamples a site
    uniformly with probability 1 - exp(-length)
  * "dense" alignment: every tip i.i.d. uniform (worst case for per-node
    compression: P ~ L at every node with more than a few tips below it)

State codes are indices into the model alphabet (T, C, A, G / 61 sense codons).
"""

from __future__ import annotations

import numpy as np

from .tree import Topology, random_join_tree

SEED = 20251210


def evolved_alignment(topo: Topology, lengths, n_columns, n_states, rng, pi=None, dtype=np.uint8):
    """returns codes[n_tips, L] in ``topo.tip_names`` order."""
    pi = np.full(n_states, 1.0 / n_states) if pi is None else np.asarray(pi, float)
    state = [None] * topo.n_nodes
    state[topo.root] = rng.choice(n_states, size=n_columns, p=pi).astype(dtype)
    # pre-order = reversed post-order
    for n in range(topo.n_nodes - 1, -1, -1):
        for k in topo.children[n]:
            p_change = 1.0 - np.exp(-lengths[k])
            resample = rng.random(n_columns) < p_change
            fresh = rng.integers(0, n_states, size=n_columns, dtype=dtype)
            state[k] = np.where(resample, fresh, state[n]).astype(dtype)
        if topo.children[n]:
            state[n] = None  # free
    return np.stack([state[n] for n in topo.tips])


def dense_alignment(topo: Topology, n_columns, n_states, rng, dtype=np.uint8):
    return rng.integers(0, n_states, size=(len(topo.tips), n_columns), dtype=dtype)


def make_workload(n_tips, n_columns, n_states, kind="evolved", seed=SEED, shard=0):
    """(topo, lengths, codes) for one shard; the tree depends only on ``seed``
    (all shards of a multi-GPU run share it), the columns on (seed, shard)."""
    topo, lengths = random_join_tree(n_tips, np.random.default_rng(seed))
    rng = np.random.default_rng([seed, 1 + shard])
    if kind == "evolved":
        codes = evolved_alignment(topo, lengths, n_columns, n_states, rng)
    elif kind == "dense":
        codes = dense_alignment(topo, n_columns, n_states, rng)
    else:
        raise ValueError(kind)
    return topo, lengths, codes
