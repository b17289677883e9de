"""Experiment: renumber every internal node's patterns so that the gather from its largest
internal child is monotone; time a full evaluation before / after (same arithmetic per row)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from cogent3_b200.patterns import compress_alignment, NodePatterns, PatternTree
from cogent3_b200.engine import LikelihoodEngine

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
cols = int(sys.argv[2]) if len(sys.argv) > 2 else None
problem = bench.build_problem(cfg, cols, "evolved", 0)
model_name, topo, lengths, codes, rows, bins, setup = problem
pt = compress_alignment(topo, codes, rows)

def permuted(pt):
    perm = [None] * topo.n_nodes  # new position of old row p
    nodes = list(pt.nodes)
    for n, kids in enumerate(topo.children):
        nd = pt.nodes[n]
        if not kids:
            perm[n] = np.arange(nd.n_rows)
            continue
        ix = nd.indexes.copy()  # [C, rows]
        # remap child rows to the children's new numbering
        for c, k in enumerate(kids):
            ix[c] = perm[k][ix[c]]
        internal = [c for c, k in enumerate(kids) if topo.children[k]]
        P = nd.n_rows - 1
        if internal and n != topo.root:
            c0 = max(internal, key=lambda c: pt.nodes[kids[c]].n_rows)
            order = np.argsort(ix[c0][:P], kind="stable")  # old rows in new order
        else:
            order = np.arange(P)
        pi = np.empty(nd.n_rows, np.int64)
        pi[order] = np.arange(P)
        pi[P] = P
        perm[n] = pi
        new_ix = np.empty_like(ix)
        new_ix[:, pi] = ix
        counts = np.empty_like(nd.counts); counts[pi] = nd.counts
        nodes[n] = NodePatterns(nd.name, False, None, counts, pi[nd.index], indexes=new_ix)
    return PatternTree(topo, nodes, pt.n_states, pt.n_columns)

def run(patterns, label):
    lf, _ = bench.make_lf(problem, patterns=patterns)
    eng = lf.engine
    base, rates = lf.lengths.copy(), lf.rates()
    v = lf.get_log_likelihood()
    import torch
    for k in range(5):
        eng.update(np.nan_to_num(bench.step_lengths(base, k))[:, None] * rates[None, :])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    N = 200
    for k in range(N):
        eng.update(np.nan_to_num(bench.step_lengths(base, k))[:, None] * rates[None, :])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / N
    eng.set_profiling(True)
    eng.update(np.nan_to_num(bench.step_lengths(base, 7))[:, None] * rates[None, :])
    prof = eng.launch_profile()
    print(label, "lnL", v, "ms/step", dt * 1e3, [(r["kind"], round(r["ms"], 4), round(r["bytes"] / max(r["ms"], 1e-9) / 1e6)) for r in prof])

run(pt, "reference order")
run(permuted(pt), "permuted      ")
