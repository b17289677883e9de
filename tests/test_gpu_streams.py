"""Independent root subtrees on their own streams (C3_STREAMS=1, opt-in): the sweep is launched per (root subtree,
level) with one stream per subtree, so that one subtree's launch ramps up while another's drains.  Same kernels,
same arguments per node: every partial row, per-pattern likelihood and lnL BIT-identical to the single-stream
sweep -- for full sweeps, dirty paths, 4 states and codons, and when the evaluation is replayed from a CUDA graph
(the subtree streams are forked and joined inside the capture)."""

import numpy as np
import pytest

from cogent3_b200 import models as hostmodels
from cogent3_b200.likelihood import LikelihoodFunction
from cogent3_b200.synthetic import make_workload

pytestmark = pytest.mark.gpu


def build(monkeypatch, model, topo, lengths, codes, bins, streams):
    monkeypatch.setenv("C3_STREAMS", "1" if streams else "0")
    m = hostmodels.get_model(model)
    lf = LikelihoodFunction(m, topo, bins=bins)
    lf.set_alignment(codes.astype(np.int64), np.eye(m.n_states))
    lf.set_motif_probs_from_data()
    lf.lengths[:-1] = np.asarray(lengths)[:-1]
    for k, p in enumerate(m.parameter_order):
        lf.set_param_rule(p, value=0.5 + 0.37 * k)
    if bins > 1:
        lf.set_param_rule("rate_shape", value=0.6)
    return lf


@pytest.mark.parametrize(
    "model,n_tips,L,bins",
    [("GTR", 40, 300_000, 4), ("HKY85", 64, 400_000, 1), ("CNFGTR", 16, 30_000, 1), ("HKY85", 12, 2_000, 2)],
)
def test_subtree_streams_are_bit_identical(model, n_tips, L, bins, monkeypatch):
    m = hostmodels.get_model(model)
    topo, lengths, codes = make_workload(n_tips, L, m.n_states, kind="evolved", seed=29)
    one = build(monkeypatch, model, topo, lengths, codes, bins, streams=False)
    many = build(monkeypatch, model, topo, lengths, codes, bins, streams=True)
    base = one.lengths.copy()
    values = []
    for k in range(40):  # the same structure 40 times: plain launches first, then a captured graph
        for lf in (one, many):
            lf.lengths[:] = base * (1.0 + 0.01 * (k % 4))
            lf._dirty_model = True
        a, b = one.get_log_likelihood(), many.get_log_likelihood()
        assert a == b, k
        values.append(b)
    for k in range(4, 40):
        assert values[k] == values[k % 4]
    if L >= 30_000:
        assert many.engine.stats()["graph_replays"] > 0
    for n in range(topo.n_nodes):
        if topo.is_tip[n]:
            continue
        for k in range(bins):
            np.testing.assert_array_equal(one.engine.get_partials(n, k), many.engine.get_partials(n, k))
    rng = np.random.default_rng(5)
    for n in rng.choice(topo.edges, size=5, replace=False):
        v = float(rng.uniform(0.01, 0.5))
        for lf in (one, many):
            lf.set_param_rule("length", edge=topo.names[n], value=v)
        assert one.get_log_likelihood() == many.get_log_likelihood()
    for k in range(bins):
        np.testing.assert_array_equal(one.engine.get_lh(k), many.engine.get_lh(k))
