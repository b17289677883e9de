"""The recomputing launch (prune4_rec_kernel: several consecutive small / medium tree levels in ONE launch without
a barrier -- a unit recomputes the rows of children that are computed in the same launch instead of waiting for
them) against the level-by-level launches: every partial row, per-pattern likelihood and lnL BIT-identical, for
full sweeps and dirty paths, K = 1, 2, 4, fixed motifs, a polytomy inside the recomputed levels."""

import numpy as np
import pytest

from cogent3_b200 import models as hostmodels
from cogent3_b200.likelihood import LikelihoodFunction
from cogent3_b200.synthetic import make_workload

pytestmark = pytest.mark.gpu


def build(monkeypatch, model, topo, lengths, codes, bins, rec, rec_units=None):
    for k in ("C3_REC", "C3_REC_UNITS", "C3_WAVE", "C3_PATH", "C3_GRID_SMALL"):
        monkeypatch.delenv(k, raising=False)
    if rec:
        monkeypatch.setenv("C3_REC", "1")
    if rec_units:
        monkeypatch.setenv("C3_REC_UNITS", str(rec_units))
    m = hostmodels.get_model(model)
    lf = LikelihoodFunction(m, topo, bins=bins)
    lf.set_alignment(codes.astype(np.int64), np.eye(4))
    lf.set_motif_probs_from_data()
    lf.lengths[:-1] = np.asarray(lengths)[:-1]
    for k, p in enumerate(m.parameter_order):
        lf.set_param_rule(p, value=0.5 + 0.37 * k)
    if bins > 1:
        lf.set_param_rule("rate_shape", value=0.6)
    return lf


def same_rows(a, b, topo, bins):
    for n in range(topo.n_nodes):
        if topo.is_tip[n]:
            continue
        for k in range(bins):
            np.testing.assert_array_equal(a.engine.get_partials(n, k), b.engine.get_partials(n, k))
    for k in range(bins):
        np.testing.assert_array_equal(a.engine.get_lh(k), b.engine.get_lh(k))


@pytest.mark.parametrize(
    "model,n_tips,L,bins,kind,rec_units",
    [
        ("GTR", 40, 300_000, 4, "evolved", None),
        ("HKY85", 64, 200_000, 1, "evolved", None),
        ("HKY85", 24, 60_000, 2, "dense", None),
        ("GTR", 30, 40_000, 4, "evolved", 10_000_000),  # every level recomputed, the root's included
        ("HKY85", 100, 20_000, 1, "evolved", None),
    ],
)
def test_rec_launch_is_bit_identical_to_the_level_launches(model, n_tips, L, bins, kind, rec_units, monkeypatch):
    m = hostmodels.get_model(model)
    topo, lengths, codes = make_workload(n_tips, L, m.n_states, kind=kind, seed=19)
    plain = build(monkeypatch, model, topo, lengths, codes, bins, rec=False)
    rec = build(monkeypatch, model, topo, lengths, codes, bins, rec=True, rec_units=rec_units)
    base = plain.lengths.copy()
    for k in range(3):
        for lf in (plain, rec):
            lf.lengths[:] = base * (1.0 + 0.01 * k)
            lf._dirty_model = True
        assert plain.get_log_likelihood() == rec.get_log_likelihood()
    assert rec.engine.stats()["last_launches"] < plain.engine.stats()["last_launches"]
    same_rows(plain, rec, topo, bins)
    rng = np.random.default_rng(5)
    for n in rng.choice(topo.edges, size=6, replace=False):
        v = float(rng.uniform(0.01, 0.5))
        for lf in (plain, rec):
            lf.set_param_rule("length", edge=topo.names[n], value=v)
        assert plain.get_log_likelihood() == rec.get_log_likelihood()
    same_rows(plain, rec, topo, bins)
    internal = [n for n in range(topo.n_nodes - 1) if not topo.is_tip[n]]
    for motif in (1, 2):
        fixed = np.full(topo.n_nodes, -1)
        fixed[internal[0]] = motif  # a node of the lowest level: inside the recomputed part
        fixed[internal[len(internal) // 2]] = 3 - motif
        for lf in (plain, rec):
            lf.engine.set_fixed_motifs(fixed)
            lf.engine.invalidate()
        assert plain.engine.evaluate() == rec.engine.evaluate()
        same_rows(plain, rec, topo, bins)


def test_rec_launch_with_a_polytomy_in_the_lower_levels(monkeypatch):
    from cogent3_b200.tree import Topology

    topo, lengths = Topology.from_newick(
        "(((a:0.1,b:0.2,c:0.15):0.1,(d:0.3,e:0.2):0.05):0.1,((f:0.1,g:0.1):0.2,(h:0.2,(i:0.1,j:0.3):0.1):0.1):0.2,"
        "((k:0.1,l:0.2):0.1,(m:0.2,(n:0.1,(o:0.2,p:0.1):0.1):0.1):0.2):0.1);"
    )
    rng = np.random.default_rng(3)
    codes = rng.integers(0, 4, size=(int(np.sum(topo.is_tip)), 60_000))
    plain = build(monkeypatch, "HKY85", topo, lengths, codes, 4, rec=False)
    rec = build(monkeypatch, "HKY85", topo, lengths, codes, 4, rec=True)
    base = plain.lengths.copy()
    for k in range(3):
        for lf in (plain, rec):
            lf.lengths[:] = base * (1.0 + 0.01 * k)
            lf._dirty_model = True
        assert plain.get_log_likelihood() == rec.get_log_likelihood()
    same_rows(plain, rec, topo, 4)
