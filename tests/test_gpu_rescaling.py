"""Per-pattern underflow rescaling (c3_set_rescaling).

The reference has no rescaling on this path (evolve/likelihood_tree.py:10,151-153: plain FP64
products, warnings ignored), so there are two things to prove:

* wherever the reference does NOT underflow, rescaling changes nothing: per-pattern likelihoods are
  bit-identical (powers of two are exact), lnL agrees to the last few ulps;
* where it does (a few hundred divergent taxa: lh = 0, lnL = -inf), the rescaled engine returns the
  value an extended-range evaluation of the same products gives.  The checker is the oracle's
  pruning restated in numpy ``longdouble`` (x87 80-bit: 15-bit exponent) over the oracle's own
  tables and P matrices.
"""

import numpy as np
import pytest
from goldenutil import Golden
from lfutil import lf_from_golden
from oracle_engine import OracleEngine

from cogent3_b200 import models as hostmodels
from cogent3_b200.likelihood import LikelihoodFunction
from cogent3_b200.synthetic import make_workload

pytestmark = pytest.mark.gpu


def lnL_extended(engine):
    """the oracle's pruning (oracle.partial_likelihoods / log_likelihood) in longdouble"""
    ld = np.longdouble
    topo = engine.topo
    kids_of = [list(topo.children[n]) for n in range(topo.n_nodes)]
    lht, root = engine.lht, topo.n_nodes - 1
    lhs = []
    for psub_b in engine._psubs():
        plh = [None] * topo.n_nodes
        for n, kids in enumerate(kids_of):
            if not kids:
                plh[n] = lht[n].input_likelihoods.astype(ld)
                continue
            out = None
            for c, k in enumerate(kids):
                g = (plh[k] @ psub_b[k].astype(ld).T)[lht[n].indexes[c]]
                out = g if out is None else out * g
            plh[n] = out
        lhs.append(plh[root] @ engine.pi.astype(ld))
    mix = lhs[0] if len(lhs) == 1 else sum(ld(b) * lh for b, lh in zip(engine.bprobs, lhs, strict=True))
    return float(np.sum(np.log(mix) * lht[root].counts.astype(ld))), lhs


@pytest.mark.parametrize("case", ["c1_brca1_hky85_default", "c2_gtr_gamma4_mid", "c4_cnfgtr_small", "c3_gn_pade_small"])
@pytest.mark.parametrize("tips", [2, 4])
def test_rescaling_is_exact_where_the_reference_does_not_underflow(case, tips):
    g = Golden(case)
    lf = lf_from_golden(g)
    plain = lf.get_log_likelihood()
    assert lf.engine.rescaling() == (False, 0)
    assert not lf.engine.get_root_exponents().any()
    lh_plain = [lf.engine.get_lh(b) for b in range(g.n_bins)]

    lf.engine.set_rescaling("on", tip_threshold=tips)
    scaled = lf.engine.evaluate()
    active, n_scaled = lf.engine.rescaling()
    assert active
    if tips == 2 or g.topo.n_nodes > 40:  # (the smallest golden trees have no subtree of 4 tips below the root)
        assert n_scaled > 0
        assert lf.engine.get_root_exponents().min() < 0  # rows really were rescaled
    for b in range(g.n_bins):
        assert np.array_equal(lf.engine.get_lh(b), lh_plain[b])  # bit for bit
    assert abs(scaled - plain) <= 1e-13 * abs(plain)
    assert abs(scaled - g.lnL) <= 1e-9 * abs(g.lnL)

    # partial updates keep the exponents of untouched subtrees: same value as a full sweep
    rng = np.random.default_rng(5)
    for n in rng.choice(g.topo.edges, size=3, replace=False):
        lf.set_param_rule("length", edge=g.topo.names[n], value=float(rng.uniform(0.01, 0.5)))
        part = lf.get_log_likelihood()
        lh_part = lf.engine.get_lh(0)
        lf.engine.invalidate()
        assert lf.engine.evaluate() == part
        assert np.array_equal(lf.engine.get_lh(0), lh_part)

    # and back: rescaling off reproduces the plain products
    lf.engine.set_rescaling("off")
    off = lf.engine.evaluate()
    assert lf.engine.rescaling() == (False, 0)
    lf2 = lf_from_golden(g)
    for n in g.topo.edges:
        lf2.set_param_rule("length", edge=g.topo.names[n], value=float(lf.lengths[n]))
    assert off == lf2.get_log_likelihood()


def _deep(model, n_tips, L, bins, factory, seed=11):
    m = hostmodels.get_model(model)
    topo, lengths, codes = make_workload(n_tips, L, m.n_states, kind="dense", seed=seed)
    lf = LikelihoodFunction(m, topo, bins=bins, engine_factory=factory)
    lf.set_alignment(codes.astype(np.int64), np.eye(m.n_states))
    lf.set_motif_probs_from_data()
    lf.lengths[:-1] = np.asarray(lengths)[:-1] * 4.0  # long branches: ~0.6 decades per tip
    for k, p in enumerate(m.parameter_order):
        lf.set_param_rule(p, value=0.5 + 0.37 * k)
    if bins > 1:
        lf.set_param_rule("rate_shape", value=0.6)
    return lf


@pytest.mark.parametrize("model,n_tips,L,bins", [("HKY85", 700, 600, 4), ("GTR", 640, 400, 1)])
def test_rescaling_rescues_an_underflowing_tree(model, n_tips, L, bins):
    ref = _deep(model, n_tips, L, bins, OracleEngine)
    with np.errstate(all="ignore"):
        assert not np.isfinite(ref.get_log_likelihood())  # the reference's answer: log(0)
    want, _ = lnL_extended(ref.engine)
    assert np.isfinite(want)

    # "off": the reference's arithmetic, -inf included
    lf = _deep(model, n_tips, L, bins, None)
    lf.engine.set_rescaling("off")
    assert not np.isfinite(lf.get_log_likelihood())
    assert lf.engine.rescaling() == (False, 0)

    # "auto" (the default): the non-finite result switches rescaling on and the call returns the rescued value
    lf = _deep(model, n_tips, L, bins, None)
    got = lf.get_log_likelihood()
    active, n_scaled = lf.engine.rescaling()
    assert active and n_scaled >= n_tips // 40
    assert abs(got - want) <= 1e-9 * abs(want), (got, want)


@pytest.mark.parametrize("model,n_tips,L,bins", [("HKY85", 700, 600, 4)])
def test_rescaled_per_pattern_values(model, n_tips, L, bins):
    ref = _deep(model, n_tips, L, bins, OracleEngine)
    with np.errstate(all="ignore"):
        ref.get_log_likelihood()
    _, lhs = lnL_extended(ref.engine)
    lf = _deep(model, n_tips, L, bins, None)
    lf.get_log_likelihood()
    assert lf.engine.rescaling()[0]
    # a partial update in rescued mode agrees with a full sweep
    name = lf.topo.names[int(lf.topo.edges[3])]
    lf.set_param_rule("length", edge=name, value=0.33)
    ref.set_param_rule("length", edge=name, value=0.33)
    part = lf.get_log_likelihood()
    lf.engine.invalidate()
    assert lf.engine.evaluate() == part
    with np.errstate(all="ignore"):
        ref.get_log_likelihood()
    want, lhs = lnL_extended(ref.engine)
    assert abs(part - want) <= 1e-9 * abs(want)
    # per-pattern log-likelihoods: log(lh_scaled) + E ln 2  vs  log of the extended-range value
    lf.engine.set_rescaling("on")
    E = lf.engine.get_root_exponents()
    scaled = lf.engine.get_scaled_lh()
    # The K bins of a pattern share ONE exponent (their mixture is what the likelihood needs), so a
    # bin that is > ~1e300 times less likely than the pattern's best bin reads 0 -- here the slowest
    # gamma class on i.i.d. columns.  Every bin within 500 nats of the best one must match.
    exp = np.stack([np.log(lh).astype(float) for lh in lhs])
    with np.errstate(divide="ignore"):
        got = np.log(scaled) + E * np.log(2.0)
    near = exp - exp.max(axis=0) > -500.0
    assert near.any(axis=0).all() and near.sum() > near.shape[1]
    assert np.max(np.abs(got[near] - exp[near])) <= 1e-9 * np.max(np.abs(exp))
    assert np.all(got[~near] <= exp[~near] + 1.0)  # 0, or a denormal with a few bits left
    # posterior bin probabilities (likelihood_calculation.py:247-257) from the scaled values
    bp = ref.engine.bprobs[:, None]
    post = bp * scaled / (bp * scaled).sum(axis=0)
    want_post = np.stack([(b * lh) for b, lh in zip(ref.engine.bprobs, lhs, strict=True)])
    want_post = (want_post / want_post.sum(axis=0)).astype(float)
    assert np.max(np.abs(post - want_post)) <= 1e-9


@pytest.mark.gpu
def test_non_finite_values_behave_like_the_reference():
    """NaN propagates (numpy.maximum(P, 0) keeps NaN, maths/matrix_exponentiation.py:40; the optimiser treats a
    NaN / -inf lnL as out of bounds, maths/optimisers.py:116-118) and does NOT switch per-pattern rescaling on;
    a log-likelihood of -inf that rescaling cannot help (a pattern whose likelihood is truly 0: an impossible
    fixed motif) does not leave rescaling switched on either -- the engine goes back to the plain products"""
    from cogent3_b200 import models as hostmodels
    from cogent3_b200.likelihood import LikelihoodFunction
    from cogent3_b200.synthetic import make_workload

    m = hostmodels.get_model("HKY85")
    topo, lengths, codes = make_workload(10, 2000, 4, kind="evolved", seed=9)
    lf = LikelihoodFunction(m, topo, bins=1)
    lf.set_alignment(codes.astype(np.int64), np.eye(4))
    lf.set_motif_probs_from_data()
    lf.lengths[:-1] = np.asarray(lengths)[:-1]
    good = lf.get_log_likelihood()
    assert np.isfinite(good) and lf.engine.rescaling() == (False, 0)
    # a NaN distance: NaN out, rescaling stays off, the next good point is evaluated normally
    d = np.nan_to_num(lf.lengths.copy())[:, None].copy()
    bad = d.copy()
    bad[3] = np.nan
    assert np.isnan(lf.engine.update(bad))
    assert lf.engine.rescaling() == (False, 0)
    assert lf.engine.update(d) == good
    # a zero in P (supplied matrix) under the observed data: likelihood exactly 0 for some patterns -> -inf,
    # which rescaling cannot change; afterwards the engine is back on the plain path and still right
    P = np.eye(4)  # no substitutions at all on a tip branch with differing states below/above
    tip = int(topo.tips[0])
    lf.engine.set_psub(tip, 0, P)
    v = lf.engine.evaluate()
    if np.isinf(v) and v < 0:
        assert lf.engine.rescaling()[0] is False
        v2 = lf.engine.evaluate()
        assert np.isinf(v2) and v2 < 0
        assert lf.engine.rescaling()[0] is False
