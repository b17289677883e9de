"""Pin the oracle (oracle/oracle.py) to golden vectors from the live reference."""

import numpy as np
import pytest
from goldenutil import Golden, golden_cases

from oracle import oracle

CASES = golden_cases()


def build_oracle_lht(g, loop=False):
    leaves = {}
    for n in g.topo.tips:
        name = g.topo.names[n]
        leaves[name] = oracle.make_leaf(g.codes_by_tip()[name], g.symbol_rows, name, loop=loop)
    return oracle.build_lht(g.topo.children, g.topo.names, leaves, loop=loop)


@pytest.mark.parametrize("case", CASES)
def test_compression_bit_exact(case):
    g = Golden(case)
    loop = g.codes.shape[1] <= 500  # dict-loop restatement for the small cases too
    for use_loop in {False, loop}:
        lht = build_oracle_lht(g, loop=use_loop)
        for n, nd in enumerate(lht):
            assert np.array_equal(nd.counts, g.node(n, "counts")), (case, n)
            assert np.array_equal(nd.index, g.node(n, "index")), (case, n)
            assert np.array_equal(nd.ambig, g.node(n, "ambig")), (case, n)
            if nd.is_tip:
                assert np.array_equal(nd.input_likelihoods, g.node(n, "table")), (case, n)
            else:
                assert np.array_equal(nd.uniq, g.node(n, "uniq")), (case, n)
                assert np.array_equal(nd.indexes, g.node(n, "indexes")), (case, n)


@pytest.mark.parametrize("case", CASES)
def test_exponentiators_match_reference_psubs(case):
    g = Golden(case)
    expm = g.meta["expm"]
    for b in range(g.n_bins):
        for n in g.topo.edges:
            ex = oracle.make_exponentiator(g.Q[b, n], expm)
            P = ex(g.distances[b, n])
            np.testing.assert_allclose(P, g.psubs[b, n], rtol=0, atol=1e-15)


@pytest.mark.parametrize("case", CASES)
def test_pruning_and_lnL(case):
    g = Golden(case)
    lht = build_oracle_lht(g)
    psubs = [[g.psubs[b, n] for n in range(g.topo.n_nodes)] for b in range(g.n_bins)]
    lnL, lhs, _ = oracle.log_likelihood(
        g.topo.children, lht, psubs, g.mprobs, g.bprobs, return_parts=True
    )
    for b in range(g.n_bins):
        np.testing.assert_allclose(lhs[b], g.lh[b], rtol=0, atol=1e-15)
    assert abs(lnL - g.lnL) <= 1e-12 * abs(g.lnL)
    if "pinned_lnL" in g.meta:
        # constants hard-coded in the reference's own test-suite
        np.testing.assert_allclose(lnL, g.meta["pinned_lnL"], rtol=1e-7)


def test_pade_matches_eigen():
    g = Golden("c2_gtr_gamma4_small")
    Q = g.Q[0, 0]
    for t in (0.0, 1e-6, 0.01, 0.3, 2.0, 9.5):
        Pe = oracle.checked_exponentiator(Q)(t)
        Pp = oracle.PadeExponentiator(Q)(t)
        np.testing.assert_allclose(Pe, Pp, atol=1e-12)


def test_gamma_rates_weighted_mean_one():
    r = oracle.gamma_rates([0.25] * 4, 0.5)
    assert abs(np.sum(r * 0.25) - 1.0) < 1e-14
    assert np.all(np.diff(r) > 0)


def test_indexed_doc_example():
    u, c, i = oracle.indexed(["a", "b", "c", "a", "a"])
    assert (u, c, list(i)) == (["a", "b", "c"], [3, 1, 1], [0, 1, 2, 0, 0])
