"""CPU tests of the host side: pattern compression, host models (Q), the standalone
likelihood function's parameter plumbing -- with the oracle standing in for the GPU
engine (tests/oracle_engine.py)."""

import numpy as np
import pytest
from goldenutil import Golden, golden_cases
from lfutil import lf_from_golden
from oracle_engine import OracleEngine

from cogent3_b200 import models as hostmodels
from cogent3_b200 import patterns as pt
from cogent3_b200.tree import Topology, random_join_tree
from oracle import oracle

CASES = golden_cases()


@pytest.mark.parametrize("case", CASES)
def test_vectorised_compression_bit_exact_vs_reference(case):
    g = Golden(case)
    tree = pt.compress_alignment(g.topo, g.codes_by_tip(), g.symbol_rows)
    for n, nd in enumerate(tree.nodes):
        assert nd.counts.dtype == np.float64 and nd.index.dtype == np.int64
        assert np.array_equal(nd.counts, g.node(n, "counts"))
        assert np.array_equal(nd.index, g.node(n, "index"))
        assert np.array_equal(nd.ambig, g.node(n, "ambig"))
        if nd.is_tip:
            assert np.array_equal(nd.table, g.node(n, "table"))
        else:
            assert nd.indexes.dtype == np.int64 and nd.indexes.flags["C_CONTIGUOUS"]
            assert np.array_equal(nd.uniq, g.node(n, "uniq"))
            assert np.array_equal(nd.indexes, g.node(n, "indexes"))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_first_seen_unique_matches_dict_loop(seed):
    rng = np.random.default_rng(seed)
    for span in (3, 50, 10**6, 2**40):
        keys = rng.integers(0, span, size=2000)
        first, counts, index = pt.first_seen_unique(keys)
        u, c, i = oracle.indexed([int(k) for k in keys])
        assert [int(keys[f]) for f in first] == u
        assert list(counts) == c
        assert np.array_equal(index, i)


def test_first_seen_unique_empty_and_single():
    f, c, i = pt.first_seen_unique(np.zeros(0, np.int64))
    assert len(f) == len(c) == len(i) == 0
    f, c, i = pt.first_seen_unique(np.array([7]))
    assert list(f) == [0] and list(c) == [1.0] and list(i) == [0]


def test_key_overflow_folding():
    # three children with ~2^31 rows each would overflow a naive mixed-radix key
    a = np.array([0, 5, 5, 2**31 - 2, 0])
    key = pt._combine_keys([a, a[::-1].copy(), a], [2**31, 2**31, 2**31])
    first, _, index = pt.first_seen_unique(key)
    u, _, i = oracle.indexed(list(zip(a, a[::-1], a, strict=True)))
    assert np.array_equal(index, i) and len(first) == len(u)


@pytest.mark.parametrize("case", [c for c in CASES if not c.startswith("kat")])
def test_host_models_reproduce_reference_Q(case):
    g = Golden(case)
    model = hostmodels.get_model(g.meta["model"])
    assert list(model.alphabet) == g.alphabet
    assert np.array_equal(model.inst_mask, g.z["inst_mask"])
    for pname, mask in model.predicates.items():
        key = "mask_" + pname.replace("/", "_").replace(">", "to")
        assert np.array_equal(mask, g.z[key]), pname
    for n in g.topo.edges:
        params = dict(g.meta.get("params", {}))
        params.update(g.meta.get("edge_params", {}).get(g.topo.names[n], {}))
        Q = model.calcQ(g.mprobs, params)
        np.testing.assert_allclose(Q, g.Q[0, n], rtol=1e-14, atol=1e-16)


def test_gamma_rates_match_reference_distances():
    g = Golden("c2_gtr_gamma4_small")
    rates = hostmodels.gamma_rates(4, g.meta["rate_shape"], g.bprobs)
    n = 0
    np.testing.assert_allclose(g.lengths[n] * rates, g.distances[:, n], rtol=1e-15)


@pytest.mark.parametrize("case", CASES)
def test_likelihood_function_host_logic_via_oracle_engine(case):
    g = Golden(case)
    if g.case.startswith("c1_brca1") and "default" not in g.case:
        pytest.skip("one brca1 point is enough for the (slow) numpy engine")
    lf = lf_from_golden(g, engine_factory=OracleEngine)
    lnL = lf.get_log_likelihood()
    assert abs(lnL - g.lnL) <= 1e-10 * abs(g.lnL)
    for b in range(g.n_bins):
        np.testing.assert_allclose(lf.engine.get_lh(b), g.lh[b], rtol=0, atol=1e-13)


def test_motif_probs_from_data_match_reference():
    g = Golden("c2_gtr_gamma4_ambig")
    lf = lf_from_golden(g, engine_factory=OracleEngine)
    lf.set_motif_probs_from_data()
    np.testing.assert_allclose(lf.mprobs, g.mprobs, rtol=1e-14)


def test_newick_roundtrip_and_names():
    topo, lengths = Topology.from_newick("((a:0.1,b:0.2):0.05,c:0.3,(d:0.1,e:0.2):0.3);")
    assert topo.names == ["a", "b", "edge.0", "c", "d", "e", "edge.1", "root"]
    assert topo.children[-1] == [2, 3, 6]
    topo2, l2 = Topology.from_newick(topo.to_newick(lengths))
    assert topo2.names == topo.names and topo2.children == topo.children
    np.testing.assert_array_equal(l2[:-1], lengths[:-1])


def test_random_join_tree_shape():
    topo, lengths = random_join_tree(64, np.random.default_rng(1))
    assert len(topo.tips) == 64 and topo.n_nodes == 2 * 64 - 2
    assert len(topo.children[topo.root]) == 3
    assert np.all((lengths[:-1] >= 0.02) & (lengths[:-1] <= 0.3))
