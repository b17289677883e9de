"""Device-side pattern compression (c3_compress_alignment) against the golden dumps of the live
reference and against the host restatement (cogent3_b200.patterns, itself pinned to the
reference in test_oracle_golden / test_host_logic).  Integer work: everything bit-exact."""

import numpy as np
import pytest
from goldenutil import Golden, golden_cases

from cogent3_b200 import models as hostmodels
from cogent3_b200.engine import LikelihoodEngine
from cogent3_b200.likelihood import LikelihoodFunction
from cogent3_b200.patterns import compress_alignment
from cogent3_b200.synthetic import make_workload
from cogent3_b200.tree import Topology

pytestmark = pytest.mark.gpu
CASES = golden_cases()


def codes_in_node_order(g):
    by_tip = g.codes_by_tip()
    return np.stack([by_tip[g.topo.names[n]] for n in g.topo.tips])


def assert_same_tables(engine, host, topo):
    dev = engine.patterns
    np.testing.assert_array_equal(dev.rows(), host.rows())
    for n in range(topo.n_nodes):
        d, h = dev.nodes[n], host.nodes[n]
        np.testing.assert_array_equal(d.index, h.index)
        if h.is_tip:
            np.testing.assert_array_equal(d.table, h.table)
        else:
            np.testing.assert_array_equal(d.indexes, h.indexes)
            np.testing.assert_array_equal(d.uniq, h.uniq)
    np.testing.assert_array_equal(dev.root.counts, host.root.counts)


@pytest.mark.parametrize("case", CASES)
def test_device_compression_matches_reference_tables(case):
    g = Golden(case)
    eng = LikelihoodEngine.from_alignment(g.topo, codes_in_node_order(g), g.symbol_rows, n_bins=g.n_bins,
                                          keep_index=True)  # fmt: skip
    dev = eng.patterns
    for n in range(g.topo.n_nodes):
        np.testing.assert_array_equal(dev.nodes[n].index, g.node(n, "index"))
        np.testing.assert_array_equal(dev.nodes[n].counts, g.node(n, "counts"))
        if g.topo.is_tip[n]:
            np.testing.assert_array_equal(dev.nodes[n].table, g.node(n, "table"))
        else:
            np.testing.assert_array_equal(dev.nodes[n].indexes, g.node(n, "indexes"))
            np.testing.assert_array_equal(dev.nodes[n].uniq, g.node(n, "uniq"))
    eng.close()


@pytest.mark.parametrize(
    "n_tips,L,n_states,kind,ambiguity",
    [
        (24, 200_000, 4, "evolved", 0.0),
        (16, 120_000, 4, "dense", 0.0),
        (12, 50_000, 4, "evolved", 0.05),
        (8, 20_000, 61, "evolved", 0.0),
        (5, 1, 4, "dense", 0.0),
        (5, 0, 4, "dense", 0.0),
        (6, 33, 4, "dense", 0.0),
    ],
)
def test_device_compression_matches_host(n_tips, L, n_states, kind, ambiguity):
    topo, _lengths, codes = make_workload(n_tips, max(L, 1), n_states, kind=kind, seed=5)
    codes = codes[:, :L]
    rows = hostmodels.dna_symbol_rows()[1] if n_states == 4 else hostmodels.codon_symbol_rows()[1]
    if ambiguity:
        rng = np.random.default_rng(1)
        amb = rng.random(codes.shape) < ambiguity
        codes = codes.copy()
        codes[amb] = rng.integers(n_states, rows.shape[0], size=int(amb.sum()))
    host = compress_alignment(topo, codes.astype(np.int64), rows)
    eng = LikelihoodEngine.from_alignment(topo, codes, rows, keep_index=True)
    assert_same_tables(eng, host, topo)
    eng.close()


def test_device_compression_polytomies():
    topo, _ = Topology.from_newick("((a:0.1,b:0.2,c:0.1,d:0.3):0.1,e:0.2,(f:0.1,g:0.2):0.1,h:0.1,i:0.3);")
    rng = np.random.default_rng(3)
    codes = rng.integers(0, 4, size=(9, 5000))
    rows = hostmodels.dna_symbol_rows()[1]
    host = compress_alignment(topo, codes.astype(np.int64), rows)
    eng = LikelihoodEngine.from_alignment(topo, codes, rows, keep_index=True)
    assert_same_tables(eng, host, topo)
    eng.close()


def test_device_and_host_compression_give_identical_likelihoods():
    m = hostmodels.get_model("GTR")
    topo, lengths, codes = make_workload(20, 60_000, 4, kind="evolved", seed=17)
    lfs = []
    for how in ("device", "host"):
        lf = LikelihoodFunction(m, topo, bins=4)
        lf.set_alignment(codes, compress=how)
        lf.set_motif_probs_from_data()
        lf.lengths[:-1] = np.asarray(lengths)[:-1]
        lf.set_param_rule("rate_shape", value=0.7)
        lfs.append(lf)
    np.testing.assert_array_equal(lfs[0].mprobs, lfs[1].mprobs)
    assert lfs[0].get_log_likelihood() == lfs[1].get_log_likelihood()
    for b in range(4):
        site = lfs[0].get_full_length_likelihoods(bin=f"bin{b}")
        want = lfs[1].engine.get_lh(b)[lfs[1].patterns.root.index]
        np.testing.assert_array_equal(site, want)
        np.testing.assert_array_equal(lfs[1].get_full_length_likelihoods(bin=f"bin{b}"), want)
    np.testing.assert_array_equal(lfs[0].get_bin_probs(), lfs[1].get_bin_probs())
    # only the root's column numbering stays on the device by default
    with pytest.raises(Exception, match="no column numbering"):
        lfs[0].engine.fetch_column_index(0)


def test_symbol_code_out_of_range_is_an_error():
    topo, _ = Topology.from_newick("(a:0.1,b:0.2,c:0.1);")
    codes = np.array([[0, 1, 2], [0, 1, 200], [3, 3, 3]], np.uint8)
    with pytest.raises(Exception, match="out of range"):
        LikelihoodEngine.from_alignment(topo, codes, hostmodels.dna_symbol_rows()[1])
