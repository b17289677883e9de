"""GPU parity: the CUDA path (through the C ABI) against the golden vectors of the
live reference and against the CPU oracle on the same inputs.

Tolerances are the north-star's: pattern tables bit-exact (checked on the CPU side),
total lnL <= 1e-9 relative, per-pattern likelihoods <= 1e-12 absolute."""

import numpy as np
import pytest
from goldenutil import Golden, golden_cases
from lfutil import lf_from_golden
from oracle_engine import OracleEngine

from cogent3_b200 import models as hostmodels
from cogent3_b200.likelihood import LikelihoodFunction
from cogent3_b200.synthetic import make_workload

pytestmark = pytest.mark.gpu
CASES = golden_cases()
REL_LNL = 1e-9
ABS_LH = 1e-12


@pytest.mark.parametrize("case", CASES)
def test_golden_lnL_lh_psubs(case):
    g = Golden(case)
    lf = lf_from_golden(g)
    lnL = lf.get_log_likelihood()
    assert abs(lnL - g.lnL) <= REL_LNL * abs(g.lnL), (lnL, g.lnL)
    for b in range(g.n_bins):
        lh = lf.engine.get_lh(b)
        assert np.max(np.abs(lh - g.lh[b])) <= ABS_LH
        for n in g.topo.edges:
            P = lf.engine.get_psub(n, b)
            np.testing.assert_allclose(P, g.psubs[b, n], rtol=0, atol=1e-13)
    if "pinned_lnL" in g.meta:
        np.testing.assert_allclose(lnL, g.meta["pinned_lnL"], rtol=1e-7)
    st = lf.engine.stats()
    assert st["last_launches"] >= 2  # expm + >=1 level launch (+ reduction unless the small-levels launch did it): the CUDA path ran


@pytest.mark.parametrize("case", ["c2_gtr_gamma4_small", "c4_cnfgtr_small", "c3_gn_pade_small"])
def test_partials_match_oracle(case):
    g = Golden(case)
    lf = lf_from_golden(g)
    ref = lf_from_golden(g, engine_factory=OracleEngine)
    lf.get_log_likelihood()
    ref.get_log_likelihood()
    for n in range(g.topo.n_nodes):
        for b in range(g.n_bins):
            got = lf.engine.get_partials(n, b)
            want = ref.engine.get_partials(n, b)
            # entries are compared relative to their row's largest value: tiny P entries
            # carry ~1e-16 absolute error from the different summation order
            # (eigen-path P entries of ~1e-7 are only good to ~1e-16 ABSOLUTE in either
            # implementation, so rows reached only through such entries differ ~1e-9 relative)
            scale = np.maximum(want.max(axis=1, keepdims=True), 1e-300)
            assert np.max(np.abs(got - want)) <= 1e-12
            assert np.max(np.abs(got - want) / scale) <= 1e-8


@pytest.mark.parametrize("case", ["c1_brca1_hky85_default", "c2_gtr_gamma4_mid", "c4_cnfgtr_small"])
def test_partial_update_equals_full_recompute(case):
    g = Golden(case)
    lf = lf_from_golden(g)
    ref = lf_from_golden(g, engine_factory=OracleEngine)
    lf.get_log_likelihood()
    rng = np.random.default_rng(3)
    for n in rng.choice(g.topo.edges, size=4, replace=False):
        name = g.topo.names[n]
        v = float(rng.uniform(0.01, 0.5))
        lf.set_param_rule("length", edge=name, value=v)
        ref.set_param_rule("length", edge=name, value=v)
        got = lf.get_log_likelihood()
        st = lf.engine.stats()
        assert st["last_expm_jobs"] == g.n_bins  # only that edge was re-exponentiated
        # only its ancestors re-pruned -- plus, for M != 4 (edge-major: a node's rows include the P
        # of the branch above it), the internal node below the changed branch
        below = 1 if (g.n_states != 4 and not g.topo.is_tip[n]) else 0
        assert st["last_nodes"] == len(g.topo.ancestors(n)) + below
        want = ref.get_log_likelihood()
        assert abs(got - want) <= REL_LNL * abs(want)
        lf.engine.invalidate()
        full = lf.get_log_likelihood()
        assert got == full  # dirty-path result is bit-identical to a full sweep


def test_rate_param_and_mprobs_changes():
    g = Golden("c2_gtr_gamma4_small")
    lf = lf_from_golden(g)
    ref = lf_from_golden(g, engine_factory=OracleEngine)
    for setter in (
        lambda f: f.set_param_rule("A/G", value=2.2),
        lambda f: f.set_param_rule("rate_shape", value=1.3),
        lambda f: f.set_motif_probs(np.array([0.1, 0.2, 0.3, 0.4])),
        lambda f: f.set_param_rule("bprobs", value=np.array([0.4, 0.3, 0.2, 0.1])),
    ):
        setter(lf)
        setter(ref)
        got, want = lf.get_log_likelihood(), ref.get_log_likelihood()
        assert abs(got - want) <= REL_LNL * abs(want)


def test_fixed_motif_masks_like_reference():
    g = Golden("c2_gtr_gamma4_small")
    lf = lf_from_golden(g)
    ref = lf_from_golden(g, engine_factory=OracleEngine)
    lf.get_log_likelihood()
    ref.get_log_likelihood()
    node = g.topo.n_nodes - 2 if not g.topo.is_tip[g.topo.n_nodes - 2] else g.topo.root
    for motif in range(4):
        fixed = np.full(g.topo.n_nodes, -1)
        fixed[node] = motif
        lf.engine.set_fixed_motifs(fixed)
        ref.engine.set_fixed_motifs(list(fixed))
        lf.engine.evaluate()
        ref.engine.evaluate()
        for b in range(g.n_bins):
            assert np.max(np.abs(lf.engine.get_lh(b) - ref.engine.get_lh(b))) <= ABS_LH


def test_direct_psub_edges():
    """P supplied by the host (discrete-time edges) bypasses expm"""
    g = Golden("c5_hky85_gamma4_small")
    lf = lf_from_golden(g)
    ref = lf_from_golden(g, engine_factory=OracleEngine)
    lf.get_log_likelihood()
    ref.get_log_likelihood()
    rng = np.random.default_rng(0)
    P = rng.random((4, 4))
    P /= P.sum(axis=1, keepdims=True)
    for b in range(g.n_bins):
        lf.engine.set_psub(2, b, P)
        ref.engine.set_psub(2, b, P)
    got, want = lf.engine.evaluate(), ref.engine.evaluate()
    assert abs(got - want) <= REL_LNL * abs(want)
    np.testing.assert_array_equal(lf.engine.get_psub(2, 1), P)


@pytest.mark.parametrize("force_dmma", [False, True])
def test_four_state_through_both_kernels(force_dmma, monkeypatch):
    """the 4-state register kernel and the generic DMMA kernel (MP=8) agree"""
    if force_dmma:
        monkeypatch.setenv("C3_FORCE_DMMA", "1")
    g = Golden("c2_gtr_gamma4_mid")
    lf = lf_from_golden(g)
    lnL = lf.get_log_likelihood()
    assert lf.engine.stats()["padded_states"] == (8 if force_dmma else 4)
    assert abs(lnL - g.lnL) <= REL_LNL * abs(g.lnL)


@pytest.mark.parametrize(
    "model,n_tips,L,bins,kind",
    [
        ("GTR", 24, 50_000, 4, "evolved"),
        ("HKY85", 16, 30_000, 1, "dense"),
        ("GN", 12, 20_000, 1, "evolved"),
        ("CNFGTR", 10, 4_000, 1, "evolved"),
        ("CNFGTR", 8, 2_000, 2, "dense"),
    ],
)
def test_synthetic_against_oracle(model, n_tips, L, bins, kind):
    m = hostmodels.get_model(model)
    topo, lengths, codes = make_workload(n_tips, L, m.n_states, kind=kind, seed=99)
    rows = np.eye(m.n_states)
    kw = {"expm": "pade"} if model == "GN" else {}
    lfs = []
    for factory in (None, OracleEngine):
        lf = LikelihoodFunction(m, topo, bins=bins, engine_factory=factory, **kw)
        lf.set_alignment(codes.astype(np.int64), rows)
        lf.set_motif_probs_from_data()
        lf.lengths[:-1] = lengths[:-1]
        for k, p in enumerate(m.parameter_order):
            lf.set_param_rule(p, value=0.5 + 0.37 * k)
        if bins > 1:
            lf.set_param_rule("rate_shape", value=0.6)
        lfs.append(lf)
    got, want = lfs[0].get_log_likelihood(), lfs[1].get_log_likelihood()
    assert abs(got - want) <= REL_LNL * abs(want), (got, want)
    for b in range(bins):
        assert np.max(np.abs(lfs[0].engine.get_lh(b) - lfs[1].engine.get_lh(b))) <= ABS_LH


@pytest.mark.parametrize(
    "model,n_tips,L,bins,kind,rooted",
    [
        ("GTR", 24, 50_000, 4, "evolved", False),
        ("HKY85", 16, 30_000, 1, "dense", False),
        ("HKY85", 20, 40_000, 2, "evolved", False),
        ("HKY85", 9, 60_000, 4, "dense", True),
    ],
)
def test_four_state_kernel_families_bit_identical(model, n_tips, L, bins, kind, rooted, monkeypatch):
    """regpipe2 (default), regpipe, dataflow, cp.async ring and thread-per-unit kernels do the
    same arithmetic in the same order: every partial row, lh and lnL are bit-identical; so is
    a dirty-path update inside the dataflow launch"""
    from cogent3_b200.tree import Topology

    m = hostmodels.get_model(model)
    topo, lengths, codes = make_workload(n_tips, L, m.n_states, kind=kind, seed=7)
    if rooted:
        # bifurcating root: the root itself is computed inside the dataflow launch
        topo, lengths = Topology.from_newick("(((a:0.1,b:0.2):0.1,(c:0.3,d:0.2):0.05):0.1,((e:0.1,f:0.1):0.2,(g:0.2,(h:0.1,i:0.3):0.1):0.1):0.2);")
        codes = codes[: int(np.sum(topo.is_tip))]

    def build(env):
        for k in ("C3_FLOW", "C3_RING", "C3_NO_STRIP", "C3_NO_SMALL", "C3_REG1", "C3_REG2_U2", "C3_ROOT_U", "C3_DYN"):
            monkeypatch.delenv(k, raising=False)
        for k in env:
            name, _, value = k.partition("=")
            monkeypatch.setenv(name, value or "1")
        lf = LikelihoodFunction(m, topo, bins=bins)
        lf.set_alignment(codes.astype(np.int64), np.eye(4))
        lf.set_motif_probs_from_data()
        lf.lengths[:-1] = np.asarray(lengths)[:-1]
        for k, p in enumerate(m.parameter_order):
            lf.set_param_rule(p, value=0.5 + 0.37 * k)
        if bins > 1:
            lf.set_param_rule("rate_shape", value=0.6)
        return lf

    variants = [(), ("C3_REG1",), ("C3_FLOW",), ("C3_RING",), ("C3_NO_STRIP",), ("C3_NO_SMALL",), ("C3_FLOW", "C3_NO_SMALL"),
                ("C3_REG2_U2",), ("C3_ROOT_U=2",), ("C3_DYN",)]
    lfs = [build(v) for v in variants]
    vals = [lf.get_log_likelihood() for lf in lfs]
    assert len(set(vals)) == 1, vals
    for n in range(topo.n_nodes):
        if topo.is_tip[n]:
            continue
        for b in range(bins):
            want = lfs[0].engine.get_partials(n, b)
            for lf in lfs[1:]:
                np.testing.assert_array_equal(lf.engine.get_partials(n, b), want)
    # dirty path (one branch length at a time) through each family
    rng = np.random.default_rng(11)
    for n in rng.choice(topo.edges, size=3, replace=False):
        v = float(rng.uniform(0.01, 0.5))
        got = []
        for lf in lfs:
            lf.set_param_rule("length", edge=topo.names[n], value=v)
            got.append(lf.get_log_likelihood())
        assert len(set(got)) == 1, got
        lfs[0].engine.invalidate()
        assert lfs[0].get_log_likelihood() == got[0]


@pytest.mark.parametrize("bins,n_tips,L", [(4, 24, 60_000), (2, 14, 40_000), (4, 9, 300)])
def test_lean_root_is_bit_identical_and_lh_comes_back_on_demand(bins, n_tips, L, monkeypatch):
    """lean root (default): the regpipe2 root launch writes the class mixture mix[p] instead of lh[p][K] and the
    reduction reads mix + counts.  lnL is bit-identical to C3_LEAN_ROOT=0; an accessor gets exactly the lh of the
    full launch (the root launch runs again with the lh output); the next evaluations keep lh materialised, later
    ones go lean again; a change of the bin probabilities alone re-mixes; graph replays included."""
    m = hostmodels.get_model("GTR")
    topo, lengths, codes = make_workload(n_tips, L, 4, kind="evolved", seed=21)

    def build(lean):
        monkeypatch.setenv("C3_LEAN_ROOT", "1" if lean else "0")
        monkeypatch.setenv("C3_GRAPH_AFTER", "3")
        lf = LikelihoodFunction(m, topo, bins=bins)
        lf.set_alignment(codes.astype(np.int64), np.eye(4))
        lf.set_motif_probs_from_data()
        lf.lengths[:-1] = np.asarray(lengths)[:-1]
        for k, p in enumerate(m.parameter_order):
            lf.set_param_rule(p, value=0.5 + 0.37 * k)
        lf.set_param_rule("rate_shape", value=0.6)
        return lf

    lean, full = build(True), build(False)
    rng = np.random.default_rng(5)
    edges = list(topo.edges)
    for step in range(40):
        if step % 3 == 0:  # every branch length
            v = rng.uniform(0.01, 0.4, size=len(edges))
            for lf in (lean, full):
                for n, x in zip(edges, v, strict=True):
                    lf.set_param_rule("length", edge=topo.names[n], value=float(x))
        else:  # one branch
            n = edges[int(rng.integers(len(edges)))]
            v = float(rng.uniform(0.01, 0.5))
            for lf in (lean, full):
                lf.set_param_rule("length", edge=topo.names[n], value=v)
        a, b = lean.get_log_likelihood(), full.get_log_likelihood()
        assert a == b, (step, a, b)
        if step % 7 == 6:
            # the bin probabilities ALONE (engine level: no branch, no rate class moves): only the mixture changes
            bp = rng.dirichlet(np.ones(bins) * 5.0)
            got = []
            for lf in (lean, full):
                lf.engine.set_bin_probs(bp)
                got.append(lf.engine.evaluate())
            assert got[0] == got[1], (step, got)
        if step in (4, 5, 20, 39):
            for bin_ in range(bins):
                np.testing.assert_array_equal(lean.engine.get_lh(bin_), full.engine.get_lh(bin_))
            assert lean.engine.evaluate() == full.engine.evaluate()  # nothing pending was committed or lost
    n_lean = lean.engine.stats()["lean_evaluations"]
    assert full.engine.stats()["lean_evaluations"] == 0
    if L >= 10_000:
        # most evaluations were lean; the few after an lh request (and the request's own root launch) were not
        assert 0.5 * lean.engine.stats()["evaluations"] < n_lean < lean.engine.stats()["evaluations"]
    else:
        assert n_lean == 0  # (small problems reduce inside the small-levels launch and always write lh)


@pytest.mark.parametrize("env", [{"C3_EDGE_MAJOR": "0"}, {"C3_EDGE_MAJOR": "0", "C3_DMMA_REG": "0"}])
def test_dmma_parent_major_kernels_match(env, monkeypatch):
    """the parent-major DMMA kernels (register pipeline; first shared-memory-tile version) give the
    same numbers as the default edge-major kernel, to the tolerances (the k order of the first
    differs)"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for case in ("c4_cnfgtr_mid", "c4_cnfgtr_small", "kat_codon", "kat_dinucleotide", "kat_protein"):
        g = Golden(case)
        lf = lf_from_golden(g)
        lnL = lf.get_log_likelihood()
        assert abs(lnL - g.lnL) <= REL_LNL * abs(g.lnL)
        assert np.max(np.abs(lf.engine.get_lh(0) - g.lh[0])) <= ABS_LH


@pytest.mark.parametrize("force_dmma", [False, True])
def test_polytomy_and_many_bins(force_dmma, monkeypatch):
    """5-way polytomy at the root, K=3 (generic-K kernel path; with the DMMA kernels: more
    children than the register-pipeline kernel takes -> first version for that node)"""
    from cogent3_b200.tree import Topology

    if force_dmma:
        monkeypatch.setenv("C3_FORCE_DMMA", "1")

    topo, _ = Topology.from_newick("(a:0.1,b:0.2,c:0.15,(d:0.1,e:0.3):0.05,f:0.2);")
    rng = np.random.default_rng(5)
    codes = rng.integers(0, 4, size=(6, 700))
    lfs = []
    for factory in (None, OracleEngine):
        lf = LikelihoodFunction("HKY85", topo, bins=3, engine_factory=factory)
        lf.set_alignment(codes, np.eye(4))
        lf.lengths[:-1] = [0.1, 0.2, 0.15, 0.1, 0.3, 0.05, 0.2]
        lf.set_param_rule("kappa", value=2.5)
        lf.set_param_rule("rate_shape", value=0.8)
        lfs.append(lf)
    got, want = lfs[0].get_log_likelihood(), lfs[1].get_log_likelihood()
    assert abs(got - want) <= REL_LNL * abs(want)


def test_empty_and_single_column_alignments():
    from cogent3_b200.tree import Topology

    topo, _ = Topology.from_newick("((a:0.1,b:0.2):0.1,c:0.3,d:0.2);")
    for L in (0, 1):
        codes = np.zeros((4, L), np.int64)
        lfs = []
        for factory in (None, OracleEngine):
            lf = LikelihoodFunction("HKY85", topo, engine_factory=factory)
            lf.set_alignment(codes, np.eye(4))
            lf.lengths[:-1] = 0.2
            lfs.append(lf)
        got, want = lfs[0].get_log_likelihood(), lfs[1].get_log_likelihood()
        assert got == pytest.approx(want, rel=1e-12, abs=1e-15)


def test_determinism_repeated_and_recreated():
    g = Golden("c2_gtr_gamma4_mid")
    vals = set()
    for _ in range(3):
        lf = lf_from_golden(g)
        vals.add(lf.get_log_likelihood())
        lf.engine.invalidate()
        vals.add(lf.get_log_likelihood())
    assert len(vals) == 1


def test_optimise_improves_and_matches_oracle_path():
    g = Golden("c2_gtr_gamma4_small")
    lf = lf_from_golden(g)
    before = lf.get_log_likelihood()
    after = lf.optimise(max_evaluations=400)
    assert after > before
    # the optimum found on the GPU evaluates to the same lnL on the oracle
    ref = lf_from_golden(g, engine_factory=OracleEngine)
    ref.lengths[:] = lf.lengths
    ref.params = dict(lf.params)
    ref.rate_shape = lf.rate_shape
    ref._dirty_q = ref._dirty_model = True
    want = ref.get_log_likelihood()
    assert abs(after - want) <= REL_LNL * abs(want)


def test_evaluate_many_matches_individual_evaluations():
    """c3_evaluate_many: several independent engines in flight at once give exactly the numbers
    of one-at-a-time evaluation, for full sweeps and for dirty-path updates"""
    from cogent3_b200.likelihood import evaluate_many

    cases = ["c1_brca1_hky85_default", "c2_gtr_gamma4_small", "c5_hky85_gamma4_small", "c4_cnfgtr_small",
             "c3_gn_pade_small", "kat_protein", "c2_gtr_gamma4_mid", "c1_brca1_hky85_point2"]  # fmt: skip
    goldens = [Golden(c) for c in cases]
    many = [lf_from_golden(g) for g in goldens]
    single = [lf_from_golden(g) for g in goldens]
    got = evaluate_many(many)
    for g, v, lf in zip(goldens, got, single, strict=True):
        assert v == lf.get_log_likelihood()
        assert abs(v - g.lnL) <= REL_LNL * abs(g.lnL)
    rng = np.random.default_rng(2)
    for _ in range(3):
        for g, a, b in zip(goldens, many, single, strict=True):
            n = int(rng.choice(g.topo.edges))
            v = float(rng.uniform(0.01, 0.4))
            for lf in (a, b):
                lf.set_param_rule("length", edge=g.topo.names[n], value=v)
        got = evaluate_many(many)
        for v, lf in zip(got, single, strict=True):
            assert v == lf.get_log_likelihood()
    # nothing changed: served from the cached values, no launch
    again = evaluate_many(many)
    np.testing.assert_array_equal(again, got)
    assert all(lf.engine.stats()["last_launches"] == 0 for lf in many)
    assert len(evaluate_many([])) == 0


def test_fused_allreduce_between_engines_in_one_process():
    """c3_comm_*: column shards in separate engines; the reduction kernel's last CTA exchanges the
    shard values through the peers' mailboxes and every engine returns the TOTAL (summed in rank
    order).  Here the "ranks" are three engines of one process (mailboxes shared by device
    pointer), evaluated together with c3_evaluate_many; across processes the mailboxes are
    mapped through CUDA IPC handles (tools/check_fused_allreduce.py, run under torchrun)."""
    from cogent3_b200.engine import evaluate_many as engine_many
    from cogent3_b200.sharded import column_block

    m = hostmodels.get_model("HKY85")
    topo, lengths, codes = make_workload(14, 30_000, 4, kind="evolved", seed=23)
    world = 3

    def make(cols):
        lf = LikelihoodFunction(m, topo, bins=4)
        lf.set_alignment(cols)
        lf.set_motif_probs(np.array([0.2, 0.3, 0.25, 0.25]))
        lf.lengths[:-1] = np.asarray(lengths)[:-1]
        lf.set_param_rule("kappa", value=3.0)
        lf.set_param_rule("rate_shape", value=0.8)
        return lf

    shards = [make(codes[:, slice(*column_block(codes.shape[1], world, r))]) for r in range(world)]
    local = [lf.get_log_likelihood() for lf in shards]
    want = 0.0
    for v in local:  # rank order, like the kernel
        want += v
    whole = make(codes).get_log_likelihood()
    assert abs(want - whole) <= 1e-12 * abs(whole)

    ptrs = [lf.engine.comm_mailbox(world)[0] for lf in shards]
    for r, lf in enumerate(shards):
        lf.engine.comm_attach(world, r, device_ptrs=ptrs)
    for lf in shards:
        lf._sync()
    got = engine_many([lf.engine for lf in shards])
    assert list(got) == [want] * world
    # more evaluations (both mailbox parities), dirty-path updates included
    rng = np.random.default_rng(4)
    for _ in range(4):
        n = int(rng.choice(topo.edges))
        v = float(rng.uniform(0.02, 0.4))
        for lf in shards:
            lf.set_param_rule("length", edge=topo.names[n], value=v)
            lf._sync()
        got = engine_many([lf.engine for lf in shards])
        assert len(set(got)) == 1
        ref = make(codes)
        ref.lengths[:] = shards[0].lengths
        assert abs(got[0] - ref.get_log_likelihood()) <= 1e-12 * abs(got[0])
    # nothing changed: every rank serves the cached TOTAL, no launch
    again = engine_many([lf.engine for lf in shards])
    np.testing.assert_array_equal(again, got)
    # detached engines give shard values again
    for lf in shards:
        lf.engine.comm_detach()
    back = [lf.engine.evaluate() for lf in shards]
    assert abs(sum(back) - got[0]) <= 1e-12 * abs(got[0]) and len(set(back)) == world


def test_multi_locus_on_one_gpu_matches_the_oracle_sum():
    """whole loci as a second axis (SumDefn over loci, likelihood_calculation.py:161-163): the
    engines of a rank's loci are launched together (c3_evaluate_many) and summed in locus order"""
    from cogent3_b200.sharded import MultiLocusLikelihoodFunction

    topo, lengths, _ = make_workload(12, 10, 4, seed=31)
    alns = [make_workload(12, L, 4, seed=31, shard=k)[2].astype(np.int64) for k, L in enumerate((4000, 911, 20_000))]
    names = ["a", "b", "c"]
    got_lf = MultiLocusLikelihoodFunction("GTR", topo, names, bins=4)
    want_lf = MultiLocusLikelihoodFunction("GTR", topo, names, bins=4, engine_factory=OracleEngine)
    for lf in (got_lf, want_lf):
        lf.set_alignment(alns, np.eye(4))
        lf.set_motif_probs_from_data()
        lf.set_lengths(lengths)
        for k, p in enumerate(hostmodels.get_model("GTR").parameter_order):
            lf.set_param_rule(p, value=0.6 + 0.3 * k)
        lf.set_param_rule("rate_shape", value=0.7)
        lf.set_param_rule("rate_shape", value=1.4, locus="b")
    got, want = got_lf.get_log_likelihood(), want_lf.get_log_likelihood()
    assert abs(got - want) <= REL_LNL * abs(want)
    per_g, per_w = got_lf.locus_log_likelihoods(), want_lf.locus_log_likelihoods()
    for i in range(3):
        assert abs(per_g[i] - per_w[i]) <= REL_LNL * abs(per_w[i])
    got_lf.set_param_rule("length", edge=topo.names[2], value=0.4, locus="c")
    want_lf.set_param_rule("length", edge=topo.names[2], value=0.4, locus="c")
    got, want = got_lf.get_log_likelihood(), want_lf.get_log_likelihood()
    assert abs(got - want) <= REL_LNL * abs(want)


@pytest.mark.parametrize("case", ["c1_brca1_hky85_default", "c2_gtr_gamma4_mid", "c4_cnfgtr_small"])
def test_graph_replay_is_bit_identical(case, monkeypatch):
    """an evaluation structure (dirty nodes, expm jobs) seen twice is captured into a CUDA graph and
    replayed; the numbers travel in the parameter block, so replays must return the bits of plain launches"""
    g = Golden(case)

    def run(graphs):
        monkeypatch.setenv("C3_GRAPH", "1" if graphs else "0")
        monkeypatch.setenv("C3_GRAPH_AFTER", "2")  # (default 16: only structures that keep coming up)
        lf = lf_from_golden(g)
        out = []
        rng = np.random.default_rng(8)
        edges = [g.topo.names[int(n)] for n in rng.choice(g.topo.edges, size=3, replace=False)]
        for rep in range(5):
            # full sweeps: every length moves
            lf.lengths[:-1] = g.lengths[:-1] * (1.0 + 0.01 * (rep + 1))
            lf._dirty_model = True
            out.append(lf.get_log_likelihood())
            # the same three single-branch updates every round: three structures, revisited
            for k, name in enumerate(edges):
                lf.set_param_rule("length", edge=name, value=0.05 + 0.03 * k + 0.01 * rep)
                out.append(lf.get_log_likelihood())
        out.append(np.array(lf.engine.get_lh(0)))
        return out, lf.engine.stats()

    plain, st0 = run(False)
    graphed, st1 = run(True)
    assert st0["graph_replays"] == 0
    assert st1["graph_replays"] >= 12  # 4 structures x 5 rounds: launched plainly, captured, then replayed 3 times
    assert st0["kernel_launches"] == st1["kernel_launches"]
    for a, b in zip(plain[:-1], graphed[:-1], strict=True):
        assert a == b
    assert np.array_equal(plain[-1], graphed[-1])


def test_batched_graph_of_many_engines_matches_individual_evaluations(monkeypatch):
    """c3_update_many: a set of engines that keeps being evaluated together is captured into ONE graph
    (every engine on its own stream inside it); results are the bits of one-at-a-time evaluations,
    and a batch with a different structure (one branch only) falls back without harm"""
    from cogent3_b200.engine import update_many

    monkeypatch.setenv("C3_GRAPH_AFTER", "3")
    g = Golden("c1_brca1_hky85_default")
    lfs = [lf_from_golden(g) for _ in range(6)]
    solo = lf_from_golden(g)
    for f in [*lfs, solo]:
        f.get_log_likelihood()
    engines = [f.engine for f in lfs]
    rates = solo.rates()
    base = np.nan_to_num(g.lengths)
    rng = np.random.default_rng(2)
    for step in range(12):
        if step == 8:  # a different structure in the middle: only two branches move
            d = np.array(last)
            d[int(g.topo.edges[3])] *= 1.1
            d[int(g.topo.edges[9])] *= 0.9
        else:
            d = (base * (1.0 + 0.01 * (step + 1)))[:, None] * rates[None, :]
        # every engine its own table on odd steps, one shared table on even ones
        tables = np.stack([d * (1.0 + 1e-3 * i) for i in range(len(engines))]) if step % 2 else d
        got = update_many(engines, tables)
        for i in range(len(engines)):
            want = solo.engine.update(tables[i] if step % 2 else d)
            assert got[i] == want, (step, i)
        last = tables[0] if step % 2 else d
    replays = [e.stats()["graph_replays"] for e in engines]
    assert min(replays) >= 5, replays  # captured at the third batch, replayed for the full sweeps after it


@pytest.mark.gpu
@pytest.mark.parametrize("n_tips,L,rescale", [(10, 30_000, False), (24, 120_000, False), (16, 5_000, True)])
def test_phylo_hmm_reduction_over_the_sites(n_tips, L, rescale):
    """c3_hmm_reduce (a parallel product of per-site 2 x 2 matrices) against the oracle's restatement of the
    reference's sequential loop (log_dot_reduce, evolve/likelihood_tree.py:168-176) on the same per-bin root
    likelihoods; with per-pattern rescaling active the exponents are folded in"""
    from oracle import oracle

    m = hostmodels.get_model("HKY85")
    topo, lengths, codes = make_workload(n_tips, L, 4, kind="evolved", seed=23)
    lf = LikelihoodFunction(m, topo, bins=4)
    lf.set_alignment(codes.astype(np.int64), np.eye(4))
    lf.set_motif_probs_from_data()
    lf.lengths[:-1] = np.asarray(lengths)[:-1]
    lf.set_param_rule("kappa", value=2.7)
    lf.set_param_rule("rate_shape", value=0.6)
    if rescale:
        lf.engine.set_rescaling("on", 4)
    lf.get_log_likelihood()
    assert lf.engine.rescaling()[0] == rescale
    K = 4
    alloc = [0, 0, 1, 1]  # PatchSiteDistribution.__init__ (evolve/likelihood_calculation.py:198-207)
    bprobs = np.array([0.1, 0.3, 0.35, 0.25])
    pprobs = np.array([bprobs[:2].sum(), bprobs[2:].sum()])
    weights = [bprobs[i] / pprobs[alloc[i]] for i in range(K)]
    s = 0.2  # bin_switch: SiteClassTransitionMatrix (evolve/likelihood_calculation.py:188-194 analogue)
    switch = np.array([[1 - s * pprobs[1], s * pprobs[1]], [s * pprobs[0], 1 - s * pprobs[0]]]).T
    got = lf.engine.hmm_reduce(alloc, weights, switch, pprobs)
    lhs = [lf.engine.get_lh(b) for b in range(K)]
    plhs = np.ascontiguousarray(oracle.patch_weighted_sum_lhs(alloc, weights, lhs).T)
    index = lf.patterns.root.index
    want = oracle.log_dot_reduce(np.asarray(index), pprobs, switch, plhs)
    assert abs(got - want) <= 1e-10 * abs(want), (got, want)


def rng_bp(k, bins):
    return np.random.default_rng(100 + k).dirichlet(np.ones(bins) * 6.0)


@pytest.mark.parametrize(
    "model,n_tips,L,bins,rooted",
    [("HKY85", 24, 3000, 1, False), ("GTR", 17, 1800, 4, False), ("HKY85", 9, 700, 2, True), ("GTR", 40, 2000, 4, False)],
)
def test_chain_kernel_single_branch_updates_bit_identical(model, n_tips, L, bins, rooted, monkeypatch):
    """single-branch updates of a small problem: the chain kernel (one launch, every root row walks the dirty
    path through composed gather tables, no barrier between the levels) against the small-levels launch
    (C3_CHAIN=0): lnL, every partial row and lh bit-identical, for every branch of the tree -- tip edges,
    edges under the root, with a fixed motif on the path, through graph replays"""
    from cogent3_b200.tree import Topology

    m = hostmodels.get_model(model)
    topo, lengths, codes = make_workload(n_tips, L, 4, kind="evolved", seed=31)
    if rooted:
        topo, lengths = Topology.from_newick("(((a:0.1,b:0.2):0.1,(c:0.3,d:0.2):0.05):0.1,((e:0.1,f:0.1):0.2,(g:0.2,(h:0.1,i:0.3):0.1):0.1):0.2);")
        codes = codes[: int(np.sum(topo.is_tip))]

    # ambiguity codes (R, N) in 3 % of the cells: tip rows that are not one-hot take the summed form of the
    # pre-multiplied tip table, also inside the fused launch
    rows = np.vstack([np.eye(4), [[1.0, 0.0, 1.0, 0.0], [1.0, 1.0, 1.0, 1.0]]])
    codes = codes.copy()
    amb = np.random.default_rng(17).random(codes.shape)
    codes[amb < 0.015] = 4
    codes[amb > 0.985] = 5

    def build(chain, fused=True):
        monkeypatch.setenv("C3_CHAIN", "1" if chain else "0")
        monkeypatch.setenv("C3_CHAIN_FUSED", "1" if fused else "0")
        monkeypatch.setenv("C3_GRAPH_AFTER", "2")
        lf = LikelihoodFunction(m, topo, bins=bins)
        lf.set_alignment(codes.astype(np.int64), rows)
        lf.set_motif_probs_from_data()
        lf.lengths[:-1] = np.asarray(lengths)[:-1]
        for k, p in enumerate(m.parameter_order):
            lf.set_param_rule(p, value=0.5 + 0.37 * k)
        if bins > 1:
            lf.set_param_rule("rate_shape", value=0.6)
        return lf

    # a: ONE launch per update (parameters by value, the branch's P computed in the launch); c: expm launch + chain
    # launch behind a parameter-block copy (graph-replayed); b: expm + small-levels launch
    a, b, c = build(True), build(False), build(True, fused=False)
    assert a.get_log_likelihood() == b.get_log_likelihood() == c.get_log_likelihood()
    rng = np.random.default_rng(3)
    internal = [n for n in range(topo.n_nodes) if not topo.is_tip[n]]
    edges = list(topo.edges)
    for rep in range(3):
        for n in edges:
            v = float(rng.uniform(0.01, 0.6))
            for lf in (a, b, c):
                lf.set_param_rule("length", edge=topo.names[n], value=v)
            va, vb, vc = a.get_log_likelihood(), b.get_log_likelihood(), c.get_log_likelihood()
            assert va == vb == vc, (rep, topo.names[n], va, vb, vc)
            if rep == 1:
                # the P (and, below a tip, the tip table) the fused launch left in memory is the expm launch's
                for bin_ in range(bins):
                    np.testing.assert_array_equal(a.engine.get_psub(n, bin_), b.engine.get_psub(n, bin_))
        for n in internal:
            for bin_ in range(bins):
                np.testing.assert_array_equal(a.engine.get_partials(n, bin_), b.engine.get_partials(n, bin_))
                np.testing.assert_array_equal(c.engine.get_partials(n, bin_), b.engine.get_partials(n, bin_))
        for bin_ in range(bins):
            np.testing.assert_array_equal(a.engine.get_lh(bin_), b.engine.get_lh(bin_))
            np.testing.assert_array_equal(c.engine.get_lh(bin_), b.engine.get_lh(bin_))
        if rep == 0:
            # a fixed motif on an internal node of some paths (ancestral-state style masks)
            fm = np.full(topo.n_nodes, -1, np.int32)
            fm[internal[len(internal) // 2]] = 2
            for lf in (a, b, c):
                lf.engine.set_fixed_motifs(fm)
            assert a.engine.evaluate() == b.engine.evaluate() == c.engine.evaluate()
    # VALUES that are not structure -- the root motif probabilities, the bin probabilities -- must reach a launch
    # that is replayed from a graph (a change of pi dirties the root alone: a one-node path, the same structure
    # every time; this is what the reference's own test_discrete_vs_general1 found)
    for k in range(6):
        pi = rng.dirichlet(np.ones(4) * 8.0)
        for lf in (a, b, c):
            lf.engine.set_root_mprobs(pi)
            if bins > 1 and k % 2:
                lf.engine.set_bin_probs(rng_bp(k, bins))
        va, vb, vc = a.engine.evaluate(), b.engine.evaluate(), c.engine.evaluate()
        assert va == vb == vc, (k, va, vb, vc)
        n = edges[k % len(edges)]
        for lf in (a, b, c):
            lf.set_param_rule("length", edge=topo.names[n], value=0.07 + 0.01 * k)
        # (the mirror API re-sets its own pi / bprobs when it evaluates: compare at engine level again afterwards)
        va, vb, vc = a.get_log_likelihood(), b.get_log_likelihood(), c.get_log_likelihood()
        assert va == vb == vc, (k, va, vb, vc)
    # every branch moved: the next full sweep reads the tables the fused launches wrote back
    for lf in (a, b, c):
        lf.set_param_rule("rate_shape" if bins > 1 else m.parameter_order[0], value=0.77)
    assert a.get_log_likelihood() == b.get_log_likelihood() == c.get_log_likelihood()
    sa, sb, sc = a.engine.stats(), b.engine.stats(), c.engine.stats()
    assert sb["chain_launches"] == 0
    assert sa["chain_launches"] >= len(edges)  # (paths longer than 16 nodes keep the level launches)
    assert sc["chain_launches"] == sa["chain_launches"] and sc["graph_replays"] > 0
    assert sa["kernel_launches"] < sc["kernel_launches"]  # (no expm launch in front of a fused update)
    # the full sweep after invalidation agrees with the last chain update
    last = a.get_log_likelihood()
    a.engine.invalidate()
    assert a.engine.evaluate() == last


@pytest.mark.parametrize(
    "model,n_tips,L,bins,rooted",
    [("HKY85", 24, 3000, 1, False), ("GTR", 17, 1800, 4, False), ("HKY85", 9, 700, 2, True), ("GTR", 60, 1500, 4, False)],
)
def test_walk_kernel_bit_identical(model, n_tips, L, bins, rooted, monkeypatch):
    """every evaluation of a small problem that is not a single path -- full sweeps, several branches at once, a
    rate-matrix parameter -- by the tree-walk kernel (one launch, every root row evaluates the dirty nodes itself,
    dirty children from a per-thread value stack) against the small-levels launch (C3_WALK=0, C3_CHAIN=0): lnL,
    every partial row and lh bit-identical, through graph replays, with fixed motifs and ambiguity codes"""
    from cogent3_b200.tree import Topology

    m = hostmodels.get_model(model)
    topo, lengths, codes = make_workload(n_tips, L, 4, kind="evolved", seed=33)
    if rooted:
        topo, lengths = Topology.from_newick("(((a:0.1,b:0.2):0.1,(c:0.3,d:0.2):0.05):0.1,((e:0.1,f:0.1):0.2,(g:0.2,(h:0.1,i:0.3):0.1):0.1):0.2);")
        codes = codes[: int(np.sum(topo.is_tip))]
    rows = np.vstack([np.eye(4), [[1.0, 0.0, 1.0, 0.0], [1.0, 1.0, 1.0, 1.0]]])
    codes = codes.copy()
    amb = np.random.default_rng(19).random(codes.shape)
    codes[amb < 0.015] = 4
    codes[amb > 0.985] = 5

    def build(walk):
        monkeypatch.setenv("C3_WALK", "1" if walk else "0")
        monkeypatch.setenv("C3_CHAIN", "1" if walk else "0")
        monkeypatch.setenv("C3_GRAPH_AFTER", "2")
        lf = LikelihoodFunction(m, topo, bins=bins)
        lf.set_alignment(codes.astype(np.int64), rows)
        lf.set_motif_probs_from_data()
        lf.lengths[:-1] = np.asarray(lengths)[:-1]
        for k, p in enumerate(m.parameter_order):
            lf.set_param_rule(p, value=0.5 + 0.37 * k)
        if bins > 1:
            lf.set_param_rule("rate_shape", value=0.6)
        return lf

    a, b = build(True), build(False)
    assert a.get_log_likelihood() == b.get_log_likelihood()  # the first full sweep
    rng = np.random.default_rng(4)
    internal = [n for n in range(topo.n_nodes) if not topo.is_tip[n]]
    edges = list(topo.edges)

    def same_state():
        for n in internal:
            for bin_ in range(bins):
                np.testing.assert_array_equal(a.engine.get_partials(n, bin_), b.engine.get_partials(n, bin_))
        for bin_ in range(bins):
            np.testing.assert_array_equal(a.engine.get_lh(bin_), b.engine.get_lh(bin_))

    same_state()
    for step in range(24):
        if step % 4 == 0:  # a rate-matrix parameter: everything moves
            for lf in (a, b):
                lf.set_param_rule(m.parameter_order[0], value=0.4 + 0.05 * step)
        else:  # two or three branches at once: the dirty set is a union of paths, the same structures revisited
            pick = [edges[(7 * step + 3 * k) % len(edges)] for k in range(2 + step % 2)]
            for n in pick:
                v = float(rng.uniform(0.01, 0.6))
                for lf in (a, b):
                    lf.set_param_rule("length", edge=topo.names[n], value=v)
        va, vb = a.get_log_likelihood(), b.get_log_likelihood()
        assert va == vb, (step, va, vb)
        if step == 9:
            fm = np.full(topo.n_nodes, -1, np.int32)
            fm[internal[len(internal) // 3]] = 1
            for lf in (a, b):
                lf.engine.set_fixed_motifs(fm)
            assert a.engine.evaluate() == b.engine.evaluate()
        if step in (3, 10, 23):
            same_state()
    sa, sb = a.engine.stats(), b.engine.stats()
    assert sb["walk_launches"] == 0 and sb["chain_launches"] == 0
    assert sa["walk_launches"] >= 12 and sa["graph_replays"] > 0  # (some picks of a small tree are a single path: chain launches)
