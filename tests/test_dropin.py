"""The drop-in glue under the REAL cogent3 objects (build container only: needs
/root/reference).  The engine is the CPU oracle test double here (no GPU), so what is
tested is the host logic of cogent3_b200/dropin.py: the replaced Defn sub-graph, slot /
distance / fixed-motif plumbing, accessors, optimiser protocol.  The same tests run with
the CUDA engine when a GPU and the reference are both present (never on the GPU box)."""

import functools

import numpy as np
import pytest
from refutil import import_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="needs /root/reference")

if reference_available():
    cogent3 = import_reference()
    from cogent3 import get_dataset, get_model, make_aligned_seqs, make_tree

    from cogent3_b200.dropin import make_likelihood_function
    from oracle.engine import OracleEngine

    FACTORY = functools.partial(OracleEngine, mode="port")


def small_problem(n_tips=7, L=240, n_states=4, seed=5):
    from cogent3_b200 import models as hostmodels
    from cogent3_b200.synthetic import evolved_alignment
    from cogent3_b200.tree import random_join_tree

    rng = np.random.default_rng(seed)
    topo, lengths = random_join_tree(n_tips, rng)
    codes = evolved_alignment(topo, lengths, L, n_states, rng)
    symbols = hostmodels.dna_symbol_rows()[0] if n_states == 4 else hostmodels.codon_symbol_rows()[0]
    data = {n: "".join(symbols[c] for c in row) for n, row in zip(topo.tip_names, codes, strict=True)}
    return make_tree(topo.to_newick(lengths)), make_aligned_seqs(data, moltype="dna")


def pair(model_factory, tree, aln, **kw):
    ref = model_factory().make_likelihood_function(tree, **kw)
    ref.set_alignment(aln)
    lf = make_likelihood_function(model_factory(), tree, engine_factory=FACTORY, **kw)
    lf.set_alignment(aln)
    return ref, lf


def assert_same(ref, lf, rel=1e-12):
    a, b = ref.get_log_likelihood(), lf.get_log_likelihood()
    assert abs(a - b) <= rel * abs(a), (a, b)


def test_brca1_hky85_matches_reference_and_survey_value():
    aln, tree = get_dataset("brca1"), get_dataset("mammal-tree")
    ref, lf = pair(lambda: get_model("HKY85"), tree, aln)
    assert lf.get_log_likelihood() == pytest.approx(-193293.46430542142, rel=1e-12)
    assert_same(ref, lf)
    for f in (ref, lf):
        f.set_param_rule("kappa", value=3.3)
        f.set_param_rule("length", edge="Human", value=0.07)
    assert_same(ref, lf)
    assert lf.get_num_free_params() == ref.get_num_free_params() == 108
    np.testing.assert_allclose(
        lf.get_full_length_likelihoods(), ref.get_full_length_likelihoods(), rtol=0, atol=1e-15
    )
    np.testing.assert_array_equal(
        lf.get_psub_for_edge("Human").to_array(), ref.get_psub_for_edge("Human").to_array()
    )


def test_gtr_gamma_bins_and_posteriors():
    tree, aln = small_problem()
    mk = lambda: get_model("GTR", ordered_param="rate", distribution="gamma")  # noqa: E731
    ref, lf = pair(mk, tree, aln, bins=4)
    for f in (ref, lf):
        f.set_param_rule("bprobs", is_constant=True)
        f.set_param_rule("rate_shape", value=0.5)
        f.set_param_rule("A/G", value=3.2)
    assert_same(ref, lf)
    np.testing.assert_allclose(lf.get_bin_probs().array, ref.get_bin_probs().array, atol=1e-13)
    np.testing.assert_allclose(
        lf.get_full_length_likelihoods(), ref.get_full_length_likelihoods(), atol=1e-15
    )
    assert lf.get_G_statistic() == pytest.approx(ref.get_G_statistic(), rel=1e-10)


def test_gn_time_heterogeneous_pade():
    tree, aln = small_problem(n_tips=6, L=150, seed=8)
    ref, lf = pair(lambda: get_model("GN"), tree, aln, expm="pade")
    rng = np.random.default_rng(0)
    for f in (ref, lf):
        f.set_time_heterogeneity(is_independent=True, upper=50)
    names = [e.name for e in tree.get_edge_vector(include_root=False)]
    for name in names:
        for p in lf.model.get_param_list():
            v = float(rng.uniform(0.3, 3.0))
            for f in (ref, lf):
                f.set_param_rule(p, edge=name, value=v)
    assert_same(ref, lf, rel=1e-11)


def test_codon_model():
    tree, aln = small_problem(n_tips=5, L=40, n_states=61, seed=3)
    ref, lf = pair(lambda: get_model("CNFGTR"), tree, aln)
    for f in (ref, lf):
        f.set_param_rule("omega", value=0.4)
        f.set_param_rule("C/T", value=3.6)
    assert_same(ref, lf, rel=1e-11)


def test_discrete_edges_use_supplied_psubs():
    tree, aln = small_problem(n_tips=5, L=200, seed=11)
    tip = tree.get_tip_names()[0]
    ref, lf = pair(lambda: get_model("HKY85"), tree, aln, discrete_edges=[tip])
    assert_same(ref, lf, rel=1e-11)


def test_ancestral_reconstruction_uses_fixed_motifs():
    tree, aln = small_problem(n_tips=5, L=60, seed=2)
    ref, lf = pair(lambda: get_model("F81"), tree, aln)
    a = ref.reconstruct_ancestral_seqs()
    b = lf.reconstruct_ancestral_seqs()
    assert set(a) == set(b)
    for edge in a:
        np.testing.assert_allclose(b[edge].array, a[edge].array, atol=1e-13)


def test_optimise_powell_follows_the_same_path():
    tree, aln = small_problem(n_tips=6, L=300, seed=4)
    ref, lf = pair(lambda: get_model("HKY85"), tree, aln)
    for f in (ref, lf):
        f.optimise(local=True, show_progress=False, max_evaluations=250, limit_action="ignore")
    assert_same(ref, lf, rel=1e-9)
    assert lf.get_param_value("kappa") == pytest.approx(ref.get_param_value("kappa"), rel=1e-6)


def test_multi_locus_sum():
    tree, aln = small_problem(n_tips=5, L=200, seed=6)
    half = len(aln) // 2
    alns = [aln[:half], aln[half:]]
    ref = get_model("HKY85").make_likelihood_function(tree, loci=["a", "b"])
    ref.set_alignment(alns)
    lf = make_likelihood_function(get_model("HKY85"), tree, loci=["a", "b"], engine_factory=FACTORY)
    lf.set_alignment(alns)
    assert_same(ref, lf)


def test_set_alignment_twice_reuses_the_function():
    tree, aln = small_problem(n_tips=5, L=100, seed=7)
    _, aln2 = small_problem(n_tips=5, L=130, seed=7)
    ref, lf = pair(lambda: get_model("HKY85"), tree, aln)
    assert_same(ref, lf)
    ref.set_alignment(aln2)
    lf.set_alignment(aln2)
    assert_same(ref, lf)


def test_product_default_engine_has_no_cpu_fallback():
    from cogent3_b200 import _lib

    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    tree, aln = small_problem(n_tips=4, L=30, seed=1)
    lf = make_likelihood_function(get_model("HKY85"), tree)  # default engine = CUDA
    with pytest.raises((_lib.C3Error, RuntimeError)):
        lf.set_alignment(aln)


# ---- the vectorised pattern compression behind lf.set_alignment ----------------
def _same_lht(a, b):
    """recursively: every table of the reference's lht and of the drop-in's, bit for bit"""
    assert type(a) is type(b) and a.edge_name == b.edge_name
    assert np.array_equal(np.asarray(a.index), np.asarray(b.index))
    assert np.asarray(a.index).dtype == np.asarray(b.index).dtype
    assert np.array_equal(a.counts, b.counts) and a.counts.dtype == b.counts.dtype
    assert np.array_equal(a.ambig, b.ambig)
    assert list(a.shape) == list(b.shape)
    if hasattr(a, "input_likelihoods"):  # a leaf
        assert list(a.uniq) == list(b.uniq)
        assert np.array_equal(a.input_likelihoods, b.input_likelihoods)
        return 1
    assert np.array_equal(a.uniq, b.uniq) and a.uniq.dtype == b.uniq.dtype
    assert np.array_equal(a.indexes, b.indexes) and a.indexes.dtype == b.indexes.dtype
    assert b.indexes.flags["C_CONTIGUOUS"]
    n = 1
    for (ia, ca), (ib, cb) in zip(a._indexed_children, b._indexed_children, strict=True):
        assert np.array_equal(ia, ib)
        n += _same_lht(ca, cb)
    return n


@pytest.mark.parametrize(
    "model,data",
    [("HKY85", "brca1"), ("GTR", "ambiguous"), ("CNFGTR", "codon"), ("GN", "small"), ("dinuc", "small")],
)
def test_fast_compression_builds_the_reference_tables(model, data):
    from cogent3.evolve import substitution_model

    if data == "brca1":
        aln, tree = get_dataset("brca1"), get_dataset("mammal-tree")
        tree = tree.get_sub_tree(aln.names)
    elif data == "codon":
        tree, aln = small_problem(n_tips=6, L=120, n_states=61)
    elif data == "ambiguous":
        tree = make_tree("((a,b),c,(d,e));")
        aln = make_aligned_seqs(
            {"a": "ACGTACGTAAGTNNRY-A", "b": "ACGTACGGAAGT-RYYCA", "c": "ACCTACGGTAGCAAGG-A",
             "d": "ACCTATGGTAGCAAGGTA", "e": "NCCTATGG--GCAWGGTA"}, moltype="dna")  # fmt: skip
    else:
        tree, aln = small_problem(n_tips=9, L=400)

    def factory():
        if model == "dinuc":
            return substitution_model.TimeReversibleDinucleotide(equal_motif_probs=True)
        return get_model(model)

    ref, lf = pair(factory, tree, aln)
    a, b = ref.get_param_value("lht"), lf.get_param_value("lht")
    assert _same_lht(a, b) == len(tree.get_edge_vector())
    assert_same(ref, lf)
    np.testing.assert_array_equal(ref.get_motif_probs().to_array(), lf.get_motif_probs().to_array())


def test_fast_compression_is_what_set_alignment_runs():
    """a 40-taxon x 30k-column alignment: the reference's dict loops take seconds per pass"""
    import time

    tree, aln = small_problem(n_tips=40, L=30_000, seed=9)
    t0 = time.perf_counter()
    ref = get_model("HKY85").make_likelihood_function(tree)
    ref.set_alignment(aln)
    t_ref = time.perf_counter() - t0
    t0 = time.perf_counter()
    lf = make_likelihood_function(get_model("HKY85"), tree, engine_factory=FACTORY)
    lf.set_alignment(aln)
    t_fast = time.perf_counter() - t0
    assert _same_lht(ref.get_param_value("lht"), lf.get_param_value("lht")) == 2 * 40 - 2
    assert t_fast < t_ref, (t_fast, t_ref)
    print(f"set_alignment 40 x 30k: reference {t_ref:.2f} s, drop-in {t_fast:.2f} s")


# ---- the engine Defn fed with the factors of the psubs (Qd, length, rate) -----------------------
def _psub_cases():
    tree, aln = small_problem(n_tips=8, L=300, seed=3)
    yield "HKY85", (lambda: get_model("HKY85")), tree, aln, {}, None
    yield "GTR+G4", (lambda: get_model("GTR", ordered_param="rate", distribution="gamma")), tree, aln, {"bins": 4}, None
    yield "GN het", (lambda: get_model("GN")), tree, aln, {"expm": "pade"}, "het"


@pytest.mark.parametrize("factored", [True, False])
def test_psub_accessors_match_the_reference(factored):
    for tag, mk, tree, aln, kw, extra in _psub_cases():
        ref = mk().make_likelihood_function(tree, **kw)
        ref.set_alignment(aln)
        lf = make_likelihood_function(mk(), tree, engine_factory=FACTORY, factored=factored, **kw)
        lf.set_alignment(aln)
        assert (lf._factors is not None) == factored, tag
        for f in (ref, lf):
            if "bins" in kw:
                f.set_param_rule("bprobs", is_constant=True)
                f.set_param_rule("rate_shape", value=0.7)
            if extra == "het":
                f.set_time_heterogeneity(is_independent=True, upper=50)
            f.set_param_rule("length", edge=tree.get_tip_names()[0], value=0.37)
        assert_same(ref, lf)
        a, b = ref.get_all_psubs(), lf.get_all_psubs()
        assert list(a) == list(b) or set(a) == set(b), (tag, list(a)[:3], list(b)[:3])
        for key in a:
            np.testing.assert_array_equal(a[key].to_array(), b[key].to_array())
        name = tree.get_tip_names()[1]
        scope = {"bin": "bin2"} if "bins" in kw else {}
        np.testing.assert_array_equal(
            ref.get_psub_for_edge(name, **scope).to_array(), lf.get_psub_for_edge(name, **scope).to_array()
        )
        np.testing.assert_array_equal(
            np.asarray(ref.get_param_value("psubs", edge=name, **scope)),
            np.asarray(lf.get_param_value("psubs", edge=name, **scope)),
        )
        assert ref.all_psubs_DLC() == lf.all_psubs_DLC()
        np.testing.assert_allclose(
            ref.get_lengths_as_ens()[name], lf.get_lengths_as_ens()[name], rtol=1e-12
        )


def test_factored_defn_has_no_distance_or_psub_cells():
    tree, aln = small_problem(n_tips=12, L=200, seed=4)
    mk = lambda: get_model("GTR", ordered_param="rate", distribution="gamma")  # noqa: E731
    counts = {}
    for factored in (True, False):
        lf = make_likelihood_function(mk(), tree, bins=4, engine_factory=FACTORY, factored=factored)
        lf.set_alignment(aln)
        calc = lf.make_calculator()
        counts[factored] = len(calc._cells)
        names = {d.name for d in lf.defns}
        assert ("psubs" in names) == (not factored)
    n_edges = len(tree.get_edge_vector()) - 1
    # K*E distance cells and K*E psubs cells are gone
    assert counts[False] - counts[True] >= 2 * 4 * n_edges - 8, counts


def test_factored_optimise_and_partial_updates():
    aln, tree = get_dataset("brca1"), get_dataset("mammal-tree")
    tree = tree.get_sub_tree(aln.names)
    ref = get_model("HKY85").make_likelihood_function(tree)
    ref.set_alignment(aln)
    lf = make_likelihood_function(get_model("HKY85"), tree, engine_factory=FACTORY)
    lf.set_alignment(aln)
    assert lf._factors is not None
    assert_same(ref, lf)
    rc, lc = ref.make_calculator(), lf.make_calculator()
    x = np.array(rc.get_value_array())
    assert np.array_equal(x, lc.get_value_array())
    rng = np.random.default_rng(0)
    for _ in range(25):  # one parameter at a time, like Powell's line searches
        i = int(rng.integers(len(x)))
        x[i] *= 1.0 + 0.05 * rng.standard_normal()
        a, b = rc(x), lc(x)
        assert abs(a - b) <= 1e-12 * abs(a)
