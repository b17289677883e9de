"""The drop-in glue under the REAL cogent3 objects.

* build container (no GPU): cogent3 comes from /root/reference/src and the engine is the CPU oracle
  test double, so what is tested is the host logic of cogent3_b200/dropin.py: the replaced Defn
  sub-graph, slot / distance / fixed-motif plumbing, accessors, optimiser protocol, the product
  hook ``cogent3_b200.install()``;
* GPU box: tests/test_gpu_dropin.py re-exports every test of this module under the ``gpu`` marker;
  cogent3 then comes from ``baseline/_ref`` (unmodified copy) and FACTORY is None,
  i.e. the product's own default engine -- libc3cuda.  Same bodies, same tolerances: the composite
  real cogent3 + drop-in + CUDA engine against the stock class in the same process and against the
  reference's own known answers."""

import contextlib
import functools
import json
import os
import pickle

import numpy as np
import pytest
from refutil import import_reference, reference_available, reference_tests_dir

pytestmark = pytest.mark.skipif(not reference_available(), reason="needs the reference package")


def _gpu_present():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


ON_GPU = _gpu_present()
HERE = os.path.dirname(os.path.abspath(__file__))

if reference_available():
    cogent3 = import_reference()
    from cogent3 import get_dataset, get_model, load_aligned_seqs, load_tree, make_aligned_seqs, make_tree

    import cogent3_b200
    from cogent3_b200.dropin import make_likelihood_function

    if ON_GPU:
        FACTORY = None  # the product default: LikelihoodEngine over libc3cuda (no CPU fallback)
    else:
        from oracle.engine import OracleEngine

        FACTORY = functools.partial(OracleEngine, mode="port")
    DATA = os.path.join(reference_tests_dir() or "", "data")


@contextlib.contextmanager
def installed(**kw):
    """the product hook: inside, ``sm.make_likelihood_function(tree)`` itself builds the CUDA class"""
    cogent3_b200.install(engine_factory=FACTORY, **kw)
    try:
        yield
    finally:
        cogent3_b200.uninstall()


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
        lf.get_full_length_likelihoods(), ref.get_full_length_likelihoods(), rtol=1e-12, atol=1e-15
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
        lf.get_full_length_likelihoods(), ref.get_full_length_likelihoods(), rtol=1e-12, atol=1e-15
    )
    assert lf.get_G_statistic() == pytest.approx(ref.get_G_statistic(), rel=1e-10)


def test_fast_gamma_rates_bit_identical():
    """a5: the drop-in's GammaDefn computes the rate classes without scipy's distribution wrapper --
    the same bits as the reference's calc (recalculation/definition.py:236-244), inside the likelihood
    function too (the Defn's value is what both classes hand to ``distance``)"""
    from cogent3.recalculation.definition import GammaDefn

    from cogent3_b200.dropin import fast_gamma_rates

    rng = np.random.default_rng(11)
    for i in range(600):
        k = int(rng.integers(2, 9))
        w = rng.random(k) + 0.01 if i % 3 else np.ones(k) / k
        a = float(np.exp(rng.uniform(np.log(1e-2), np.log(300.0))))
        a = np.float64(a) if i % 2 else a
        want, got = GammaDefn.calc(None, w, a), fast_gamma_rates(w, a)
        assert want.dtype == got.dtype and want.tobytes() == got.tobytes(), (w, a)
    # outside the plain case the reference's own code answers (an empty last bin puts a percentile at 1)
    for w in (np.array([0.5, 0.5, 0.0]), np.array([0.3, 0.3, 0.1, 0.0]), np.array([1.7, 0.9, 0.0])):
        np.testing.assert_array_equal(fast_gamma_rates(w, 0.7), GammaDefn.calc(None, w, 0.7))
    tree, aln = small_problem()
    mk = lambda: get_model("GTR", ordered_param="rate", distribution="gamma")  # noqa: E731
    ref, lf = pair(mk, tree, aln, bins=4)
    gamma = [d for d in lf.defns if type(d) is GammaDefn]
    assert gamma and all(d.calc is fast_gamma_rates for d in gamma)
    for shape in (0.05, 0.5, 1.7, 40.0):
        for f in (ref, lf):
            f.set_param_rule("rate_shape", value=shape)
        for b in lf.bin_names:
            assert lf.get_param_value("rate", bin=b) == ref.get_param_value("rate", bin=b)
        assert_same(ref, lf)
    pickle.loads(pickle.dumps(lf))


def test_accessors_between_streaming_evaluations():
    """an alignment large enough for the streaming kernels (more than 8192 root patterns): the root launch then only
    writes the class mixture (lean root) and per-pattern likelihoods are recomputed when an accessor asks --
    lnL, per-column likelihoods, posterior bin probabilities and the G statistic against the stock class between
    optimiser-style updates, and again after them"""
    tree, aln = small_problem(n_tips=10, L=30_000, seed=12)
    mk = lambda: get_model("GTR", ordered_param="rate", distribution="gamma")  # noqa: E731
    ref, lf = pair(mk, tree, aln, bins=4)
    for f in (ref, lf):
        f.set_param_rule("rate_shape", value=0.7)
    assert_same(ref, lf)
    names = [e.name for e in tree.postorder() if e.name != "root"]
    for k in range(12):  # more evaluations than an lh request keeps lh materialised for
        for f in (ref, lf):
            f.set_param_rule("length", edge=names[k % len(names)], value=0.03 + 0.01 * k)
            if k % 5 == 4:
                f.set_param_rule("A/G", value=2.0 + 0.1 * k)
        assert_same(ref, lf)
        if k in (0, 1, 11):
            np.testing.assert_allclose(
                lf.get_full_length_likelihoods(), ref.get_full_length_likelihoods(), rtol=1e-12, atol=1e-15
            )
            assert_same(ref, lf)  # (the request neither committed nor lost anything)
    np.testing.assert_allclose(lf.get_bin_probs().array, ref.get_bin_probs().array, atol=1e-13)
    assert lf.get_G_statistic() == pytest.approx(ref.get_G_statistic(), rel=1e-10)
    if ON_GPU:
        st = lf._engine().stats()
        assert st["lean_evaluations"] > 0, st


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


@pytest.mark.parametrize("threads", [None, "3"])
@pytest.mark.parametrize(
    "model,data",
    [("HKY85", "brca1"), ("GTR", "ambiguous"), ("CNFGTR", "codon"), ("GN", "small"), ("dinuc", "small")],
)
def test_fast_compression_builds_the_reference_tables(model, data, threads, monkeypatch):
    """every table of every node equal to the reference's; also when the nodes of a tree level (and the leaves) are
    built in a thread pool, as they are from 10^5 columns on (C3_TABLE_THREADS forces it here)"""
    from cogent3.evolve import substitution_model

    from cogent3_b200 import dropin

    if threads is None:
        monkeypatch.delenv("C3_TABLE_THREADS", raising=False)
    else:
        monkeypatch.setenv("C3_TABLE_THREADS", threads)
    dropin._LEAF_CACHE.clear()

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


# ---- the reference's own known answers, through the PRODUCT hook ---------------------------------
# (constants and set-ups: /root/reference/tests/test_evolve/test_likelihood_function.py; the stock entry
# point sm.make_likelihood_function(tree) is what builds the CUDA class here)
def _kat_rules(name):
    with open(os.path.join(HERE, "golden", "kat_rules.json")) as f:
        return json.load(f)[name]


def test_install_routes_the_stock_entry_point():
    from cogent3.evolve.parameter_controller import AlignmentLikelihoodFunction

    tree, aln = small_problem(n_tips=5, L=80, seed=12)
    stock_before = get_model("HKY85").make_likelihood_function(tree)
    with installed():
        lf = get_model("HKY85").make_likelihood_function(tree)
        cpu = get_model("HKY85").make_likelihood_function(tree, backend="cpu")
        assert type(lf).__name__ == "CudaAlignmentLikelihoodFunction"
        assert isinstance(lf, AlignmentLikelihoodFunction)
        assert type(cpu) is AlignmentLikelihoodFunction is type(stock_before)
        with pytest.raises(ValueError):
            get_model("HKY85").make_likelihood_function(tree, backend="tpu")
        # out of the engine's scope: stays with the reference class (pair-HMM alignment)
        unaligned = get_model("HKY85").make_likelihood_function(make_tree(tip_names=["a", "b"]), aligned=False)
        assert type(unaligned).__name__ == "SequenceLikelihoodFunction"
        lf.set_alignment(aln)
        cpu.set_alignment(aln)
        assert_same(cpu, lf)
    after = get_model("HKY85").make_likelihood_function(tree)
    assert type(after) is AlignmentLikelihoodFunction


def test_kat_nucleotide_and_codon():
    """LikelihoodCalcs.test_nucleotide / test_codon (:245-270)"""
    from cogent3.evolve import substitution_model

    aln = load_aligned_seqs(os.path.join(DATA, "brca1.fasta"), moltype="dna")
    names = ["Human", "Mouse", "HowlerMon"]
    aln = aln.take_seqs(names)[0:42]
    tree = make_tree(tip_names=names)
    with installed():
        sm = substitution_model.TimeReversibleNucleotide(
            equal_motif_probs=True, motif_probs=None, predicates={"kappa": "transition"}
        )
        lf = sm.make_likelihood_function(tree)
        lf.set_alignment(aln)
        lf.set_param_rule("length", value=1.0, is_constant=True)
        lf.set_param_rule("kappa", value=3.0, is_constant=True)
        assert lf.get_num_free_params() == 0
        assert lf.get_log_likelihood() == pytest.approx(-148.6455087258624, rel=1e-12)
        sm = substitution_model.TimeReversibleCodon(
            equal_motif_probs=True, motif_probs=None,
            predicates={"kappa": "transition", "omega": "replacement"}, mprob_model="tuple",
        )  # fmt: skip
        lf = sm.make_likelihood_function(tree)
        lf.set_alignment(aln)
        lf.set_param_rule("length", value=1.0, is_constant=True)
        lf.set_param_rule("kappa", value=3.0, is_constant=True)
        lf.set_param_rule("omega", value=0.5, is_constant=True)
        assert lf.get_log_likelihood() == pytest.approx(-103.05742415448259, rel=1e-12)


def test_kat_calclikelihood_g_statistic_ancestral():
    """LikelihoodFunctionTests.test_calclikelihood / test_g_statistic / test_ancestralsequences (:411-439)"""
    from cogent3.evolve import substitution_model

    sm = substitution_model.TimeReversibleNucleotide(equal_motif_probs=True, predicates={"beta": "transition"})
    aln = load_aligned_seqs(os.path.join(DATA, "brca1_5.paml"), moltype=sm.moltype)
    tree = load_tree(os.path.join(DATA, "brca1_5.tree"))
    with installed():
        lf = sm.make_likelihood_function(tree)
        lf.set_param_rule("beta", is_independent=True)
        lf.set_alignment(aln)
        for species, length in [("DogFaced", 0.1), ("NineBande", 0.2), ("Human", 0.3), ("HowlerMon", 0.4), ("Mouse", 0.5)]:
            lf.set_param_rule("length", value=length, edge=species, is_constant=True)
        for s1, s2, length in [("Human", "HowlerMon", 0.7), ("Human", "Mouse", 0.6)]:
            lca = tree.get_connecting_node(s1, s2).name
            lf.set_param_rule("length", value=length, edge=lca, is_constant=True)
        lf.set_param_rule("beta", value=4.0, is_constant=True)
        assert type(lf).__name__ == "CudaAlignmentLikelihoodFunction"
        assert lf.get_log_likelihood() == pytest.approx(-250.686745262, abs=1e-9)
        assert lf.get_G_statistic() == pytest.approx(230.77670557, abs=1e-6)
        result = lf.reconstruct_ancestral_seqs()["edge.0"]
        assert result[-1][2] == pytest.approx(2.28460181711e-05, abs=1e-8)


def test_kat_cnfgtr_rate_het_and_time_rate_het():
    """ComparisonTests.test_codon_rate_het (-6755.451985437475, nfp 19, posterior bin probabilities,
    :1522-1772) and test_time_rate_het (-6753.45662012937, rtol 1e-5, nfp 21, :1774-2044)"""
    aln = load_aligned_seqs(os.path.join(DATA, "primate_brca1.fasta"), moltype="dna")
    tree = load_tree(os.path.join(DATA, "primate_brca1.tree"))
    with installed():
        lf = get_model("CNFGTR").make_likelihood_function(tree, bins=["neutral", "adaptive"], digits=2, space=3)
        lf.set_alignment(aln)
        lf.apply_param_rules(_kat_rules("test_codon_rate_het"))
        assert type(lf).__name__ == "CudaAlignmentLikelihoodFunction"
        assert lf.lnL == pytest.approx(-6755.451985437475, rel=1e-9)
        assert lf.nfp == 19
        got = lf.get_bin_probs()
        np.testing.assert_allclose(
            got["adaptive"][:3], [0.8138636520270726, 0.8723917957174725, 0.9922405018465282], rtol=5e-6
        )
        assert got.shape == (2, len(aln) // 3)

        lf = get_model("CNFGTR").make_likelihood_function(tree, bins=["0", "1", "2a", "2b"], digits=2, space=3)
        lf.set_alignment(aln)
        epsilon = 1e-6
        lf.set_param_rule("omega", bins=["0", "2a"], upper=1.0, init=1 - epsilon)
        lf.set_param_rule("omega", bins=["1", "2b"], is_constant=True, value=1.0)
        lf.set_param_rule("omega", bins=["2a", "2b"], edges=["Chimpanzee", "Human"], init=99, lower=1.0,
                          upper=100.0, is_constant=False)  # fmt: skip
        lf.apply_param_rules(_kat_rules("test_time_rate_het"))
        assert lf.lnL == pytest.approx(-6753.45662012937, rel=1e-5)
        assert lf.nfp == 21
        # and against the stock class in the same process, to the product tolerance
        ref = get_model("CNFGTR").make_likelihood_function(tree, bins=["0", "1", "2a", "2b"], backend="cpu")
        ref.set_alignment(aln)
        ref.set_param_rule("omega", bins=["0", "2a"], upper=1.0, init=1 - epsilon)
        ref.set_param_rule("omega", bins=["1", "2b"], is_constant=True, value=1.0)
        ref.set_param_rule("omega", bins=["2a", "2b"], edges=["Chimpanzee", "Human"], init=99, lower=1.0,
                           upper=100.0, is_constant=False)  # fmt: skip
        ref.apply_param_rules(_kat_rules("test_time_rate_het"))
        assert_same(ref, lf, rel=1e-9)


def _three_loci(tree_names, aln):
    third = len(aln) // 3
    return [aln[:third], aln[third : 2 * third], aln[2 * third :]]


def test_kat_loci_and_engine_cache():
    """ComparisonTests.test_loci (-9168.333116203343, nfp 2, :2046-2188); and with THREE loci the engines are
    built once, not once per evaluation (the per-Defn engine cache holds one engine per locus)"""
    from cogent3.recalculation.scope import ALL

    aln = load_aligned_seqs(os.path.join(DATA, "long_testseqs.fasta"), moltype="dna")
    half = len(aln) // 2
    tree = make_tree(tip_names=aln.names)
    with installed():
        lf = get_model("HKY85").make_likelihood_function(tree, loci=["1st-half", "2nd-half"], digits=2, space=3)
        lf.set_param_rule("length", is_independent=False)
        lf.set_param_rule("kappa", loci=ALL)
        lf.set_alignment([aln[:half], aln[half:]])
        lf.apply_param_rules(_kat_rules("test_loci"))
        assert lf.lnL == pytest.approx(-9168.333116203343, rel=1e-10)
        assert lf.nfp == 2

        names = ["a", "b", "c"]
        loci = _three_loci(aln.names, aln)
        lf = get_model("HKY85").make_likelihood_function(tree, loci=names)
        ref = get_model("HKY85").make_likelihood_function(tree, loci=names, backend="cpu")
        for f in (ref, lf):
            f.set_param_rule("length", is_independent=False)
            f.set_alignment(loci)
        assert_same(ref, lf)
        defn = lf._gpu_defn()
        created = defn.engines_created
        assert created == 3
        for k in range(4):
            for f in (ref, lf):
                f.set_param_rule("kappa", value=2.0 + 0.3 * k)
                f.set_param_rule("length", value=0.1 + 0.01 * k, is_independent=False)
            assert_same(ref, lf)
        calc, rcalc = lf.make_calculator(), ref.make_calculator()
        x = np.array(calc.get_value_array())
        for k in range(4):
            xk = x * (1.0 + 0.01 * (k + 1))
            a, b = rcalc(xk), calc(xk)
            assert abs(a - b) <= 1e-12 * abs(a)
        assert defn.engines_created == created, "engines were rebuilt between evaluations"


def test_brca1_optimise_and_per_pattern_values():
    """SURVEY §8c smoke vectors: C1 initial lnL -193293.46430542142, optimised -57533.56733281018; per-pattern
    root likelihoods against the stock class in the same process (<= 1e-12 absolute)"""
    aln, tree = get_dataset("brca1"), get_dataset("mammal-tree")
    tree = tree.get_sub_tree(aln.names)
    with installed():
        lf = get_model("HKY85").make_likelihood_function(tree)
        ref = get_model("HKY85").make_likelihood_function(tree, backend="cpu")
        for f in (ref, lf):
            f.set_alignment(aln)
        assert lf.get_log_likelihood() == pytest.approx(-193293.46430542142, rel=1e-12)
        lh, lh_ref = np.asarray(lf.get_param_value("lh")), np.asarray(ref.get_param_value("lh"))
        assert lh.shape == lh_ref.shape
        np.testing.assert_allclose(lh, lh_ref, rtol=0, atol=1e-12)
        np.testing.assert_allclose(lf.get_full_length_likelihoods(), ref.get_full_length_likelihoods(), rtol=0, atol=1e-12)
        lf.optimise(local=True, show_progress=False)
        assert lf.get_log_likelihood() == pytest.approx(-57533.56733281018, rel=1e-9)
        # the same point on the stock class
        ref.apply_param_rules(lf.get_param_rules())
        assert_same(ref, lf, rel=1e-10)
        # relative error of every per-pattern value at the optimum (values ~1e-20 .. 1e-2)
        a, b = np.asarray(lf.get_param_value("lh")), np.asarray(ref.get_param_value("lh"))
        assert np.max(np.abs(a - b) / np.maximum(b, 1e-300)) < 1e-10


def test_calculator_undo_after_a_rejected_step():
    """Calculator.change's 1-deep undo (recalculation/calculation.py:409-439): revisiting the previous point
    presents the previous objects again; the engine's identity diffing must return the previous value"""
    tree, aln = small_problem(n_tips=9, L=500, seed=21)
    mk = lambda: get_model("GTR", ordered_param="rate", distribution="gamma")  # noqa: E731
    ref, lf = pair(mk, tree, aln, bins=4)
    for f in (ref, lf):
        f.set_param_rule("bprobs", is_constant=True)
    rc, lc = ref.make_calculator(), lf.make_calculator()
    # the two Defn graphs list their optimisable parameters in different orders
    where = {(p.name, repr(p.scope)): i for i, p in enumerate(rc.opt_pars)}
    perm = np.array([where[(p.name, repr(p.scope))] for p in lc.opt_pars])
    x0 = np.array(rc.get_value_array())
    assert np.array_equal(x0[perm], lc.get_value_array())
    rng = np.random.default_rng(3)
    f0 = lc(x0[perm])
    assert abs(rc(x0) - f0) <= 1e-12 * abs(f0)
    for _ in range(12):
        i = int(rng.integers(len(x0)))
        x1 = x0.copy()
        x1[i] = x1[i] * (1.0 + 0.2 * rng.standard_normal()) + 0.01
        a, b = rc(x1), lc(x1[perm])
        assert abs(a - b) <= 1e-12 * abs(a)
        if rng.random() < 0.5:
            # "rejected": back to the previous point -- bit-identical to what it returned before
            assert lc(x0[perm]) == f0
            assert abs(rc(x0) - f0) <= 1e-12 * abs(f0)
        else:
            x0, f0 = x1, b


def test_pickle_and_rich_dict_round_trip():
    from cogent3.util.deserialise import deserialise_object

    tree, aln = small_problem(n_tips=6, L=200, seed=13)
    with installed():
        lf = get_model("HKY85").make_likelihood_function(tree)
        lf.set_alignment(aln)
        lf.set_param_rule("kappa", value=2.7)
        lnl = lf.get_log_likelihood()
        # serialised results stay interchangeable with stock cogent3 (no GPU needed to read them)
        data = lf.to_rich_dict()
        assert data["type"].endswith("parameter_controller.AlignmentLikelihoodFunction")
        json.dumps(data)
    back = deserialise_object(data)
    assert type(back).__name__ == "AlignmentLikelihoodFunction"
    assert back.get_log_likelihood() == pytest.approx(lnl, rel=1e-12)
    # pickle: the engine (device memory, ctypes handles) is rebuilt on demand in the new object
    clone = pickle.loads(pickle.dumps(lf))
    assert type(clone).__name__ == "CudaAlignmentLikelihoodFunction"
    assert clone.get_log_likelihood() == pytest.approx(lnl, rel=1e-13)
    clone.set_param_rule("kappa", value=3.1)
    lf.set_param_rule("kappa", value=3.1)
    assert clone.get_log_likelihood() == pytest.approx(lf.get_log_likelihood(), rel=1e-13)


def test_solved_models_and_measure_evals_per_second():
    """a9: ``rate_matrix_required=False`` models (closed-form calc_TN93_P, evolve/solved_models.py:43-49);
    and the reference's own throughput probe runs through the drop-in (recalculation/calculation.py:547-583)"""
    tree, aln = small_problem(n_tips=8, L=400, seed=17)
    for name in ("F81", "HKY85", "TN93"):
        ref, lf = pair(lambda: get_model(name, rate_matrix_required=False), tree, aln)  # noqa: B023
        for f in (ref, lf):
            if name != "F81":
                f.set_param_rule("kappa" if name == "HKY85" else "kappa_r", value=2.3)
            f.set_param_rule("length", edge=tree.get_tip_names()[2], value=0.21)
        assert_same(ref, lf, rel=1e-12)
        np.testing.assert_allclose(lf.get_full_length_likelihoods(), ref.get_full_length_likelihoods(), rtol=1e-12, atol=1e-300)
        # P(t) comes from ONE closed-form slot of the engine (c3_set_tn93), not from a host matrix per (edge, bin)
        st = next(iter(lf._gpu_defn()._engines.values()))
        assert [type(ex).__name__ for ex in st["slot_obj"].values()] == ["SolvedExponentiator"]
        assert not st["psub_obj"]
        np.testing.assert_allclose(lf.get_psub_for_edge(tree.get_tip_names()[2]).to_array(),
                                   ref.get_psub_for_edge(tree.get_tip_names()[2]).to_array(), rtol=0, atol=1e-15)  # fmt: skip
        eng = lf._engine()
        if hasattr(eng, "get_psub"):
            node = st["topo"].names.index(tree.get_tip_names()[2])
            np.testing.assert_allclose(eng.get_psub(node, 0), ref.get_psub_for_edge(tree.get_tip_names()[2]).to_array(),
                                       rtol=0, atol=1e-14)  # fmt: skip
    calc = lf.make_calculator()
    rate = calc.measure_evals_per_second(time_limit=0.2, wall=True)
    assert rate > 0


def test_patterns_sharded_over_devices_in_one_process():
    """``devices=[...]``: the root's pattern rows are dealt in contiguous blocks to one engine per device;
    lnL, per-pattern and per-column values are those of the unsharded function.  On a box
    with one GPU all shards live on device 0 (the host logic and the fused mailbox sum are the same)."""
    from cogent3_b200 import _lib

    n_dev = max(1, _lib.device_count()) if ON_GPU else 1
    tree, aln = small_problem(n_tips=10, L=900, seed=23)
    mk = lambda: get_model("GTR", ordered_param="rate", distribution="gamma")  # noqa: E731
    for n_shards in (2, 3, 8):
        devices = [i % n_dev for i in range(n_shards)]
        with installed():
            lf = mk().make_likelihood_function(tree, bins=4, devices=devices)
            one = mk().make_likelihood_function(tree, bins=4)
            ref = mk().make_likelihood_function(tree, bins=4, backend="cpu")
        for f in (ref, one, lf):
            f.set_alignment(aln)
            f.set_param_rule("bprobs", is_constant=True)
            f.set_param_rule("rate_shape", value=0.6)
            f.set_param_rule("A/G", value=2.9)
        assert_same(ref, lf, rel=1e-12)
        assert abs(lf.get_log_likelihood() - one.get_log_likelihood()) <= 1e-12 * abs(one.get_log_likelihood())
        eng = lf._engine()
        assert type(eng).__name__ == "MultiDeviceEngine" and len(eng.engines) == n_shards
        np.testing.assert_allclose(np.asarray(lf.get_param_value("lh", bin="bin1")),
                                   np.asarray(one.get_param_value("lh", bin="bin1")), rtol=0, atol=1e-15)  # fmt: skip
        np.testing.assert_allclose(lf.get_full_length_likelihoods(), ref.get_full_length_likelihoods(), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(lf.get_bin_probs().array, ref.get_bin_probs().array, atol=1e-13)
        for f in (ref, lf):
            f.set_param_rule("length", edge=tree.get_tip_names()[3], value=0.33)
        assert_same(ref, lf, rel=1e-12)
        calc, rcalc = lf.make_calculator(), ref.make_calculator()
        where = {(p.name, repr(p.scope)): i for i, p in enumerate(rcalc.opt_pars)}
        perm = np.array([where[(p.name, repr(p.scope))] for p in calc.opt_pars])
        x = np.array(rcalc.get_value_array())
        for k in range(3):
            xk = x * (1.0 + 0.01 * (k + 1))
            a, b = rcalc(xk), calc(xk[perm])
            assert abs(a - b) <= 1e-12 * abs(a)


def test_phylo_hmm_prunes_on_the_engine():
    """``sites_independent=False`` (PatchSiteDistribution / SiteHmm,
    evolve/likelihood_calculation.py:197-296): the per-bin root likelihoods come from the engine, the reduction
    over the sites is the reference's own hidden Markov model; lnL, posterior bin probabilities and the
    optimiser's path are the stock class's"""
    tree, aln = small_problem(n_tips=8, L=600, seed=31)
    mk = lambda: get_model("HKY85", ordered_param="rate", distribution="gamma")  # noqa: E731
    with installed():
        lf = mk().make_likelihood_function(tree, bins=4, sites_independent=False)
        ref = mk().make_likelihood_function(tree, bins=4, sites_independent=False, backend="cpu")
    assert type(lf).__name__ == "CudaAlignmentLikelihoodFunction"
    for f in (ref, lf):
        f.set_alignment(aln)
        f.set_param_rule("bprobs", is_constant=True)
        f.set_param_rule("rate_shape", value=0.7)
        f.set_param_rule("kappa", value=3.1)
        f.set_param_rule("bin_switch", value=0.3)
    assert_same(ref, lf, rel=1e-12)
    assert lf._engine().stats()["evaluations"] > 0
    np.testing.assert_allclose(lf.get_bin_probs().array, ref.get_bin_probs().array, rtol=1e-10, atol=1e-13)
    for f in (ref, lf):
        f.set_param_rule("bin_switch", value=0.05)
        f.set_param_rule("length", edge=tree.get_tip_names()[1], value=0.4)
    assert_same(ref, lf, rel=1e-12)
    assert lf.get_num_free_params() == ref.get_num_free_params()
    # the optimiser's view: the same points through both Calculators (their parameter orders differ)
    rc, lc = ref.make_calculator(), lf.make_calculator()
    where = {(p.name, repr(p.scope)): i for i, p in enumerate(rc.opt_pars)}
    perm = np.array([where[(p.name, repr(p.scope))] for p in lc.opt_pars])
    x = np.array(rc.get_value_array())
    rng = np.random.default_rng(1)
    for _ in range(6):
        xk = x * (1.0 + 0.05 * rng.standard_normal(len(x)))
        a, b = rc(xk), lc(xk[perm])
        assert abs(a - b) <= 1e-11 * abs(a)


def test_many_likelihood_functions_optimised_in_lockstep():
    """N independent cogent3 optimisations (model fits over many alignments, bootstrap replicates)
    run concurrently, their evaluations batched into one device call per round (cogent3_b200/batch.py); every
    function ends where its own sequential ``lf.optimise()`` ends"""
    from cogent3_b200.batch import log_likelihoods, optimise_many

    problems = [small_problem(n_tips=6, L=150 + 10 * k, seed=40 + k) for k in range(6)]

    def build():
        out = []
        for tree, aln in problems:
            lf = make_likelihood_function(get_model("HKY85"), tree, engine_factory=FACTORY)
            lf.set_alignment(aln)
            out.append(lf)
        return out

    seq, par = build(), build()
    for lf in seq:
        lf.optimise(local=True, show_progress=False, max_evaluations=400, limit_action="ignore")
    _, stats = optimise_many(par, return_stats=True, local=True, max_evaluations=400, limit_action="ignore")
    assert stats["mean_batch"] > 2.0, stats  # the evaluations really met in batches
    for a, b in zip(seq, par, strict=True):
        assert abs(a.get_log_likelihood() - b.get_log_likelihood()) <= 1e-9 * abs(a.get_log_likelihood())
        assert b.get_param_value("kappa") == pytest.approx(a.get_param_value("kappa"), rel=1e-6)
    np.testing.assert_allclose(log_likelihoods(par), [lf.get_log_likelihood() for lf in par], rtol=1e-13)


def test_pairwise_distances_in_lockstep():
    """SURVEY §8 f-4, the pairwise path: ``EstimateDistances`` (evolve/distance.py:125-234) fits one two-taxon
    likelihood function per pair; with the hook installed each is a CUDA function, and
    ``batch.run_estimate_distances`` optimises the pairs concurrently -- same distances as the stock run"""
    from cogent3.evolve.distance import EstimateDistances

    from cogent3_b200.batch import run_estimate_distances

    _tree, aln = small_problem(n_tips=6, L=300, seed=51)
    stock = EstimateDistances(aln, get_model("HKY85"))
    stock.run(show_progress=False)
    want = stock.get_pairwise_distances()
    with installed():
        d = EstimateDistances(aln, get_model("HKY85"))
        stats = run_estimate_distances(d)
        got = d.get_pairwise_distances()
        one = EstimateDistances(aln, get_model("HKY85"))
        one.run(show_progress=False)  # the stock loop, every fit through the drop-in
        seq = one.get_pairwise_distances()
    assert stats["mean_batch"] > 3.0, stats
    names = aln.names
    for a in names:
        for b in names:
            if a != b:
                assert got[(a, b)] == pytest.approx(want[(a, b)], rel=1e-5, abs=1e-7)
                assert seq[(a, b)] == pytest.approx(want[(a, b)], rel=1e-5, abs=1e-7)


def test_oracle_hmm_restatement_is_pinned_to_the_live_reference():
    """oracle.log_dot_reduce / patch_weighted_sum_lhs (the checker of c3_hmm_reduce in tests/test_gpu_parity.py)
    against the reference's SiteHmm on the stock class: the same bits"""
    if ON_GPU:
        pytest.skip("pins the CPU oracle: build container")
    from oracle import oracle

    tree, aln = small_problem(n_tips=6, L=300, seed=3)
    sm = get_model("HKY85", ordered_param="rate", distribution="gamma")
    lf = sm.make_likelihood_function(tree, bins=4, sites_independent=False)
    lf.set_alignment(aln)
    lf.set_param_rule("bin_switch", value=0.2)
    lf.set_param_rule("rate_shape", value=0.8)
    want = lf.get_log_likelihood()
    blh = lf.get_param_value("bindex")
    d = blh.distrib
    lhs = [np.asarray(lf.get_param_value("lh", bin=b)) for b in lf.bin_names]
    plhs = np.ascontiguousarray(oracle.patch_weighted_sum_lhs(d.alloc, d.bprobs, lhs).T)
    tm = d.transition_matrix
    assert oracle.log_dot_reduce(blh.root.index, tm.StationaryProbs, tm.Matrix, plhs) == want
