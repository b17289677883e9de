"""Generate golden vectors from the LIVE reference (cogent3 under /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    PYTHONPATH=/root/reference/src:tests/_shims NUMBA_CACHE_DIR=/tmp/numba_cache \
        python tests/golden/make_golden.py

Writes tests/golden/<case>.npz.  Each file holds, for one (model, tree,
alignment, parameter vector): the tree topology as the reference walks it, the
alignment as symbol codes, every per-node pattern table of the reference's
``lht`` (uniq / indexes / counts / index / tip indicator rows), Q, every
P[bin, edge], per-pattern ``lh`` per bin and the total log-likelihood.
The fixtures are the parity anchor for oracle/ and, through it, for the CUDA path.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from cogent3 import get_dataset, get_model, make_aligned_seqs, make_tree  # noqa: E402

from cogent3_b200 import models as hostmodels  # noqa: E402
from cogent3_b200.synthetic import evolved_alignment  # noqa: E402
from cogent3_b200.tree import random_join_tree  # noqa: E402


def walk_tree(tree):
    names, children, ids = [], [], {}
    for e in tree.postorder():
        ids[e.name] = len(names)
        names.append(e.name)
        children.append([ids[c.name] for c in e.children])
    return names, children


def dump_lht(lht, names, out):
    for n, name in enumerate(names):
        nd = lht.get_edge(name)
        out[f"n{n}_counts"] = np.asarray(nd.counts, float)
        out[f"n{n}_index"] = np.asarray(nd.index, np.int64)
        if hasattr(nd, "indexes"):
            out[f"n{n}_uniq"] = np.asarray(nd.uniq, np.int64)
            out[f"n{n}_indexes"] = np.asarray(nd.indexes, np.int64)
            assert nd.indexes.dtype == np.int64
        else:
            out[f"n{n}_table"] = np.asarray(nd.input_likelihoods, float)
            out[f"n{n}_motifs"] = np.array([str(u) for u in nd.uniq])
        out[f"n{n}_ambig"] = np.asarray(nd.ambig, float)


def dump_case(case, lf, sm, codes, symbols, extra_meta):
    tree = lf.tree
    names, children = walk_tree(tree)
    out = {}
    lht = lf.get_param_value("lht")
    dump_lht(lht, names, out)
    bins = list(lf.bin_names)
    M = len(sm.get_alphabet())
    n_nodes = len(names)
    psubs = np.zeros((len(bins), n_nodes, M, M))
    Qs = np.zeros((len(bins), n_nodes, M, M))
    lengths = np.full(n_nodes, np.nan)
    dists = np.full((len(bins), n_nodes), np.nan)
    multi = len(bins) > 1
    for n, name in enumerate(names[:-1]):
        lengths[n] = lf.get_param_value("length", edge=name)
        for b, bn in enumerate(bins):
            kw = {"bin": bn} if multi else {}
            psubs[b, n] = lf.get_param_value("psubs", edge=name, **kw)
            Qs[b, n] = lf.get_param_value("Q", edge=name, **kw)
            if multi and "rate" in lf.get_param_names():
                dists[b, n] = lengths[n] * lf.get_param_value("rate", bin=bn)
            else:
                dists[b, n] = lengths[n]
    lh = np.stack(
        [np.asarray(lf.get_param_value("lh", **({"bin": bn} if multi else {}))) for bn in bins]
    )
    if list(symbols) == hostmodels.dna_symbol_rows()[0]:
        codes = hostmodels.recode_dna_gaps(codes)
    out.update(psubs=psubs, Q=Qs, lengths=lengths, distances=dists, lh=lh, codes=codes)
    out["inst_mask"] = np.asarray(sm._instantaneous_mask_f, float)
    for pname, mask in getattr(sm, "predicate_masks", {}).items():
        out["mask_" + str(pname).replace("/", "_").replace(">", "to")] = np.asarray(mask, float)
    mprobs = np.asarray(lf.get_param_value("mprobs"), float)
    meta = {
        "case": case,
        "names": names,
        "children": children,
        "bins": len(bins),
        "alphabet": [str(m) for m in sm.get_alphabet()],
        "symbols": list(symbols),
        "mprobs": mprobs.tolist(),
        "bprobs": (np.asarray(lf.get_param_value("bprobs"), float).tolist() if multi else [1.0]),
        "lnL": float(lf.get_log_likelihood()),
        "parameter_order": [str(p) for p in getattr(sm, "parameter_order", [])],
        "cogent3_version": __import__("cogent3").__version__,
        "numpy_version": np.__version__,
    }
    meta.update(extra_meta)
    out["meta"] = np.array(json.dumps(meta))
    path = os.path.join(HERE, f"{case}.npz")
    np.savez_compressed(path, **out)
    print(f"{case}: lnL={meta['lnL']!r} root_patterns={lh.shape[1]} -> {os.path.getsize(path)} B")


def codes_to_aln(codes, symbols, tip_names, moltype="dna"):
    data = {
        name: "".join(symbols[c] for c in row) for name, row in zip(tip_names, codes, strict=True)
    }
    return make_aligned_seqs(data, moltype=moltype)


def synthetic(n_tips, L, n_states, seed, ambiguity=0.0, symbols=None):
    rng = np.random.default_rng(seed)
    topo, lengths = random_join_tree(n_tips, rng)
    codes = evolved_alignment(topo, lengths, L, n_states, rng).astype(np.int64)
    if ambiguity:
        amb = rng.random(codes.shape) < ambiguity
        codes[amb] = rng.integers(n_states, len(symbols), size=int(amb.sum()))
        # the stored codes are what the reference sees after recode_gaps=True
        # ('-'/'?' -> 'N'); the strings handed to cogent3 below still hold the
        # gap characters, so the reference does that recoding itself
    return topo, lengths, codes


def set_lengths(lf, topo_names, lengths):
    with lf.updates_postponed():
        for name, ln in zip(topo_names, lengths, strict=True):
            if name != "root":
                lf.set_param_rule("length", edge=name, value=float(ln))


GTR_RATES = {"A/C": 1.1, "A/G": 3.2, "A/T": 0.7, "C/G": 0.9, "C/T": 3.6}


def case_brca1_hky85():
    """C1: HKY85 on the bundled brca1 alignment + mammal tree."""
    aln = get_dataset("brca1")
    tree = get_dataset("mammal-tree")
    sm = get_model("HKY85")
    lf = sm.make_likelihood_function(tree)
    lf.set_alignment(aln)
    symbols, _ = hostmodels.dna_symbol_rows()
    names, _ = walk_tree(lf.tree)
    tips = [n for n in names if n in aln.names]
    codes = np.stack([hostmodels.encode_dna(str(aln.get_gapped_seq(t))) for t in tips])
    dump_case("c1_brca1_hky85_default", lf, sm, codes, symbols, {"model": "HKY85", "tips": tips,
              "params": {"kappa": float(lf.get_param_value("kappa"))}, "expm": "either"})
    # a second parameter point (non-default kappa and lengths from the tree file)
    lf.set_param_rule("kappa", value=4.3052)
    rng = np.random.default_rng(7)
    with lf.updates_postponed():
        for name in names[:-1]:
            lf.set_param_rule("length", edge=name, value=float(rng.uniform(0.005, 0.2)))
    dump_case("c1_brca1_hky85_point2", lf, sm, codes, symbols, {"model": "HKY85", "tips": tips,
              "params": {"kappa": 4.3052}, "expm": "either"})


def case_gtr_gamma(case, n_tips, L, seed, ambiguity=0.0):
    symbols, _ = hostmodels.dna_symbol_rows()
    topo, lengths, codes = synthetic(n_tips, L, 4, seed, ambiguity, symbols)
    tree = make_tree(topo.to_newick(lengths))
    aln = codes_to_aln(codes, symbols, topo.tip_names)
    sm = get_model("GTR", ordered_param="rate", distribution="gamma")
    lf = sm.make_likelihood_function(tree, bins=4)
    lf.set_alignment(aln)
    lf.set_param_rule("bprobs", is_constant=True)
    lf.set_param_rule("rate_shape", value=0.5)
    for p, v in GTR_RATES.items():
        lf.set_param_rule(p, value=v)
    names, _ = walk_tree(lf.tree)
    dump_case(case, lf, sm, codes, symbols, {"model": "GTR", "tips": topo.tip_names,
              "params": GTR_RATES, "rate_shape": 0.5, "expm": "either"})


def case_hky_gamma(case, n_tips, L, seed):
    symbols, _ = hostmodels.dna_symbol_rows()
    topo, lengths, codes = synthetic(n_tips, L, 4, seed)
    tree = make_tree(topo.to_newick(lengths))
    aln = codes_to_aln(codes, symbols, topo.tip_names)
    sm = get_model("HKY85", ordered_param="rate", distribution="gamma")
    lf = sm.make_likelihood_function(tree, bins=4)
    lf.set_alignment(aln)
    lf.set_param_rule("bprobs", is_constant=True)
    lf.set_param_rule("rate_shape", value=0.7)
    lf.set_param_rule("kappa", value=4.0)
    dump_case(case, lf, sm, codes, symbols, {"model": "HKY85", "tips": topo.tip_names,
              "params": {"kappa": 4.0}, "rate_shape": 0.7, "expm": "either"})


def case_gn(case, n_tips, L, seed, expm):
    symbols, _ = hostmodels.dna_symbol_rows()
    topo, lengths, codes = synthetic(n_tips, L, 4, seed)
    tree = make_tree(topo.to_newick(lengths))
    aln = codes_to_aln(codes, symbols, topo.tip_names)
    sm = get_model("GN")
    lf = sm.make_likelihood_function(tree, expm=expm)
    lf.set_alignment(aln)
    lf.set_time_heterogeneity(is_independent=True, upper=50)
    rng = np.random.default_rng(seed + 1)
    names, _ = walk_tree(lf.tree)
    per_edge = {}
    pnames = [str(p) for p in sm.parameter_order]
    with lf.updates_postponed():
        for name in names[:-1]:
            per_edge[name] = {}
            for p in pnames:
                v = float(rng.uniform(0.3, 3.0))
                lf.set_param_rule(p, edge=name, value=v)
                per_edge[name][p] = v
    dump_case(case, lf, sm, codes, symbols, {"model": "GN", "tips": topo.tip_names,
              "edge_params": per_edge, "expm": expm})


def case_cnfgtr(case, n_tips, L, seed):
    symbols, _ = hostmodels.codon_symbol_rows()
    topo, lengths, codes = synthetic(n_tips, L, 61, seed)
    tree = make_tree(topo.to_newick(lengths))
    aln = codes_to_aln(codes, symbols, topo.tip_names)
    sm = get_model("CNFGTR")
    assert [str(m) for m in sm.get_alphabet()] == list(hostmodels.SENSE_CODONS)
    lf = sm.make_likelihood_function(tree)
    lf.set_alignment(aln)
    params = dict(GTR_RATES, omega=0.4)
    for p, v in params.items():
        lf.set_param_rule(p, value=v)
    dump_case(case, lf, sm, codes, symbols, {"model": "CNFGTR", "tips": topo.tip_names,
              "params": params, "expm": "either"})


def codes_from_leaves(lf, tips):
    """symbol table + codes recovered from the reference's own tip tables (for
    alignments whose motifs are not in the canned encoders, e.g. partial codons)."""
    leaves = lf.get_param_value("leaf_likelihoods")
    symbols, rows, codes = [], [], []
    for t in tips:
        leaf = leaves[t]
        ids = []
        for u, motif in enumerate(leaf.uniq[:-1]):
            motif = str(motif)
            if motif not in symbols:
                symbols.append(motif)
                rows.append(np.asarray(leaf.input_likelihoods[u], float))
            ids.append(symbols.index(motif))
        codes.append(np.asarray(ids, np.int64)[np.asarray(leaf.index)])
    return symbols, np.stack(rows), np.stack(codes)


def case_known_answers():
    """the reference test-suite's pinned lnL constants, re-derived live:
    tests/test_evolve/test_likelihood_function.py:145-160 (setUp),
    :245-270 (test_codon -103.05742415448259, test_nucleotide -148.6455087258624)."""
    from cogent3 import load_aligned_seqs
    from cogent3.evolve import substitution_model

    otus = ["Human", "Mouse", "HowlerMon"]
    aln = load_aligned_seqs("/root/reference/tests/data/brca1.fasta", moltype="dna")
    aln = aln.take_seqs(otus)[0:42]
    tree = make_tree(tip_names=otus)

    sm = substitution_model.TimeReversibleNucleotide(
        equal_motif_probs=True, motif_probs=None, predicates={"kappa": "transition"}
    )
    lf = sm.make_likelihood_function(tree)
    lf.set_alignment(aln)
    lf.set_param_rule("length", value=1.0, is_constant=True)
    lf.set_param_rule("kappa", value=3.0, is_constant=True)
    assert abs(lf.get_log_likelihood() - -148.6455087258624) < 1e-9
    names, _ = walk_tree(lf.tree)
    tips = [n for n in names if n in otus]
    symbols, rows, codes = codes_from_leaves(lf, tips)
    dump_case("kat_nucleotide", lf, sm, codes, symbols, {"model": "HKY85", "tips": tips,
              "params": {"kappa": 3.0}, "expm": "either", "symbol_rows": rows.tolist(),
              "pinned_lnL": -148.6455087258624})

    sm = substitution_model.TimeReversibleCodon(
        equal_motif_probs=True, motif_probs=None,
        predicates={"kappa": "transition", "omega": "replacement"}, mprob_model="tuple",
    )
    lf = sm.make_likelihood_function(tree)
    lf.set_alignment(aln)
    lf.set_param_rule("length", value=1.0, is_constant=True)
    lf.set_param_rule("kappa", value=3.0, is_constant=True)
    lf.set_param_rule("omega", value=0.5, is_constant=True)
    assert abs(lf.get_log_likelihood() - -103.05742415448259) < 1e-9
    symbols, rows, codes = codes_from_leaves(lf, tips)
    dump_case("kat_codon", lf, sm, codes, symbols, {"model": "GY94", "tips": tips,
              "params": {"kappa": 3.0, "omega": 0.5}, "expm": "either",
              "symbol_rows": rows.tolist(), "pinned_lnL": -103.05742415448259})


def case_known_answers_large_alphabets():
    """16- and 20-state known answers of the reference test-suite
    (tests/test_evolve/test_likelihood_function.py:298-316: test_dinucleotide
    -118.35045332768402, test_protein -91.35162044257062).  The host side of this repo does
    not restate those models; the fixtures carry Q, so the tests drive the engine with the
    reference's Q (eigendecomposed on the host like the reference does) -- "model": "supplied-Q"."""
    from cogent3 import load_aligned_seqs
    from cogent3.evolve import substitution_model

    otus = ["Human", "Mouse", "HowlerMon"]
    aln = load_aligned_seqs("/root/reference/tests/data/brca1.fasta", moltype="dna")
    aln = aln.take_seqs(otus)[0:42]
    tree = make_tree(tip_names=otus)

    sm = substitution_model.TimeReversibleDinucleotide(
        equal_motif_probs=True, predicates={"kappa": "transition"}, mprob_model="tuple"
    )
    lf = sm.make_likelihood_function(tree)
    lf.set_alignment(aln)
    lf.set_param_rule("length", value=1.0, is_constant=True)
    lf.set_param_rule("kappa", value=3.0, is_constant=True)
    assert abs(lf.get_log_likelihood() - -118.35045332768402) < 1e-9
    names, _ = walk_tree(lf.tree)
    tips = [n for n in names if n in otus]
    symbols, rows, codes = codes_from_leaves(lf, tips)
    dump_case("kat_dinucleotide", lf, sm, codes, symbols, {"model": "supplied-Q", "tips": tips,
              "params": {}, "expm": "either", "symbol_rows": rows.tolist(),
              "pinned_lnL": -118.35045332768402})

    sm = substitution_model.TimeReversibleProtein(equal_motif_probs=True)
    lf = sm.make_likelihood_function(tree)
    lf.set_alignment(aln.get_translation())
    lf.set_param_rule("length", value=1.0, is_constant=True)
    assert abs(lf.get_log_likelihood() - -91.35162044257062) < 1e-9
    symbols, rows, codes = codes_from_leaves(lf, tips)
    dump_case("kat_protein", lf, sm, codes, symbols, {"model": "supplied-Q", "tips": tips,
              "params": {}, "expm": "either", "symbol_rows": rows.tolist(),
              "pinned_lnL": -91.35162044257062})


def main():
    os.makedirs(HERE, exist_ok=True)
    if "--large-alphabets-only" in sys.argv:
        case_known_answers_large_alphabets()
        return
    case_brca1_hky85()
    case_gtr_gamma("c2_gtr_gamma4_small", 8, 300, 11)
    case_gtr_gamma("c2_gtr_gamma4_ambig", 7, 200, 12, ambiguity=0.08)
    case_gtr_gamma("c2_gtr_gamma4_mid", 16, 4000, 13)
    case_hky_gamma("c5_hky85_gamma4_small", 12, 500, 21)
    case_gn("c3_gn_pade_small", 6, 200, 31, "pade")
    case_gn("c3_gn_either_small", 9, 400, 32, "either")
    case_cnfgtr("c4_cnfgtr_small", 6, 120, 41)
    case_cnfgtr("c4_cnfgtr_mid", 10, 600, 42)
    case_known_answers()
    case_known_answers_large_alphabets()


if __name__ == "__main__":
    main()
