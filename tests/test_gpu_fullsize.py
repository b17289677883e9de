"""Parity at BASELINE.json's FULL sizes, where the oracle cannot run the whole problem in seconds:
size-independent properties of the likelihood, plus a direct oracle check on a random sample of
the columns.

* additivity: columns are independent given the parameters (the "sites independent" path,
  evolve/likelihood_calculation.py:125-167), so lnL(whole) = lnL(first block) + lnL(rest) even
  though the three engines build entirely different pattern tables;
* column order does not matter: a permuted alignment has the same lnL (every node's first-seen
  pattern numbering changes);
* sum over columns of log(per-column likelihood) = lnL: ties the accessor path
  (get_full_length_likelihoods, evolve/likelihood_tree.py:93-94) to the reduction kernel;
* one branch changed: the dirty-path update is bit-identical to a full sweep;
* a random sample of columns: the oracle (numpy restatement of the reference) evaluates just
  those columns with the same parameters; per-column likelihoods must match the engine's
  values for the same columns of the full problem to 1e-12 absolute (the north-star bar) and
  1e-12 relative for 4 states (1e-7 for codons, see the comment in the test).
"""

import numpy as np
import pytest
from oracle_engine import OracleEngine

from cogent3_b200 import models as hostmodels
from cogent3_b200.likelihood import LikelihoodFunction
from cogent3_b200.synthetic import make_workload

pytestmark = pytest.mark.gpu

GTR_RATES = {"A/C": 1.1, "A/G": 3.2, "A/T": 0.7, "C/G": 0.9, "C/T": 3.6}
# (model, taxa, columns, bins, parameters, expm): the shapes of BASELINE.json configs[1..3]
FULL = {
    "c2": ("GTR", 64, 1_000_000, 4, GTR_RATES, "either"),
    "c3": ("GN", 128, 2_000_000, 1, None, "pade"),
    "c4": ("CNFGTR", 32, 200_000, 1, dict(GTR_RATES, omega=0.4), "either"),
}
SEED = 20251210


def _configure(lf, cfg, lengths, mprobs):
    model, _n, _L, bins, params, _expm = FULL[cfg]
    lf.set_motif_probs(mprobs)
    lf.lengths[:-1] = np.asarray(lengths)[:-1]
    if params is None:  # GN: every edge its own rate terms (set_time_heterogeneity)
        rng = np.random.default_rng(SEED + 3)
        for n in lf.topo.edges:
            for p in lf.model.parameter_order:
                lf.set_param_rule(p, edge=n, value=float(rng.uniform(0.3, 3.0)))
    else:
        for p, v in params.items():
            lf.set_param_rule(p, value=v)
    if bins > 1:
        lf.set_param_rule("rate_shape", value=0.5)
    return lf


def _build(cfg, topo, lengths, codes, mprobs, factory=None):
    model, _n, _L, bins, _params, expm = FULL[cfg]
    m = hostmodels.get_model(model)
    lf = LikelihoodFunction(m, topo, bins=bins, expm=expm, engine_factory=factory)
    lf.set_alignment(codes, np.eye(m.n_states))
    return _configure(lf, cfg, lengths, mprobs)


@pytest.mark.parametrize("cfg", ["c2", "c4", "c3"])
def test_full_size_properties(cfg):
    model, n_tips, L, bins, _params, _expm = FULL[cfg]
    m = hostmodels.get_model(model)
    topo, lengths, codes = make_workload(n_tips, L, m.n_states, kind="evolved", seed=SEED)
    mprobs = np.bincount(codes.ravel(), minlength=m.n_states).astype(float)
    mprobs /= mprobs.sum()

    whole = _build(cfg, topo, lengths, codes, mprobs)
    total = whole.get_log_likelihood()
    assert np.isfinite(total)
    assert whole.engine.rescaling() == (False, 0)  # the plain products, as in the reference
    assert whole.patterns.n_columns == L

    # --- sum over columns of log(per-column likelihood) == lnL --------------------------------
    site = np.stack([whole.engine.get_site_lh(b) for b in range(bins)])  # [K, L]
    mix = site[0] if bins == 1 else (whole.bprobs[:, None] * site).sum(axis=0)
    assert abs(float(np.sum(np.log(mix))) - total) <= 1e-11 * abs(total)

    # --- a random sample of columns against the oracle -----------------------------------------
    rng = np.random.default_rng(7)
    cols = np.sort(rng.choice(L, size=1500 if m.n_states == 4 else 300, replace=False))
    ref = _build(cfg, topo, lengths, codes[:, cols].astype(np.int64), mprobs, factory=OracleEngine)
    ref.get_log_likelihood()
    # the north-star bar is 1e-12 ABSOLUTE per pattern; relative is the stricter statement at these
    # magnitudes (1e-30 .. 1e-60).  Codon columns whose likelihood flows through eigen-path P entries
    # of ~1e-7 -- good to ~1e-16 absolute in either implementation -- agree to ~1e-9 relative only.
    rel = 1e-12 if m.n_states == 4 else 1e-7
    for b in range(bins):
        want = ref.engine.get_lh(b)[ref.patterns.root.index]  # per column of the sample
        got = site[b][cols]
        assert np.max(np.abs(got - want)) <= 1e-12
        assert np.max(np.abs(got - want) / want) <= rel, (cfg, b)
    sample_lnL = float(np.sum(np.log(mix[cols])))
    assert abs(sample_lnL - ref.get_log_likelihood()) <= 1e-9 * abs(sample_lnL)

    # --- one branch changed: dirty path == full sweep, bit for bit -----------------------------
    name = topo.names[int(topo.edges[len(topo.edges) // 2])]
    whole.set_param_rule("length", edge=name, value=0.123)
    part = whole.get_log_likelihood()
    st = whole.engine.stats()
    assert 0 < st["last_nodes"] < topo.n_nodes - n_tips
    whole.engine.invalidate()
    assert whole.engine.evaluate() == part
    whole.set_param_rule("length", edge=name, value=float(lengths[topo.index(name)]))
    assert whole.get_log_likelihood() == total
    del whole, site, mix

    # --- additivity over column blocks and invariance to column order --------------------------
    cut = L // 3 + 17
    a = _build(cfg, topo, lengths, codes[:, :cut], mprobs).get_log_likelihood()
    b = _build(cfg, topo, lengths, codes[:, cut:], mprobs).get_log_likelihood()
    assert abs((a + b) - total) <= 1e-12 * abs(total), (a, b, total)
    perm = np.random.default_rng(3).permutation(L)
    shuffled = _build(cfg, topo, lengths, np.ascontiguousarray(codes[:, perm]), mprobs).get_log_likelihood()
    assert abs(shuffled - total) <= 1e-12 * abs(total)
