"""One-off validation: the drop-in glue with the CUDA engine under the REAL cogent3.

Needs an unmodified copy of the reference package importable (baseline/_ref, git-ignored;
created with `cp -r /root/reference/src/cogent3 baseline/_ref/`), the import shims of
tests/_shims, and a GPU.  Prints reference-vs-CUDA numbers for the four model families.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
os.environ.setdefault("COGENT3_REFERENCE_ROOT", os.path.join(ROOT, "baseline", "_ref"))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "_shims"), os.path.join(ROOT, "baseline", "_ref"),
                os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
from cogent3 import get_dataset, get_model, make_aligned_seqs, make_tree  # noqa: E402

from cogent3_b200.dropin import make_likelihood_function  # noqa: E402


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



def report(tag, ref, lf, check_lh=True):
    a, b = ref.get_log_likelihood(), lf.get_log_likelihood()
    d = np.abs(ref.get_full_length_likelihoods() - lf.get_full_length_likelihoods()).max()
    print(f"{tag}: ref lnL {a!r}  cuda lnL {b!r}  rel {abs(a - b) / abs(a):.2e}  max|dlh| {d:.2e}")
    assert abs(a - b) <= 1e-9 * abs(a) and (not check_lh or d <= 1e-12)


aln, tree = get_dataset("brca1"), get_dataset("mammal-tree")
ref = get_model("HKY85").make_likelihood_function(tree)
ref.set_alignment(aln)
lf = make_likelihood_function(get_model("HKY85"), tree)
lf.set_alignment(aln)
report("C1 brca1 HKY85 initial", ref, lf)
t = time.time()
lf.optimise(local=True, show_progress=False)
t_gpu = time.time() - t
t = time.time()
ref.optimise(local=True, show_progress=False)
t_cpu = time.time() - t
print(f"C1 optimise: cuda {t_gpu:.2f}s lnL {lf.get_log_likelihood()!r} | reference {t_cpu:.2f}s "
      f"lnL {ref.get_log_likelihood()!r} (survey: -57533.56733281018)")
report("C1 brca1 HKY85 optimised (independent optimiser runs)", ref, lf, check_lh=False)
for f in (ref, lf):
    print("  evals/s:", f.measure_evals_per_second(time_limit=2, wall=True))

tree, aln = small_problem(n_tips=20, L=5000, seed=9)
mk = lambda: get_model("GTR", ordered_param="rate", distribution="gamma")  # noqa: E731
ref = mk().make_likelihood_function(tree, bins=4)
ref.set_alignment(aln)
lf = make_likelihood_function(mk(), tree, bins=4)
lf.set_alignment(aln)
for f in (ref, lf):
    f.set_param_rule("bprobs", is_constant=True)
    f.set_param_rule("rate_shape", value=0.5)
report("C2-like GTR+G4 20x5000", ref, lf)

tree, aln = small_problem(n_tips=12, L=3000, seed=10)
ref = get_model("GN").make_likelihood_function(tree, expm="pade")
ref.set_alignment(aln)
lf = make_likelihood_function(get_model("GN"), tree, expm="pade")
lf.set_alignment(aln)
for f in (ref, lf):
    f.set_time_heterogeneity(is_independent=True, upper=50)
report("C3-like GN per-edge Q Pade 12x3000", ref, lf)

tree, aln = small_problem(n_tips=10, L=900, n_states=61, seed=12)
ref = get_model("CNFGTR").make_likelihood_function(tree)
ref.set_alignment(aln)
lf = make_likelihood_function(get_model("CNFGTR"), tree)
lf.set_alignment(aln)
for f in (ref, lf):
    f.set_param_rule("omega", value=0.4)
report("C4-like CNFGTR 10x900 codons", ref, lf)
# ---- set_alignment / make_calculator wall time at a size where the reference's Python loops hurt ----
tree, aln = small_problem(n_tips=64, L=100_000, seed=13)
times = {}
for tag, make in (("reference", lambda: mk().make_likelihood_function(tree, bins=4)),
                  ("drop-in", lambda: make_likelihood_function(mk(), tree, bins=4))):
    t0 = time.perf_counter()
    f = make()
    f.set_alignment(aln)
    t1 = time.perf_counter()
    f.set_param_rule("bprobs", is_constant=True)
    calc = f.make_calculator()
    t2 = time.perf_counter()
    x = calc.get_value_array()
    calc(x)
    n, t3 = 0, time.perf_counter()
    while time.perf_counter() - t3 < 2.0:
        calc([v * (1 + 1e-3 * ((n % 5) + 1)) for v in x])
        n += 1
    t4 = time.perf_counter()
    times[tag] = (t1 - t0, t2 - t1, (t4 - t3) / n, f.get_log_likelihood())
    print(f"64 x 100k GTR+G4, {tag}: set_alignment {t1 - t0:.2f} s, make_calculator {t2 - t1:.2f} s, "
          f"full evaluation through Calculator {1e3 * (t4 - t3) / n:.2f} ms, lnL {times[tag][3]!r}")
a, b = times["reference"][3], times["drop-in"][3]
assert abs(a - b) <= 1e-9 * abs(a)

# ---- a tree deep enough to underflow FP64: the reference returns -inf, the engine rescales ------
from cogent3_b200.synthetic import dense_alignment  # noqa: E402
from cogent3_b200.tree import random_join_tree  # noqa: E402

rng = np.random.default_rng(11)
topo, lengths = random_join_tree(700, rng)
codes = dense_alignment(topo, 300, 4, rng)
data = {n: "".join("TCAG"[c] for c in row) for n, row in zip(topo.tip_names, codes, strict=True)}
tree, aln = make_tree(topo.to_newick(np.asarray(lengths) * 4.0)), make_aligned_seqs(data, moltype="dna")
ref = get_model("HKY85").make_likelihood_function(tree)
ref.set_alignment(aln)
lf = make_likelihood_function(get_model("HKY85"), tree)
lf.set_alignment(aln)
a, b = ref.get_log_likelihood(), lf.get_log_likelihood()
print(f"700 taxa x 300 i.i.d. columns, HKY85: reference lnL {a!r}, drop-in lnL {b!r} "
      f"(per-pattern rescaling switched on automatically: {lf._engine().rescaling()})")
assert not np.isfinite(a) and np.isfinite(b)
print("drop-in + CUDA engine + real cogent3: OK")
