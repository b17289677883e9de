"""Developer tool: two full evaluations of a bench config (for ncu captures)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--columns", type=int, default=None)
ap.add_argument("--kind", default="evolved")
ap.add_argument("--evals", type=int, default=3)
a = ap.parse_args()
problem = bench.build_problem(a.config, a.columns, a.kind, 0)
lf, _ = bench.make_lf(problem)
base, rates = lf.lengths.copy(), lf.rates()
print("lnL", lf.get_log_likelihood())
for k in range(a.evals):
    v = lf.engine.update(np.nan_to_num(bench.step_lengths(base, k))[:, None] * rates[None, :])
print("lnL", v, lf.engine.stats())
