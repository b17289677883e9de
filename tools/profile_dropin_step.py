"""Developer tool (GPU box): where a full-change Calculator step through the drop-in spends its host time."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

assert bench.import_cogent3() is not None
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
cols = int(sys.argv[2]) if len(sys.argv) > 2 else None
problem = bench.build_problem(cfg, cols, "evolved", 0)
lf, _ = bench.build_cogent3_lf(problem, cfg, backend="cuda")
lf.get_log_likelihood()
calc = lf.make_calculator()
x = np.array(calc.get_value_array())
for k in range(30):
    calc(bench.calculator_points(x, k))
N = 300
t0 = time.perf_counter()
for k in range(N):
    calc(bench.calculator_points(x, k))
print("us per full-change step:", (time.perf_counter() - t0) / N * 1e6)
is_length = np.array([p.name == "length" for p in calc.opt_pars])
t0 = time.perf_counter()
for k in range(N):
    xk = x.copy()
    xk[is_length] = bench.calculator_points(x, k)[is_length]
    calc(xk)
print("us per lengths-only step:", (time.perf_counter() - t0) / N * 1e6)
pr = cProfile.Profile()
pr.enable()
for k in range(N):
    calc(bench.calculator_points(x, k))
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(45)
