"""Developer tool: device time per full evaluation when evaluations are queued without a host
sync in between (c3_evaluate_device) against the blocking c3_update -- the difference is GPU idle
time while the host prepares the next step."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--steps", type=int, default=300)
a = ap.parse_args()
problem = bench.build_problem(a.config, None, "evolved", 0)
stream = torch.cuda.current_stream()
lf, _ = bench.make_lf(problem, stream=stream.cuda_stream)
eng = lf.engine
base, rates = lf.lengths.copy(), lf.rates()
lf.get_log_likelihood()
buf = torch.zeros(1, dtype=torch.float64, device="cuda")
ds = [np.nan_to_num(bench.step_lengths(base, k))[:, None] * rates[None, :] for k in range(7)]
for mode in ("blocking", "queued", "blocking", "queued"):
    for k in range(5):
        eng.update(ds[k % 7])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for k in range(a.steps):
        if mode == "blocking":
            eng.update(ds[k % 7])
        else:
            eng.set_distances(ds[k % 7])
            eng.evaluate_device(buf.data_ptr())
    e1.record(stream)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"{a.config} {mode}: {e0.elapsed_time(e1) / a.steps * 1e3:.1f} us per evaluation on the device, "
          f"host loop {t_host / a.steps * 1e6:.1f} us per call")
