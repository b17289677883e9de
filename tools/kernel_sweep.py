"""Developer tool: build libc3cuda variants with different compile-time knobs and time the
pruning kernels of a full evaluation with each (CUDA events via c3_get_launch_profile).

    python tools/kernel_sweep.py --config c2 [--columns L] [--kind evolved]
"""
import argparse
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "c2": [{"P4S_WARPS": w, "P4S_STAGES_NC2": st, "P4S_WARPS_NC3": w3} for w, st, w3 in
           [(8, 2, 8), (6, 2, 6), (10, 2, 5), (7, 2, 4), (5, 2, 7), (9, 2, 8)]],
}


def build_variant(defs, tag):
    from cogent3_b200 import build as b

    out = os.path.join(ROOT, "build", "variants", f"libc3cuda_{tag}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [b.nvcc_path(), *b.NVCC_FLAGS, "-I", os.path.join(ROOT, "include"),
           *[f"-D{k}={v}" for k, v in defs.items()], "-o", out, *b.SOURCES]
    subprocess.run(cmd, check=True, capture_output=True)
    return out


def run_one(lib, args):
    env = dict(os.environ, C3CUDA_LIB=lib)
    code = f"""
import sys, json, numpy as np
sys.path.insert(0, {ROOT!r})
import bench
class A: pass
problem = bench.build_problem({args.config!r}, {args.columns!r}, {args.kind!r}, 0)
lf, _ = bench.make_lf(problem)
eng = lf.engine
lf.get_log_likelihood()
base = lf.lengths.copy(); rates = lf.rates()
eng.set_profiling(True)
tot = []
for k in range(6):
    eng.update(np.nan_to_num(bench.step_lengths(base, k))[:, None] * rates[None, :])
    pr = eng.launch_profile()
    tot.append((sum(r['ms'] for r in pr if r['kind'].startswith('prune')), sum(r['bytes'] for r in pr if r['kind'].startswith('prune')), [round(r['ms'],4) for r in pr]))
best = min(tot[2:])
print(json.dumps({{'prune_ms': best[0], 'GBps': best[1]/best[0]/1e6, 'per_launch': best[2]}}))
"""
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    if res.returncode != 0:
        return {"error": res.stderr[-500:]}
    return json.loads(res.stdout.strip().splitlines()[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--columns", type=int, default=None)
    ap.add_argument("--kind", default="evolved")
    ap.add_argument("--build-only", action="store_true")
    args = ap.parse_args()
    for i, defs in enumerate(VARIANTS.get(args.config, [{}])):
        tag = "_".join(f"{k.split('_')[-1].lower()}{v}" for k, v in defs.items()) or "default"
        lib = os.path.join(ROOT, "build", "variants", f"libc3cuda_{tag}.so")
        if not os.path.exists(lib):
            lib = build_variant(defs, tag)
        if args.build_only:
            continue
        print(tag, json.dumps(run_one(lib, args)), flush=True)


if __name__ == "__main__":
    main()
