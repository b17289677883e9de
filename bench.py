#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the cogent3 likelihood hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config c2|c1|c3|c4|c5] [--columns L] [--kind evolved|dense]

One "step" = one FULL log-likelihood evaluation (every branch length changed, so
every P matrix is re-exponentiated and every internal node re-pruned) of the
BASELINE.json config named by --config (default c2: GTR+Gamma4, 64 taxa x 1M
columns, SURVEY §8d "evolved" synthetic alignment).

metric   pattern.edge partial updates/s = U_gp / time, with U_gp = K * sum over edges of
         the pattern rows at the edge's PARENT node, counted from the reference's own
         per-node compression tables -- the same numerator for GPU and CPU.
value    steps timed on the device (CUDA events on the engine's stream), pattern tables
         and partials resident in HBM, parameters arriving from the host each step.
e2e      the same steps through the public API (LikelihoodFunction.set_param_rule /
         get_log_likelihood): host Q build + eigendecomposition, H2D of the parameter
         block, all kernels, D2H of the scalar; wall-clock with a device sync each side.
roofline the pruning kernels of one step: algorithmic bytes / CUDA-event time per launch
         (c3_get_launch_profile), against MEASURED_PEAKS.json.
cpu_baseline  the oracle port (oracle/engine.py) on the host cores, bounded sample.

With N>1 (torchrun) every rank holds its own column shard of the same size (weak
scaling); the only collective is the NCCL all-reduce of the scalar lnL.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cogent3_b200 import models as hostmodels  # noqa: E402
from cogent3_b200.likelihood import LikelihoodFunction  # noqa: E402
from cogent3_b200.patterns import compress_alignment  # noqa: E402
from cogent3_b200.synthetic import SEED, make_workload  # noqa: E402

METRIC = "pattern_edge_partial_updates_per_s"
UNIT = "pattern·edge·class updates/s"

GTR_RATES = {"A/C": 1.1, "A/G": 3.2, "A/T": 0.7, "C/G": 0.9, "C/T": 3.6}

CONFIGS = {
    # name: (model, taxa, columns, bins, description)
    "c1": ("HKY85", 55, 3009, 1, "HKY85 brca1 55x3009 (golden fixture)"),
    "c2": ("GTR", 64, 1_000_000, 4, "GTR+Gamma4 64 taxa x 1M columns"),
    "c3": ("GN", 128, 2_000_000, 1, "GN per-edge Q (Pade) 128 taxa x 2M columns"),
    "c4": ("CNFGTR", 32, 200_000, 1, "CNFGTR 61-state 32 taxa x 200k codons"),
    "c5": ("HKY85", 256, 10_000_000, 4, "HKY85+Gamma4 256 taxa x 10M columns"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------
def build_problem(cfg, columns, kind, shard):
    """(model, topo, lengths, codes, symbol_rows, bins, setup) for one column shard"""
    model_name, n_tips, L, bins, _ = CONFIGS[cfg]
    L = columns or L
    if cfg == "c1":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from goldenutil import Golden

        g = Golden("c1_brca1_hky85_point2")
        return ("HKY85", g.topo, g.lengths.copy(), g.codes, g.symbol_rows, 1,
                {"params": {"kappa": 4.3052}, "mprobs": g.mprobs})  # fmt: skip
    m = hostmodels.get_model(model_name)
    topo, lengths, codes = make_workload(n_tips, L, m.n_states, kind=kind, seed=SEED, shard=shard)
    rows = np.eye(m.n_states)
    setup = {}
    if cfg == "c2":
        setup = {"params": GTR_RATES, "rate_shape": 0.5}
    elif cfg == "c3":
        rng = np.random.default_rng(SEED + 3)
        setup = {"edge_params": {n: {p: float(rng.uniform(0.3, 3.0)) for p in m.parameter_order}
                                 for n in topo.edges}, "expm": "pade"}  # fmt: skip
    elif cfg == "c4":
        setup = {"params": dict(GTR_RATES, omega=0.4)}
    elif cfg == "c5":
        setup = {"params": {"kappa": 4.0}, "rate_shape": 0.5}
    return model_name, topo, lengths, codes, rows, bins, setup


def make_lf(problem, engine_factory=None, device=0, stream=None, patterns=None):
    model_name, topo, lengths, codes, rows, bins, setup = problem
    lf = LikelihoodFunction(model_name, topo, bins=bins, expm=setup.get("expm", "either"),
                            engine_factory=engine_factory, device=device, stream=stream)  # fmt: skip
    t0 = time.perf_counter()
    if patterns is None and engine_factory is None:
        # the CUDA engine compresses the alignment itself (c3_compress_alignment)
        lf.set_alignment(codes, rows, compress="device")
        t1 = t2 = time.perf_counter()
    else:
        if patterns is None:
            patterns = compress_alignment(topo, codes, rows)
        t1 = time.perf_counter()
        lf.set_alignment(None, patterns=patterns)
        t2 = time.perf_counter()
    if "mprobs" in setup:
        lf.set_motif_probs(setup["mprobs"])
    else:
        lf.set_motif_probs_from_data()
    lf.lengths[:-1] = lengths[:-1]
    for p, v in setup.get("params", {}).items():
        lf.set_param_rule(p, value=v)
    for n, pv in setup.get("edge_params", {}).items():
        for p, v in pv.items():
            lf.set_param_rule(p, edge=n, value=v)
    if "rate_shape" in setup:
        lf.set_param_rule("rate_shape", value=setup["rate_shape"])
    return lf, {"compress_s": t1 - t0, "upload_s": t2 - t1,
                "compress": "device" if (t1 == t2) else "host"}


def step_lengths(base, k):
    """every branch length changes each step => a full evaluation"""
    return base * (1.0 + 1e-3 * ((k % 7) + 1))


# ----------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", os.environ.get("C3_BENCH_SMI_MS", "50"), "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[5:9], strict=False):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(smax)) if smax else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


# ----------------------------------------------------------------------------------
# CPU arm (oracle port) -- also the "--impl reference" arm
# ----------------------------------------------------------------------------------
def cpu_arm(problem, budget_s=20.0, sample_columns=None, steps=20, warmup=1):
    """time full evaluations of the CPU oracle on a bounded sample of the workload.
    Tries the reference-structured single-thread port and the fused all-cores C loop and
    returns the faster (numpy.inner is slower with BLAS threads).  Each mode runs
    ``warmup`` untimed and up to ``steps`` timed evaluations (fewer if ``budget_s`` runs out);
    the value is units / MEAN step time, like the GPU arm's."""
    import functools

    from oracle.engine import OracleEngine, host_threads

    model_name, topo, lengths, codes, rows, bins, setup = problem
    L = codes.shape[1]
    M = rows.shape[1]
    Ls = min(L, sample_columns or (200_000 if M <= 4 else 20_000))
    sub = (model_name, topo, lengths, codes[:, :Ls], rows, bins, setup)
    patterns = compress_alignment(topo, sub[3], rows)
    _, u_gp = patterns.work_units(bins)
    best = None
    threads = host_threads()
    for mode, cores in (("fused", threads), ("port", 1)):
        lf, _ = make_lf(sub, engine_factory=functools.partial(OracleEngine, mode=mode, n_threads=cores),
                        patterns=patterns)  # fmt: skip
        base = lf.lengths.copy()
        for w in range(max(1, warmup)):  # untimed
            lf.lengths[:] = step_lengths(base, 10_000 + w)
            lf._dirty_model = True
            lf.get_log_likelihood()
        times, k, t_start = [], 0, time.perf_counter()
        while True:
            lf.lengths[:] = step_lengths(base, k)
            lf._dirty_model = True
            t0 = time.perf_counter()
            lnl = lf.get_log_likelihood()
            times.append(time.perf_counter() - t0)
            k += 1
            if k >= steps or (k >= 3 and time.perf_counter() - t_start > budget_s / 2):
                break
        t = float(np.mean(times))
        rec = {"value": u_gp / t, "unit": UNIT, "cores": cores, "kind": "port", "mode": mode,
               "steps_timed": k, "ms_per_step": t * 1e3, "best_ms": float(np.min(times)) * 1e3,
               "sample": f"{Ls} of {L} columns, same tree/model/parameters; mean of {k} full "
                         f"evaluations ({t * 1e3:.2f} ms each, U_gp={u_gp})",
               "lnL_sample": lnl}  # fmt: skip
        if best is None or rec["value"] > best["value"]:
            best = rec
    return best


def cogent3_arm(problem, cfg, budget_s=120.0, sample_columns=None, steps=20, warmup=1):
    """time the UNMODIFIED cogent3 (baseline/_ref, a plain copy of the reference package; it is
    pure Python + numba) on a bounded sample of the workload: build the likelihood function
    through cogent3's own public API, ``calc = lf.make_calculator()``, then full evaluations
    ``calc(x * (1 + eps_k))`` (every parameter changes: SURVEY §8d).  Returns None when the
    package is not importable on this machine (then the oracle port is the CPU arm)."""
    ref_root = os.environ.get("COGENT3_REFERENCE_ROOT", os.path.join(ROOT, "baseline", "_ref"))
    if not os.path.isdir(os.path.join(ref_root, "cogent3")):
        return None
    os.environ["COGENT3_REFERENCE_ROOT"] = ref_root
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    sys.path[:0] = [os.path.join(ROOT, "tests", "_shims"), ref_root]
    try:
        from cogent3 import get_model, make_aligned_seqs, make_tree
    except Exception:  # noqa: BLE001 - fall back to the port
        return None
    from threadpoolctl import threadpool_limits

    from cogent3_b200.patterns import from_reference_lht

    model_name, topo, lengths, codes, rows, bins, setup = problem
    L = codes.shape[1]
    M = rows.shape[1]
    Ls = min(L, sample_columns or (50_000 if M <= 4 else 5_000))
    if cfg == "c1":
        symbols = hostmodels.dna_symbol_rows()[0]
    else:
        symbols = hostmodels.dna_symbol_rows()[0] if M == 4 else hostmodels.codon_symbol_rows()[0]
    t0 = time.perf_counter()
    data = {n: "".join(symbols[c] for c in row[:Ls]) for n, row in zip(topo.tip_names, codes, strict=True)}
    aln = make_aligned_seqs(data, moltype="dna")
    tree = make_tree(topo.to_newick(np.nan_to_num(lengths)))
    kw = {}
    if bins > 1:
        sm = get_model(model_name, ordered_param="rate", distribution="gamma")
        kw["bins"] = bins
    else:
        sm = get_model(model_name)
    if setup.get("expm"):
        kw["expm"] = setup["expm"]
    lf = sm.make_likelihood_function(tree, **kw)
    lf.set_alignment(aln)
    t_aln = time.perf_counter() - t0
    if bins > 1:
        lf.set_param_rule("bprobs", is_constant=True)
        lf.set_param_rule("rate_shape", value=setup.get("rate_shape", 0.5))
    if "edge_params" in setup:
        lf.set_time_heterogeneity(is_independent=True, upper=50)
        with lf.updates_postponed():
            for n, pv in setup["edge_params"].items():
                for pname, v in pv.items():
                    lf.set_param_rule(pname, edge=topo.names[n], value=v)
    for pname, v in setup.get("params", {}).items():
        lf.set_param_rule(pname, value=v)
    lnl0 = float(lf.get_log_likelihood())
    _, pt = from_reference_lht(lf.get_param_value("lht"))
    _, u_gp = pt.work_units(bins)
    t0 = time.perf_counter()
    calc = lf.make_calculator()
    t_calc = time.perf_counter() - t0
    x = np.array(calc.get_value_array(), float)
    best = None
    for label, limit in (("1 BLAS thread", 1), ("default BLAS threads", None)):
        with threadpool_limits(limits=limit):
            for w in range(max(1, warmup)):
                calc(x * (1.0 - 1e-4 * (1 + w)))
            times, k, t_start = [], 0, time.perf_counter()
            while True:
                xk = x * (1.0 + 1e-4 * (1 + k % 97))  # never the previous point: nothing is cached
                t1 = time.perf_counter()
                lnl = calc(xk)
                times.append(time.perf_counter() - t1)
                k += 1
                if k >= steps or (k >= 3 and time.perf_counter() - t_start > budget_s / 2):
                    break
        t = float(np.mean(times))
        rec = {"value": u_gp / t, "unit": UNIT, "cores": 1 if limit == 1 else os.cpu_count(),
               "kind": "reference", "mode": f"cogent3 Calculator, {label}",
               "steps_timed": k, "ms_per_step": t * 1e3, "best_ms": float(np.min(times)) * 1e3,
               "sample": f"{Ls} of {L} columns, same tree/model/parameters; mean of {k} full evaluations "
                         f"calc(x*(1+eps)) ({t * 1e3:.2f} ms each, U_gp={u_gp}); set_alignment {t_aln:.1f} s, "
                         f"make_calculator {t_calc:.1f} s",
               "lnL_sample": lnl0, "lnL_last": float(lnl)}  # fmt: skip
        if best is None or rec["value"] > best["value"]:
            best = rec
    return best


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    problem = build_problem(args.config, args.columns, args.kind, 0)
    # bounded sample: steps+warmup evaluations must finish within a few minutes
    # (both CPU modes get half of a 4-minute budget; `steps` reports what was actually timed)
    rec = None
    if not args.port_only:
        try:
            rec = cogent3_arm(problem, args.config, budget_s=120.0, sample_columns=args.cpu_sample,
                              steps=args.steps, warmup=args.warmup)  # fmt: skip
        except Exception as err:  # noqa: BLE001 - the port below always works
            print(f"cogent3 arm failed ({err!r}); timing the oracle port", file=sys.stderr)
    port = cpu_arm(problem, budget_s=60.0 if rec else 240.0, sample_columns=args.cpu_sample,
                   steps=args.steps, warmup=args.warmup)  # fmt: skip
    if rec is None:
        rec = port
    else:
        rec["port"] = port  # the C restatement (oracle/), a stronger CPU baseline than the reference
    line = {
        "impl": "reference", "metric": METRIC, "value": rec["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": rec["steps_timed"], "steps_requested": args.steps,
        "warmup": args.warmup, "ms_per_step": rec["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, problem),
        "cpu_baseline": rec,
        "e2e": {"value": rec["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": ("unmodified cogent3 from baseline/_ref through its own public API"
                 if rec["kind"] == "reference" else
                 "cogent3's CPU algorithm restated (oracle/, parity pinned to the live reference); "
                 "baseline/_ref (a copy of the reference package) was not importable here"),
    }  # fmt: skip
    print(json.dumps(line))


def workload_config(args, problem):
    model_name, topo, _, codes, _, bins, setup = problem
    return {
        "workload": f"{args.config}: {CONFIGS[args.config][4]}",
        "substitution_model": model_name, "taxa": len(topo.tips), "columns_per_gpu": int(codes.shape[1]),
        "rate_classes": bins, "alignment": args.kind, "expm": setup.get("expm", "either"),
        "l2_policy": "working set (resident partials + gather tables) larger than the 126 MB L2",
        "step": "full evaluation: all branch lengths changed",
    }  # fmt: skip


# ----------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------
def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the likelihood engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    problem = build_problem(args.config, args.columns, args.kind, shard=rank)
    # a stream of our own (torch's legacy default stream cannot be captured into the engine's CUDA graphs);
    # it is torch's current stream from here on, so NCCL orders after the engine's kernels
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    lf, setup_t = make_lf(problem, device=local_rank, stream=stream)
    eng = lf.engine
    bins = lf.n_bins
    u_mv, u_gp = lf.patterns.work_units(bins)
    base = lf.lengths.copy()
    rates = lf.rates()
    lnl_dev = torch.zeros(1, dtype=torch.float64, device="cuda")

    fused = False
    if world > 1:
        from cogent3_b200.sharded import attach_fused_allreduce

        fused = attach_fused_allreduce(eng)  # the sum over ranks inside the reduction kernel

    # the parameter vectors of the timed steps are inputs: prepared before the timed region (the
    # cycle of step_lengths has 7 members), like the alignment tables
    step_dist = [np.ascontiguousarray(np.nan_to_num(step_lengths(base, k))[:, None] * rates[None, :])
                 for k in range(7)]  # fmt: skip

    def device_step(k):
        d = step_dist[k % 7]
        if world == 1 or fused:
            return eng.update(d)
        eng.set_distances(d)
        eng.evaluate_device(lnl_dev.data_ptr())
        dist.all_reduce(lnl_dev)  # the one collective: 8 bytes over NVLink
        return float(lnl_dev.item())

    def api_step(k):
        lf.lengths[:] = step_lengths(base, k)
        lf._dirty_model = True
        if lf.model.parameter_order and not lf.edge_params:
            p = lf.model.parameter_order[0]
            lf.set_param_rule(p, value=lf.params[p] * (1.0 + 1e-4 * ((k % 3) - 1)))
        v = lf.get_log_likelihood()
        if world > 1 and not fused:
            lnl_dev.fill_(v)
            dist.all_reduce(lnl_dev)
            v = float(lnl_dev.item())
        return v

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lf.get_log_likelihood()
    for k in range(max(args.warmup, 3)):
        device_step(k)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = eng.stats()["kernel_launches"]

    # ---- device-timed steps ------------------------------------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for k in range(args.steps):
        lnl = device_step(k)
    ev1.record(stream)
    barrier()
    dev_s = ev0.elapsed_time(ev1) * 1e-3
    launches = eng.stats()["kernel_launches"] - launches0

    # ---- end-to-end steps through the public API --------------------------------------
    for k in range(2):
        api_step(k)
    h2d = 0
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        api_step(k)
        h2d += eng.stats()["last_h2d_bytes"]
    barrier()
    e2e_s = time.perf_counter() - t0
    clock_rec = clocks.stop() if rank == 0 else None

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    dev_s = max_over_ranks(dev_s)
    e2e_s = max_over_ranks(e2e_s)
    total_units = sum_over_ranks(float(u_gp))
    total_umv = sum_over_ranks(float(u_mv))

    # ---- roofline of the pruning kernels (per-launch CUDA events) -------------------
    eng.set_profiling(True)
    prof_rows = []
    for k in range(6):
        # alternate two length vectors so that every profiled step is a full evaluation
        eng.update(np.nan_to_num(base * (1.0 + 1e-3 * (1 + k % 2)))[:, None] * rates[None, :])
        rows = eng.launch_profile()
        if rows and k >= 1:
            prof_rows.append(rows)
    eng.set_profiling(False)
    prof = prof_rows[-1]
    prune = [[r for r in rows if r["kind"].startswith("prune")] for rows in prof_rows]
    all_ms = float(np.median([sum(r["ms"] for r in rows) for rows in prune]))
    all_bytes = sum(r["bytes"] for r in prune[-1])
    all_flops = sum(r["flops"] for r in prune[-1])
    # the roofline is that of the DOMINANT kernel: the launches of prune4_reg2_kernel (levels and root
    # instantiations) for 4 states, of the contraction instance of prune_dmma_edge_kernel otherwise;
    # the aggregate over every pruning launch (small-levels launch, codon root) is reported next to it
    M_states = lf.model.n_states
    dom_kinds = ("prune", "prune_root") if M_states == 4 else ("prune",)
    dom = [[r for r in rows if r["kind"] in dom_kinds] for rows in prune]
    dom_name = "prune4_reg2_kernel" if M_states == 4 else "prune_dmma_edge_kernel"
    if not dom[-1]:
        dom, dom_name = prune, "prune4_small_kernel"
    prune_ms = float(np.median([sum(r["ms"] for r in rows) for rows in dom]))
    prune_bytes = sum(r["bytes"] for r in dom[-1])
    prune_flops = sum(r["flops"] for r in dom[-1])
    step_ms = float(np.median([sum(r["ms"] for r in rows) for rows in prof_rows]))
    hbm_peak, peak_src = peaks()
    st = eng.stats()
    M = lf.model.n_states
    if M == 4:
        achieved = prune_bytes / (prune_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src}  # fmt: skip
    else:
        from cogent3_b200.engine import measure_dmma_peak

        dmma_peak = measure_dmma_peak(local_rank)
        # ALGORITHMIC flops (2 M^2 per internal-child row); the kernel executes 2 MP^2 (M padded to
        # a multiple of 8: 61 -> 64), reported next to it
        MP = st["padded_states"]
        alg_flops = prune_flops * (M * M) / (MP * MP)
        achieved = alg_flops / (prune_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s",
                    "frac": achieved / dmma_peak, "traffic": None,
                    "peak_source": "FP64 DMMA issue-loop microbenchmark on this GPU (c3_measure_dmma_peak)",
                    "executed_tflops": prune_flops / (prune_ms * 1e-3) / 1e12,
                    "executed_frac": prune_flops / (prune_ms * 1e-3) / 1e12 / dmma_peak,
                    "hbm_gbs_achieved": prune_bytes / (prune_ms * 1e-3) / 1e9}  # fmt: skip
    # DRAM traffic of the dominant launch from the committed ncu capture (per launch, like
    # `achieved` is per launch); null when no capture exists for this config
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            tr = json.load(f).get(args.config)
        if tr and args.kind == "evolved" and not args.columns:
            roofline["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            roofline["traffic_note"] = (f"{tr['kernel']}: algorithmic {tr['algorithmic_bytes']} B per launch; "
                                        f"{tr.get('profile', 'profiles/')}")
    except (OSError, ValueError, KeyError):
        pass
    if M == 4:
        agg = all_bytes / (all_ms * 1e-3) / 1e9
    else:
        agg = all_flops * (M * M) / (st["padded_states"] ** 2) / (all_ms * 1e-3) / 1e12
    roofline["all_pruning_launches"] = {"achieved": agg, "frac": agg / roofline["peak"], "launches": len(prune[-1]),
                                        "ms_per_step": all_ms}  # fmt: skip
    roofline.update({
        "kernel": dom_name,
        "launches_per_step": len(dom[-1]), "kernel_ms_per_step": prune_ms,
        "kernel_share_of_step": prune_ms / step_ms if step_ms else None,
        "alg_bytes_per_step": prune_bytes, "executed_flops_per_step": prune_flops,
        "per_launch": [{"kind": r["kind"], "ms": round(r["ms"], 5), "bytes": r["bytes"],
                        "GBps": round(r["bytes"] / r["ms"] / 1e6, 1) if r["ms"] > 0 else None,
                        **({"exec_TFLOPs": round(r["flops"] / r["ms"] / 1e9, 2)} if M != 4 and r["ms"] > 0 else {})}
                       for r in prof],
    })  # fmt: skip

    # ---- partial updates: one branch length at a time (the optimiser's inner loop) --
    partial = None
    if world == 1:
        edges = lf.topo.edges
        sel = edges if len(edges) <= 128 else list(np.random.default_rng(0).choice(edges, 128, replace=False))
        d = np.nan_to_num(base)[:, None] * rates[None, :]
        eng.update(d)
        units = 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for n in sel:
            d2 = d.copy()
            d2[n] *= 1.01
            eng.update(d2)
            units += lf.patterns.work_units(bins, nodes=set(lf.topo.ancestors(n)))[1]
        t_part = time.perf_counter() - t0
        partial = {"evals_per_s": len(sel) / t_part, "updates_per_s": units / t_part,
                   "edges_sampled": len(sel), "mean_units_per_eval": units / len(sel)}  # fmt: skip

    # ---- N > 1: the fused all-reduce against an NCCL all-reduce of the shard values ----------
    allreduce_check = None
    if world > 1 and fused:
        d = np.nan_to_num(step_lengths(base, 3))[:, None] * rates[None, :]
        total = eng.update(d)
        eng.comm_detach()
        eng.invalidate()
        t = torch.tensor([eng.update(d)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        allreduce_check = {"fused_total": total, "nccl_sum_of_shards": float(t.item()),
                           "rel_diff": abs(total - float(t.item())) / abs(total)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- many small alignments at once: B independent likelihood functions --
    many = None
    if world == 1 and args.config == "c1":
        from cogent3_b200.engine import update_many as engine_update_many

        B = 64
        lfs = [make_lf(problem, device=local_rank)[0] for _ in range(B)]
        engines = [f.engine for f in lfs]
        for f in lfs:
            f.get_log_likelihood()

        def many_step(k):
            return engine_update_many(engines, step_dist[k % 7])

        for k in range(20):  # (past the engines' graph-capture threshold: steady state)
            many_step(k)
        n_rep = 100
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(n_rep):
            vals = many_step(k)
        t_many = (time.perf_counter() - t0) / n_rep
        t_seq = 0.0
        for k in range(5):  # the same work, one blocking evaluation after the other
            d = np.nan_to_num(step_lengths(base, 1000 + k))[:, None] * rates[None, :]
            for e_ in engines:
                e_.set_distances(d)
            t0 = time.perf_counter()
            one_by_one = [e_.evaluate() for e_ in engines]
            t_seq += (time.perf_counter() - t0) / 5
        many = {"engines": B, "evals_per_s": B / t_many, "ms_per_batch": t_many * 1e3,
                "one_at_a_time_evals_per_s": B / t_seq, "all_equal": bool(len(set(vals)) == 1),
                "api": "c3_update_many: every engine on its own stream; a batch that keeps coming up is ONE CUDA graph"}
        del one_by_one
        for f in lfs:
            f.engine.close()

    cpu_rec = None
    if world == 1 and not args.no_cpu:
        # in a fresh process: torch's OpenMP runtime settings must not throttle the CPU arm
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", args.config,
               "--kind", args.kind, "--steps", "20", "--warmup", "1"]  # fmt: skip  (bounded: ~10-30 s)
        if args.columns:
            cmd += ["--columns", str(args.columns)]
        if args.cpu_sample:
            cmd += ["--cpu-sample", str(args.cpu_sample)]
        env = {k: v for k, v in os.environ.items() if not k.startswith("OMP_")}
        try:
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
            cpu_rec = json.loads(res.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as err:  # noqa: BLE001 - the GPU line must still be printed
            cpu_rec = {"error": f"cpu baseline failed: {err!r}"}

    value = total_units * args.steps / dev_s
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": dev_s / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, problem),
        "e2e": {"value": total_units * args.steps / e2e_s, "unit": UNIT,
                "ms_per_step": e2e_s / args.steps * 1e3,
                "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": 8,
                "api": "LikelihoodFunction.set_param_rule + get_log_likelihood (host Q/eig, "
                       "c3_set_eigen, c3_set_distances, c3_evaluate)"},
        "gpu_launches": int(launches),
        "clocks": clock_rec,
        "roofline": roofline,
        "cpu_baseline": cpu_rec,
        "units": {"U_gp_per_step": total_units, "U_mv_per_step": total_umv,
                  "lf_evals_per_s": args.steps / dev_s,
                  "root_patterns_per_gpu": int(lf.patterns.root.n_rows - 1),
                  "node_rows_per_step": int(st["last_node_rows"])},
        "partial_update": partial,
        "many_small": many,
        "allreduce": ({"kind": "fused into the reduction kernel (peer-memory mailboxes over NVLink)",
                       "check": allreduce_check} if fused else
                      ({"kind": "NCCL all-reduce of one double per step"} if world > 1 else None)),
        "setup": {**setup_t, "device_bytes": int(st["device_bytes"]),
                  "partial_bytes": int(st["partial_bytes"]), "levels": int(st["n_levels"])},
        "lnL": lnl,
        "host": {"cpu_count": os.cpu_count()},
    }  # fmt: skip
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--columns", type=int, default=None, help="override the column count per GPU")
    ap.add_argument("--kind", default="evolved", choices=["evolved", "dense"])
    ap.add_argument("--cpu-sample", type=int, default=None, help="columns in the CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--port-only", action="store_true", help="reference arm: time only the oracle port")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port", "29533",
               os.path.abspath(__file__), *sys.argv[1:]]  # fmt: skip
        raise SystemExit(subprocess.call(cmd))
    # Everything libraries print (NCCL's version banner, warnings) goes to stderr; the real
    # stdout receives exactly one JSON line.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    captured = []
    builtin_print = print

    def emit(line):
        captured.append(line)

    globals()["print"] = lambda *a, **k: emit(" ".join(str(x) for x in a))
    try:
        if args.impl == "reference":
            run_reference_arm(args)
        else:
            run_ours(args)
    finally:
        globals()["print"] = builtin_print
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for line in captured:
        builtin_print(line, flush=True)


if __name__ == "__main__":
    main()
