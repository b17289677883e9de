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
         and partials resident in HBM, parameters arriving from the host each step; every
         step is a BLOCKING evaluation (the result is back on the host before the next
         step is issued, as in an optimiser).  `pipelined`: the same steps issued back to
         back (c3_evaluate_device), one synchronisation at the end.
e2e      the same metric END TO END THROUGH THE PRODUCT: unmodified cogent3 (baseline/_ref)
         with cogent3_b200.install(); `calc = lf.make_calculator(); calc(x_k)` with every
         optimisable parameter changed -- the call the reference arm times: cogent3's own
         Q build, eigendecomposition and Calculator sweep on the host, host buffers into
         c3_set_eigen / c3_update, all kernels, the scalar back through mapped pinned
         memory; wall clock with a device sync each side.  (`e2e_mirror`: the same steps
         through the repo's cogent3-free mirror API.)
roofline the pruning kernels of one step: algorithmic bytes / CUDA-event time per launch
         (c3_get_launch_profile), against MEASURED_PEAKS.json.
cpu_baseline  unmodified cogent3 (kind "reference") on a bounded sample of the columns, the
         oracle port next to it; `--impl reference` times cogent3 on the FULL C2 workload.
c4       (default config, N = 1) the CNFGTR 61-state config measured the same way, roofline
         bound "tensor" against max(DMMA issue-loop peak, cuBLAS DGEMM 8192^3) of this run.
c5_strong  (default config) the ONE 10M-column C5 alignment split over the N ranks.
c1       (default config, N = 1) brca1: full evaluations, single-branch updates, 64 engines at
         once, and `e2e.optimise`: lf.optimise(local=True) through the drop-in and through
         stock cogent3 in the same process.

With N>1 (torchrun) every rank holds its own column shard of the same size (weak
scaling); the sum over ranks of the scalar lnL is fused into the reduction
kernel (peer-memory mailboxes), NCCL all-reduce as the fallback.
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


def import_cogent3():
    """the UNMODIFIED cogent3 from baseline/_ref (a plain copy of the reference package) with the
    three import shims of tests/_shims; None when it is not on this machine"""
    ref_root = os.environ.get("COGENT3_REFERENCE_ROOT", os.path.join(ROOT, "baseline", "_ref"))
    if not os.path.isdir(os.path.join(ref_root, "cogent3")):
        return None
    os.environ["COGENT3_REFERENCE_ROOT"] = ref_root
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    for pth in (os.path.join(ROOT, "tests", "_shims"), ref_root):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    try:
        import cogent3
        import cogent3.maths.util  # noqa: F401  (import order fixes the reference's numpy error state)
        from cogent3.evolve import substitution_model  # noqa: F401
    except Exception:  # noqa: BLE001
        return None
    return cogent3


def build_cogent3_lf(problem, cfg, columns=None, backend="cpu", device=0):
    """the likelihood function of `problem` through cogent3's OWN public API: get_model(...)
    .make_likelihood_function(tree), lf.set_alignment(aln), parameter rules.  backend "cuda": the same calls
    with cogent3_b200.install() active, i.e. sm.make_likelihood_function itself returns the CUDA class."""
    from cogent3 import get_model, make_aligned_seqs, make_tree

    model_name, topo, lengths, codes, rows, bins, setup = problem
    L = codes.shape[1]
    M = rows.shape[1]
    Ls = min(L, columns or L)
    symbols = np.array(hostmodels.dna_symbol_rows()[0] if (cfg == "c1" or M == 4) else hostmodels.codon_symbol_rows()[0])
    t0 = time.perf_counter()
    if symbols.dtype.itemsize == np.dtype("U1").itemsize:  # single characters: vectorised
        table = np.frombuffer("".join(symbols).encode("ascii"), np.uint8)
        data = {n: table[row[:Ls]].tobytes().decode("ascii") for n, row in zip(topo.tip_names, codes, strict=True)}
    else:
        data = {n: "".join(symbols[row[:Ls]]) for n, row in zip(topo.tip_names, codes, strict=True)}
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
    if backend == "cuda":
        import cogent3_b200

        cogent3_b200.install(device=device)
        try:
            lf = sm.make_likelihood_function(tree, device=device, **kw)
        finally:
            cogent3_b200.uninstall()
        assert type(lf).__name__ == "CudaAlignmentLikelihoodFunction"
    else:
        lf = sm.make_likelihood_function(tree, **kw)
    t1 = time.perf_counter()
    lf.set_alignment(aln)
    t_aln = time.perf_counter() - t1
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
    if "mprobs" in setup:
        lf.set_motif_probs(dict(zip(symbols[:4], setup["mprobs"], strict=True)))
    return lf, {"columns": Ls, "make_objects_s": t1 - t0, "set_alignment_s": t_aln}


def calculator_points(x, k):
    """the parameter vector of timed step k: EVERY optimisable parameter moves (a full evaluation), never the
    previous point (nothing is cached) -- the same sequence for the reference and for the drop-in"""
    return x * (1.0 + 1e-4 * (1 + k % 97))


def single_parameter_rate(calc, x, budget_s=1.0, max_evals=4000):
    """evaluations / s of an each-parameter-in-turn optimiser making REAL changes through
    Calculator.change (what measure_evals_per_second, recalculation/calculation.py:547-583, probes with
    not-a-real-change updates, which an engine that diffs its inputs answers from its cache)"""
    n, t0, k = 0, time.perf_counter(), 0
    while True:
        for i in range(len(x)):
            k += 1
            calc.change([(i, x[i] * (1.0 + 1e-6 * (1 + k % 89)))])
            n += 1
            if n >= max_evals or time.perf_counter() - t0 > budget_s:
                return n / (time.perf_counter() - t0)


def cogent3_arm(problem, cfg, budget_s=120.0, sample_columns=None, steps=20, warmup=1):
    """time the UNMODIFIED cogent3 (baseline/_ref) on the workload, or on a bounded sample of its columns:
    build the likelihood function through cogent3's own public API, ``calc = lf.make_calculator()``, then full
    evaluations ``calc(x * (1 + eps_k))`` (every parameter changes: SURVEY §8d).  Returns None when the
    package is not importable on this machine (then the oracle port is the CPU arm)."""
    if import_cogent3() is None:
        return None
    from threadpoolctl import threadpool_limits

    from cogent3_b200.patterns import from_reference_lht

    model_name, topo, lengths, codes, rows, bins, setup = problem
    L = codes.shape[1]
    M = rows.shape[1]
    Ls = min(L, sample_columns or (50_000 if M <= 4 else 5_000))
    lf, t_setup = build_cogent3_lf(problem, cfg, columns=Ls)
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
                xk = calculator_points(x, k)
                t1 = time.perf_counter()
                lnl = calc(xk)
                times.append(time.perf_counter() - t1)
                k += 1
                if k >= steps or (k >= 3 and time.perf_counter() - t_start > budget_s / 2):
                    break
            calc(x)
            # the reference's own probe needs ~2 rounds over all parameters before it returns: small samples only
            evals = {"measure_evals_per_second": (float(calc.measure_evals_per_second(time_limit=0.5, wall=True))
                                                  if Ls <= 60_000 else None),
                     "single_parameter_real_changes_per_s": single_parameter_rate(calc, x, budget_s=2.0, max_evals=400)}  # fmt: skip
        t = float(np.mean(times))
        rec = {"value": u_gp / t, "unit": UNIT, "cores": 1 if limit == 1 else os.cpu_count(),
               "kind": "reference", "mode": f"cogent3 Calculator, {label}",
               "steps_timed": k, "ms_per_step": t * 1e3, "best_ms": float(np.min(times)) * 1e3,
               "sample": (f"{Ls} of {L} columns" if Ls < L else f"all {L} columns") +
                         f", same tree/model/parameters; mean of {k} full evaluations "
                         f"calc(x*(1+eps)) ({t * 1e3:.2f} ms each, U_gp={u_gp}); set_alignment "
                         f"{t_setup['set_alignment_s']:.1f} s, make_calculator {t_calc:.1f} s",
               "columns": Ls, "set_alignment_s": t_setup["set_alignment_s"], "make_calculator_s": t_calc,
               "lf_evals": evals, "lnL_sample": lnl0, "lnL_last": float(lnl)}  # fmt: skip
        if best is None or rec["value"] > best["value"]:
            best = rec
    return best


def host_memory_available_gb():
    try:
        import psutil

        return psutil.virtual_memory().available / 1e9
    except Exception:  # noqa: BLE001
        return 0.0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    problem = build_problem(args.config, args.columns, args.kind, 0)
    # the headline for C2 is the reference on the FULL workload (all 1M columns: ~2 min of set_alignment +
    # make_calculator in stock cogent3, ~30 GB of host memory); the bounded 50k-column sample of round 1 is
    # timed next to it.  Other configs / small hosts: the sample only (and the line says so).
    rec = sample = None
    L, M = problem[3].shape[1], problem[4].shape[1]
    full_ok = (args.config == "c2" and args.cpu_sample is None and not args.port_only and args.full_reference
               and host_memory_available_gb() > 24.0)  # fmt: skip
    if not args.port_only:
        try:
            sample = cogent3_arm(problem, args.config, budget_s=60.0 if full_ok else 120.0,
                                 sample_columns=args.cpu_sample, steps=args.steps, warmup=args.warmup)  # fmt: skip
        except Exception as err:  # noqa: BLE001 - the port below always works
            print(f"cogent3 arm failed ({err!r}); timing the oracle port", file=sys.stderr)
        if sample is not None and full_ok:
            try:
                rec = cogent3_arm(problem, args.config, budget_s=240.0, sample_columns=L,
                                  steps=args.steps, warmup=args.warmup)  # fmt: skip
                rec["sampled_run"] = sample
            except Exception as err:  # noqa: BLE001
                print(f"full-size cogent3 run failed ({err!r}); reporting the sample", file=sys.stderr)
        if rec is None:
            rec = sample
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
        "note": ("unmodified cogent3 from baseline/_ref through its own public API (Calculator.__call__)"
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
class Ctx:
    """process-wide handles of the GPU arm"""

    def __init__(self):
        import torch

        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the likelihood engine has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        # a stream of our own (torch's legacy default stream cannot be captured into the engine's CUDA graphs);
        # it is torch's current stream from here on, so NCCL orders after the engine's kernels
        self.stream = torch.cuda.Stream(device=self.local_rank)
        torch.cuda.set_stream(self.stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op=None):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if op is None:
            self.dist.all_reduce(t)
        else:
            self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX) if self.world > 1 else x

    def sum_over_ranks(self, x):
        return self._reduce(x)


def dgemm_peak_tflops(ctx, n=8192, reps=3):
    """FP64 calibration SURVEY §8d asks for: a cuBLAS DGEMM n^3 on this GPU (2 n^3 flop, best of `reps`)"""
    torch = ctx.torch
    try:
        a = torch.randn(n, n, dtype=torch.float64, device="cuda")
        b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        torch.matmul(a, b)
        best = None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        del a, b
        torch.cuda.empty_cache()
        return 2.0 * n**3 / (best * 1e-3) / 1e12
    except Exception:  # noqa: BLE001
        return None


def roofline_rows_of_step(eng, base, rates):
    """per-launch records (kind, algorithmic bytes, ...) of one profiled full evaluation"""
    eng.set_profiling(True)
    eng.update(np.nan_to_num(base * (1.0 + 1e-3 * 5))[:, None] * rates[None, :])
    rows = eng.launch_profile()
    eng.set_profiling(False)
    return rows


def roofline_of(ctx, lf, eng, base, rates, cfg, kind, columns):
    """per-launch CUDA events of full evaluations (c3_get_launch_profile) -> the roofline object"""
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
    # the roofline is that of the DOMINANT kernel: the launches of the 4-state streaming kernel (levels and
    # root instantiations), of the contraction instance of prune_dmma_edge_kernel otherwise; the aggregate
    # over every pruning launch (small-levels launch, codon root) is reported next to it
    M = lf.model.n_states
    dom_kinds = ("prune", "prune_root") if M == 4 else ("prune",)
    dom = [[r for r in rows if r["kind"] in dom_kinds] for rows in prune]
    dom_name = eng.stats().get("kernel_name", "prune4_reg2_kernel") if M == 4 else "prune_dmma_edge_kernel"
    if not dom[-1]:
        dom, dom_name = prune, "prune4_small_kernel"
    prune_ms = float(np.median([sum(r["ms"] for r in rows) for rows in dom]))
    prune_bytes = sum(r["bytes"] for r in dom[-1])
    prune_flops = sum(r["flops"] for r in dom[-1])
    step_ms = float(np.median([sum(r["ms"] for r in rows) for rows in prof_rows]))
    hbm_peak, peak_src = peaks()
    st = eng.stats()
    if M == 4:
        achieved = prune_bytes / (prune_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src}  # fmt: skip
    else:
        from cogent3_b200.engine import measure_dmma_peak

        dmma_peak = measure_dmma_peak(ctx.local_rank)
        dgemm = dgemm_peak_tflops(ctx)
        # ALGORITHMIC flops (2 M^2 per contraction row); the kernel executes 2 MP^2 (M padded to a multiple
        # of 8: 61 -> 64), reported next to it.  Denominator: the higher of the two FP64 tensor calibrations.
        MP = st["padded_states"]
        alg_flops = prune_flops * (M * M) / (MP * MP)
        achieved = alg_flops / (prune_ms * 1e-3) / 1e12
        peak = max(dmma_peak, dgemm or 0.0)
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": None,
                    "peak_source": "max(FP64 DMMA issue-loop microbenchmark c3_measure_dmma_peak, cuBLAS DGEMM 8192^3), "
                                   "both measured in this run on this GPU",
                    "dmma_issue_loop_tflops": dmma_peak, "cublas_dgemm_8192_tflops": dgemm,
                    "executed_tflops": prune_flops / (prune_ms * 1e-3) / 1e12,
                    "executed_frac": prune_flops / (prune_ms * 1e-3) / 1e12 / peak,
                    "hbm_gbs_achieved": prune_bytes / (prune_ms * 1e-3) / 1e9}  # fmt: skip
    # DRAM traffic of the dominant launch from the committed ncu capture (per launch, like `achieved` is
    # per launch); null when no capture exists for this config
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                tr = json.load(f).get(cfg)
            if tr and kind == "evolved" and not columns:
                roofline["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
                roofline["traffic_note"] = (f"{tr['kernel']}: algorithmic {tr['algorithmic_bytes']} B per launch; "
                                            f"{tr.get('profile', 'profiles/')}")
                break
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
    return roofline


def dropin_leg(ctx, problem, cfg, steps, warmup, expect_units):
    """END TO END through the product: unmodified cogent3 (baseline/_ref) + cogent3_b200.install() + libc3cuda.
    `calc = lf.make_calculator(); calc(x_k)` -- the very call the reference arm times (cogent3_arm) -- with
    every optimisable parameter changed each step: cogent3's own Q build, eigendecomposition and Calculator
    sweep on the host, c3_set_eigen / c3_update from host buffers, all kernels, the scalar back through mapped
    pinned memory.  With N ranks every rank drives its own column shard and the sum over ranks is fused into
    the reduction kernel, so every rank's calc(x) returns the total.  None when cogent3 is not on this box."""
    if import_cogent3() is None:
        return None
    torch = ctx.torch
    from cogent3_b200.patterns import from_reference_lht

    bins = problem[5]
    lf, t_setup = build_cogent3_lf(problem, cfg, backend="cuda", device=ctx.local_rank)
    t0 = time.perf_counter()
    lnl0 = float(lf.get_log_likelihood())
    t_first = time.perf_counter() - t0
    t0 = time.perf_counter()
    calc = lf.make_calculator()
    t_calc = time.perf_counter() - t0
    x = np.array(calc.get_value_array(), float)
    eng = lf._engine()
    _, pt = from_reference_lht(lf.get_param_value("lht"))
    _, u_gp = pt.work_units(bins)
    assert int(u_gp) == int(expect_units), (u_gp, expect_units)  # the same tables as the device compression
    calc(x)
    fused = None
    if ctx.world > 1:
        from cogent3_b200.sharded import attach_fused_allreduce

        fused = bool(attach_fused_allreduce(eng))
        buf = torch.zeros(1, dtype=torch.float64, device="cuda")

    def step(xk):
        v = calc(xk)
        if ctx.world > 1 and not fused:
            buf.fill_(v)
            ctx.dist.all_reduce(buf)
            v = float(buf.item())
        return v

    for w in range(max(2, warmup)):
        step(x * (1.0 - 1e-4 * (1 + w)))
    h2d = 0
    ctx.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        lnl = step(calculator_points(x, k))
        h2d += eng.stats()["last_h2d_bytes"]
    ctx.barrier()
    wall = time.perf_counter() - t0
    # the same steps with only the branch lengths moving (no new rate matrix, so no eigendecomposition on the
    # host: the other kind of step an optimiser makes; at 61 states numpy's eig is most of a full step's host time)
    is_length = np.array([p.name == "length" for p in calc.opt_pars])
    lengths_only = None
    if is_length.any() and not is_length.all():
        def length_point(k):
            xk = x.copy()
            xk[is_length] = calculator_points(x, k)[is_length]
            return xk

        for w in range(2):
            step(length_point(200 + w))
        ctx.barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            step(length_point(k))
        ctx.barrier()
        lengths_only = ctx.max_over_ranks(time.perf_counter() - t0) / steps * 1e3
    evals = None
    if ctx.world == 1:
        calc(x)
        evals = {"measure_evals_per_second": float(calc.measure_evals_per_second(time_limit=0.5, wall=True)),
                 "single_parameter_real_changes_per_s": single_parameter_rate(calc, x, budget_s=1.0),
                 "note": "measure_evals_per_second (recalculation/calculation.py:547-583) re-presents UNCHANGED "
                         "values, which the engine answers from its input diff (no kernel runs); the second "
                         "figure makes a real one-parameter change per evaluation, for both arms"}  # fmt: skip
    if fused:
        eng.comm_detach()
    st = eng.stats()
    optimise = None
    if cfg == "c1" and ctx.world == 1:
        # BASELINE configs[0]: lf.optimise() on brca1, the reference's own Powell (local=True) driving each arm --
        # two fresh likelihood functions from the same starting point, wall-clock of the whole call
        optimise = {}
        for backend in ("cuda", "cpu"):
            lf_o, _ = build_cogent3_lf(problem, cfg, backend=backend, device=ctx.local_rank)
            lf_o.get_log_likelihood()
            t0 = time.perf_counter()
            lf_o.optimise(local=True, show_progress=False)
            optimise[backend] = {"seconds": time.perf_counter() - t0, "lnL": float(lf_o.get_log_likelihood())}
        optimise["speedup"] = optimise["cpu"]["seconds"] / optimise["cuda"]["seconds"]
        optimise["api"] = "lf.optimise(local=True) of sm.make_likelihood_function(tree): cuda = cogent3_b200.install(), cpu = stock cogent3, same process"
    return {"wall_s": wall, "h2d": h2d, "lnL0": lnl0, "lnL_last": float(lnl), "fused": fused, "lf_evals": evals,
            "lengths_only_ms": lengths_only, "optimise": optimise,
            "graph_replays": int(st["graph_replays"]), "n_parameters": int(len(x)),
            "setup": {**t_setup, "first_evaluation_s": t_first, "make_calculator_s": t_calc}}  # fmt: skip


def measure_workload(ctx, args, cfg, columns, kind, steps, warmup, shard, clocks=None, with_dropin=True,
                     with_partial=True):  # fmt: skip
    """device-timed value, end-to-end legs, roofline, partial updates of ONE config on this rank's shard"""
    torch, dist, world, rank = ctx.torch, ctx.dist, ctx.world, ctx.rank
    problem = build_problem(cfg, columns, kind, shard=shard)
    lf, setup_t = make_lf(problem, device=ctx.local_rank, stream=ctx.stream)
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

    lf.get_log_likelihood()
    for k in range(max(warmup, 3)):
        device_step(k)
    if clocks is not None:
        clocks.start()
    launches0 = eng.stats()["kernel_launches"]
    replays0 = eng.stats()["graph_replays"]

    # ---- device-timed steps ------------------------------------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    ev0.record(ctx.stream)
    for k in range(steps):
        lnl = device_step(k)
    ev1.record(ctx.stream)
    ctx.barrier()
    dev_s = ev0.elapsed_time(ev1) * 1e-3
    launches = eng.stats()["kernel_launches"] - launches0
    replays = eng.stats()["graph_replays"] - replays0

    # ---- the same steps PIPELINED: issued back to back without waiting for a result (c3_evaluate_device leaves lnL in
    # device memory), one synchronisation at the end -- the host plans step k + 1 while the kernels of step k run
    # (a replayed graph frees the parameter staging buffer right behind its copy node).  The headline `value` stays
    # the blocking loop above, which pays the launch / completion round trip every step like an optimiser does.
    pipelined = None
    if world == 1:
        for k in range(3):
            eng.set_distances(step_dist[k % 7])
            eng.evaluate_device(lnl_dev.data_ptr())
        torch.cuda.synchronize()
        ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev2.record(ctx.stream)
        for k in range(steps):
            eng.set_distances(step_dist[k % 7])
            eng.evaluate_device(lnl_dev.data_ptr())
        ev3.record(ctx.stream)
        torch.cuda.synchronize()
        pipe_s = ev2.elapsed_time(ev3) * 1e-3
        pipelined = {"ms_per_step": pipe_s / steps * 1e3, "value": float(u_gp) * steps / pipe_s, "steps": steps,
                     "lnL_last": float(lnl_dev.item()), "same_bits_as_blocking": float(lnl_dev.item()) == float(lnl),
                     "note": "c3_evaluate_device back to back, results in device memory, one synchronisation at the end"}

    # ---- end-to-end steps through the mirror API (no cogent3 needed) ------------------
    for k in range(2):
        api_step(k)
    h2d = 0
    ctx.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        api_step(k)
        h2d += eng.stats()["last_h2d_bytes"]
    ctx.barrier()
    mirror_s = time.perf_counter() - t0
    if clocks is not None:
        clock_rec = clocks.stop()
    else:
        clock_rec = None

    dev_s = ctx.max_over_ranks(dev_s)
    mirror_s = ctx.max_over_ranks(mirror_s)
    total_units = ctx.sum_over_ranks(float(u_gp))
    total_umv = ctx.sum_over_ranks(float(u_mv))

    roofline = roofline_of(ctx, lf, eng, base, rates, cfg, kind, columns)

    # ---- partial updates: one branch length at a time (the optimiser's inner loop) --
    partial = None
    if world == 1 and with_partial:
        edges = lf.topo.edges
        sel = edges if len(edges) <= 128 else list(np.random.default_rng(0).choice(edges, 128, replace=False))
        d = np.nan_to_num(base)[:, None] * rates[None, :]
        eng.update(d)
        # the inputs of the timed loop are prepared outside it: one distance table per update (ONE branch differs
        # from the previous table), the unit count of every update
        tables, units = [], 0
        for rep in range(3 if len(sel) < 512 else 1):
            for n in sel:
                d2 = d.copy()
                d2[n] *= 1.01 + 0.001 * rep
                tables.append(d2)
                units += lf.patterns.work_units(bins, nodes=set(lf.topo.ancestors(n)))[1]
        for d2 in tables[: len(sel)]:  # warm-up: every structure has been planned (and graph-captured) once
            eng.update(d2)
        eng.update(d)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for d2 in tables:
            eng.update(d2)
        t_part = time.perf_counter() - t0
        eng.set_profiling(True)
        eng.update(tables[0])
        eng.update(tables[1])
        launches_of_one = [{"kind": r["kind"], "us": round(r["ms"] * 1e3, 2)} for r in eng.launch_profile()]
        eng.set_profiling(False)
        partial = {"evals_per_s": len(tables) / t_part, "updates_per_s": units / t_part, "us_per_eval": t_part / len(tables) * 1e6,
                   "edges_sampled": len(sel), "evals_timed": len(tables), "mean_units_per_eval": units / len(tables),
                   "launches_of_one_update": launches_of_one}  # fmt: skip

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

    st = eng.stats()
    root_patterns = int(lf.patterns.root.n_rows - 1)
    # the WHOLE device-timed step against the roofline: every launch's algorithmic bytes (pruning launches and
    # the reduction) over the step's own time -- launches on several streams overlap, so the sum of per-launch
    # durations above no longer adds up to the step
    if lf.model.n_states == 4:
        step_bytes = sum(r["bytes"] for r in roofline_rows_of_step(eng, base, rates))
        t_step = dev_s / steps
        roofline["whole_step"] = {"alg_bytes_per_step": step_bytes, "ms_per_step": t_step * 1e3,
                                  "achieved": step_bytes / t_step / 1e9, "frac": step_bytes / t_step / 1e9 / roofline["peak"],
                                  "note": "all launches of a device-timed step (per-rank), algorithmic bytes / step time"}  # fmt: skip
    out = {
        "problem": problem, "lf": lf,
        "value": total_units * steps / dev_s, "ms_per_step": dev_s / steps * 1e3, "steps": steps,
        "gpu_launches": int(launches), "graph_replays": int(replays),
        "clocks": clock_rec, "roofline": roofline, "pipelined": pipelined, "partial_update": partial,
        "units": {"U_gp_per_step": total_units, "U_mv_per_step": total_umv, "lf_evals_per_s": steps / dev_s,
                  "root_patterns_per_gpu": root_patterns, "node_rows_per_step": int(st["last_node_rows"])},
        "allreduce": ({"kind": "fused into the reduction kernel (peer-memory mailboxes over NVLink); the sharded "
                               "evaluation replays from a CUDA graph (evaluation number in device memory)",
                       "check": allreduce_check} if fused else
                      ({"kind": "NCCL all-reduce of one double per step"} if world > 1 else None)),
        "setup": {**setup_t, "device_bytes": int(st["device_bytes"]),
                  "partial_bytes": int(st["partial_bytes"]), "levels": int(st["n_levels"])},
        "lnL": lnl,
        "e2e_mirror": {"value": total_units * steps / mirror_s, "unit": UNIT, "ms_per_step": mirror_s / steps * 1e3,
                       "h2d_bytes_per_step": h2d / steps, "d2h_bytes_per_step": 16,
                       "api": "cogent3_b200.likelihood.LikelihoodFunction.set_param_rule + get_log_likelihood "
                              "(host Q/eig, c3_set_eigen, c3_set_distances, c3_evaluate)"},
    }  # fmt: skip

    # ---- end to end through the PRODUCT: real cogent3 + install() + CUDA engine ----------------------
    dropin = None
    if with_dropin:
        eng_keep = eng  # (both engines stay resident: the drop-in builds its own from cogent3's tables)
        try:
            dropin = dropin_leg(ctx, problem, cfg, steps=min(steps, 400), warmup=warmup, expect_units=u_gp)
        except Exception as err:  # noqa: BLE001 - the line is still printed, with the reason
            dropin = {"error": repr(err)}
        del eng_keep
    have = ctx.sum_over_ranks(1.0 if (dropin and "wall_s" in dropin) else 0.0)
    if dropin and "wall_s" in dropin and have == world:
        n = min(steps, 400)
        wall = ctx.max_over_ranks(dropin["wall_s"])
        out["e2e"] = {"value": total_units * n / wall, "unit": UNIT, "ms_per_step": wall / n * 1e3, "steps": n,
                      "h2d_bytes_per_step": dropin["h2d"] / n, "d2h_bytes_per_step": 16,
                      "api": "cogent3 Calculator.__call__ of sm.make_likelihood_function(tree) with "
                             "cogent3_b200.install(): calc(x_k), every optimisable parameter changed (the call the "
                             "reference arm times); unmodified cogent3 from baseline/_ref on the host, libc3cuda on "
                             "the device",
                      "n_parameters": dropin["n_parameters"], "lf_evals": dropin["lf_evals"],
                      "lengths_only": ({"ms_per_step": dropin["lengths_only_ms"],
                                        "value": total_units / (dropin["lengths_only_ms"] * 1e-3),
                                        "step": "every branch length changed, rate-matrix parameters unchanged"}
                                       if dropin.get("lengths_only_ms") else None),
                      "fused_allreduce": dropin["fused"], "graph_replays": dropin["graph_replays"],
                      "setup": dropin["setup"], "lnL_first": dropin["lnL0"]}  # fmt: skip
        if dropin.get("optimise"):
            out["e2e"]["optimise"] = dropin["optimise"]
    else:
        why = (dropin or {}).get("error", "cogent3 (baseline/_ref) is not on this machine")
        out["e2e"] = dict(out["e2e_mirror"], note=f"drop-in leg unavailable ({why}): the mirror API instead")
    return out


def c5_strong_block(ctx, steps, warmup):
    """C5 as `north_star` states it: ONE 256-taxon x 10M-column HKY85+Gamma4 alignment, its columns split over the
    N ranks (strong scaling).  The alignment is the concatenation of 8 blocks of 1.25M columns (block j = the
    synthetic generator's shard j), rank r holds blocks [8r/N, 8(r+1)/N): the same 10M columns at every N; at
    N = 1 one GPU holds them all (35.9 GB resident)."""
    torch, dist, world, rank = ctx.torch, ctx.dist, ctx.world, ctx.rank
    model_name, n_tips, L, bins, _ = CONFIGS["c5"]
    n_blocks = 8
    if n_blocks % world != 0:
        return {"skipped": f"{world} ranks do not divide the {n_blocks} column blocks"}
    per = n_blocks // world
    m = hostmodels.get_model(model_name)
    t0 = time.perf_counter()
    parts = []
    topo = lengths = None
    for j in range(rank * per, (rank + 1) * per):
        topo, lengths, codes = make_workload(n_tips, L // n_blocks, m.n_states, kind="evolved", seed=SEED, shard=j)
        parts.append(codes)
    codes = np.concatenate(parts, axis=1) if len(parts) > 1 else parts[0]
    del parts
    t_gen = time.perf_counter() - t0
    # (motif probabilities fixed, not counted from the rank's own columns: every rank, and every N, must
    #  evaluate the SAME likelihood function -- the total lnL is then the same number at N = 1, 2, 4, 8)
    problem = (model_name, topo, lengths, codes, np.eye(m.n_states), bins,
               {"params": {"kappa": 4.0}, "rate_shape": 0.5, "mprobs": np.full(m.n_states, 1.0 / m.n_states)})  # fmt: skip
    try:
        lf, setup_t = make_lf(problem, device=ctx.local_rank, stream=ctx.stream)
    except MemoryError as err:
        return {"skipped": f"does not fit: {err}"}
    eng = lf.engine
    u_mv, u_gp = lf.patterns.work_units(bins)
    base, rates = lf.lengths.copy(), lf.rates()
    fused = False
    if world > 1:
        from cogent3_b200.sharded import attach_fused_allreduce

        fused = attach_fused_allreduce(eng)
    lnl_dev = torch.zeros(1, dtype=torch.float64, device="cuda")
    step_dist = [np.ascontiguousarray(np.nan_to_num(step_lengths(base, k))[:, None] * rates[None, :])
                 for k in range(7)]  # fmt: skip

    def device_step(k):
        d = step_dist[k % 7]
        if world == 1 or fused:
            return eng.update(d)
        eng.set_distances(d)
        eng.evaluate_device(lnl_dev.data_ptr())
        dist.all_reduce(lnl_dev)
        return float(lnl_dev.item())

    lf.get_log_likelihood()
    for k in range(max(3, warmup)):
        device_step(k)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    ev0.record(ctx.stream)
    t0 = time.perf_counter()
    for k in range(steps):
        lnl = device_step(k)
    ev1.record(ctx.stream)
    ctx.barrier()
    wall = time.perf_counter() - t0
    lnl = device_step(0)  # (the reported value: always the same parameter point, whatever the step count)
    dev_s = ctx.max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
    wall = ctx.max_over_ranks(wall)
    total_units = ctx.sum_over_ranks(float(u_gp))
    rows = ctx.sum_over_ranks(float(lf.patterns.root.n_rows - 1))
    check = None
    if fused:
        d = step_dist[3]
        total = eng.update(d)
        eng.comm_detach()
        eng.invalidate()
        t = torch.tensor([eng.update(d)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        check = {"fused_total": total, "nccl_sum_of_shards": float(t.item()),
                 "rel_diff": abs(total - float(t.item())) / abs(total)}
    st = eng.stats()
    hbm_peak, _ = peaks()
    alg_bytes = ctx.sum_over_ranks(float(st["last_alg_bytes"]))
    out = {
        "workload": "c5: HKY85+Gamma4 256 taxa x 10M columns, ONE alignment, columns split over the ranks",
        "scaling": "strong", "n_gpus": world, "columns_total": int(L), "columns_per_gpu": int(codes.shape[1]),
        "value": total_units * steps / dev_s, "unit": UNIT, "ms_per_step": dev_s / steps * 1e3, "steps": steps,
        "wall_ms_per_step": wall / steps * 1e3,
        "lnL": lnl, "lnL_note": "total over all ranks at the parameter point of step 0: the same number at every N",
        "U_gp_per_step_all_ranks": total_units, "root_pattern_rows_all_ranks": rows,
        "pattern_note": "per-shard compression: every rank compresses its own columns, so the N-rank total of "
                        "pattern rows / updates is >= the 1-rank (global compression) figure",
        "hbm_gbs_algorithmic_all_ranks": alg_bytes / (dev_s / steps) / 1e9,
        "hbm_frac_per_gpu": alg_bytes / (dev_s / steps) / 1e9 / world / hbm_peak,
        "allreduce": {"fused": bool(fused), "check": check} if world > 1 else None,
        "setup": {**setup_t, "generate_s": t_gen, "device_bytes": int(st["device_bytes"])},
    }  # fmt: skip
    eng.close()
    return out


def slim(block):
    """what of a measure_workload record goes into a secondary block of the JSON line"""
    keep = ("value", "ms_per_step", "steps", "gpu_launches", "graph_replays", "roofline", "units", "e2e", "e2e_mirror",
            "partial_update", "pipelined", "setup", "lnL", "cpu_baseline")  # fmt: skip
    return {k: block[k] for k in keep if k in block}


def cpu_baseline_subprocess(args, cfg, columns=None, kind="evolved", cpu_sample=None):
    # in a fresh process: torch's OpenMP runtime settings must not throttle the CPU arm
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", cfg,
           "--kind", kind, "--steps", "20", "--warmup", "1", "--no-full-reference"]  # fmt: skip  (bounded: ~10-30 s)
    if columns:
        cmd += ["--columns", str(columns)]
    if cpu_sample:
        cmd += ["--cpu-sample", str(cpu_sample)]
    env = {k: v for k, v in os.environ.items() if not k.startswith("OMP_")}
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        return json.loads(res.stdout.strip().splitlines()[-1])["cpu_baseline"]
    except Exception as err:  # noqa: BLE001 - the GPU line must still be printed
        return {"error": f"cpu baseline failed: {err!r}"}


def many_small_block(ctx, problem, lf, B=64):
    """many small alignments at once: B independent engines of the same size, all branch lengths
    changed, through c3_update_many (every engine on its own stream; a batch that keeps coming up is ONE graph)"""
    from cogent3_b200.engine import update_many as engine_update_many

    torch = ctx.torch
    base, rates = lf.lengths.copy(), lf.rates()
    step_dist = [np.ascontiguousarray(np.nan_to_num(step_lengths(base, k))[:, None] * rates[None, :])
                 for k in range(7)]  # fmt: skip
    lfs = [make_lf(problem, device=ctx.local_rank)[0] for _ in range(B)]
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
    del one_by_one
    for f in lfs:
        f.engine.close()
    return {"engines": B, "evals_per_s": B / t_many, "ms_per_batch": t_many * 1e3,
            "one_at_a_time_evals_per_s": B / t_seq, "all_equal": bool(len(set(vals)) == 1),
            "api": "c3_update_many: every engine on its own stream; a batch that keeps coming up is ONE CUDA graph"}


def lockstep_block(ctx, problem, B=16, n_evals=300):
    """the same idea THROUGH cogent3: B likelihood functions built by sm.make_likelihood_function (hook installed),
    each driven by its own thread making real one-parameter changes through Calculator.change (an optimiser's
    inner loop); cogent3_b200.batch lets their evaluations meet in one c3_evaluate_many call per round"""
    if import_cogent3() is None:
        return None
    from cogent3_b200.batch import run_lockstep

    calcs = []
    for _ in range(B):
        lf, _t = build_cogent3_lf(problem, "c1", backend="cuda", device=ctx.local_rank)
        lf.get_log_likelihood()
        calcs.append(lf.make_calculator())
    xs = [np.array(c.get_value_array(), float) for c in calcs]

    def drive(calc, x, n):
        k = 0
        for _ in range(n):
            i = k % len(x)
            k += 1
            calc.change([(i, x[i] * (1.0 + 1e-6 * (1 + k % 89)))])

    for c, x in zip(calcs, xs, strict=True):
        drive(c, x, 40)  # warm-up (graph capture thresholds)
    t0 = time.perf_counter()
    for c, x in zip(calcs, xs, strict=True):
        drive(c, x, n_evals)
    t_seq = time.perf_counter() - t0
    run_lockstep([(lambda c=c, x=x: drive(c, x, 40)) for c, x in zip(calcs, xs, strict=True)])
    t0 = time.perf_counter()
    _res, batch = run_lockstep([(lambda c=c, x=x: drive(c, x, n_evals)) for c, x in zip(calcs, xs, strict=True)])
    t_par = time.perf_counter() - t0
    return {"likelihood_functions": B, "evals_each": n_evals,
            "one_after_the_other_evals_per_s": B * n_evals / t_seq, "lockstep_evals_per_s": B * n_evals / t_par,
            "mean_batch": batch.evaluations / max(1, batch.batches),
            "api": "cogent3 Calculator.change in B threads, evaluations batched by cogent3_b200.batch (c3_evaluate_many)"}


def run_ours(args):
    ctx = Ctx()
    torch, dist, world, rank = ctx.torch, ctx.dist, ctx.world, ctx.rank
    clocks = ClockSampler(ctx.local_rank) if rank == 0 else None
    main_rec = measure_workload(ctx, args, args.config, args.columns, args.kind, steps=args.steps,
                                warmup=max(args.warmup, 3), shard=rank, clocks=clocks,
                                with_dropin=not args.no_dropin and (args.columns or CONFIGS[args.config][2]) <= 2_500_000)  # fmt: skip
    problem, lf = main_rec.pop("problem"), main_rec.pop("lf")

    many = lockstep = None
    if world == 1 and args.config == "c1":
        many = many_small_block(ctx, problem, lf)
        if not args.no_dropin:
            try:
                lockstep = lockstep_block(ctx, problem)
            except Exception as err:  # noqa: BLE001
                lockstep = {"error": repr(err)}
    lf.engine.close()
    del lf

    # ---- what else `north_star` names, in the same line (default config only) ------------------------
    extras = args.config == "c2" and not args.columns and args.kind == "evolved" and not args.no_extra
    c4 = c5 = c1 = None
    if extras and world == 1:
        # C1: the one BASELINE config on real data (brca1, 55 taxa x 3009 columns): latency, not bandwidth
        try:
            rec = measure_workload(ctx, args, "c1", None, "evolved", steps=min(10 * args.steps, 2000), warmup=20, shard=0,
                                   with_dropin=not args.no_dropin)  # fmt: skip
            p1, lf1 = rec.pop("problem"), rec.pop("lf")
            c1 = {"workload": f"c1: {CONFIGS['c1'][4]}", "us_per_full_evaluation": rec["ms_per_step"] * 1e3,
                  "full_evals_per_s": rec["units"]["lf_evals_per_s"], "partial_update": rec["partial_update"],
                  "e2e": rec["e2e"], "many_small": many_small_block(ctx, p1, lf1)}
            if not args.no_dropin:
                c1["lockstep"] = lockstep_block(ctx, p1)
            lf1.engine.close()
        except Exception as err:  # noqa: BLE001
            c1 = {"error": repr(err)}
    if extras and world == 1:
        try:
            rec = measure_workload(ctx, args, "c4", None, "evolved", steps=min(args.steps, 300), warmup=3, shard=0,
                                   with_dropin=not args.no_dropin, with_partial=False)  # fmt: skip
            rec.pop("problem")
            rec.pop("lf").engine.close()
            if not args.no_cpu:
                rec["cpu_baseline"] = cpu_baseline_subprocess(args, "c4")
            c4 = slim(rec)
            c4["workload"] = f"c4: {CONFIGS['c4'][4]}"
        except Exception as err:  # noqa: BLE001
            c4 = {"error": repr(err)}
        torch.cuda.empty_cache()
    if extras:
        try:
            c5 = c5_strong_block(ctx, steps=min(args.steps, 50 if world == 1 else 200), warmup=3)
        except Exception as err:  # noqa: BLE001
            c5 = {"error": repr(err)}
        ok = ctx.sum_over_ranks(1.0 if (c5 and "value" in c5) else 0.0)
        if ok != world and c5 and "value" in c5:
            c5 = {"error": "another rank failed"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu_rec = None
    if world == 1 and not args.no_cpu:
        cpu_rec = cpu_baseline_subprocess(args, args.config, args.columns, args.kind, args.cpu_sample)

    line = {
        "metric": METRIC, "value": main_rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": main_rec["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, problem),
        "e2e": main_rec["e2e"], "e2e_mirror": main_rec["e2e_mirror"],
        "gpu_launches": main_rec["gpu_launches"], "graph_replays": main_rec["graph_replays"],
        "clocks": main_rec["clocks"],
        "roofline": main_rec["roofline"],
        "cpu_baseline": cpu_rec,
        "units": main_rec["units"],
        "partial_update": main_rec["partial_update"], "pipelined": main_rec.get("pipelined"),
        "many_small": many, "lockstep": lockstep,
        "c1": c1,
        "allreduce": main_rec["allreduce"],
        "setup": main_rec["setup"],
        "lnL": main_rec["lnL"],
        "c4": c4,
        "c5_strong": c5,
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
    ap.add_argument("--no-full-reference", dest="full_reference", action="store_false",
                    help="reference arm: only the bounded sample, not the full-size run")
    ap.add_argument("--no-extra", action="store_true", help="our arm: skip the c4 and c5_strong blocks")
    ap.add_argument("--no-dropin", action="store_true", help="our arm: skip the end-to-end leg through real cogent3")
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
