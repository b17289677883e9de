"""Many small likelihood functions at once through cogent3's own optimisers.

The app layer's real workload is hundreds of independent small alignments (``app/evo.py:317-386`` model fits
over ``apply_to``, ``evolve/bootstrap.py:68`` replicates, ``phylo/maximum_likelihood.py:33-57`` candidate
trees).  A brca1-sized evaluation keeps a B200 busy for a few tens of microseconds spread over a handful of SMs
and costs as much again in launch latency: one likelihood function at a time cannot use the GPU.  The engine's
answer is ``c3_evaluate_many`` (every engine on its own stream, all launched before any is waited for; a
batch that keeps coming up becomes one CUDA graph) -- but cogent3's optimisers are sequential programs that
call ``f(x)`` and wait.

This module runs N such programs (``lf.optimise()``, or any callable) in N threads and lets their likelihood
evaluations MEET: the terminal Defn of the drop-in hands its engine to the batch instead of evaluating it;
when every live thread has arrived, one thread issues a single ``c3_evaluate_many`` for all of them and wakes
the others with their values.  Each optimiser still sees plain ``f(x) -> float``, makes its own decisions and
finishes on its own; Powell, simulated annealing, ``Calculator`` are unchanged.  (The Python of the N
optimisers is serialised by the GIL; what overlaps is the device work and its launch / wait latency.)

    from cogent3_b200.batch import optimise_many, fit_many
    optimise_many(lfs, local=True, show_progress=False)          # lfs built through the drop-in
    lfs = fit_many(get_model("HKY85"), tree, alignments, local=True)
"""

from __future__ import annotations

import threading

import numpy as np

_LOCAL = threading.local()


def current():
    """the batch the calling thread evaluates in, or None"""
    return getattr(_LOCAL, "batch", None)


class _Seat:
    """one waiting thread: its own lock to sleep on (no shared condition to fight over on wake-up)"""

    __slots__ = ("lock", "result")

    def __init__(self):
        self.lock = threading.Lock()
        self.lock.acquire()
        self.result = None


class Lockstep:
    """rendezvous of the engine evaluations of several threads (see the module docstring).

    A thread that arrives takes a seat and sleeps on the seat's own lock; the thread that completes the round
    (every live thread is seated) runs the batched call outside the mutex and releases every seat with its
    value.  No condition variable: a notify_all would send all sleepers after one mutex at once."""

    def __init__(self, n_threads: int):
        self.mutex = threading.Lock()
        self.live = int(n_threads)
        self.waiting: list = []  # [(seat, engine)]
        self.batches = 0
        self.evaluations = 0
        self._seats = threading.local()

    # -- called by the drop-in's terminal Defn (dropin.GpuLogLikelihoodDefn._finish) ---------------------
    def evaluate(self, engine) -> float:
        seat = getattr(self._seats, "seat", None)
        if seat is None:
            seat = self._seats.seat = _Seat()
        with self.mutex:
            self.waiting.append((seat, engine))
            batch = None
            if len(self.waiting) >= self.live:
                batch, self.waiting = self.waiting, []
        if batch is not None:
            self._run(batch)
        seat.lock.acquire()  # released by whoever ran the round (possibly this very thread)
        out = seat.result
        if isinstance(out, BaseException):
            raise out
        return out

    def _run(self, batch):
        """one batched call for every seated thread; every seat gets its own outcome"""
        engines = [e for _s, e in batch]
        try:
            values = _evaluate_engines(engines)
        except Exception:  # noqa: BLE001 - find out whose evaluation failed
            values = []
            for e in engines:
                try:
                    values.append(e.evaluate())
                except Exception as err:  # noqa: BLE001
                    values.append(err)
        self.batches += 1
        self.evaluations += len(batch)
        for (seat, _e), v in zip(batch, values, strict=True):
            seat.result = v
            seat.lock.release()

    def leave(self):
        """a thread is through: the others must not wait for it any more"""
        with self.mutex:
            self.live -= 1
            batch = None
            if self.waiting and len(self.waiting) >= self.live:
                batch, self.waiting = self.waiting, []
        if batch is not None:
            self._run(batch)


def _evaluate_engines(engines):
    if len(engines) == 1:
        return [engines[0].evaluate()]
    plain = [e for e in engines if hasattr(e, "_h")]
    if len(plain) == len(engines):
        from .engine import evaluate_many

        return [float(v) for v in evaluate_many(engines)]  # ONE FFI crossing: c3_evaluate_many
    return [e.evaluate() for e in engines]  # (test doubles, multi-device engines)


def run_lockstep(jobs):
    """run the callables ``jobs`` concurrently, one thread each, their engine evaluations in lock-step batches.
    Returns ([result or exception per job], Lockstep) -- an exception is re-raised by ``optimise_many``."""
    jobs = list(jobs)
    batch = Lockstep(len(jobs))
    results: list = [None] * len(jobs)

    def worker(i, job):
        _LOCAL.batch = batch
        try:
            results[i] = job()
        except BaseException as err:  # noqa: BLE001 - handed to the caller
            results[i] = err
        finally:
            _LOCAL.batch = None
            batch.leave()

    threads = [threading.Thread(target=worker, args=(i, job), daemon=True) for i, job in enumerate(jobs)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return results, batch


def optimise_many(lfs, return_stats=False, **opt_args):
    """``lf.optimise(**opt_args)`` for every likelihood function of ``lfs`` (built through the drop-in, alignments
    set), concurrently: per optimiser step ONE batched device call for all of them.  Each function ends at the
    point its own optimiser found -- the same point a sequential ``lf.optimise()`` finds."""
    opt_args.setdefault("show_progress", False)
    for lf in lfs:
        lf.get_log_likelihood()  # engines built (and CUDA context touched) in the calling thread
    results, batch = run_lockstep([(lambda lf=lf: lf.optimise(**opt_args)) for lf in lfs])
    for r in results:
        if isinstance(r, BaseException):
            raise r
    stats = {"batches": batch.batches, "evaluations": batch.evaluations,
             "mean_batch": batch.evaluations / max(1, batch.batches)}  # fmt: skip
    return (results, stats) if return_stats else results


def fit_many(model, tree, alignments, lf_args=None, configure=None, **opt_args):
    """one likelihood function per alignment (same model and tree: ``app/evo.py``'s model app over many
    alignments), all optimised together; returns the fitted likelihood functions"""
    from .dropin import make_likelihood_function

    lfs = []
    for aln in alignments:
        lf = make_likelihood_function(model, tree, **(lf_args or {}))
        lf.set_alignment(aln)
        if configure is not None:
            configure(lf)
        lfs.append(lf)
    optimise_many(lfs, **opt_args)
    return lfs


def log_likelihoods(lfs) -> np.ndarray:
    """current lnL of every function, evaluated together"""
    results, _ = run_lockstep([lf.get_log_likelihood for lf in lfs])
    for r in results:
        if isinstance(r, BaseException):
            raise r
    return np.array(results, float)


def run_estimate_distances(dist_calc, dist_opt_args=None, aln_opt_args=None, max_concurrent=64):
    """``EstimateDistances.run`` (evolve/distance.py:182-234) with the n (n - 1) / 2 two-taxon (or three-taxon)
    fits optimised CONCURRENTLY, their evaluations in lock-step batches (the pairwise path).
    ``dist_calc`` is a stock ``cogent3.evolve.distance.EstimateDistances``; ``cogent3_b200.install()`` must be
    active so that the fits build CUDA likelihood functions.  Results land where ``run`` puts them
    (``dist_calc.get_pairwise_distances()`` etc. work afterwards)."""
    from cogent3.evolve.distance import get_name_combinations

    dist_opt_args = dict(dist_opt_args or {})
    aln_opt_args = dict(aln_opt_args or {})
    dist_opt_args["local"] = dist_opt_args.get("local", True)
    aln_opt_args["local"] = aln_opt_args.get("local", True)
    dist_opt_args.setdefault("show_progress", False)
    combos = get_name_combinations(dist_calc._seq_collection.names, 3 if dist_calc._threeway else 2)
    stats = {"batches": 0, "evaluations": 0}
    for start in range(0, len(combos), max_concurrent):
        chunk = combos[start : start + max_concurrent]
        jobs = [(lambda comp=comp: dist_calc._doset(comp, dict(dist_opt_args), dict(aln_opt_args))) for comp in chunk]
        results, batch = run_lockstep(jobs)
        for comp, value in zip(chunk, results, strict=True):
            if isinstance(value, BaseException):
                raise value
            dist_calc._param_ests[comp] = value
        stats["batches"] += batch.batches
        stats["evaluations"] += batch.evaluations
    stats["mean_batch"] = stats["evaluations"] / max(1, stats["batches"])
    return stats
