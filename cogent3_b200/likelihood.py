"""Standalone likelihood function over the CUDA engine.

A small host-side mirror of the part of cogent3's ``AlignmentLikelihoodFunction``
API the hot path is reached through (evolve/parameter_controller.py:511-570,
evolve/likelihood_function.py:269-270, 415-437, 514-520): ``set_alignment``,
``set_param_rule``, ``set_motif_probs``, ``get_log_likelihood``, ``get_psub_for_edge``,
``get_full_length_likelihoods``, ``get_bin_probs``, ``optimise``.  It exists so the
engine can be driven, tested and benchmarked where cogent3 is not installed (the
GPU box); with cogent3 present use :mod:`cogent3_b200.dropin`, which plugs the
same engine under the real cogent3 objects.

Host work per evaluation is O(#parameters): build Q (tiny), eigendecompose on the
host with numpy exactly as the reference does (bit-identical eigenpairs,
maths/matrix_exponentiation.py:132-146), hand distances to the engine.  All
O(patterns) arithmetic is on the device.
"""

from __future__ import annotations

import numpy as np

from . import models as hostmodels
from .patterns import PatternTree, compress_alignment
from .tree import Topology


def _default_engine_factory(patterns, n_bins, device=0, stream=None):
    from .engine import LikelihoodEngine  # fails loudly without libc3cuda / a GPU

    return LikelihoodEngine(patterns, n_bins=n_bins, device=device, stream=stream)


class LikelihoodFunction:
    def __init__(self, model, topo: Topology, bins=1, expm="either", device=0, stream=None,
                 engine_factory=None, default_length=1.0):  # fmt: skip
        self.model = hostmodels.get_model(model) if isinstance(model, str) else model
        self.topo = topo
        self.n_bins = int(bins)
        self.bin_names = [f"bin{i}" for i in range(self.n_bins)] if bins > 1 else ["bin"]
        self.expm = expm
        self.device = device
        self.stream = stream
        self._engine_factory = engine_factory or _default_engine_factory
        self.engine = None
        self.patterns: PatternTree | None = None
        M = self.model.n_states
        self.mprobs = np.full(M, 1.0 / M)
        self.lengths = np.full(topo.n_nodes, float(default_length))
        self.lengths[topo.root] = np.nan
        self.params = self.model.default_params()  # global values
        self.edge_params: dict[int, dict[str, float]] = {}  # per-edge overrides (time-het)
        self.rate_shape = 1.0
        self.bprobs = np.full(self.n_bins, 1.0 / self.n_bins)
        self._dirty_q = True
        self._dirty_model = True
        self._slot_of_edge = np.zeros(topo.n_nodes, np.int32)
        self._slot_keys: list = []
        self._slot_map_sent = None
        self._n_evals = 0

    # ------------------------------------------------------------------ data
    def set_alignment(self, codes_by_tip, symbol_rows=None, patterns: PatternTree | None = None,
                      compress="auto"):  # fmt: skip
        """``codes_by_tip``: {tip name: int codes[L]} or array [n_tips, L] in
        ``topo.tip_names`` order; ``symbol_rows[S, M]`` indicator row per code.
        Runs the per-node pattern compression (a2/a3) -- on the device for the CUDA engine
        (``compress="device"``, the default there; c3_compress_alignment), on the host with
        numpy otherwise (``"host"``) -- and makes the tables resident."""
        if self.engine is not None:
            self.engine.close()
            self.engine = None
        self._tip_codes = None
        if patterns is None:
            if symbol_rows is None:
                M = self.model.n_states
                symbol_rows = (
                    hostmodels.dna_symbol_rows()[1] if M == 4 else hostmodels.codon_symbol_rows()[1]
                )
            symbol_rows = np.asarray(symbol_rows, np.float64)
            on_device = compress == "device" or (
                compress == "auto" and self._engine_factory is _default_engine_factory
                and symbol_rows.shape[0] <= 256
            )  # fmt: skip
            if on_device:
                if isinstance(codes_by_tip, dict):
                    arr = np.stack([np.asarray(codes_by_tip[nm]) for nm in self.topo.tip_names])
                else:
                    arr = np.asarray(codes_by_tip)
                if arr.size and (arr.min() < 0 or arr.max() >= symbol_rows.shape[0]):
                    raise ValueError("symbol code out of range")
                from .engine import LikelihoodEngine

                self._tip_codes = arr.astype(np.uint8, copy=False)
                self._symbol_rows = symbol_rows
                self.engine = LikelihoodEngine.from_alignment(
                    self.topo, self._tip_codes, symbol_rows, n_bins=self.n_bins, device=self.device,
                    stream=self.stream,
                )  # fmt: skip
                self.patterns = self.engine.patterns
            else:
                patterns = compress_alignment(self.topo, codes_by_tip, symbol_rows)
        if self.engine is None:
            self.patterns = patterns
            self.engine = self._engine_factory(
                patterns, self.n_bins, device=self.device, stream=self.stream
            )
        self._dirty_q = self._dirty_model = True
        self._slot_keys, self._slot_map_sent = [], None
        return self

    def set_motif_probs(self, mprobs):
        mprobs = np.asarray(mprobs, float)
        assert mprobs.shape == (self.model.n_states,)
        self.mprobs = mprobs
        self._dirty_q = self._dirty_model = True

    def state_counts(self):
        """unambiguous-state counts over all tips (``count_motifs``,
        evolve/motif_prob_model.py:48-59, include_ambiguity=False)."""
        counts = np.zeros(self.model.n_states)
        if getattr(self, "_tip_codes", None) is not None:
            # device-compressed alignment: the same sum straight from the symbol codes
            rows = self._symbol_rows
            unambig = rows.sum(axis=1) == 1.0
            for codes in self._tip_codes:
                per_symbol = np.bincount(codes, minlength=rows.shape[0]).astype(np.float64)
                counts += (rows[unambig] * per_symbol[unambig, None]).sum(axis=0)
            return counts
        for n in self.topo.tips:
            nd = self.patterns.nodes[n]
            unambig = nd.ambig == 1.0
            counts += (nd.table[unambig] * nd.counts[unambig, None]).sum(axis=0)
        return counts

    def set_motif_probs_from_data(self):
        """unambiguous-state frequencies over all tips, like
        ``set_motif_probs_from_data`` (evolve/parameter_controller.py:135-162)."""
        counts = self.state_counts()
        self.set_motif_probs(counts / counts.sum())

    # ------------------------------------------------------------ parameters
    def set_param_rule(self, par_name, value=None, edge=None, edges=None, **_ignored):
        names = [edge] if edge is not None else edges
        ids = None if names is None else [self.topo.index(n) if isinstance(n, str) else n for n in names]
        if par_name == "length":
            if ids is None:
                self.lengths[:-1] = value
            else:
                self.lengths[ids] = value
            self._dirty_model = True
        elif par_name == "rate_shape":
            self.rate_shape = float(value)
            self._dirty_model = True
        elif par_name == "bprobs":
            if value is not None:
                self.bprobs = np.asarray(value, float)
                self._dirty_model = True
        elif par_name in self.model.parameter_order:
            if ids is None:
                self.params[par_name] = float(value)
                for ov in self.edge_params.values():
                    ov.pop(par_name, None)
            else:
                for i in ids:
                    self.edge_params.setdefault(i, {})[par_name] = float(value)
            self._dirty_q = True
        else:
            raise KeyError(par_name)

    def set_time_heterogeneity(self, is_independent=True):
        """every edge gets its own copy of the rate parameters (per-edge Q), cf.
        ``set_time_heterogeneity`` (evolve/parameter_controller.py:246-345)."""
        if is_independent:
            for n in self.topo.edges:
                self.edge_params.setdefault(n, dict(self.params))
            self._dirty_q = True

    def get_param_value(self, par_name, edge=None, bin=None):  # noqa: A002
        if par_name == "length":
            return float(self.lengths[self.topo.index(edge)])
        if par_name == "rate_shape":
            return self.rate_shape
        if par_name == "bprobs":
            return self.bprobs.copy()
        if par_name == "mprobs":
            return self.mprobs.copy()
        if par_name == "rate":
            return float(self.rates()[self._bin_index(bin)])
        if par_name == "lh":
            self.get_log_likelihood()
            return self.engine.get_lh(self._bin_index(bin))
        if par_name == "psubs":
            return self.get_psub_for_edge(edge, bin=bin)
        if par_name == "Q":
            n = self.topo.index(edge)
            return self.model.calcQ(self.mprobs, self._params_of_edge(n))
        if par_name in self.model.parameter_order:
            if edge is not None:
                return self._params_of_edge(self.topo.index(edge))[par_name]
            return self.params[par_name]
        raise KeyError(par_name)

    def _bin_index(self, bin_):
        if bin_ is None:
            return 0
        return self.bin_names.index(bin_) if isinstance(bin_, str) else int(bin_)

    def _params_of_edge(self, n):
        ov = self.edge_params.get(n)
        return self.params if not ov else {**self.params, **ov}

    def rates(self):
        if self.n_bins == 1:
            return np.ones(1)
        key = (self.rate_shape, self.bprobs.tobytes())
        if getattr(self, "_rates_key", None) != key:  # scipy's gamma.ppf is slow: cache
            self._rates = hostmodels.gamma_rates(self.n_bins, self.rate_shape, self.bprobs)
            self._rates_key = key
        return self._rates

    # ------------------------------------------------------------ evaluation
    def _sync_q(self):
        """(re)build the distinct rate matrices and their exponentiator state."""
        order = self.model.parameter_order
        keys, slot_of_edge = {}, np.zeros(self.topo.n_nodes, np.int32)
        if not self.edge_params:  # one Q for the whole tree (the common case)
            keys[tuple(self.params[k] for k in order)] = 0
        else:
            for n in self.topo.edges:
                p = self._params_of_edge(n)
                slot_of_edge[n] = keys.setdefault(tuple(p[k] for k in order), len(keys))
        mp = self.mprobs.tobytes()
        for key, slot in keys.items():
            if slot < len(self._slot_keys) and self._slot_keys[slot] == (key, mp):
                continue
            Q = self.model.calcQ(self.mprobs, dict(zip(order, key, strict=True)))
            self._upload_exponentiator(slot, Q)
        self._slot_keys = [(k, mp) for k in keys]
        if self._slot_map_sent is None or not np.array_equal(slot_of_edge, self._slot_map_sent):
            self.engine.set_edge_qslots(np.repeat(slot_of_edge[:, None], self.n_bins, axis=1))
            self._slot_map_sent = slot_of_edge
        self._slot_of_edge = slot_of_edge
        self._dirty_q = False

    def _upload_exponentiator(self, slot, Q):
        """``ExpDefn.calc`` + ``_EigenPade`` (evolve/substitution_calculation.py:48-86)."""
        allow_eigen, check_eigen, allow_pade = {
            "eigen": (True, False, False),
            "checked": (True, True, False),
            "pade": (False, False, True),
            "either": (True, True, True),
        }[str(self.expm)]
        if allow_eigen:
            try:
                es = hostmodels.EigenSystem(Q, check=check_eigen)
                self.engine.set_eigen(slot, es.roots, es.evT, es.evI)
                return
            except (ArithmeticError, np.linalg.LinAlgError):
                if not allow_pade:
                    raise
        self.engine.set_Q(slot, Q)

    def _sync(self):
        if self.engine is None:
            raise RuntimeError("set_alignment first")
        if self._dirty_q:
            self._sync_q()
        if self._dirty_model:
            self.engine.set_root_mprobs(self.mprobs)
            if self.n_bins > 1:
                self.engine.set_bin_probs(self.bprobs)
            lengths = self.lengths.copy()
            lengths[self.topo.root] = 0.0  # the root has no branch (NaN placeholder)
            dist = lengths[:, None] * self.rates()[None, :]
            self.engine.set_distances(dist)
            self._dirty_model = False

    def get_log_likelihood(self) -> float:
        self._sync()
        self._n_evals += 1
        return self.engine.evaluate()

    @property
    def lnL(self):
        return self.get_log_likelihood()

    # -------------------------------------------------------------- accessors
    def get_psub_for_edge(self, name, bin=None):  # noqa: A002
        self.get_log_likelihood()
        return self.engine.get_psub(self.topo.index(name), self._bin_index(bin))

    def get_full_length_likelihoods(self, bin=None):  # noqa: A002
        """per-column likelihoods (``get_full_length_likelihoods``,
        evolve/likelihood_tree.py:93-94): lh[root.index]"""
        if hasattr(self.engine, "get_site_lh"):
            self.get_log_likelihood()
            return self.engine.get_site_lh(self._bin_index(bin))  # gathered on the device
        lh = self.get_param_value("lh", bin=bin)
        return lh[self.patterns.root.index]

    def get_bin_probs(self):
        """posterior bin probabilities per column (``BinnedLikelihood.get_posterior_probs``,
        evolve/likelihood_calculation.py:247-257)."""
        self.get_log_likelihood()
        if hasattr(self.engine, "get_site_lh"):
            res = np.array([b * self.engine.get_site_lh(k) for k, b in enumerate(self.bprobs)])
        else:
            idx = self.patterns.root.index
            res = np.array([b * self.engine.get_lh(k)[idx] for k, b in enumerate(self.bprobs)])
        res /= res.sum(axis=0)
        return res

    def get_num_free_params(self):
        return len(self._free_vector()[0])

    # -------------------------------------------------------------- optimise
    def _free_vector(self):
        """free parameters: every branch length, every (global / per-edge) rate term,
        rate_shape when K>1 -- what cogent3 optimises by default."""
        spec, x = [], []
        for n in self.topo.edges:
            spec.append(("length", n, None))
            x.append(self.lengths[n])
        if self.edge_params:
            for n in sorted(self.edge_params):
                for p in self.model.parameter_order:
                    spec.append(("edge_param", n, p))
                    x.append(self._params_of_edge(n)[p])
        else:
            for p in self.model.parameter_order:
                spec.append(("param", None, p))
                x.append(self.params[p])
        if self.n_bins > 1:
            spec.append(("rate_shape", None, None))
            x.append(self.rate_shape)
        return spec, np.array(x, float)

    _BOUNDS = {"length": (1e-6, 10.0), "param": (1e-6, 1e6), "edge_param": (1e-6, 1e6),
               "rate_shape": (1e-2, 1e3)}  # fmt: skip

    def _apply_vector(self, spec, x):
        for (kind, n, p), v in zip(spec, x, strict=True):
            if kind == "length":
                if self.lengths[n] != v:
                    self.lengths[n] = v
                    self._dirty_model = True
            elif kind == "param":
                if self.params[p] != v:
                    self.params[p] = v
                    self._dirty_q = True
            elif kind == "edge_param":
                if self.edge_params[n].get(p) != v:
                    self.edge_params[n][p] = v
                    self._dirty_q = True
            elif self.rate_shape != v:
                self.rate_shape = v
                self._dirty_model = True

    def optimise(self, max_evaluations=None, tolerance=1e-6, show_progress=False, **_kw):
        """local (Powell) maximisation of lnL in log-parameter space.  In the drop-in
        cogent3's own optimisers are used unchanged; this is the standalone stand-in."""
        from scipy.optimize import minimize

        spec, x0 = self._free_vector()
        lo = np.array([self._BOUNDS[k][0] for k, _, _ in spec])
        hi = np.array([self._BOUNDS[k][1] for k, _, _ in spec])
        evals = [0]

        def f(z):
            x = np.clip(np.exp(z), lo, hi)
            self._apply_vector(spec, x)
            try:
                v = self.get_log_likelihood()
            except ArithmeticError:
                return np.inf
            evals[0] += 1
            return -v if np.isfinite(v) else np.inf

        opts = {"xtol": tolerance, "ftol": tolerance}
        if max_evaluations:
            opts["maxfev"] = int(max_evaluations)
        res = minimize(f, np.log(np.clip(x0, lo, hi)), method="Powell", options=opts,
                       bounds=list(zip(np.log(lo), np.log(hi), strict=True)))  # fmt: skip
        self._apply_vector(spec, np.clip(np.exp(res.x), lo, hi))
        self.last_optimise_evaluations = evals[0]
        return self.get_log_likelihood()


def evaluate_many(lfs) -> np.ndarray:
    """log-likelihoods of several independent likelihood functions, evaluated concurrently on
    the GPU (many small alignments).  Host-side parameter syncing stays per
    function; the device work of all of them is in flight at once."""
    from .engine import evaluate_many as engine_many

    for lf in lfs:
        lf._sync()
    return engine_many([lf.engine for lf in lfs])
