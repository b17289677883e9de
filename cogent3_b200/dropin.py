"""Drop-in glue: the CUDA engine underneath cogent3's own likelihood-function objects.

cogent3 has no plugin group for likelihood back-ends; the operator API
the hot path sits behind is the recalculation Defn / cell protocol.  This module
replaces exactly one thing -- the sub-DAG that
``likelihood_calculation.make_total_loglikelihood_defn`` builds
(evolve/likelihood_calculation.py:125-167: one ``inner`` Defn per edge, one ``plh``
Defn per internal node, ``lh``, ``bdist``/``bindex``, ``tll``/``logsum``) -- by ONE
terminal Defn whose ``calc`` drives libc3cuda.  Everything upstream (alignment
conversion, per-node pattern compression ``lht``, motif probabilities, Q, the
eigendecomposition ``Qd``, gamma rates, ``distance``) and everything downstream
(``ParameterController``, ``Calculator``, Powell / simulated annealing) is cogent3's
unchanged code, and keeps returning the same numbers.

    from cogent3 import get_model
    from cogent3_b200.dropin import make_likelihood_function

    lf = make_likelihood_function(get_model("GTR"), tree, bins=4)   # instead of
    lf.set_alignment(aln)                 # sm.make_likelihood_function(tree, bins=4)
    lf.get_log_likelihood(); lf.optimise()

Protocol honoured (recalculation/definition.py:111-154, calculation.py:402-506):
``calc(*arg_values)`` is pure in its arguments, returns a Python float, raises
``ArithmeticError`` for a recoverable bad point; unchanged inputs keep object identity
between calls (calculation.py:431-438), which is how dirty edges are found.  Device
buffers are owned by the engine object cached on the Defn, never by cell values.
"""

from __future__ import annotations

import numpy as np

from .patterns import from_reference_lht

__all__ = ["CudaAlignmentLikelihoodFunction", "make_likelihood_function"]


def _default_engine_factory(patterns, n_bins, device=0, stream=None):
    from .engine import LikelihoodEngine  # no CPU fallback: raises without libc3cuda / a GPU

    return LikelihoodEngine(patterns, n_bins=n_bins, device=device, stream=stream)


class LazyPsub:
    """value of the ``psubs`` Defn in the CUDA graph: the exponentiator and the distance
    instead of the exponentiated matrix, so that P(t) is computed on the device
    (batched over all dirty edges and rate classes) and not by K*E numpy calls per
    evaluation (10-15 % of a CPU evaluation).  Materialises to exactly
    the reference's matrix when something reads it (``get_psub_for_edge`` etc.)."""

    __slots__ = ("exponentiator", "t")

    def __init__(self, exponentiator, t):
        self.exponentiator = exponentiator
        self.t = t

    def materialise(self):
        return self.exponentiator(self.t)

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self.materialise())
        return a.astype(dtype) if dtype is not None else a

    @property
    def shape(self):
        return self.materialise().shape

    def __getitem__(self, item):
        return self.materialise()[item]

    def copy(self):
        return self.materialise().copy()


def _build(cogent3_modules):
    """create the Defn / LF classes lazily so that importing this module does not
    require cogent3"""
    (likelihood_calculation, parameter_controller, definition, matrix_exponentiation) = cogent3_modules
    CalculationDefn = definition.CalculationDefn
    NonParamDefn = definition.NonParamDefn
    SumDefn = definition.SumDefn
    EigenExponentiator = matrix_exponentiation.EigenExponentiator
    PadeExponentiator = matrix_exponentiation.PadeExponentiator

    class LazyPsubDefn(CalculationDefn):
        """``psubs = CallDefn(Qd, distance)`` (evolve/substitution_model.py:702-710)
        without the call"""

        name = "psubs"

        def calc(self, exponentiator, distance):
            return LazyPsub(exponentiator, distance)

    class GpuLogLikelihoodDefn(CalculationDefn):
        """total log-likelihood of one locus, computed by libc3cuda.

        args: lht, root (LhtEdgeLookupDefn), root mprobs, [bprobs], then one
        fixed_motif per internal node (post-order), then one psub per (edge, bin)
        in edge-major order (post-order edges)."""

        name = "gpu_lnL"

        def setup(self, edge_names, internal_names, n_bins, engine_factory, device):
            self.edge_names = list(edge_names)
            self.internal_names = list(internal_names)
            self.n_bins = n_bins
            self.engine_factory = engine_factory
            self.device = device
            self._engines = {}  # id(lht) -> state

        # engines hold device memory and ctypes handles: never pickled / deep-copied
        def __getstate__(self):
            d = dict(self.__dict__)
            d["_engines"] = {}
            return d

        def __deepcopy__(self, memo):
            import copy

            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            for k, v in self.__dict__.items():
                new.__dict__[k] = {} if k == "_engines" else copy.deepcopy(v, memo)
            return new

        def _state_for(self, lht):
            st = self._engines.get(id(lht))
            if st is not None and st["lht"] is lht:
                return st
            # a new alignment (or locus): flatten cogent3's own pattern tables -- they stay
            # the source of truth -- and make the tables resident on the device
            topo, patterns = from_reference_lht(lht)
            engine = self.engine_factory(patterns, self.n_bins, device=self.device)
            node_of = {name: i for i, name in enumerate(topo.names)}
            st = {
                "lht": lht,
                "engine": engine,
                "topo": topo,
                "edge_ids": [node_of[n] for n in self.edge_names],
                "internal_ids": [node_of[n] for n in self.internal_names],
                "slot_of": {},  # id(exponentiator) -> slot
                "slot_obj": {},  # slot -> exponentiator (strong ref: ids are not recycled)
                "psub_obj": {},  # (node, bin) -> last directly supplied P object
                "qslots": np.full((topo.n_nodes, self.n_bins), -1, np.int32),
                "dist": np.zeros((topo.n_nodes, self.n_bins), np.float64),
            }
            # keep at most two alignments' worth of device memory per Defn (Calculator
            # keeps a 1-deep undo, calculation.py:409-439)
            while len(self._engines) >= 2:
                _, old = self._engines.popitem()
                old["engine"].close()
            self._engines[id(lht)] = st
            return st

        def _slot_for(self, st, ex):
            slot = st["slot_of"].get(id(ex))
            if slot is not None and st["slot_obj"].get(slot) is ex:
                return slot
            used = set(st["slot_obj"])
            slot = next(i for i in range(len(used) + 1) if i not in used)
            eng = st["engine"]
            if isinstance(ex, EigenExponentiator):
                eng.set_eigen(slot, ex.roots, ex.evT, ex.evI)
            elif isinstance(ex, PadeExponentiator):
                eng.set_Q(slot, np.asarray(ex.Q, float))
            else:
                return None
            st["slot_of"][id(ex)] = slot
            st["slot_obj"][slot] = ex
            return slot

        def calc(self, lht, root, mprobs, *rest):
            K = self.n_bins
            if K > 1:
                bprobs, rest = rest[0], rest[1:]
            else:
                bprobs = None
            n_int = len(self.internal_names)
            fixed, psubs = rest[:n_int], rest[n_int:]
            st = self._state_for(lht)
            eng = st["engine"]
            live = set()
            qslots, dist = st["qslots"], st["dist"]
            k = 0
            for node in st["edge_ids"]:
                for b in range(K):
                    p = psubs[k]
                    k += 1
                    slot = None
                    if isinstance(p, LazyPsub):
                        slot = self._slot_for(st, p.exponentiator)
                        if slot is None:
                            p = p.materialise()
                    if slot is not None:
                        live.add(slot)
                        if st["psub_obj"].pop((node, b), None) is not None:
                            qslots[node, b] = -1  # was a supplied P: re-enable expm below
                        if qslots[node, b] != slot:
                            qslots[node, b] = slot
                            st["qslots_dirty"] = True
                        dist[node, b] = p.t
                    else:
                        # P supplied directly (discrete-time edges, solved models, ...)
                        qslots[node, b] = -1
                        if st["psub_obj"].get((node, b)) is not p:
                            eng.set_psub(node, b, np.asarray(p, float))
                            st["psub_obj"][(node, b)] = p
            if st.pop("qslots_dirty", False):
                eng.set_edge_qslots(qslots)  # negative entries are left alone by the engine
            # release the slots no psub refers to any more
            for slot in [s for s in st["slot_obj"] if s not in live]:
                ex = st["slot_obj"].pop(slot)
                st["slot_of"].pop(id(ex), None)
            eng.set_distances(dist)
            eng.set_root_mprobs(np.asarray(mprobs, float))
            if bprobs is not None:
                eng.set_bin_probs(np.asarray(bprobs, float))
            fm = np.full(st["topo"].n_nodes, -1, np.int32)
            for node, f in zip(st["internal_ids"], fixed, strict=True):
                fm[node] = -1 if f in (None, -1) else int(f)
            eng.set_fixed_motifs(fm)
            return eng.evaluate()

        def engine_for(self, lht):
            st = self._engines.get(id(lht))
            return None if st is None or st["lht"] is not lht else st["engine"]

    def make_total_loglikelihood_defn(tree, leaves, psubs, mprobs, bprobs, bin_names, locus_names,
                                      sites_independent, engine_factory, device):  # fmt: skip
        """same signature and role as the reference function of this name
        (evolve/likelihood_calculation.py:125-134), plus the engine selection"""
        if not sites_independent:
            raise NotImplementedError(
                "the phylo-HMM (sites_independent=False) is out of scope (SURVEY §8a)"
            )
        fixed_motifs = NonParamDefn("fixed_motif", ["edge"])
        lht = likelihood_calculation.LikelihoodTreeDefn(leaves, tree=tree)
        root = likelihood_calculation.LhtEdgeLookupDefn(lht, edge_name="root")
        edges = [e for e in tree.postorder() if e.name != "root"]
        internal = [e.name for e in tree.postorder() if not e.istip()]
        edge_names = [e.name for e in edges]
        args = [lht, root, mprobs.select_from_dimension("edge", "root")]
        multi_bin = len(bin_names) > 1
        if multi_bin:
            args.append(bprobs)
        args.extend(fixed_motifs.select_from_dimension("edge", name) for name in internal)
        for name in edge_names:
            p_e = psubs.select_from_dimension("edge", name)
            if multi_bin:
                args.extend(p_e.across_dimension("bin", bin_names))
            else:
                args.append(p_e.select_from_dimension("bin", bin_names[0]))
        tll = GpuLogLikelihoodDefn(
            *args, edge_names=edge_names, internal_names=internal, n_bins=len(bin_names),
            engine_factory=engine_factory, device=device,
        )  # fmt: skip
        if len(locus_names) > 1:
            tll = SumDefn(*tll.across_dimension("locus", locus_names))
        else:
            tll = tll.select_from_dimension("locus", locus_names[0])
        return tll

    class CudaAlignmentLikelihoodFunction(parameter_controller.AlignmentLikelihoodFunction):
        """``AlignmentLikelihoodFunction`` (evolve/parameter_controller.py:511-570) whose
        likelihood sub-graph runs on the GPU.  Extra constructor keywords:
        ``engine_factory`` (test seam) and ``device``."""

        def make_likelihood_defn(self, sites_independent=True, discrete_edges=None,
                                 engine_factory=None, device=0):  # fmt: skip
            # constructor kwargs land in _serialisable (parameter_controller.py:62-66);
            # a callable cannot be serialised, so it is dropped from there
            self._serialisable.pop("engine_factory", None)
            defns = self.model.make_param_controller_defns(bin_names=self.bin_names)
            psubs = defns["psubs"]
            if discrete_edges is not None:
                from cogent3.evolve.discrete_markov import PartialyDiscretePsubsDefn

                psubs = PartialyDiscretePsubsDefn(self.motifs, psubs, discrete_edges)
            elif type(psubs) is definition.CallDefn and len(psubs.args) == 2:
                # psubs = CallDefn(Qd, distance): keep (exponentiator, distance) symbolic.
                # The CallDefn the model built is detached from its inputs so that the
                # controller does not see a value that is never used (scope.py:148-162).
                original = psubs
                psubs = LazyPsubDefn(*original.args)
                for arg in original.args:
                    if original in arg.clients:
                        arg.clients.remove(original)
            self._gpu_defn_factory = None
            return make_total_loglikelihood_defn(
                self.tree, defns["align"], psubs, defns["word_probs"], defns["bprobs"],
                self.bin_names, self.locus_names, sites_independent,
                engine_factory or _default_engine_factory, device,
            )  # fmt: skip

        # -- values the rest of cogent3 looks up by name -------------------
        def _gpu_defn(self):
            for d in self.defns:
                if isinstance(d, GpuLogLikelihoodDefn):
                    return d
            raise KeyError("gpu_lnL")

        def _engine(self, locus=None):
            kw = {} if locus is None else {"locus": locus}
            lht = super().get_param_value("lht", **kw)
            self.get_log_likelihood()  # make sure the engine state is current
            eng = self._gpu_defn().engine_for(lht)
            if eng is None:
                raise RuntimeError("no engine for this locus: set_alignment first")
            return eng

        def get_param_value(self, par_name, *args, **kw):
            if par_name == "lh":
                # per-pattern root likelihoods of one bin, straight from the device
                bin_ = kw.get("bin")
                b = 0 if bin_ is None else list(self.bin_names).index(bin_)
                return self._engine(kw.get("locus")).get_lh(b)
            if par_name in ("bindex", "bdist"):
                locus = kw.get("locus")
                lkw = {} if locus is None else {"locus": locus}
                dist = likelihood_calculation.BinnedSiteDistribution(
                    np.asarray(super().get_param_value("bprobs", **lkw))
                )
                return dist if par_name == "bdist" else dist(super().get_param_value("root", **lkw))
            value = super().get_param_value(par_name, *args, **kw)
            if isinstance(value, LazyPsub):
                value = value.materialise()
            return value

        def get_all_psubs(self):
            defn = self.defn_for["psubs"]
            saved = list(defn.values)
            try:
                defn.values = [v.materialise() if isinstance(v, LazyPsub) else v for v in saved]
                return super().get_all_psubs()
            finally:
                defn.values = saved

    return CudaAlignmentLikelihoodFunction, GpuLogLikelihoodDefn, LazyPsubDefn


_CLASSES = None


def _classes():
    global _CLASSES
    if _CLASSES is None:
        from cogent3.evolve import likelihood_calculation, parameter_controller
        from cogent3.maths import matrix_exponentiation
        from cogent3.recalculation import definition

        _CLASSES = _build(
            (likelihood_calculation, parameter_controller, definition, matrix_exponentiation)
        )
    return _CLASSES


def __getattr__(name):
    if name == "CudaAlignmentLikelihoodFunction":
        return _classes()[0]
    raise AttributeError(name)


def make_likelihood_function(model, tree, motif_probs_from_align=None, optimise_motif_probs=None,
                             expm=None, digits=None, space=None, default_length=1.0, warn=False,
                             **kw):  # fmt: skip
    """``model.make_likelihood_function(tree, ...)`` (evolve/substitution_model.py:313-391)
    with the CUDA likelihood function class in place of ``AlignmentLikelihoodFunction``;
    same arguments, same defaults, same post-construction steps."""
    klass = _classes()[0]
    if motif_probs_from_align is None:
        motif_probs_from_align = model.motif_probs_from_align
    if optimise_motif_probs is None:
        optimise_motif_probs = model._optimise_motif_probs
    kw["optimise_motif_probs"] = optimise_motif_probs
    kw["motif_probs_from_align"] = motif_probs_from_align
    result = klass(model, tree, default_length=default_length, **kw)
    if model.motif_probs is not None:
        result.set_motif_probs(model.motif_probs, is_constant=not optimise_motif_probs, warn=warn)
    if expm is None:
        expm = model._default_expm_setting
    if expm is not None:
        result.set_expm(expm)
    if digits or space:
        result.set_tables_format(digits=digits, space=space)
    return result
