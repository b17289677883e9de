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

from collections import OrderedDict

import numpy as np

from . import batch as _batch
from .patterns import from_reference_lht

__all__ = ["CudaAlignmentLikelihoodFunction", "make_likelihood_function", "install", "uninstall", "installed"]


def _default_engine_factory(patterns, n_bins, device=0, stream=None):
    from .engine import LikelihoodEngine  # no CPU fallback: raises without libc3cuda / a GPU

    return LikelihoodEngine(patterns, n_bins=n_bins, device=device, stream=stream)


class SolvedExponentiator:
    """the "exponentiator" of a model built with ``rate_matrix_required=False`` (F81 / HKY85 / TN93,
    ``PredefinedNucleotide``, evolve/solved_models.py:17-58): there is no Q and no eigendecomposition, P(t) has a
    closed form in (pi, kappa_y, kappa_r).  Held by ``LazyPsub`` like the eigen / Pade objects, so that the engine
    computes every P on the device (c3_set_tn93) -- and materialises to exactly the reference's matrix
    (``model.calc_psub_matrix``) when cogent3 reads a psub."""

    __slots__ = ("model", "pi", "kappas")

    def __init__(self, model, pi, kappas):
        self.model, self.pi, self.kappas = model, pi, tuple(kappas)

    def __call__(self, t):
        return self.model.calc_psub_matrix(self.pi, t, *self.kappas)

    def alphas(self):
        kappa_y = self.kappas[0] if len(self.kappas) > 0 else 1.0
        kappa_r = self.kappas[1] if len(self.kappas) > 1 else kappa_y
        return float(kappa_y), float(kappa_r)


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


# ---------------------------------------------------------------------------------------
# a5: the gamma rate classes.  ``GammaDefn.calc`` (recalculation/definition.py:236-244) asks
# ``scipy.stats.gamma.ppf`` for one median per bin; each call walks scipy's generic distribution
# wrapper (argument broadcasting, support checks, ``place``) around ONE ``gammaincinv``: ~80 us per
# bin, 0.32 ms of a 0.86 ms full-change step at C2 (profiles/r2_dropin_host_profile_c2.txt) -- as long
# as the whole device sweep.  ``rv_continuous.ppf`` computes ``_ppf(q, a) * scale + loc`` with
# ``gamma._ppf = scipy.special.gammaincinv(a, q)``, ``scale = 1 / a``, ``loc = 0``: the same three
# floating-point operations are done here on the whole vector of percentiles.  Bit-identical
# (tests/test_dropin.py::test_fast_gamma_rates_bit_identical); anything outside the plain case
# (a percentile not strictly inside (0, 1), a non-positive or non-finite shape) goes to the reference.
# ---------------------------------------------------------------------------------------
def fast_gamma_rates(weights, a):
    from scipy.special import gammaincinv

    given = weights
    weights = weights / np.sum(weights)
    percentiles = np.add.accumulate(weights) - weights * 0.5
    shape = np.asarray(a)
    if (shape.ndim != 0 or not np.isfinite(shape) or not shape > 0
            or not ((percentiles > 0) & (percentiles < 1)).all()):  # fmt: skip
        from cogent3.recalculation.definition import GammaDefn

        return GammaDefn.calc(None, given, a)  # (the weights as they came: the reference normalises them itself)
    medians = gammaincinv(shape, percentiles) * np.asarray(1 / a) + 0
    scale = np.sum(medians * weights)
    return medians / scale


def _use_fast_gamma(defns, definition):
    """every ``GammaDefn`` reachable from the model's Defns computes its rates with ``fast_gamma_rates``
    (an instance attribute: ``make_calc_function`` returns ``self.calc``, recalculation/definition.py:130-131)"""
    seen, stack = set(), [d for d in defns.values() if d is not None]
    while stack:
        d = stack.pop()
        if id(d) in seen:
            continue
        seen.add(id(d))
        if type(d) is definition.GammaDefn and "calc" not in d.__dict__:
            d.calc = fast_gamma_rates
        stack.extend(getattr(d, "args", ()))

# ---------------------------------------------------------------------------------------
# cogent3's per-node pattern compression, vectorised
#
# ``lf.set_alignment`` spends its time in ``_indexed`` (evolve/likelihood_tree.py:186-203): a Python
# dict loop over every column of every node -- ~7 s per 64 taxa x 100k columns, and it runs three
# times per set_alignment + make_calculator.  The two functions below build the SAME
# objects (real ``LikelihoodTreeLeaf`` / ``LikelihoodTreeEdge`` instances with identical ``uniq``,
# ``index``, ``counts``, ``indexes``, ``ambig``, ``shape``) from numpy's first-occurrence ranking
# (``patterns.first_seen_unique``, checked bit-exact against the dict loop in
# tests/test_host_logic.py and against the live reference here in tests/test_dropin.py).
# ---------------------------------------------------------------------------------------
def fast_likelihood_tree_leaf(likelihood_tree, sequence, alphabet, seq_name):
    """``make_likelihood_tree_leaf`` (evolve/likelihood_tree.py:223-278) without the per-column loop"""
    from .patterns import first_seen_unique

    motif_len = alphabet.motif_len
    try:
        seq = str(sequence)
        raw = np.frombuffer(seq.encode("ascii"), np.uint8)
        n = len(raw) // motif_len
        if motif_len > 7 or n == 0:
            raise ValueError("use the reference loop")
        words = raw[: n * motif_len].reshape(n, motif_len).astype(np.int64)
        key = words[:, 0].copy()
        for k in range(1, motif_len):
            key = key * 256 + words[:, k]
        first, counts, index = first_seen_unique(key)
        uniq_motifs = [seq[int(c) * motif_len : (int(c) + 1) * motif_len] for c in first]
        uniq_motifs.append("?" * motif_len)  # extra column for gap
        counts = np.append(counts, 0.0).astype(float)
        moltype = getattr(sequence, "moltype", None) or alphabet.moltype
        likelihoods = likelihood_tree.get_matched_array(alphabet, moltype, uniq_motifs, dtype=float)
        return likelihood_tree.LikelihoodTreeLeaf(
            uniq_motifs, likelihoods, counts, index.astype(int), seq_name, alphabet, sequence
        )
    except Exception:  # noqa: BLE001 - anything unusual: the reference's own code and its error messages
        return likelihood_tree.make_likelihood_tree_leaf(sequence, alphabet, seq_name)


# ``lf.set_alignment`` needs every tip's table twice: for the motif counts (set_motif_probs_from_data) and for
# the ``leaf_likelihoods`` Defn.  Same alignment, same sequence, same alphabet -> the same (immutable) leaf.
_LEAF_CACHE: dict = {}
_LEAF_CACHE_LOCK = __import__("threading").Lock()


def cached_likelihood_tree_leaf(likelihood_tree, align, seq_name, recode_gaps, alphabet):
    import weakref

    key = (id(align), seq_name, bool(recode_gaps))
    hit = _LEAF_CACHE.get(key)
    if hit is not None:
        ref, alpha, leaf = hit
        if ref() is align and (alpha is alphabet or alpha == alphabet):
            return leaf
    leaf = fast_likelihood_tree_leaf(likelihood_tree, align.get_gapped_seq(seq_name, recode_gaps), alphabet, seq_name)
    with _LEAF_CACHE_LOCK:  # (leaves of a large alignment are built by several threads: build_leaves)
        if len(_LEAF_CACHE) > 2048:
            for k in [k for k, (ref, _a, _l) in list(_LEAF_CACHE.items()) if ref() is None]:
                _LEAF_CACHE.pop(k, None)
            if len(_LEAF_CACHE) > 2048:
                _LEAF_CACHE.clear()
        try:
            _LEAF_CACHE[key] = (weakref.ref(align), alphabet, leaf)
        except TypeError:  # an alignment type without weak references: no caching
            pass
    return leaf


def fast_likelihood_tree_edge(likelihood_tree, children, edge_name):
    """``LikelihoodTreeEdge(children, edge_name)`` (evolve/likelihood_tree.py:13-70, pre-aligned
    children) without the per-column zip + dict loop"""
    from .patterns import _combine_keys, first_seen_unique

    e = likelihood_tree.LikelihoodTreeEdge.__new__(likelihood_tree.LikelihoodTreeEdge)
    e.edge_name = edge_name
    e.alphabet = children[0].alphabet
    M = children[0].shape[-1]
    for child in children:
        assert child.shape[-1] == M
    assignments = [np.asarray(c.index, np.int64) for c in children]
    sizes = [len(c.uniq) for c in children]
    first, counts, index = first_seen_unique(_combine_keys(assignments, sizes))
    uniq = np.stack([a[first] for a in assignments], axis=1).reshape(len(first), len(children))
    gap = np.array([[s - 1 for s in sizes]], np.int64)  # extra column for gap
    e.uniq = np.concatenate([uniq, gap]).astype(int)
    e.index = index.astype(int)
    e.indexes = np.ascontiguousarray(e.uniq.T)
    e.counts = np.append(counts, 0.0).astype(float)
    e._indexed_children = list(zip(e.indexes, children, strict=False))
    e.shape = [len(e.uniq), M]
    ambigs = [child.ambig[ix] for (ix, child) in e._indexed_children]
    e.ambig = np.prod(ambigs, axis=0)
    return e


def build_leaves(likelihood_tree, alignment, names, recode_gaps, alphabet):
    """{name: LikelihoodTreeLeaf} for the sequences of one alignment (cached per alignment object); the leaves are
    independent of each other: from 10^5 columns on they are built in a thread pool"""
    names = list(names)
    workers = _table_workers(len(alignment)) if names else 1

    def leaf(name):
        return cached_likelihood_tree_leaf(likelihood_tree, alignment, name, recode_gaps, alphabet)

    if workers > 1:
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=workers) as pool:
            return dict(zip(names, pool.map(leaf, names), strict=True))
    return {name: leaf(name) for name in names}


def _table_workers(n_columns):
    """threads for building the pattern tables of one alignment: 1 below 10^5 columns (the pool would cost more than
    it saves), else up to 8 (C3_TABLE_THREADS overrides)"""
    import os

    env = os.environ.get("C3_TABLE_THREADS")
    if env is not None:
        try:
            return max(1, int(env))
        except ValueError:
            return 1
    if n_columns < 100_000:
        return 1
    return max(1, min(8, os.cpu_count() or 1))


# identity-keyed memo of a Defn's last few results (one per locus + undo); strong references, so
# ids cannot be recycled.  Never pickled.
_MEMO_MAX = 2


def _memo_get(defn, key):
    for k, v in defn.__dict__.get("_memo", ()):
        if len(k) == len(key) and all(a is b for a, b in zip(k, key, strict=True)):
            return v
    return None


def _memo_put(defn, key, value):
    memo = defn.__dict__.setdefault("_memo", [])
    memo.append((key, value))
    cap = max(_MEMO_MAX, defn.__dict__.get("_memo_cap", 0))
    del memo[: max(0, len(memo) - cap)]
    return value


def _memo_free_state(self):
    d = dict(self.__dict__)
    d.pop("_memo", None)
    return d


def _build(cogent3_modules):
    """create the Defn / LF classes lazily so that importing this module does not
    require cogent3"""
    (likelihood_calculation, parameter_controller, definition, matrix_exponentiation) = cogent3_modules
    from cogent3.evolve import likelihood_tree, substitution_calculation

    class FastAlignmentAdaptDefn(substitution_calculation.AlignmentAdaptDefn):
        """``leaf_likelihoods``: ``model.convert_alignment`` (evolve/substitution_model.py:393-404) with
        the vectorised leaf builder"""

        def calc(self, model, alignment):
            # cogent3 evaluates this cell up to three times for one alignment (set_alignment, the
            # eager Defn update, Calculator construction: SURVEY §3.2); the tables depend only on
            # (model, alignment), both immutable here, so the same objects are handed out again --
            # which also keeps the identity of ``lht`` and with it the engine and its device tables
            hit = _memo_get(self, (model, alignment))
            if hit is not None:
                return hit
            if type(model).convert_sequence is not _reference_convert_sequence(model):
                return model.convert_alignment(alignment)  # a model with its own conversion
            alphabet = model.get_alphabet()
            leaves = build_leaves(likelihood_tree, alignment, alignment.names, model.recode_gaps, alphabet)
            return _memo_put(self, (model, alignment), leaves)

        __getstate__ = _memo_free_state

    def _is_solved_psubs(psubs):
        try:
            from cogent3.evolve.solved_models import PredefinedNucleotide
        except Exception:  # noqa: BLE001
            return False
        calc = getattr(psubs, "calc", None)
        return getattr(calc, "__func__", None) is PredefinedNucleotide.calc_psub_matrix

    def _reference_convert_sequence(model):
        from cogent3.evolve.substitution_model import _SubstitutionModel

        return _SubstitutionModel.convert_sequence

    class FastLikelihoodTreeDefn(likelihood_calculation.LikelihoodTreeDefn):
        """``lht``: ``recursive_lht_build`` (evolve/likelihood_calculation.py:103-122) with the
        vectorised edge builder"""

        def calc(self, leaves):
            hit = _memo_get(self, (leaves,))
            if hit is not None:
                return hit

            def build(edge):
                if edge.istip():
                    return leaves[edge.name]
                kids = [build(child) for child in edge.children]
                return fast_likelihood_tree_edge(likelihood_tree, kids, edge.name)

            n_columns = max((len(leaf.index) for leaf in leaves.values()), default=0)
            workers = _table_workers(n_columns)
            if workers <= 1:
                return _memo_put(self, (leaves,), build(self.tree))
            # large alignments: the nodes of one tree level are independent, and the ranking of a node's columns
            # (numpy scatter / pandas hash table over 10^5..10^7 keys) runs without the GIL -- level by level in a
            # thread pool.  Same function per node, same tables.
            height, order = {}, list(self.tree.postorder())
            for edge in order:
                height[id(edge)] = 0 if edge.istip() else 1 + max(height[id(c)] for c in edge.children)
            done = {id(e): leaves[e.name] for e in order if e.istip()}
            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(max_workers=workers) as pool:
                for h in range(1, height[id(self.tree)] + 1):
                    level = [e for e in order if height[id(e)] == h]
                    made = pool.map(
                        lambda e: fast_likelihood_tree_edge(likelihood_tree, [done[id(c)] for c in e.children], e.name),
                        level,
                    )
                    for e, lht_edge in zip(level, made, strict=True):
                        done[id(e)] = lht_edge
            return _memo_put(self, (leaves,), done[id(self.tree)])

        __getstate__ = _memo_free_state

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

    class SolvedLazyPsubDefn(CalculationDefn):
        """``psubs = CalcDefn(model.calc_psub_matrix)(word_probs, distance, *rate_params)``
        (evolve/solved_models.py:22-41) without the call: one shared ``SolvedExponentiator`` per distinct
        (pi, kappas), so all (edge, bin)s of a scope exponentiate from ONE device slot"""

        name = "psubs"

        def setup(self, model):
            self.model = model
            self._made = []  # [(pi object, kappas, exponentiator)], identity / value keyed, a handful of entries

        def calc(self, pi, time, *kappas):
            for p, k, ex in self._made:
                if p is pi and k == kappas:
                    return LazyPsub(ex, time)
            ex = SolvedExponentiator(self.model, pi, kappas)
            self._made.append((pi, kappas, ex))
            del self._made[:-8]
            return LazyPsub(ex, time)

        def __getstate__(self):
            d = dict(self.__dict__)
            d["_made"] = []
            return d

    class GpuLogLikelihoodDefn(CalculationDefn):
        """total log-likelihood of one locus, computed by libc3cuda.

        args: lht, root (LhtEdgeLookupDefn), root mprobs, [bprobs], then one
        fixed_motif per internal node (post-order), then one psub per (edge, bin)
        in edge-major order (post-order edges)."""

        name = "gpu_lnL"

        def setup(self, edge_names, internal_names, n_bins, engine_factory, device, n_loci=1, devices=None,
                  return_lh=False):
            self.edge_names = list(edge_names)
            self.internal_names = list(internal_names)
            self.n_bins = n_bins
            self.engine_factory = engine_factory
            self.device = device
            self.devices = list(devices) if devices else None
            # phylo-HMM (sites_independent=False): the LAST argument is the reference's ``SiteHmm`` (the value
            # of its "bindex" Defn); the per-bin root likelihoods the engine has just computed are reduced
            # over the SITES by that model instead of being mixed per pattern.  Everything the result
            # depends on is an argument of this ONE cell, so whatever the Calculator presents (undo
            # included) the engine re-derives its state from the inputs in front of it
            self.return_lh = bool(return_lh)
            # id(lht) -> state, least recently used first.  The Defn instance (and so this cache) is
            # shared by every locus cell behind SumDefn(*tll.across_dimension("locus", ...)): it must
            # hold one engine per locus plus one for the Calculator's 1-deep undo
            # (calculation.py:409-439), or every evaluation of a 3-locus function rebuilds engines
            self._engines = OrderedDict()
            self.max_engines = max(2, int(n_loci) + 1)
            self.engines_created = 0

        # engines hold device memory and ctypes handles: never pickled / deep-copied
        def __getstate__(self):
            d = dict(self.__dict__)
            d["_engines"] = OrderedDict()
            return d

        def __deepcopy__(self, memo):
            import copy

            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            for k, v in self.__dict__.items():
                new.__dict__[k] = OrderedDict() if k == "_engines" else copy.deepcopy(v, memo)
            return new

        def _state_for(self, lht):
            st = self._engines.get(id(lht))
            if st is not None and st["lht"] is lht:
                self._engines.move_to_end(id(lht))
                return st
            # a new alignment (or locus): flatten cogent3's own pattern tables -- they stay
            # the source of truth -- and make the tables resident on the device
            topo, patterns = from_reference_lht(lht)
            if self.devices and len(self.devices) > 1:
                # site patterns sharded over several GPUs of this process (cogent3_b200/multidevice.py)
                from .multidevice import MultiDeviceEngine

                engine = MultiDeviceEngine(patterns, self.n_bins, devices=self.devices, engine_cls=self.engine_factory)
            else:
                device = self.devices[0] if self.devices else self.device
                engine = self.engine_factory(patterns, self.n_bins, device=device)
            node_of = {name: i for i, name in enumerate(topo.names)}
            st = {
                "lht": lht,
                "engine": engine,
                "topo": topo,
                "edge_ids": [node_of[n] for n in self.edge_names],
                "edge_ids_arr": np.array([node_of[n] for n in self.edge_names], np.intp),
                "internal_ids": [node_of[n] for n in self.internal_names],
                "slot_of": {},  # id(exponentiator) -> slot
                "slot_obj": {},  # slot -> exponentiator (strong ref: ids are not recycled)
                "psub_obj": {},  # (node, bin) -> last directly supplied P object
                "qslots": np.full((topo.n_nodes, self.n_bins), -1, np.int32),
                "dist": np.zeros((topo.n_nodes, self.n_bins), np.float64),
            }
            # device memory per Defn: one engine per locus + one (the Calculator keeps a 1-deep undo,
            # calculation.py:409-439); the OLDEST entry goes first
            self._engines.pop(id(lht), None)  # (a recycled id: the old lht object is gone)
            while len(self._engines) >= self.max_engines:
                _, old = self._engines.popitem(last=False)
                old["engine"].close()
            self._engines[id(lht)] = st
            self.engines_created += 1
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
            elif isinstance(ex, SolvedExponentiator) and hasattr(eng, "set_tn93"):
                eng.set_tn93(slot, np.asarray(ex.pi, float), *ex.alphas())
            else:
                return None
            st["slot_of"][id(ex)] = slot
            st["slot_obj"][slot] = ex
            return slot

        def calc(self, lht, root, mprobs, *rest):
            K = self.n_bins
            blh = None
            if self.return_lh:
                blh, rest = rest[-1], rest[:-1]
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
            return self._finish(eng, blh)

        def _finish(self, eng, blh=None, dist=None):
            lockstep = _batch.current()
            # inside cogent3_b200.batch.run_lockstep the evaluations of several optimiser threads meet in one
            # c3_evaluate_many call; otherwise the engine is evaluated right here -- distances
            # and evaluation in ONE FFI crossing (c3_update) when the caller has not handed them over yet
            if lockstep is None:
                value = eng.evaluate() if dist is None else eng.update(dist)
            else:
                if dist is not None:
                    eng.set_distances(dist)
                value = lockstep.evaluate(eng)
            if self.return_lh:
                # SiteHmm.__call__ (evolve/likelihood_calculation.py:265-269): its reduction over the sites --
                # log_dot_reduce, a Python loop over every column (evolve/likelihood_tree.py:168-176) -- runs on
                # the device when the engine offers it (c3_hmm_reduce: a parallel product of the per-site 2 x 2
                # matrices over the likelihoods that are already there); else the reference's own loop
                distrib = getattr(blh, "distrib", None)
                if hasattr(eng, "hmm_reduce") and distrib is not None:
                    tm = distrib.transition_matrix
                    matrix, start = np.asarray(tm.Matrix, float), np.asarray(tm.StationaryProbs, float)
                    if matrix.shape == (2, 2) and start.shape == (2,):
                        return eng.hmm_reduce(distrib.alloc, distrib.bprobs, matrix, start)
                return blh(*[eng.get_lh(b) for b in range(self.n_bins)])
            return value

        def engine_for(self, lht):
            st = self._engines.get(id(lht))
            return None if st is None or st["lht"] is not lht else st["engine"]

    class GpuFactoredLogLikelihoodDefn(GpuLogLikelihoodDefn):
        """The same terminal Defn fed with the FACTORS of the psubs instead of the psubs:

            psubs[edge, bin] = Qd[edge, bin](length[edge] * rate[bin])
                (``CallDefn(Qd, distance)``, ``distance = ProductDefn(length, rate)``,
                evolve/substitution_model.py:684-710)

        args: lht, root, root mprobs, [bprobs], one fixed_motif per internal node, one Qd
        (exponentiator) per (edge, bin), one length per edge, [one rate per bin].  The K*E
        ``distance`` and K*E ``psubs`` cells disappear from the Calculator's sweep (each costs
        cogent3 ~1.5 us of Python per evaluation: 1.5 ms of a 2.1 ms full evaluation at 64 taxa x 4
        bins, and 2 K cells plus a K*E-long Python loop here for every single-branch update); the
        product is formed here with one numpy outer product -- the same IEEE multiplication
        ``ProductDefn.calc`` does (``numpy.prod((length, rate))``,
        recalculation/definition.py:679-683)."""

        name = "gpu_lnL"

        def setup(self, has_rate, **kw):
            super().setup(**kw)
            self.has_rate = has_rate

        def calc(self, lht, root, mprobs, *rest):
            K, E = self.n_bins, len(self.edge_names)
            blh = None
            if self.return_lh:
                blh, rest = rest[-1], rest[:-1]
            if K > 1:
                bprobs, rest = rest[0], rest[1:]
            else:
                bprobs = None
            n_int = len(self.internal_names)
            fixed, rest = rest[:n_int], rest[n_int:]
            qds, lengths, rates = rest[: E * K], rest[E * K : E * K + E], rest[E * K + E :]
            st = self._state_for(lht)
            eng = st["engine"]
            qslots, dist = st["qslots"], st["dist"]
            edge_ids = st["edge_ids"]
            last = st.get("last_qds")
            try:
                # tuple comparison short-cuts on identity in C; exponentiators compare by identity
                same_qds = last is not None and bool(qds == last)
            except Exception:  # noqa: BLE001 - an exotic exponentiator with an array-valued __eq__
                same_qds = last is not None and all(a is b for a, b in zip(qds, last, strict=True))
            if not same_qds:
                # (re)assign exponentiators to q slots: whenever a rate-matrix parameter moved.  The E*K
                # entries name only a few DISTINCT objects (one for a time-homogeneous model, one per edge
                # with per-edge Q), so the per-entry work is numpy's: ids -> unique -> slot of each
                live, host = set(), {}
                try:
                    homogeneous = qds.count(qds[0]) == len(qds)
                except Exception:  # noqa: BLE001 - an exotic exponentiator with an array-valued __eq__
                    homogeneous = False
                if homogeneous:
                    # time-homogeneous model: ONE exponentiator for every (edge, bin) -- tuple.count compares by
                    # identity first, a C loop; the general path below costs ~25 us of numpy per evaluation
                    slot = self._slot_for(st, qds[0])
                    if slot is not None:
                        live.add(slot)
                    per_entry = np.full((E, K), -1 if slot is None else slot, np.int32)
                else:
                    ids = np.fromiter(map(id, qds), dtype=np.int64, count=len(qds))
                    _uniq, first, inverse = np.unique(ids, return_index=True, return_inverse=True)
                    slot_of_uniq = np.empty(len(first), np.int32)
                    for u, f in enumerate(first):
                        slot = self._slot_for(st, qds[int(f)])
                        slot_of_uniq[u] = -1 if slot is None else slot
                        if slot is not None:
                            live.add(slot)
                    per_entry = slot_of_uniq[inverse.reshape(-1)].reshape(E, K)
                qslots[edge_ids, :] = per_entry
                if (per_entry < 0).any():
                    # not an eigen / Pade exponentiator: P on the host for those (edge, bin)s
                    for e_i, b in zip(*np.nonzero(per_entry < 0), strict=True):
                        host[(edge_ids[int(e_i)], int(b))] = qds[int(e_i) * K + int(b)]
                if st["psub_obj"]:
                    for key in [k for k in st["psub_obj"] if k not in host]:
                        st["psub_obj"].pop(key, None)
                eng.set_edge_qslots(qslots)  # negative entries are left alone by the engine
                for slot in [s for s in st["slot_obj"] if s not in live]:
                    ex = st["slot_obj"].pop(slot)
                    st["slot_of"].pop(id(ex), None)
                st["last_qds"] = qds
                st["host_ex"] = host
            d = np.array(lengths, np.float64)[:, None]
            if self.has_rate:
                d = d * np.array(rates, np.float64)[None, :]
            dist[st["edge_ids_arr"], :] = d
            for (node, b), ex in st["host_ex"].items():
                key = (ex, float(dist[node, b]))
                prev = st["psub_obj"].get((node, b))
                if prev is None or prev[0] is not key[0] or prev[1] != key[1]:
                    eng.set_psub(node, b, np.asarray(ex(key[1]), float))
                    st["psub_obj"][(node, b)] = key
            # unchanged inputs keep object identity between calls (calculation.py:431-438)
            if st.get("last_mprobs") is not mprobs:
                eng.set_root_mprobs(np.asarray(mprobs, float))
                st["last_mprobs"] = mprobs
            if bprobs is not None and st.get("last_bprobs") is not bprobs:
                eng.set_bin_probs(np.asarray(bprobs, float))
                st["last_bprobs"] = bprobs
            if st.get("last_fixed") != fixed:
                fm = np.full(st["topo"].n_nodes, -1, np.int32)
                for node, f in zip(st["internal_ids"], fixed, strict=True):
                    fm[node] = -1 if f in (None, -1) else int(f)
                eng.set_fixed_motifs(fm)
                st["last_fixed"] = fixed
            return self._finish(eng, blh, dist=dist)

    def make_total_loglikelihood_defn(tree, leaves, psubs, mprobs, bprobs, bin_names, locus_names,
                                      sites_independent, engine_factory, device, factors=None, devices=None):  # fmt: skip
        """same signature and role as the reference function of this name
        (evolve/likelihood_calculation.py:125-134), plus the engine selection"""
        # phylo-HMM: with several bins and sites_independent=False the per-bin root likelihoods
        # are reduced over the SITES by a hidden Markov model (PatchSiteDistribution / SiteHmm,
        # evolve/likelihood_calculation.py:197-235,260-296) instead of being mixed per pattern.  The pruning --
        # all of the O(patterns x edges) work -- runs on the device as usual; the reduction is the reference's own
        # code on the K x patterns array the engine hands back.
        hmm = (not sites_independent) and len(bin_names) > 1
        fixed_motifs = NonParamDefn("fixed_motif", ["edge"])
        lht = FastLikelihoodTreeDefn(leaves, tree=tree)
        leaves.__dict__["_memo_cap"] = lht.__dict__["_memo_cap"] = len(locus_names) + 1
        root = likelihood_calculation.LhtEdgeLookupDefn(lht, edge_name="root")
        edges = [e for e in tree.postorder() if e.name != "root"]
        internal = [e.name for e in tree.postorder() if not e.istip()]
        edge_names = [e.name for e in edges]
        args = [lht, root, mprobs.select_from_dimension("edge", "root")]
        multi_bin = len(bin_names) > 1
        tail_args = []
        if hmm:
            from cogent3.recalculation.definition import CalcDefn, CallDefn, ProbabilityParamDefn

            switch = ProbabilityParamDefn("bin_switch", dimensions=["locus"])
            site_pattern = CalcDefn(likelihood_calculation.PatchSiteDistribution, name="bdist")(switch, bprobs)
            tail_args.append(CallDefn(site_pattern, lht, name="bindex"))
        if multi_bin:
            args.append(bprobs)
        args.extend(fixed_motifs.select_from_dimension("edge", name) for name in internal)
        def select(defn, dimension, cat):
            # (a Defn that does not have the dimension is the same for every category)
            return defn.select_from_dimension(dimension, cat) if dimension in defn.valid_dimensions else defn

        if factors is not None:
            Qd, length, rate = factors
            for name in edge_names:
                q_e = select(Qd, "edge", name)
                args.extend(select(q_e, "bin", b) for b in bin_names)
            args.extend(select(length, "edge", name) for name in edge_names)
            if rate is not None:
                args.extend(select(rate, "bin", b) for b in bin_names)
            args.extend(tail_args)
            tll = GpuFactoredLogLikelihoodDefn(
                *args, edge_names=edge_names, internal_names=internal, n_bins=len(bin_names),
                engine_factory=engine_factory, device=device, has_rate=rate is not None,
                n_loci=len(locus_names), devices=devices, return_lh=hmm,
            )  # fmt: skip
        else:
            for name in edge_names:
                p_e = psubs.select_from_dimension("edge", name)
                if multi_bin:
                    args.extend(p_e.across_dimension("bin", bin_names))
                else:
                    args.append(p_e.select_from_dimension("bin", bin_names[0]))
            args.extend(tail_args)
            tll = GpuLogLikelihoodDefn(
                *args, edge_names=edge_names, internal_names=internal, n_bins=len(bin_names),
                engine_factory=engine_factory, device=device, n_loci=len(locus_names), devices=devices,
                return_lh=hmm,
            )  # fmt: skip
        if len(locus_names) > 1:
            tll = SumDefn(*tll.across_dimension("locus", locus_names))
        else:
            tll = tll.select_from_dimension("locus", locus_names[0])
        return tll

    class CudaAlignmentLikelihoodFunction(parameter_controller.AlignmentLikelihoodFunction):
        """``AlignmentLikelihoodFunction`` (evolve/parameter_controller.py:511-570) whose
        likelihood sub-graph runs on the GPU.  Extra constructor keywords: ``device`` (CUDA ordinal),
        ``devices=[0, 1, ...]`` (site patterns sharded over several GPUs of this process),
        ``engine_factory`` (test seam)."""

        def make_likelihood_defn(self, sites_independent=True, discrete_edges=None,
                                 engine_factory=None, device=0, factored=True, devices=None):  # fmt: skip
            # constructor kwargs land in _serialisable (parameter_controller.py:62-66);
            # a callable cannot be serialised, so it is dropped from there
            self._serialisable.pop("engine_factory", None)
            self._serialisable.pop("factored", None)
            factors = None
            defns = self.model.make_param_controller_defns(bin_names=self.bin_names)
            _use_fast_gamma(defns, definition)
            psubs = defns["psubs"]
            if discrete_edges is not None:
                from cogent3.evolve.discrete_markov import PartialyDiscretePsubsDefn

                psubs = PartialyDiscretePsubsDefn(self.motifs, psubs, discrete_edges)
            elif type(psubs) is definition.CallDefn and len(psubs.args) == 2:
                # psubs = CallDefn(Qd, distance).  The CallDefn the model built is detached from its
                # inputs so that the controller does not see a value that is never used
                # (scope.py:148-162).
                original = psubs

                def detach(defn):
                    for arg in defn.args:
                        if defn in arg.clients:
                            arg.clients.remove(defn)

                detach(original)
                Qd, distance = original.args
                if (type(distance) is definition.ProductDefn and len(distance.args) == 2
                        and distance.args[0].name == "length" and not distance.clients
                        and factored):  # fmt: skip
                    # distance = ProductDefn(length, rate): hand the FACTORS to the engine Defn
                    detach(distance)
                    factors = (Qd, distance.args[0], distance.args[1])
                elif distance.name == "length" and factored:
                    factors = (Qd, distance, None)  # no rate heterogeneity: distance is the length
                else:
                    # keep (exponentiator, distance) symbolic, one LazyPsub cell per (edge, bin)
                    psubs = LazyPsubDefn(*original.args)
            elif _is_solved_psubs(psubs):
                # rate_matrix_required=False: P has a closed form, computed on the device (c3_set_tn93)
                original = psubs
                for arg in original.args:
                    if original in arg.clients:
                        arg.clients.remove(original)
                psubs = SolvedLazyPsubDefn(*original.args, model=self.model)
            self._factors = factors
            self._hmm = (not sites_independent) and len(self.bin_names) > 1
            self._gpu_defn_factory = None
            leaves = defns["align"]
            if type(leaves) is substitution_calculation.AlignmentAdaptDefn:
                # same inputs (model, alignment), vectorised conversion; the Defn the model built is
                # detached like the psubs CallDefn above
                original = leaves
                leaves = FastAlignmentAdaptDefn(*original.args)
                for arg in original.args:
                    if original in arg.clients:
                        arg.clients.remove(original)
            return make_total_loglikelihood_defn(
                self.tree, leaves, psubs, defns["word_probs"], defns["bprobs"],
                self.bin_names, self.locus_names, sites_independent,
                engine_factory or _default_engine_factory, device, factors=factors, devices=devices,
            )  # fmt: skip

        def set_motif_probs_from_data(self, align, locus=None, is_constant=None, include_ambiguity=False,
                                      is_independent=None, auto=False, warn=False, pseudocount=None,
                                      **kwargs):  # fmt: skip
            """the reference method (evolve/parameter_controller.py:135-162) with the motif counts
            (``MotifProbModel.count_motifs``, evolve/motif_prob_model.py:48-59: one
            ``make_likelihood_tree_leaf`` dict loop per sequence -- most of what is left of
            ``set_alignment`` once the tables are built with numpy) taken from the vectorised leaves"""
            from cogent3.evolve import motif_prob_model, substitution_model

            mm = self.model.mprob_model
            if (type(self.model).count_motifs is not substitution_model._SubstitutionModel.count_motifs
                    or type(mm).count_motifs is not motif_prob_model.MotifProbModel.count_motifs):  # fmt: skip
                return super().set_motif_probs_from_data(
                    align, locus=locus, is_constant=is_constant, include_ambiguity=include_ambiguity,
                    is_independent=is_independent, auto=auto, warn=warn, pseudocount=pseudocount, **kwargs,
                )  # fmt: skip
            alpha = mm.get_counted_alphabet()
            counts = None
            leaves = build_leaves(likelihood_tree, align, align.names, self.model.recode_gaps, alpha)
            for seq_name in align.names:
                leaf = leaves[seq_name]
                count = leaf.get_motif_counts(include_ambiguity=include_ambiguity)
                counts = count.copy() if counts is None else counts + count
            if is_constant is None:
                is_constant = not self.optimise_motif_probs
            pseudocount = 0.0 if is_constant or (counts != 0).all() else pseudocount or 0.5
            counts += pseudocount
            mprobs = counts / (1.0 * sum(counts))
            self.set_motif_probs(mprobs, locus=locus, is_constant=is_constant, is_independent=is_independent,
                                 auto=auto, warn=warn, **kwargs)  # fmt: skip

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

        def to_rich_dict(self):
            """serialised results stay interchangeable with stock cogent3 (and deserialise without a
            GPU): the recorded type is the reference class (evolve/likelihood_function.py:968-1030,
            util/deserialise.py:118-140); the engine seams are not construction arguments"""
            data = super().to_rich_dict()
            data["type"] = "cogent3.evolve.parameter_controller.AlignmentLikelihoodFunction"
            for key in ("engine_factory", "device", "devices", "factored"):
                data.get("likelihood_construction", {}).pop(key, None)
            return data

        # -- psubs when the engine Defn is fed with their factors: there is no "psubs" Defn in the
        # graph any more, so the three ways cogent3 reads them (get_param_value("psubs", ...),
        # get_psub_for_edge -> get_param_value, get_all_psubs; evolve/likelihood_function.py:272-320)
        # compute Qd(length * rate) from the Defns that are there -- the reference's own arithmetic
        def _factored_psub(self, edge, bin=None, locus=None):  # noqa: A002
            Qd, length, rate = self._factors
            given = {"edge": edge, "bin": bin, "locus": locus}

            def value(defn):
                scope = {d: given[d] for d in defn.valid_dimensions if given.get(d) is not None}
                return super(CudaAlignmentLikelihoodFunction, self).get_param_value(defn.name, **scope)

            t = value(length)
            if rate is not None:
                t = np.prod((t, value(rate)))  # ProductDefn.calc
            return value(Qd)(t)

        def get_all_psubs(self):
            if getattr(self, "_factors", None) is None:
                return self._lazy_get_all_psubs()
            from cogent3.util.dict_array import DictArrayTemplate

            Qd, length, rate = self._factors
            # the key of the reference's dict: the used dimensions of psubs in valid-dimension
            # (alphabetical) order = those used by any of its factors
            used = set()
            for defn in (Qd, length, rate):
                if defn is not None:
                    used.update(d for d in defn.used_dimensions() if d in ("bin", "edge", "locus"))
            used.add("edge")
            dims = [d for d in ("bin", "edge", "locus") if d in used]
            cats = {"bin": list(self.bin_names), "locus": list(self.locus_names),
                    "edge": [e.name for e in self.tree.postorder(include_self=False)]}  # fmt: skip
            template = DictArrayTemplate(self._motifs, self._motifs)
            result = {}
            import itertools

            for combo in itertools.product(*(cats[d] for d in dims)):
                scope = dict(zip(dims, combo, strict=True))
                # a dimension that is not used takes its first category, like the last scope the
                # reference's loop happens to see: the values are the same by definition
                full = {d: scope.get(d, cats[d][0]) for d in ("bin", "edge", "locus")}
                key = combo[0] if len(combo) == 1 else tuple(combo)
                result[key] = template.wrap(self._factored_psub(full["edge"], full["bin"], full["locus"]))
            return result

        def get_param_value(self, par_name, *args, **kw):
            if par_name == "psubs" and getattr(self, "_factors", None) is not None and not args:
                return self._factored_psub(kw.get("edge"), kw.get("bin"), kw.get("locus"))
            if par_name == "lh":
                # per-pattern root likelihoods of one bin, straight from the device
                bin_ = kw.get("bin")
                b = 0 if bin_ is None else list(self.bin_names).index(bin_)
                return self._engine(kw.get("locus")).get_lh(b)
            if par_name in ("bindex", "bdist") and not getattr(self, "_hmm", False):
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

        def _lazy_get_all_psubs(self):
            defn = self.defn_for["psubs"]
            saved = list(defn.values)
            try:
                defn.values = [v.materialise() if isinstance(v, LazyPsub) else v for v in saved]
                return super().get_all_psubs()
            finally:
                defn.values = saved

    # picklable by reference (process-parallel apps, app/evo.py bootstrap with parallel=True, pickle
    # likelihood functions): the classes are reachable as attributes of this module
    exported = (CudaAlignmentLikelihoodFunction, GpuLogLikelihoodDefn, LazyPsubDefn,
                GpuFactoredLogLikelihoodDefn, FastAlignmentAdaptDefn, FastLikelihoodTreeDefn,
                SolvedLazyPsubDefn)  # fmt: skip
    for cls in exported:
        cls.__module__ = __name__
        cls.__qualname__ = cls.__name__
    return exported


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


_CLASS_NAMES = ("CudaAlignmentLikelihoodFunction", "GpuLogLikelihoodDefn", "LazyPsubDefn",
                "GpuFactoredLogLikelihoodDefn", "FastAlignmentAdaptDefn", "FastLikelihoodTreeDefn",
                "SolvedLazyPsubDefn")  # fmt: skip


def __getattr__(name):
    if name in _CLASS_NAMES:
        return _classes()[_CLASS_NAMES.index(name)]
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


# ---------------------------------------------------------------------------------------
# product hook: cogent3's stock entry point reaches the GPU
#
# ``_SubstitutionModel.make_likelihood_function`` picks the likelihood-function class
# (evolve/substitution_model.py:367-374: AlignmentLikelihoodFunction when ``aligned``).  cogent3 has
# no plugin group for that choice, so ``install()`` replaces the method -- the one-line
# change a maintainer would make there is shown in INTEGRATION.md -- and from then on
#
#     import cogent3_b200; cogent3_b200.install()
#     lf = get_model("GTR").make_likelihood_function(tree, bins=4)      # CUDA engine
#     lf = get_model("GTR").make_likelihood_function(tree, backend="cpu")   # the reference class
#
# every stock caller (apps, bootstrap, hypothesis tests, tree search) builds the CUDA class.  What the
# engine does not cover stays with the reference class: unaligned sequences (pair-HMM alignment).  The
# phylo-HMM (``sites_independent=False``) prunes on the device and reduces over the sites with the
# reference's own SiteHmm.
# ---------------------------------------------------------------------------------------
_STOCK_MAKE_LF = None
INSTALL_STATS = {"cuda": 0, "reference": 0}


def installed():
    return _STOCK_MAKE_LF is not None


def install(default_backend="cuda", device=0, devices=None, engine_factory=None):
    """make ``sm.make_likelihood_function(tree, ...)`` return the CUDA likelihood function.

    ``default_backend``: "cuda" or "cpu" -- what a call without ``backend=`` gets; ``device`` /
    ``devices``: the GPU(s) a call without ``device=`` / ``devices=`` uses (``devices=[0, ..., 7]``
    shards the site patterns over several GPUs in one process); ``engine_factory``: test seam.
    Idempotent; ``uninstall()`` restores the reference method."""
    global _STOCK_MAKE_LF
    from cogent3.evolve import substitution_model

    cls = substitution_model._SubstitutionModel
    if _STOCK_MAKE_LF is None:
        _STOCK_MAKE_LF = cls.make_likelihood_function
    stock = _STOCK_MAKE_LF
    if default_backend not in ("cuda", "cpu"):
        raise ValueError(f"unknown backend {default_backend!r}")

    def make_likelihood_function_with_backend(self, tree, motif_probs_from_align=None,
                                              optimise_motif_probs=None, aligned=True, expm=None,
                                              digits=None, space=None, default_length=1.0, warn=False,
                                              backend=None, **kw):  # fmt: skip
        backend = backend or default_backend
        if backend not in ("cuda", "cpu"):
            raise ValueError(f"unknown backend {backend!r}")
        if backend == "cpu" or not aligned:
            INSTALL_STATS["reference"] += 1
            for key in ("device", "devices", "engine_factory", "factored"):
                kw.pop(key, None)
            return stock(self, tree, motif_probs_from_align=motif_probs_from_align,
                         optimise_motif_probs=optimise_motif_probs, aligned=aligned, expm=expm,
                         digits=digits, space=space, default_length=default_length, warn=warn, **kw)  # fmt: skip
        INSTALL_STATS["cuda"] += 1
        if devices is not None and "devices" not in kw and "device" not in kw:
            kw["devices"] = list(devices)
        elif "device" not in kw and "devices" not in kw and device != 0:
            kw["device"] = device
        if engine_factory is not None:
            kw.setdefault("engine_factory", engine_factory)
        return make_likelihood_function(self, tree, motif_probs_from_align=motif_probs_from_align,
                                        optimise_motif_probs=optimise_motif_probs, expm=expm, digits=digits,
                                        space=space, default_length=default_length, warn=warn, **kw)  # fmt: skip

    make_likelihood_function_with_backend.__doc__ = stock.__doc__
    make_likelihood_function_with_backend.__name__ = "make_likelihood_function"
    cls.make_likelihood_function = make_likelihood_function_with_backend
    return make_likelihood_function_with_backend


def uninstall():
    global _STOCK_MAKE_LF
    if _STOCK_MAKE_LF is None:
        return
    from cogent3.evolve import substitution_model

    substitution_model._SubstitutionModel.make_likelihood_function = _STOCK_MAKE_LF
    _STOCK_MAKE_LF = None
