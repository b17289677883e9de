"""numpy restatement of cogent3's likelihood hot path (the ORACLE).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Every function cites the
reference file:line (relative to /root/reference/src/cogent3) whose arithmetic
it restates.  Parity of this module is PINNED: tests/test_oracle_golden.py
checks it against golden vectors generated from the live reference
(tests/golden/make_golden.py, run in the build container) and against the
known-answer constants of the reference's own test-suite.

Conventions (same as the reference):
  * P[i, j] = Pr(i -> j), rows sum to 1        (maths/matrix_exponentiation.py:35-40)
  * a node's pattern table has P+1 rows, the last is the all-gap row with
    count 0                                    (evolve/likelihood_tree.py:49-51,230-232)
  * children of a node are kept in tree order; ``indexes[c, p]`` is the row of
    child c that parent pattern p reads         (evolve/likelihood_tree.py:55-58)
"""

from __future__ import annotations

import numpy as np
from numpy.linalg import eig, inv, solve

# --------------------------------------------------------------------------
# (a2)/(a3) pattern compression -- integer work, must be bit exact
# --------------------------------------------------------------------------


def indexed(values):
    """first-seen-order dedupe: restates ``_indexed``
    (evolve/likelihood_tree.py:186-203) with the same Python dict loop.

    >>> indexed(['a', 'b', 'c', 'a', 'a'])
    (['a', 'b', 'c'], [3, 1, 1], array([0, 1, 2, 0, 0]))
    """
    index = np.zeros([len(values)], int)
    unique, counts, seen = [], [], {}
    for c, key in enumerate(values):
        i = seen.get(key)
        if i is None:
            i = len(unique)
            unique.append(key)
            counts.append(1)
            seen[key] = i
        else:
            counts[i] += 1
        index[c] = i
    return unique, counts, index


def indexed_rows(rows: np.ndarray):
    """vectorised ``_indexed`` for an integer array [L] or [L, C].

    Same result as :func:`indexed` on ``list(map(tuple, rows))``; used where the
    Python loop would take minutes (tests check the two agree).
    """
    rows = np.asarray(rows)
    if rows.ndim == 1:
        rows = rows[:, None]
    L = rows.shape[0]
    if L == 0:
        return rows[:0].copy(), np.zeros(0, float), np.zeros(0, int)
    _, first, inverse = np.unique(rows, axis=0, return_index=True, return_inverse=True)
    inverse = inverse.reshape(-1)
    order = np.argsort(first, kind="stable")  # sorted-unique id -> first-seen rank
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    index = rank[inverse].astype(int)
    uniq = rows[np.sort(first)]
    counts = np.bincount(index, minlength=len(order)).astype(float)
    return uniq, counts, index


class Leaf:
    """restates ``LikelihoodTreeLeaf`` as built by ``make_likelihood_tree_leaf``
    (evolve/likelihood_tree.py:223-278,281-305): ``uniq`` symbols in first-seen
    order + a trailing gap symbol with count 0; ``input_likelihoods`` are the 0/1
    indicator rows (ambiguity codes -> several 1s, gap row all ones)."""

    def __init__(self, uniq, input_likelihoods, counts, index, name):
        self.uniq = uniq
        self.input_likelihoods = np.asarray(input_likelihoods, float)
        self.counts = np.asarray(counts, float)
        self.index = np.asarray(index, int)
        self.edge_name = name
        self.shape = list(self.input_likelihoods.shape)
        self.ambig = self.input_likelihoods.sum(axis=-1)
        self.is_tip = True


def make_leaf(codes, symbol_rows, name, loop=False):
    """tip compression from integer symbol codes.

    ``codes[L]``: one integer per alignment column (index into ``symbol_rows``);
    ``symbol_rows[S, M]``: indicator row of every symbol (``get_matched_array``,
    evolve/likelihood_tree.py:206-220).  The gap row appended at the end is all
    ones ("?"*motif_len resolves to every state; :230-232)."""
    codes = np.asarray(codes)
    if loop:
        uniq, counts, index = indexed([int(c) for c in codes])
    else:
        u, counts, index = indexed_rows(codes)
        uniq, counts = [int(x) for x in u[:, 0]], list(counts)
    M = symbol_rows.shape[1]
    table = np.concatenate([symbol_rows[uniq].reshape(len(uniq), M), np.ones((1, M))])
    return Leaf([*uniq, -1], table, [*counts, 0], index, name)


class Edge:
    """restates ``_LikelihoodTreeEdge.__init__`` (evolve/likelihood_tree.py:13-70)
    for pre-aligned children (the ``alignment is None`` branch)."""

    def __init__(self, children, name, loop=False):
        self.edge_name = name
        self.children = children
        M = children[0].shape[-1]
        assignments = [c.index for c in children]
        if loop:
            uniq, counts, self.index = indexed(list(zip(*assignments, strict=False)))
        else:
            u, counts, self.index = indexed_rows(np.stack(assignments, axis=1))
            uniq, counts = [tuple(int(x) for x in r) for r in u], list(counts)
        uniq.append(tuple(c.shape[0] - 1 for c in children))  # gap row (:49-51)
        counts.append(0)
        self.uniq = np.asarray(uniq, int).reshape(len(uniq), len(children))
        self.indexes = np.ascontiguousarray(self.uniq.T)  # [C, P+1]  (:55-58)
        self.counts = np.array(counts, float)  # (:62)
        self.shape = [len(self.uniq), M]  # (:66)
        ambigs = [c.ambig[ix] for ix, c in zip(self.indexes, children, strict=False)]
        self.ambig = np.prod(ambigs, axis=0)  # (:69-70)
        self.is_tip = False


def build_lht(children_of, names, leaves, loop=False):
    """restates ``recursive_lht_build`` (evolve/likelihood_calculation.py:103-112).

    ``children_of[n]``: ordered child ids of node n (empty for tips); node ids are
    in post-order, the root is the last.  Returns one Leaf/Edge per node."""
    nodes = [None] * len(children_of)
    for n, kids in enumerate(children_of):
        if len(kids) == 0:
            nodes[n] = leaves[names[n]]
        else:
            nodes[n] = Edge([nodes[k] for k in kids], names[n], loop=loop)
    return nodes


# --------------------------------------------------------------------------
# (a5) gamma rate classes
# --------------------------------------------------------------------------


def gamma_rates(weights, a):
    """restates ``GammaDefn.calc`` (recalculation/definition.py:236-244)."""
    from scipy.stats import gamma

    weights = np.asarray(weights, float)
    weights = weights / np.sum(weights)
    percentiles = np.add.accumulate(weights) - weights * 0.5
    medians = np.array([gamma.ppf(p, a, scale=1 / a) for p in percentiles])
    scale = np.sum(medians * weights)
    return medians / scale


# --------------------------------------------------------------------------
# (a6)-(a8) exponentiators
# --------------------------------------------------------------------------


class EigenExponentiator:
    """restates ``EigenExponentiator`` (maths/matrix_exponentiation.py:23-40)."""

    def __init__(self, Q, roots, ev, evT, evI):
        self.Q, self.roots, self.ev, self.evT, self.evI = Q, roots, ev, evT, evI

    def __call__(self, t):
        exp_roots = np.exp(t * self.roots)
        result = np.inner(self.evT * exp_roots, self.evI)
        if result.dtype.kind == "c":
            result = np.asarray(result.real)
        return np.maximum(result, 0.0)


def fast_exponentiator(Q):
    """restates ``FastExponentiator`` (maths/matrix_exponentiation.py:132-135)."""
    roots, evT = eig(Q)
    ev = evT.T
    return EigenExponentiator(Q, roots, ev, evT, inv(ev))


def checked_exponentiator(Q):
    """restates ``CheckedExponentiator`` (maths/matrix_exponentiation.py:138-146)."""
    roots, evT = eig(Q)
    ev = evT.T
    evI = inv(ev)
    reQ = np.inner(ev.T * roots, evI).real
    if not np.allclose(Q, reQ):
        raise ArithmeticError("eigen failed precision test")
    return EigenExponentiator(Q, roots, ev, evT, evI)


def pade_order_and_scale(norm):
    """the data-dependent (j, q) of ``PadeExponentiator.__call__``
    (maths/matrix_exponentiation.py:100-113)."""
    j = int(np.floor(np.log(max(norm, 0.5)) / np.log(2.0))) + 1
    e, q, qf = 1.0, 0, 1.0
    while e > 1e-12:
        q += 1
        q2 = 2.0 * q
        qf *= q**2 / (q2 * (q2 - 1) * q2 * (q2 + 1))
        e = 8 * (norm / (2**j)) ** (2 * q) * qf
    return j, q


class PadeExponentiator:
    """restates ``PadeExponentiator`` (maths/matrix_exponentiation.py:92-129)."""

    def __init__(self, Q):
        self.Q = Q

    def __call__(self, t=1.0):
        A = self.Q * t
        M = A.shape[0]
        norm = np.maximum.reduce(np.sum(np.absolute(A), axis=1))
        j, q = pade_order_and_scale(norm)
        A = A / 2.0**j
        X = A
        c = 1.0 / 2
        N = np.identity(M) + c * A
        D = np.identity(M) - c * A
        for k in range(2, q + 1):
            c = c * (q - k + 1) / (k * (2 * q - k + 1))
            X = np.dot(A, X)
            cX = c * X
            N = N + cX
            D = D + cX if not k % 2 else D - cX
        F = solve(D, N)
        for _ in range(1, j + 1):
            F = np.dot(F, F)
        return F


def make_exponentiator(Q, expm="either"):
    """restates ``ExpDefn.calc`` + ``_EigenPade.__call__``
    (evolve/substitution_calculation.py:48-86)."""
    allow_eigen, check_eigen, allow_pade = {
        "eigen": (True, False, False),
        "checked": (True, True, False),
        "pade": (False, False, True),
        "either": (True, True, True),
    }[str(expm)]
    if not allow_eigen:
        return PadeExponentiator(Q)
    eigen = checked_exponentiator if check_eigen else fast_exponentiator
    if not allow_pade:
        return eigen(Q)
    try:
        return eigen(Q)
    except (ArithmeticError, np.linalg.LinAlgError):
        return PadeExponentiator(Q)


# --------------------------------------------------------------------------
# (a10)-(a15) Felsenstein pruning and the site-weighted reduction
# --------------------------------------------------------------------------


def sum_input_likelihoods(child_indexes, result, likelihoods):
    """restates numba ``sum_input_likelihoods``
    (evolve/likelihood_tree_numba.py:7-25): result[p, m] = prod_c lik_c[idx[c, p], m]."""
    for c, plhs in enumerate(likelihoods):
        g = plhs[child_indexes[c]]
        if c == 0:
            result[:] = g
        else:
            result *= g
    return result


def partial_likelihoods(children_of, lht, psub_of_edge, fixed_motif=None):
    """post-order pruning for ONE rate class: restates the Defn chain built by
    ``make_partial_likelihood_defns`` (evolve/likelihood_calculation.py:75-100):
    per child edge ``numpy.inner(child_plh, psub)`` (:86), per internal node the
    gather-product (:40-47), optional fixed-motif mask (:50-59).

    ``psub_of_edge[n]``: P matrix of the edge above node n.  Returns plh per node."""
    plh = [None] * len(children_of)
    for n, kids in enumerate(children_of):
        if len(kids) == 0:
            plh[n] = lht[n].input_likelihoods  # (:32-37)
            continue
        inners = [np.inner(plh[k], psub_of_edge[k]) for k in kids]
        out = np.ones(lht[n].shape, float)
        sum_input_likelihoods(lht[n].indexes, out, inners)
        if fixed_motif is not None and fixed_motif[n] not in (None, -1):
            keep = fixed_motif[n]
            out[:, [m for m in range(out.shape[1]) if m != keep]] = 0.0
        plh[n] = out
    return plh


def root_lh(plh_root, root_mprobs):
    """restates ``CalcDefn(numpy.inner, name="lh")`` (likelihood_calculation.py:147-148)."""
    return np.inner(plh_root, root_mprobs)


def weighted_sum_lh(bprobs, lhs):
    """restates ``BinnedSiteDistribution.get_weighted_sum_lh``
    (evolve/likelihood_calculation.py:178-185): sequential sum over bins."""
    result = np.zeros(lhs[0].shape, float)
    for bprob, lh in zip(bprobs, lhs, strict=False):
        result += lh * bprob
    return result


def log_sum_across_sites(lhs, counts):
    """restates numba ``get_log_sum_across_sites``
    (evolve/likelihood_tree_numba.py:36-42): sequential FP64 accumulation."""
    with np.errstate(all="ignore"):
        log_lhs = np.log(lhs)
    terms = log_lhs * counts
    # counts==0 rows (the gap row) contribute log(lh)*0; lh>0 there so this is 0.
    # numpy.cumsum accumulates sequentially, like the reference loop.
    return float(np.cumsum(terms)[-1]) if len(terms) else 0.0


def log_likelihood(children_of, lht, psubs, root_mprobs, bprobs=None, fixed_motif=None,
                   return_parts=False):
    """total log-likelihood of one locus.

    ``psubs[b][n]``: P of the edge above node n for rate class b (list over bins).
    Restates ``make_total_loglikelihood_defn`` for sites_independent=True
    (evolve/likelihood_calculation.py:125-167) and ``BinnedLikelihood.__call__``
    (:238-245)."""
    root = len(children_of) - 1
    lhs, plhs = [], []
    for psub_b in psubs:
        plh = partial_likelihoods(children_of, lht, psub_b, fixed_motif)
        plhs.append(plh)
        lhs.append(root_lh(plh[root], root_mprobs))
    if len(psubs) > 1:
        mix = weighted_sum_lh(bprobs, lhs)
    else:
        mix = lhs[0]
    lnL = log_sum_across_sites(mix, lht[root].counts)
    if return_parts:
        return lnL, lhs, plhs
    return lnL


# --------------------------------------------------------------------------
# per-site accessors
# --------------------------------------------------------------------------


def full_length_likelihoods(lht_root, lh):
    """restates ``get_full_length_likelihoods`` (evolve/likelihood_tree.py:93-94)."""
    return lh[lht_root.index]


def bin_posterior_probs(lht_root, bprobs, lhs):
    """restates ``BinnedLikelihood.get_posterior_probs``
    (evolve/likelihood_calculation.py:247-257)."""
    result = np.array(
        [b * full_length_likelihoods(lht_root, p) for b, p in zip(bprobs, lhs, strict=False)]
    )
    result /= result.sum(axis=0)
    return result


# --------------------------------------------------------------------------
# work-unit counters used by bench.py
# --------------------------------------------------------------------------


def work_units(children_of, lht, n_bins=1, nodes=None):
    """U_mv = K*sum_edges P_child, U_gp = K*sum_edges P_parent (gap rows included),
    over the edges whose PARENT is in ``nodes`` (default: every internal node)."""
    u_mv = u_gp = 0
    for n, kids in enumerate(children_of):
        if len(kids) == 0 or (nodes is not None and n not in nodes):
            continue
        for k in kids:
            u_mv += lht[k].shape[0]
            u_gp += lht[n].shape[0]
    return n_bins * u_mv, n_bins * u_gp


# ---- phylo-HMM reduction over the sites -----------------------------------------
def patch_weighted_sum_lhs(alloc, weights, lhs):
    """restates ``PatchSiteDistribution.get_weighted_sum_lhs``
    (evolve/likelihood_calculation.py:209-216): result[patch] += lh_b * weight_b, bins in order"""
    result = np.zeros((max(alloc) + 1, *lhs[0].shape), float)
    temp = np.empty(lhs[0].shape, float)
    for patch, weight, lh in zip(alloc, weights, lhs, strict=True):
        temp[:] = lh
        temp *= weight
        result[patch] += temp
    return result


def log_dot_reduce(index, patch_probs, switch_probs, plhs):
    """restates ``LikelihoodTreeEdge.log_dot_reduce`` (evolve/likelihood_tree.py:168-176): the forward
    recursion of the site HMM with BASE = 2**100 rescaling.  ``plhs``: [patterns, patches]"""
    BASE = 2.0**100
    exponent = 0
    state_probs = np.array(patch_probs, float)
    for site in index:
        state_probs = np.dot(switch_probs, state_probs) * plhs[site]
        while max(state_probs) < 1.0:
            state_probs *= BASE
            exponent -= 1
    return float(np.log(sum(state_probs)) + exponent * np.log(BASE))


# ---- closed-form P of the F81 / HKY85 / TN93 family ------------------------------
def calc_TN93_P(mprobs, time, alpha1, alpha2):
    """restates numba ``calc_TN93_P`` (evolve/solved_models_numba.py:9-51), alphabet order T, C, A, G
    (``PredefinedNucleotide.calc_psub_matrix``, evolve/solved_models.py:43-49: F81 / HKY85 / TN93 when
    passed 0 / 1 / 2 rate parameters)"""
    import math

    mprobs = np.asarray(mprobs, float)
    alpha = (alpha1, alpha2)
    pi_star = (mprobs[0] + mprobs[1], mprobs[2] + mprobs[3])
    mu = (alpha1 * pi_star[0] + 1.0 * pi_star[1], 1.0 * pi_star[0] + alpha2 * pi_star[1])
    scale_factor = 0.0
    for motif in range(4):
        i = motif // 2
        other = 1 - i
        scale_factor += (alpha[i] * mprobs[2 * i + 1 - motif % 2] + pi_star[other]) * mprobs[motif]
    time = time / scale_factor
    e_beta_t = math.exp(-time)
    transversion = 1 - e_beta_t
    e_mu_t = [math.exp(-mu[i] * time) for i in range(2)]
    transition = [1 + (pi_star[1 - i] * e_beta_t - e_mu_t[i]) / pi_star[i] for i in range(2)]
    result = np.empty((4, 4), float)
    for row in range(4):
        i = row // 2
        for column in range(4):
            j = column // 2
            p = transition[i] if i == j else transversion
            p *= mprobs[column]
            if row == column:
                p += e_mu_t[i]
            result[row, column] = p
    return result


class TN93Exponentiator:
    """slot object of the closed-form family: ``__call__(t)`` like the other exponentiators"""

    def __init__(self, pi, alpha1, alpha2):
        self.pi, self.alpha1, self.alpha2 = np.array(pi, float), float(alpha1), float(alpha2)

    def __call__(self, t):
        return calc_TN93_P(self.pi, t, self.alpha1, self.alpha2)
