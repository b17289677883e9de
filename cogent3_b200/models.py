"""Host-side substitution models for the standalone API (bench, tests, smoke).

In the drop-in (cogent3_b200/dropin.py) cogent3's own model classes build Q; this
module restates just enough of them so the engine can be driven without cogent3
installed (the GPU box has no cogent3).  Semantics follow the reference and are
checked against Q matrices dumped from the live reference (tests/golden):

  * alphabet order T, C, A, G (cogent3 core/moltype.py DNA alphabet); codons are
    the 61 sense codons of the standard code in that order
  * exchangeability R = instantaneous-mask with each predicate's cells scaled by
    its parameter (evolve/substitution_model.py:818-823)
  * stationary models:  Q = R * mprobs_matrix; Q -= diag(rowsum);
    Q /= sum(pi * rowsum)                      (substitution_model.py:716-722)
  * general (GN):       same without the mprobs_matrix factor (:607-612)
  * "tuple" mprobs_matrix is the pi vector (column scaling); "conditional" is
    pi_j / Pr(context of the changed position) (evolve/motif_prob_model.py:257-274)
  * gamma rate classes: medians of K equiprobable bins scaled to weighted mean 1
    (recalculation/definition.py:236-244)
"""

from __future__ import annotations

from itertools import permutations, product

import numpy as np

DNA = ("T", "C", "A", "G")
_PURINES = {"A", "G"}
_PYRIMIDINES = {"T", "C"}

# IUPAC ambiguity -> set of states.  '-' is recoded to 'N' by the reference when
# recode_gaps=True (core/alignment.py:3097-3125) and '?' resolves to everything.
IUPAC_DNA = {
    "T": "T", "C": "C", "A": "A", "G": "G",
    "R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT", "M": "AC",
    "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT", "?": "ACGT", "-": "ACGT",
}  # fmt: skip

_STANDARD_AA = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
ALL_CODONS = tuple("".join(c) for c in product(DNA, repeat=3))
STANDARD_CODE = dict(zip(ALL_CODONS, _STANDARD_AA, strict=True))
SENSE_CODONS = tuple(c for c in ALL_CODONS if STANDARD_CODE[c] != "*")


def dna_symbol_rows():
    """(symbols, rows[S, 4]): indicator row per DNA symbol, unambiguous first."""
    symbols = list(IUPAC_DNA)
    rows = np.zeros((len(symbols), 4))
    for s, sym in enumerate(symbols):
        for base in IUPAC_DNA[sym]:
            rows[s, DNA.index(base)] = 1.0
    return symbols, rows


def recode_dna_gaps(codes: np.ndarray) -> np.ndarray:
    """what ``Alignment.get_gapped_seq(name, recode_gaps=True)`` does before the
    tip tables are built (core/alignment.py:3117-3123): every gap character
    ('-' and '?') becomes the most general ambiguity code 'N'."""
    symbols, _ = dna_symbol_rows()
    n = symbols.index("N")
    codes = np.array(codes, copy=True)
    codes[(codes == symbols.index("-")) | (codes == symbols.index("?"))] = n
    return codes


def encode_dna(seq: str, recode_gaps: bool = True) -> np.ndarray:
    symbols, _ = dna_symbol_rows()
    lut = np.full(256, -1, np.int16)
    for i, s in enumerate(symbols):
        lut[ord(s)] = i
        lut[ord(s.lower())] = i
    lut[ord("U")] = lut[ord("u")] = symbols.index("T")
    codes = lut[np.frombuffer(seq.encode("ascii"), np.uint8)]
    if (codes < 0).any():
        bad = seq[int(np.flatnonzero(codes < 0)[0])]
        raise ValueError(f"{bad!r} not in alphabet")
    codes = codes.astype(np.int64)
    return recode_dna_gaps(codes) if recode_gaps else codes


def codon_symbol_rows():
    """(symbols, rows[S, 61]) for sense codons plus the all-ambiguous 'NNN'."""
    symbols = [*SENSE_CODONS, "NNN"]
    rows = np.zeros((len(symbols), 61))
    rows[np.arange(61), np.arange(61)] = 1.0
    rows[61, :] = 1.0
    return symbols, rows


def encode_codons(seq: str) -> np.ndarray:
    """codon codes; any codon containing a gap/ambiguity that is not a sense codon
    becomes the all-ambiguous symbol 61 only when it is 'NNN' / '---' / '???';
    partial ambiguity is resolved state by state."""
    symbols, rows = codon_symbol_rows()
    lut = {c: i for i, c in enumerate(SENSE_CODONS)}
    out = np.empty(len(seq) // 3, np.int64)
    for k in range(len(out)):
        c = seq[3 * k : 3 * k + 3].upper().replace("U", "T")
        if c in lut:
            out[k] = lut[c]
        elif set(c) <= {"N", "-", "?"}:
            out[k] = 61
        else:
            raise ValueError(f"{c!r} not in alphabet (partial ambiguity needs resolve_codon_rows)")
    return out


class SubstitutionModel:
    """continuous-time model with predicate-masked rate parameters."""

    def __init__(self, name, alphabet, inst_mask, predicates, stationary, mprob_model):
        self.name = name
        self.alphabet = tuple(alphabet)
        self.n_states = len(self.alphabet)
        self.inst_mask = np.asarray(inst_mask, float)
        self.predicates = dict(predicates)  # name -> 0/1 mask, in parameter order
        self.parameter_order = list(self.predicates)
        self._indices = [np.nonzero(m) for m in self.predicates.values()]
        self._flat_indices = [np.flatnonzero(np.asarray(m)) for m in self.predicates.values()]
        self.stationary = stationary
        self.mprob_model = mprob_model
        self.word_length = len(self.alphabet[0])
        if mprob_model == "conditional":
            self._init_conditional()

    # -- conditional motif-prob model (motif_prob_model.py:93-140,257-274) --
    def _init_conditional(self):
        words, w = self.alphabet, self.word_length
        ctx_ids = {}
        M = self.n_states
        self._w2c = []
        member = []
        for i, word in enumerate(words):
            for d in range(w):
                key = (d, word[:d] + word[d + 1 :])
                c = ctx_ids.setdefault(key, len(ctx_ids))
                member.append((i, c))
        self._w2c = np.zeros((M, len(ctx_ids)))
        for i, c in member:
            self._w2c[i, c] = 1.0
        self._ctx = np.zeros((M, M), int)
        for i, j in zip(*np.nonzero(self.inst_mask), strict=True):
            diffs = [d for d in range(w) if words[i][d] != words[j][d]]
            assert len(diffs) == 1
            d = diffs[0]
            self._ctx[i, j] = ctx_ids[(d, words[j][:d] + words[j][d + 1 :])]

    def mprobs_matrix(self, mprobs):
        mprobs = np.asarray(mprobs, float)
        if self.mprob_model == "tuple":
            return mprobs
        context_probs = mprobs @ self._w2c
        context_probs[context_probs == 0.0] = np.inf
        return mprobs / context_probs.take(self._ctx)

    def exchangeability(self, params):
        R = self.inst_mask.copy()
        flat = R.reshape(-1)  # a view: same multiplications in the same order, cheaper indexing
        for ix, name in zip(self._flat_indices, self.parameter_order, strict=True):
            flat[ix] *= params[name]
        return R

    def calcQ(self, mprobs, params):
        """rate matrix calibrated to one expected substitution per unit time."""
        mprobs = np.asarray(mprobs, float)
        Q = self.exchangeability(params)
        if self.stationary:
            Q *= self.mprobs_matrix(mprobs)
        row_totals = Q.sum(axis=1)
        Q -= np.diag(row_totals)
        Q *= 1.0 / (mprobs * row_totals).sum()
        return Q

    def default_params(self):
        return dict.fromkeys(self.parameter_order, 1.0)


class SuppliedQModel(SubstitutionModel):
    """a fixed, already calibrated rate matrix (any alphabet up to 64 states): lets the engine
    be driven with the reference's own Q for models this host layer does not restate
    (dinucleotide, protein: tests/golden kat_dinucleotide / kat_protein)."""

    def __init__(self, Q, alphabet):
        Q = np.array(Q, float)
        super().__init__("supplied-Q", alphabet, (Q != 0) & ~np.eye(len(Q), dtype=bool), {}, True, "tuple")
        self._Q = Q

    def calcQ(self, mprobs, params):
        return self._Q.copy()


def _nucleotide_mask():
    return 1.0 - np.eye(4)


def _pair_mask(alphabet, x, y, directed=False):
    M = len(alphabet)
    m = np.zeros((M, M))
    for i, a in enumerate(alphabet):
        for j, b in enumerate(alphabet):
            if i == j:
                continue
            diffs = [(p, q) for p, q in zip(a, b, strict=True) if p != q]
            if len(diffs) != 1:
                continue
            p, q = diffs[0]
            if (p, q) == (x, y) or (not directed and (p, q) == (y, x)):
                m[i, j] = 1.0
    return m


def _transition_mask(alphabet, which="both"):
    m = np.zeros((len(alphabet),) * 2)
    if which in ("both", "y"):
        m += _pair_mask(alphabet, "C", "T")
    if which in ("both", "r"):
        m += _pair_mask(alphabet, "A", "G")
    return m


_GTR_PAIRS = ("AC", "AG", "AT", "CG", "CT")


def _codon_masks():
    M = len(SENSE_CODONS)
    inst = np.zeros((M, M))
    repl = np.zeros((M, M))
    for i, a in enumerate(SENSE_CODONS):
        for j, b in enumerate(SENSE_CODONS):
            if sum(p != q for p, q in zip(a, b, strict=True)) == 1:
                inst[i, j] = 1.0
                if STANDARD_CODE[a] != STANDARD_CODE[b]:
                    repl[i, j] = 1.0
    return inst, repl


def get_model(name: str) -> SubstitutionModel:
    """canned models named as in cogent3 (evolve/models.py:122-234)."""
    name_u = name.upper()
    if name_u in ("F81", "JC69"):
        return SubstitutionModel(name_u, DNA, _nucleotide_mask(), {}, True, "tuple")
    if name_u in ("HKY85", "K80"):
        preds = {"kappa": _transition_mask(DNA)}
        return SubstitutionModel(name_u, DNA, _nucleotide_mask(), preds, True, "tuple")
    if name_u == "TN93":
        preds = {"kappa_y": _transition_mask(DNA, "y"), "kappa_r": _transition_mask(DNA, "r")}
        return SubstitutionModel(name_u, DNA, _nucleotide_mask(), preds, True, "tuple")
    if name_u == "GTR":
        preds = {f"{x}/{y}": _pair_mask(DNA, x, y) for x, y in _GTR_PAIRS}
        return SubstitutionModel("GTR", DNA, _nucleotide_mask(), preds, True, "conditional")
    if name_u == "GN":
        preds = {
            f"{f}>{t}": _pair_mask(DNA, f, t, directed=True)
            for f, t in permutations("ACTG", 2)
            if f != "T" or t != "G"
        }
        return SubstitutionModel("GN", DNA, _nucleotide_mask(), preds, False, "tuple")
    if name_u in ("CNFGTR", "CNFHKY", "MG94GTR", "MG94HKY", "GY94", "Y98"):
        inst, repl = _codon_masks()
        if name_u.endswith("GTR"):
            preds = {f"{x}/{y}": _pair_mask(SENSE_CODONS, x, y) for x, y in _GTR_PAIRS}
        else:
            preds = {"kappa": _transition_mask(SENSE_CODONS)}
        preds["omega"] = repl
        if name_u.startswith("CNF"):
            mp = "conditional"
        elif name_u.startswith("MG94"):
            mp = "monomer"
        else:
            mp = "tuple"
        if mp == "monomer":
            raise NotImplementedError("monomer motif-prob model is not on the hot path (SURVEY §8)")
        return SubstitutionModel(name_u, SENSE_CODONS, inst, preds, True, mp)
    raise ValueError(f"unknown model {name!r}")


def gamma_rates(n_bins, shape, bprobs=None):
    """K gamma-median rates with weighted mean 1 (recalculation/definition.py:236-244)."""
    from scipy.stats import gamma

    w = np.full(n_bins, 1.0 / n_bins) if bprobs is None else np.asarray(bprobs, float)
    w = w / np.sum(w)
    percentiles = np.add.accumulate(w) - w * 0.5
    medians = np.array([gamma.ppf(p, shape, scale=1 / shape) for p in percentiles])
    return medians / np.sum(medians * w)


class EigenSystem:
    """what the reference keeps per distinct Q (maths/matrix_exponentiation.py:132-146):
    ``roots``, ``evT`` and ``evI = inv(evT.T)`` from numpy's general ``eig``; kept on
    the host so eigenpairs are bit-identical to the reference's."""

    __slots__ = ("Q", "evI", "evT", "roots")

    def __init__(self, Q, check=True):
        roots, evT = np.linalg.eig(Q)
        ev = evT.T
        evI = np.linalg.inv(ev)
        if check:
            reQ = np.inner(ev.T * roots, evI).real
            # numpy.allclose(Q, reQ) spelled out (rtol 1e-5, atol 1e-8; NaN compares False): the
            # generic isclose machinery costs 60 us per call, a third of a parameter step's host time
            if not bool(np.all(np.abs(Q - reQ) <= 1e-8 + 1e-5 * np.abs(reQ))):
                raise ArithmeticError("eigen failed precision test")
        self.Q, self.roots, self.evT, self.evI = Q, roots, evT, evI
