"""loader for tests/golden/*.npz (written by tests/golden/make_golden.py)."""

from __future__ import annotations

import glob
import json
import os

import numpy as np

from cogent3_b200 import models as hostmodels
from cogent3_b200.tree import Topology

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden:
    def __init__(self, case):
        self.case = case
        self.z = np.load(os.path.join(GOLDEN_DIR, case + ".npz"), allow_pickle=False)
        self.meta = json.loads(str(self.z["meta"]))
        self.topo = Topology(self.meta["names"], self.meta["children"])
        self.codes = self.z["codes"]
        self.n_bins = self.meta["bins"]
        self.alphabet = self.meta["alphabet"]
        self.n_states = len(self.alphabet)
        if "symbol_rows" in self.meta:
            self.symbol_rows = np.array(self.meta["symbol_rows"], float)
        elif self.n_states == 4:
            self.symbol_rows = hostmodels.dna_symbol_rows()[1]
        else:
            self.symbol_rows = hostmodels.codon_symbol_rows()[1]
        self.lnL = self.meta["lnL"]
        self.psubs = self.z["psubs"]  # [K, n_nodes, M, M]
        self.Q = self.z["Q"]
        self.lh = self.z["lh"]  # [K, P_root+1]
        self.lengths = self.z["lengths"]
        self.distances = self.z["distances"]
        self.mprobs = np.array(self.meta["mprobs"], float)
        self.bprobs = np.array(self.meta["bprobs"], float)

    def codes_by_tip(self):
        return dict(zip(self.meta["tips"], self.codes, strict=True))

    def node(self, n, key):
        return self.z[f"n{n}_{key}"]
