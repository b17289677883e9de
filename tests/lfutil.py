"""build a cogent3_b200 LikelihoodFunction configured like a golden case."""

from __future__ import annotations

from cogent3_b200 import models as hostmodels
from cogent3_b200.likelihood import LikelihoodFunction


def lf_from_golden(g, engine_factory=None, device=0):
    model = g.meta["model"]
    if model == "supplied-Q":
        model = hostmodels.SuppliedQModel(g.Q[0, 0], g.alphabet)
    lf = LikelihoodFunction(
        model, g.topo, bins=g.n_bins, expm=g.meta["expm"], engine_factory=engine_factory,
        device=device,
    )  # fmt: skip
    lf.set_alignment(g.codes_by_tip(), g.symbol_rows)
    lf.set_motif_probs(g.mprobs)
    for n in g.topo.edges:
        lf.set_param_rule("length", edge=g.topo.names[n], value=float(g.lengths[n]))
    for p, v in g.meta.get("params", {}).items():
        lf.set_param_rule(p, value=v)
    for edge, pv in g.meta.get("edge_params", {}).items():
        for p, v in pv.items():
            lf.set_param_rule(p, edge=edge, value=v)
    if "rate_shape" in g.meta:
        lf.set_param_rule("rate_shape", value=g.meta["rate_shape"])
    if g.n_bins > 1:
        lf.set_param_rule("bprobs", value=g.bprobs)
    return lf
