"""the three CPU evaluators of oracle/engine.py agree (numpy restatement is the pinned one)"""

import functools

import numpy as np
import pytest
from goldenutil import Golden
from lfutil import lf_from_golden

from oracle.engine import OracleEngine


@pytest.mark.parametrize("case", ["c2_gtr_gamma4_mid", "c4_cnfgtr_small", "c3_gn_either_small",
                                  "c1_brca1_hky85_point2"])  # fmt: skip
@pytest.mark.parametrize("mode", ["port", "fused"])
def test_c_oracle_matches_numpy_oracle(case, mode):
    g = Golden(case)
    lf = lf_from_golden(g, engine_factory=functools.partial(OracleEngine, mode=mode))
    lnL = lf.get_log_likelihood()
    assert abs(lnL - g.lnL) <= 1e-11 * abs(g.lnL)
    for b in range(g.n_bins):
        np.testing.assert_allclose(lf.engine.get_lh(b), g.lh[b], rtol=0, atol=1e-13)
