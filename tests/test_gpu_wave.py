"""The wavefront launch (prune4_wave_kernel: every dirty binary node of a sweep in ONE persistent launch, all
tree levels streaming at once behind a static host-built strip schedule) against the per-level launches: the
arithmetic per row is the same, so every partial row, per-pattern likelihood and lnL must be BIT-identical --
for full sweeps, for dirty paths (one branch at a time), with fixed motifs, for K = 1, 2, 4, for schedules with
different lags, and it must step aside (same numbers through the level launches) for what it does not take:
polytomies below the root, rescaled rows."""

import numpy as np
import pytest

from cogent3_b200 import models as hostmodels
from cogent3_b200.likelihood import LikelihoodFunction
from cogent3_b200.synthetic import make_workload

pytestmark = pytest.mark.gpu

WAVE_ENV = ("C3_WAVE", "C3_WAVE_MIN", "C3_WAVE_LAG", "C3_WAVE_PRIO", "C3_WAVE_VERIFY", "C3_WAVE_GROUP", "C3_WAVE_PREFIX")


def build(monkeypatch, model, topo, lengths, codes, bins, wave, lag=None, prio=None, group=None, prefix=None):
    for k in WAVE_ENV:
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("C3_WAVE", "1" if wave else "0")
    monkeypatch.setenv("C3_WAVE_MIN", "1")  # also for the small sweeps of this test
    monkeypatch.setenv("C3_WAVE_VERIFY", "1")  # the host checks every schedule it builds
    if lag:
        monkeypatch.setenv("C3_WAVE_LAG", str(lag))
    if prio:
        monkeypatch.setenv("C3_WAVE_PRIO", str(prio))
    if group:
        monkeypatch.setenv("C3_WAVE_GROUP", str(group))
    if prefix:
        monkeypatch.setenv("C3_WAVE_PREFIX", str(prefix))
    m = hostmodels.get_model(model)
    lf = LikelihoodFunction(m, topo, bins=bins)
    lf.set_alignment(codes.astype(np.int64), np.eye(4))
    lf.set_motif_probs_from_data()
    lf.lengths[:-1] = np.asarray(lengths)[:-1]
    for k, p in enumerate(m.parameter_order):
        lf.set_param_rule(p, value=0.5 + 0.37 * k)
    if bins > 1:
        lf.set_param_rule("rate_shape", value=0.6)
    return lf


def same_rows(a, b, topo, bins):
    for n in range(topo.n_nodes):
        if topo.is_tip[n]:
            continue
        for k in range(bins):
            np.testing.assert_array_equal(a.engine.get_partials(n, k), b.engine.get_partials(n, k))
    for k in range(bins):
        np.testing.assert_array_equal(a.engine.get_lh(k), b.engine.get_lh(k))


@pytest.mark.parametrize(
    "model,n_tips,L,bins,kind,lag,prio,group,prefix",
    [
        ("GTR", 24, 60_000, 4, "evolved", None, None, None, None),
        ("HKY85", 16, 40_000, 1, "dense", 3, 16, None, None),
        ("HKY85", 20, 50_000, 2, "evolved", 6, 400, 4, None),
        ("HKY85", 40, 300_000, 4, "evolved", None, None, 8, None),
        ("GTR", 12, 3_000, 4, "dense", 12, 1, 16, None),
        ("GTR", 40, 300_000, 4, "evolved", None, None, None, 600),  # prefix mode: only the lower levels
        ("HKY85", 30, 200_000, 1, "evolved", 3, None, 2, 300),
    ],
)
def test_wave_launch_is_bit_identical_to_the_level_launches(model, n_tips, L, bins, kind, lag, prio, group, prefix,
                                                            monkeypatch):
    m = hostmodels.get_model(model)
    topo, lengths, codes = make_workload(n_tips, L, m.n_states, kind=kind, seed=13)
    plain = build(monkeypatch, model, topo, lengths, codes, bins, wave=False)
    wave = build(monkeypatch, model, topo, lengths, codes, bins, wave=True, lag=lag, prio=prio, group=group, prefix=prefix)
    base = plain.lengths.copy()
    # full sweeps: the schedule of a structure is built the second time it comes up
    for k in range(4):
        for lf in (plain, wave):
            lf.lengths[:] = base * (1.0 + 0.01 * k)
            lf._dirty_model = True
        assert plain.get_log_likelihood() == wave.get_log_likelihood()
    assert wave.engine.stats()["wave_launches"] >= 2
    assert plain.engine.stats()["wave_launches"] == 0
    same_rows(plain, wave, topo, bins)
    # dirty paths: one branch length at a time, each twice (the second time through the wave launch)
    rng = np.random.default_rng(5)
    before = wave.engine.stats()["wave_launches"]
    for n in rng.choice(topo.edges, size=4, replace=False):
        for rep in range(3):
            v = float(rng.uniform(0.01, 0.5))
            for lf in (plain, wave):
                lf.set_param_rule("length", edge=topo.names[n], value=v)
            assert plain.get_log_likelihood() == wave.get_log_likelihood()
    same_rows(plain, wave, topo, bins)
    assert prefix or wave.engine.stats()["wave_launches"] > before
    # fixed motifs (ancestral reconstruction): a node's rows masked
    internal = [n for n in range(topo.n_nodes - 1) if not topo.is_tip[n]]
    for motif in (0, 3):
        fixed = np.full(topo.n_nodes, -1)
        fixed[internal[len(internal) // 2]] = motif
        for _ in range(2):
            for lf in (plain, wave):
                lf.engine.set_fixed_motifs(fixed)
                lf.engine.invalidate()
            assert plain.engine.evaluate() == wave.engine.evaluate()
        same_rows(plain, wave, topo, bins)
    # and a full invalidation equals what the dirty-path updates left behind
    got = wave.engine.evaluate()
    wave.engine.invalidate()
    assert wave.engine.evaluate() == got


def test_wave_steps_aside_for_polytomies_and_rescaling(monkeypatch):
    from cogent3_b200.tree import Topology

    topo, lengths = Topology.from_newick(
        "((a:0.1,b:0.2,c:0.15):0.1,(d:0.3,e:0.2):0.05,((f:0.1,g:0.1):0.2,(h:0.2,(i:0.1,j:0.3):0.1):0.1):0.2);"
    )
    rng = np.random.default_rng(3)
    codes = rng.integers(0, 4, size=(int(np.sum(topo.is_tip)), 20_000))
    plain = build(monkeypatch, "HKY85", topo, lengths, codes, 4, wave=False)
    wave = build(monkeypatch, "HKY85", topo, lengths, codes, 4, wave=True)
    base = plain.lengths.copy()
    for k in range(3):
        for lf in (plain, wave):
            lf.lengths[:] = base * (1.0 + 0.01 * k)
            lf._dirty_model = True
        assert plain.get_log_likelihood() == wave.get_log_likelihood()
    assert wave.engine.stats()["wave_launches"] == 0  # a trifurcation below the root: level launches
    # a single branch inside the binary part: its path to the root is binary all the way -> wave
    edge = topo.names.index("i")
    for k in range(3):
        for lf in (plain, wave):
            lf.set_param_rule("length", edge="i", value=0.1 + 0.01 * k)
        assert plain.get_log_likelihood() == wave.get_log_likelihood()
    assert edge >= 0 and wave.engine.stats()["wave_launches"] >= 1
    n_wave = wave.engine.stats()["wave_launches"]
    for lf in (plain, wave):
        lf.engine.set_rescaling("on")
    for k in range(3):
        for lf in (plain, wave):
            lf.set_param_rule("length", edge="i", value=0.2 + 0.01 * k)
        assert plain.get_log_likelihood() == wave.get_log_likelihood()
    assert wave.engine.stats()["wave_launches"] == n_wave  # rescaled rows: level launches


@pytest.mark.parametrize("model,n_tips,L,bins", [("HKY85", 30, 2500, 1), ("GTR", 18, 700, 4), ("HKY85", 55, 3000, 2)])
def test_path_launch_is_bit_identical(model, n_tips, L, bins, monkeypatch):
    """single-branch updates of small problems go through prune4_path_kernel (one CTA walks the dirty path with
    the rows in shared memory and reduces): bit-identical to the small-levels launch it replaces, update after
    update, and the rows it leaves in global memory are what later updates start from"""
    m = hostmodels.get_model(model)
    topo, lengths, codes = make_workload(n_tips, L, m.n_states, kind="dense", seed=17)

    def make(no_path):
        for k in WAVE_ENV + ("C3_PATH",):
            monkeypatch.delenv(k, raising=False)
        if not no_path:
            monkeypatch.setenv("C3_PATH", "1")
        lf = LikelihoodFunction(m, topo, bins=bins)
        lf.set_alignment(codes.astype(np.int64), np.eye(4))
        lf.set_motif_probs_from_data()
        lf.lengths[:-1] = np.asarray(lengths)[:-1]
        for k, p in enumerate(m.parameter_order):
            lf.set_param_rule(p, value=0.5 + 0.37 * k)
        if bins > 1:
            lf.set_param_rule("rate_shape", value=0.6)
        return lf

    plain, path = make(True), make(False)
    assert plain.get_log_likelihood() == path.get_log_likelihood()
    rng = np.random.default_rng(2)
    launches = []
    for step in range(60):  # (past the CUDA-graph capture threshold of a structure)
        n = int(rng.choice(topo.edges)) if step % 3 else int(topo.edges[step % 5])
        v = float(rng.uniform(0.01, 0.6))
        for lf in (plain, path):
            lf.set_param_rule("length", edge=topo.names[n], value=v)
        assert plain.get_log_likelihood() == path.get_log_likelihood(), step
        launches.append((plain.engine.stats()["last_launches"], path.engine.stats()["last_launches"]))
    same_rows(plain, path, topo, bins)
    assert all(b <= a for a, b in launches)
    # a rate-matrix parameter: every edge dirty -> not a path, the usual launches; then paths again
    for lf in (plain, path):
        lf.set_param_rule(m.parameter_order[0], value=2.2)
    assert plain.get_log_likelihood() == path.get_log_likelihood()
    for lf in (plain, path):
        lf.set_motif_probs(np.array([0.1, 0.2, 0.3, 0.4]))  # only the root is dirty: a path of one node
    assert plain.get_log_likelihood() == path.get_log_likelihood()
    same_rows(plain, path, topo, bins)
