"""world_size-2 (and 3) sharding over gloo on CPU: the sum of the shards' log-likelihoods,
combined by ONE scalar all-reduce, equals the unsharded value; ranks stay in lock-step."""

import functools
import os
import socket

import numpy as np
import pytest

from cogent3_b200.sharded import column_block


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist

    from cogent3_b200.sharded import ShardedLikelihoodFunction
    from cogent3_b200.synthetic import make_workload
    from oracle.engine import OracleEngine

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        topo, lengths, codes = make_workload(9, 1001, 4, seed=5)
        lf = ShardedLikelihoodFunction("HKY85", topo, bins=4,
                                       engine_factory=functools.partial(OracleEngine, mode="port"))
        lf.set_alignment(codes.astype(np.int64), np.eye(4))
        lf.set_motif_probs_from_data()
        lf.lengths[:-1] = lengths[:-1]
        lf.set_param_rule("kappa", value=3.0)
        lf.set_param_rule("rate_shape", value=0.7)
        v1 = lf.get_log_likelihood()
        lf.set_param_rule("length", edge=topo.names[0], value=0.31)
        v2 = lf.get_log_likelihood()
        out.put((rank, v1, v2, lf.patterns.n_columns, lf.lf.mprobs.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_unsharded(world):
    import torch.multiprocessing as mp

    from cogent3_b200.likelihood import LikelihoodFunction
    from cogent3_b200.synthetic import make_workload
    from oracle.engine import OracleEngine

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    topo, lengths, codes = make_workload(9, 1001, 4, seed=5)
    ref = LikelihoodFunction("HKY85", topo, bins=4, engine_factory=OracleEngine)
    ref.set_alignment(codes.astype(np.int64), np.eye(4))
    ref.set_motif_probs_from_data()
    ref.lengths[:-1] = lengths[:-1]
    ref.set_param_rule("kappa", value=3.0)
    ref.set_param_rule("rate_shape", value=0.7)
    w1 = ref.get_log_likelihood()
    ref.set_param_rule("length", edge=topo.names[0], value=0.31)
    w2 = ref.get_log_likelihood()
    assert sum(r[3] for r in results) == 1001  # every column is on exactly one rank
    for _rank, v1, v2, _n, mprobs in results:
        assert v1 == results[0][1] and v2 == results[0][2]  # all ranks see the same scalar
        assert abs(v1 - w1) <= 1e-12 * abs(w1)
        assert abs(v2 - w2) <= 1e-12 * abs(w2)
        np.testing.assert_allclose(mprobs, ref.mprobs, rtol=1e-14)


def test_column_blocks_partition():
    for L in (0, 1, 7, 1000, 1001):
        for world in (1, 2, 3, 8):
            blocks = [column_block(L, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == L
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


# ---- whole loci dealt to the ranks (second axis) -----------------------------------
def _loci_problem():
    from cogent3_b200.synthetic import make_workload

    topo, lengths, _ = make_workload(7, 10, 4, seed=21)
    alns = [make_workload(7, L, 4, seed=21, shard=k)[2].astype(np.int64) for k, L in enumerate((300, 511, 77, 1200, 64))]
    return topo, lengths, alns


def _configure_loci(lf, lengths):
    lf.set_motif_probs_from_data()
    lf.set_lengths(lengths)
    lf.set_param_rule("kappa", value=2.5)
    lf.set_param_rule("kappa", value=5.0, locus="l3")  # one locus with its own kappa
    lf.set_param_rule("rate_shape", value=0.8)


def _loci_worker(rank, world, port, out):
    import torch.distributed as dist

    from cogent3_b200.sharded import MultiLocusLikelihoodFunction
    from oracle.engine import OracleEngine

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        topo, lengths, alns = _loci_problem()
        names = [f"l{i}" for i in range(len(alns))]
        lf = MultiLocusLikelihoodFunction("HKY85", topo, names, bins=2,
                                          engine_factory=functools.partial(OracleEngine, mode="port"))
        # a rank only needs the alignments of its own loci
        lf.set_alignment([a if i % world == rank else None for i, a in enumerate(alns)], np.eye(4))
        _configure_loci(lf, lengths)
        v1 = lf.get_log_likelihood()
        lf.set_param_rule("length", edge=topo.names[1], value=0.27)
        v2 = lf.get_log_likelihood()
        out.put((rank, v1, v2, sorted(lf.local)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_loci_dealt_to_ranks_equal_the_sum_of_single_locus_functions(world):
    import torch.multiprocessing as mp

    from cogent3_b200.likelihood import LikelihoodFunction
    from oracle.engine import OracleEngine

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_loci_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    topo, lengths, alns = _loci_problem()
    singles = []
    for a in alns:
        lf = LikelihoodFunction("HKY85", topo, bins=2, engine_factory=OracleEngine)
        lf.set_alignment(a, np.eye(4))
        singles.append(lf)
    counts = sum(lf.state_counts() for lf in singles)
    want = []
    for step in range(2):
        tot = 0.0
        for i, lf in enumerate(singles):
            lf.set_motif_probs(counts / counts.sum())
            lf.lengths[:-1] = lengths[:-1]
            lf.set_param_rule("kappa", value=5.0 if i == 3 else 2.5)
            lf.set_param_rule("rate_shape", value=0.8)
            if step == 1:
                lf.set_param_rule("length", edge=topo.names[1], value=0.27)
            tot += lf.get_log_likelihood()
        want.append(tot)
    owned = sorted(i for r in results for i in r[3])
    assert owned == list(range(len(alns)))  # every locus on exactly one rank
    for _rank, v1, v2, _loc in results:
        assert v1 == results[0][1] and v2 == results[0][2]
        assert abs(v1 - want[0]) <= 1e-12 * abs(want[0])
        assert abs(v2 - want[1]) <= 1e-12 * abs(want[1])
