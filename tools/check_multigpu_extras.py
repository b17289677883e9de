"""Multi-process checks of the two remaining multi-GPU paths (run under torchrun, one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29534 tools/check_multigpu_extras.py

1. whole loci dealt to the ranks (MultiLocusLikelihoodFunction): the NCCL-reduced total equals the
   sum of all loci evaluated by rank 0 alone on its own GPU;
2. column shards of a tree deep enough to underflow FP64: with the fused all-reduce every rank sees
   the same non-finite total, switches to per-pattern rescaling and repeats (c3_set_rescaling,
   automatic mode); the NCCL path switches in ShardedLikelihoodFunction.  Both totals must equal the
   unsharded, rescaled value."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cogent3_b200.likelihood import LikelihoodFunction  # noqa: E402
from cogent3_b200.sharded import MultiLocusLikelihoodFunction, ShardedLikelihoodFunction  # noqa: E402
from cogent3_b200.synthetic import make_workload  # noqa: E402

local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl")
rank, world = dist.get_rank(), dist.get_world_size()

# ---- 1. loci -----------------------------------------------------------------------------
topo, lengths, _ = make_workload(16, 10, 4, seed=41)
sizes = (30_000, 5_000, 120_000, 64, 77_777)
alns = [make_workload(16, L, 4, seed=41, shard=k)[2] for k, L in enumerate(sizes)]
names = [f"locus{i}" for i in range(len(alns))]


def configure(lf):
    lf.set_motif_probs_from_data()
    lf.set_lengths(lengths)
    lf.set_param_rule("kappa", value=3.0)
    lf.set_param_rule("kappa", value=6.0, locus="locus2")
    lf.set_param_rule("rate_shape", value=0.6)


ml = MultiLocusLikelihoodFunction("HKY85", topo, names, bins=4)
ml.set_alignment([a if i % world == rank else None for i, a in enumerate(alns)], np.eye(4))
configure(ml)
totals = []
for step in range(3):
    if step:
        ml.set_param_rule("length", edge=topo.names[step], value=0.1 * step)
    totals.append(ml.get_log_likelihood())
if rank == 0:
    # all loci on this one GPU, outside the process group's view
    solo = MultiLocusLikelihoodFunction.__new__(MultiLocusLikelihoodFunction)
    solo._torch, solo._dist, solo.group, solo.rank, solo.world = torch, dist, None, 0, 1
    solo.locus_names, solo.local, solo.on_gpu = names, list(range(len(names))), True
    solo.lfs = {i: LikelihoodFunction("HKY85", topo, bins=4, device=local_rank) for i in solo.local}
    solo._buf = torch.zeros(1, dtype=torch.float64, device=f"cuda:{local_rank}")
    solo.set_alignment(alns, np.eye(4))
    # pooled motif probabilities: the same counts the ranks all-reduced
    counts = sum(lf.state_counts() for lf in solo.lfs.values())
    for lf in solo.lfs.values():
        lf.set_motif_probs(counts / counts.sum())
    solo.set_lengths(lengths)
    solo.set_param_rule("kappa", value=3.0)
    solo.set_param_rule("kappa", value=6.0, locus="locus2")
    solo.set_param_rule("rate_shape", value=0.6)
    for step in range(3):
        if step:
            solo.set_param_rule("length", edge=topo.names[step], value=0.1 * step)
        want = solo.get_log_likelihood()
        assert abs(totals[step] - want) <= 1e-12 * abs(want), (step, totals[step], want)
    print(f"loci: {len(names)} loci on {world} ranks, totals {totals} == single-GPU sums (rel 1e-12)")
vals = [None] * world
dist.all_gather_object(vals, totals)
assert all(v == vals[0] for v in vals)

# ---- 2. automatic rescaling under sharding ----------------------------------------------------
topo, lengths, codes = make_workload(700, 4_000, 4, kind="dense", seed=11)


def deep(cls, **kw):
    lf = cls("HKY85", topo, bins=4, **kw)
    lf.set_alignment(codes.astype(np.int64), np.eye(4))
    lf.set_motif_probs(np.array([0.2, 0.3, 0.25, 0.25]))
    lf.lengths[:-1] = np.asarray(lengths)[:-1] * 4.0
    lf.set_param_rule("kappa", value=2.0)
    lf.set_param_rule("rate_shape", value=0.6)
    return lf


os.environ.pop("C3_SHARD_NCCL", None)
fused = deep(ShardedLikelihoodFunction)
os.environ["C3_SHARD_NCCL"] = "1"
nccl = deep(ShardedLikelihoodFunction)
os.environ.pop("C3_SHARD_NCCL", None)
a, b = fused.get_log_likelihood(), nccl.get_log_likelihood()
assert fused.fused and not nccl.fused
assert fused.lf.engine.rescaling()[0] and nccl.lf.engine.rescaling()[0]
assert np.isfinite(a) and abs(a - b) <= 1e-12 * abs(b), (a, b)
whole = deep(LikelihoodFunction, device=local_rank)
want = whole.get_log_likelihood()
assert whole.engine.rescaling()[0]
assert abs(a - want) <= 1e-11 * abs(want), (a, want)
first = (a, b, want)
# keep going in lock-step after the switch
for it in range(4):
    for f in (fused, nccl):
        f.set_param_rule("length", edge=topo.names[it], value=0.2 + 0.1 * it)
    a, b = fused.get_log_likelihood(), nccl.get_log_likelihood()
    assert np.isfinite(a) and abs(a - b) <= 1e-12 * abs(b), (it, a, b)
vals = [None] * world
dist.all_gather_object(vals, a)
assert len(set(vals)) == 1
if rank == 0:
    print(f"rescaling: 700 taxa x 4000 columns over {world} ranks, rescued automatically: fused {first[0]!r} == nccl {first[1]!r}; unsharded {first[2]!r}")
dist.barrier()
dist.destroy_process_group()
