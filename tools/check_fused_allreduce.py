"""Multi-process check of the fused all-reduce (run under torchrun, one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_fused_allreduce.py

Every rank holds a column shard; with the mailboxes attached (CUDA IPC handles all-gathered
through the process group) ``get_log_likelihood`` returns the total on every rank and must equal
the NCCL all-reduce of the shard values."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cogent3_b200 import models as hostmodels  # noqa: E402
from cogent3_b200.sharded import ShardedLikelihoodFunction  # noqa: E402
from cogent3_b200.synthetic import make_workload  # noqa: E402

local_rank = int(os.environ.get("LOCAL_RANK", "0"))
world_env = int(os.environ.get("WORLD_SIZE", "1"))
n_dev = torch.cuda.device_count()
# fewer GPUs than ranks (the 1-GPU test box): the ranks share devices -- CUDA IPC maps a mailbox of
# another PROCESS on the same device just as well -- and the process group is gloo (NCCL refuses two
# ranks on one GPU); the reference value is then the host sum of the all-gathered shard values
SHARED = n_dev < world_env
torch.cuda.set_device(local_rank % n_dev)
dist.init_process_group("gloo" if SHARED else "nccl")
rank, world = dist.get_rank(), dist.get_world_size()
m = hostmodels.get_model("GTR")
topo, lengths, codes = make_workload(20, 200_000, 4, kind="evolved", seed=31)


def build(force_nccl):
    if force_nccl:
        os.environ["C3_SHARD_NCCL"] = "1"
    else:
        os.environ.pop("C3_SHARD_NCCL", None)
    lf = ShardedLikelihoodFunction(m, topo, bins=4)
    if SHARED and force_nccl:
        # no NCCL here: the shard value comes back to the host and the ranks' values are gathered
        def host_sum(self=lf):
            self.lf._sync()
            vals = [None] * world
            dist.all_gather_object(vals, self.lf.engine.evaluate())
            total = 0.0
            for v in vals:
                total += v
            return total

        lf.get_log_likelihood = host_sum
    lf.set_alignment(codes)
    lf.set_motif_probs(np.array([0.2, 0.3, 0.25, 0.25]))
    lf.lengths[:-1] = np.asarray(lengths)[:-1]
    lf.set_param_rule("rate_shape", value=0.7)
    return lf


fused, nccl = build(False), build(True)
assert fused.fused or world == 1, "fused all-reduce did not attach"
assert not nccl.fused
rng = np.random.default_rng(5)
for it in range(6):
    a, b = fused.get_log_likelihood(), nccl.get_log_likelihood()
    assert abs(a - b) <= 1e-12 * abs(b), (rank, it, a, b)
    vals = [None] * world
    dist.all_gather_object(vals, a)
    assert len(set(vals)) == 1, vals  # identical on every rank
    n = int(rng.choice(topo.edges))
    v = float(rng.uniform(0.02, 0.4))
    for f in (fused, nccl):
        f.set_param_rule("length", edge=topo.names[n], value=v)
# every branch length changes (the bench's step), many times: both mailbox parities, no stale value
base = fused.lf.lengths.copy()
for it in range(40):
    for f in (fused, nccl):
        f.lf.lengths[:] = base * (1.0 + 1e-3 * ((it % 7) + 1))
        f.lf._dirty_model = True
    a, b = fused.get_log_likelihood(), nccl.get_log_likelihood()
    assert abs(a - b) <= 1e-12 * abs(b), (rank, it, a, b)
if rank == 0:
    print(f"fused all-reduce == NCCL all-reduce on {world} ranks: lnL {a!r}")
dist.barrier()
dist.destroy_process_group()
