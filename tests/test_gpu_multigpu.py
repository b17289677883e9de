"""More than one engine / rank at a time on the GPU box.

* the fused all-reduce BETWEEN PROCESSES (CUDA IPC mailboxes, one process per GPU under torchrun) against
  the NCCL sum of the shard values.  With fewer GPUs than ranks (the driver's 1-GPU test box) the two
  ranks share the GPU: the IPC mapping, the epoch protocol in device memory and the CUDA-graph replay of
  the sharded evaluation are exercised all the same; NCCL (which refuses two ranks on one device) is
  replaced by gloo + a host sum;
* graph replay with the fused all-reduce attached (the evaluation number lives in device memory)."""

import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_fused_allreduce_between_processes():
    env = dict(os.environ)
    env.setdefault("MASTER_ADDR", "127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29547",
           os.path.join(ROOT, "tools", "check_fused_allreduce.py")]  # fmt: skip
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    tail = res.stdout[-1500:] + res.stderr[-2500:]
    assert res.returncode == 0, tail
    assert "fused all-reduce == " in res.stdout, tail


@pytest.mark.gpu
def test_graph_replay_with_the_fused_allreduce_attached():
    """three engines ("ranks") of one process exchange shard values through their mailboxes; after the
    structure has been seen often enough the sharded evaluation is ONE cudaGraphLaunch per engine
    (c3_stats.graph_replays grows) and still returns the total, bit for bit what plain launches return"""
    from cogent3_b200 import models as hostmodels
    from cogent3_b200.engine import evaluate_many
    from cogent3_b200.likelihood import LikelihoodFunction
    from cogent3_b200.synthetic import make_workload

    m = hostmodels.get_model("HKY85")
    topo, lengths, codes = make_workload(9, 6000, 4, seed=11)
    blocks = [(0, 2000), (2000, 4000), (4000, 6000)]
    lfs = []
    for a, b in blocks:
        lf = LikelihoodFunction(m, topo, bins=1)
        lf.set_alignment(codes[:, a:b].astype(np.int64), np.eye(4))
        lf.set_motif_probs(np.array([0.3, 0.2, 0.25, 0.25]))
        lf.lengths[:-1] = lengths[:-1]
        lf.set_param_rule("kappa", value=2.5)
        lfs.append(lf)
    shard_sum = 0.0
    for lf in lfs:
        shard_sum += lf.get_log_likelihood()
    engines = [lf.engine for lf in lfs]
    ptrs = [e.comm_mailbox(3)[0] for e in engines]
    for r, e in enumerate(engines):
        e.comm_attach(3, r, device_ptrs=ptrs)
    base = np.nan_to_num(lfs[0].lengths.copy())
    replays0 = engines[0].stats()["graph_replays"]
    totals = []
    for k in range(40):  # the same structure (a full sweep) 40 times: captured after 16 sightings
        d = (base * (1.0 + 1e-3 * (k % 5)))[:, None]
        for e in engines:
            e.set_distances(d)
        vals = evaluate_many(engines)
        assert vals[0] == vals[1] == vals[2]
        totals.append(vals[0])
    assert engines[0].stats()["graph_replays"] > replays0, "the sharded evaluation was never replayed from a graph"
    assert abs(totals[0] - shard_sum) <= 1e-12 * abs(shard_sum)
    # the same five distance tables came round eight times: plain launches and graph replays agree bit for bit
    for k in range(5, 40):
        assert totals[k] == totals[k % 5]
    for e in engines:
        e.comm_detach()
