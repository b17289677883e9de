"""Site-pattern sharding across GPUs.

Sites are independent given the parameters, so the alignment's COLUMNS are split into
contiguous blocks, one per rank (one process per GPU).  Every rank runs the per-node
pattern compression on its own block (pattern sets are local; duplicates across shards
only cost a little redundancy and keep the counts exact), holds the full -- tiny --
parameter state, and evaluates its block's log-likelihood on its own device.  The only
communication per evaluation is ONE all-reduce (sum) of a single double: NCCL over
NVLink on GPUs (the scalar is written by ``c3_evaluate_device`` straight into the tensor
that is reduced, on the same CUDA stream), gloo in the CPU tests.  This is the
site-parallel scheme PyCogent's removed MPI mode used
(cogent3 evolve/likelihood_calculation.py:140-145: "sum over ... other sites (ie: cpus)").
"""

from __future__ import annotations

import numpy as np

from .likelihood import LikelihoodFunction


def column_block(n_columns: int, world_size: int, rank: int):
    """contiguous [start, stop) of this rank; sizes differ by at most one column"""
    base, extra = divmod(n_columns, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def attach_fused_allreduce(engine, group=None) -> bool:
    """Fuse the sum over ranks into the engine's reduction kernel (one node, one process per GPU):
    every rank allocates a mailbox, the 64-byte CUDA IPC handles are all-gathered through the
    process group, and from then on ``engine.evaluate()`` returns the TOTAL on every rank -- the
    shard values travel as peer-memory stores over NVLink inside the kernel, no collective call
    and no extra device-to-host copy per evaluation.  Returns False (caller keeps the NCCL
    all-reduce) when it is disabled (C3_SHARD_NCCL=1) or peer mapping fails on any rank."""
    import os

    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1 or os.environ.get("C3_SHARD_NCCL"):
        return False
    ok = 1
    try:
        _ptr, handle = engine.comm_mailbox(world)
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        engine.comm_attach(world, rank, ipc_handles=handles)
    except Exception:  # noqa: BLE001 - every rank must agree on the outcome
        ok = 0
    flag = torch.tensor([ok], device="cuda" if dist.get_backend(group) == "nccl" else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        if ok:
            engine.comm_detach()
        return False
    return True


class ShardedLikelihoodFunction:
    """``LikelihoodFunction`` whose alignment is sharded over a process group.

    All ranks must apply the same parameter changes (the optimiser runs replicated, as
    every rank sees the same total lnL)."""

    def __init__(self, model, topo, bins=1, group=None, device=None, engine_factory=None, **kw):
        import torch
        import torch.distributed as dist

        self._torch, self._dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.on_gpu = engine_factory is None
        if self.on_gpu:
            device = torch.cuda.current_device() if device is None else device
            self.lf = LikelihoodFunction(model, topo, bins=bins, device=device,
                                         stream=torch.cuda.current_stream(device), **kw)  # fmt: skip
            self._buf = torch.zeros(1, dtype=torch.float64, device=f"cuda:{device}")
        else:
            self.lf = LikelihoodFunction(model, topo, bins=bins, engine_factory=engine_factory, **kw)
            self._buf = torch.zeros(1, dtype=torch.float64)

    def set_alignment(self, codes, symbol_rows=None, already_sharded=False):
        """``codes[n_tips, L]`` in ``topo.tip_names`` order: the FULL alignment (this rank
        takes its column block) or, with ``already_sharded``, just this rank's block."""
        codes = np.asarray(codes)
        if not already_sharded:
            a, b = column_block(codes.shape[1], self.world, self.rank)
            codes = codes[:, a:b]
        self.lf.set_alignment(codes, symbol_rows)
        self.fused = bool(self.on_gpu and self.world > 1 and attach_fused_allreduce(self.lf.engine, self.group))
        return self

    def set_motif_probs_from_data(self):
        """global (all shards) unambiguous-state frequencies: one all-reduce of M counts"""
        lf = self.lf
        counts = lf.state_counts()
        t = self._torch.as_tensor(counts, dtype=self._torch.float64, device=self._buf.device)
        if self.world > 1:
            self._dist.all_reduce(t, group=self.group)
        counts = t.cpu().numpy()
        lf.set_motif_probs(counts / counts.sum())

    def __getattr__(self, name):  # parameter API is the wrapped function's
        if name in ("set_param_rule", "set_motif_probs", "set_time_heterogeneity", "get_param_value",
                    "lengths", "params", "model", "topo", "n_bins", "patterns", "engine", "rates"):  # fmt: skip
            return getattr(self.lf, name)
        raise AttributeError(name)

    _rescale_tried = False

    def get_log_likelihood(self) -> float:
        lf = self.lf
        lf._sync()
        if getattr(self, "fused", False):
            return lf.engine.evaluate()  # the total: the all-reduce happened inside the reduction kernel
        if self.on_gpu:
            lf.engine.evaluate_device(self._buf.data_ptr())  # async, same stream as NCCL's input
        else:
            self._buf[0] = lf.engine.evaluate()
        if self.world > 1:
            self._dist.all_reduce(self._buf, group=self.group)
        total = float(self._buf.item())
        if self.on_gpu and not np.isfinite(total) and not lf.engine.rescaling()[0] and not self._rescale_tried:
            # c3_evaluate_device is asynchronous, so the engine's automatic switch to per-pattern
            # rescaling (c3_set_rescaling) cannot see the value: every rank sees the same total
            # and switches here, then the evaluation is repeated once
            self._rescale_tried = True
            lf.engine.set_rescaling("on")
            return self.get_log_likelihood()
        return total


def loci_of_rank(n_loci: int, world_size: int, rank: int):
    """whole loci are dealt round-robin: locus i lives on rank i % world_size"""
    return list(range(rank, n_loci, world_size))


class MultiLocusLikelihoodFunction:
    """Several alignments ("loci") on one tree, whole loci dealt to the ranks.

    The reference builds one likelihood tree per locus and sums the per-locus log-likelihoods
    (``make_likelihood_function(tree, loci=[...])``; ``SumDefn`` over the locus dimension,
    evolve/likelihood_calculation.py:161-163, recalculation/definition.py:672-676).  Loci are a
    second independent axis next to the columns: every rank owns the engines of its
    loci, launches them together (``c3_evaluate_many``: each on its own stream), adds its values in
    locus order and the ranks' sums meet in ONE all-reduce of a double per evaluation (NCCL on
    GPUs, gloo in the CPU tests).  Parameters are global unless a ``locus`` is named
    (cogent3's ``set_param_rule(..., locus=...)`` / ``is_independent`` over loci)."""

    def __init__(self, model, topo, loci, bins=1, group=None, device=None, engine_factory=None, **kw):
        import torch
        import torch.distributed as dist

        self._torch, self._dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.locus_names = list(loci)
        self.local = loci_of_rank(len(self.locus_names), self.world, self.rank)
        self.on_gpu = engine_factory is None
        self.lfs = {}
        for i in self.local:
            if self.on_gpu:
                dev = torch.cuda.current_device() if device is None else device
                # stream=None: every engine creates its own stream, so the loci of a rank overlap
                self.lfs[i] = LikelihoodFunction(model, topo, bins=bins, device=dev, stream=None, **kw)
            else:
                self.lfs[i] = LikelihoodFunction(model, topo, bins=bins, engine_factory=engine_factory, **kw)
        if self.on_gpu:
            dev = torch.cuda.current_device() if device is None else device
            self._buf = torch.zeros(1, dtype=torch.float64, device=f"cuda:{dev}")
        else:
            self._buf = torch.zeros(1, dtype=torch.float64)

    def _index(self, locus):
        return self.locus_names.index(locus) if not isinstance(locus, int) else locus

    def set_alignment(self, alignments, symbol_rows=None):
        """``alignments``: one ``codes[n_tips, L_i]`` per locus, in ``loci`` order (entries of
        loci that live on other ranks may be None)"""
        assert len(alignments) == len(self.locus_names)
        for i in self.local:
            self.lfs[i].set_alignment(np.asarray(alignments[i]), symbol_rows)
        return self

    def set_motif_probs_from_data(self, pooled=True):
        """``pooled``: one set of motif probabilities from all loci (an all-reduce of M counts);
        otherwise every locus uses its own"""
        if not pooled:
            for lf in self.lfs.values():
                lf.set_motif_probs_from_data()
            return
        M = next(iter(self.lfs.values())).model.n_states if self.lfs else None
        counts = np.zeros(M if M else 0)
        for lf in self.lfs.values():
            counts = counts + lf.state_counts()
        t = self._torch.as_tensor(counts, dtype=self._torch.float64, device=self._buf.device)
        if self.world > 1:
            self._dist.all_reduce(t, group=self.group)
        counts = t.cpu().numpy()
        for lf in self.lfs.values():
            lf.set_motif_probs(counts / counts.sum())

    def set_param_rule(self, par_name, value=None, edge=None, edges=None, locus=None, **kw):
        """every rank makes the same calls; a rank ignores loci it does not own"""
        targets = self.local if locus is None else [self._index(locus)]
        for i in targets:
            if i in self.lfs:
                self.lfs[i].set_param_rule(par_name, value=value, edge=edge, edges=edges, **kw)

    def set_lengths(self, lengths, locus=None):
        targets = self.local if locus is None else [self._index(locus)]
        for i in targets:
            if i in self.lfs:
                self.lfs[i].lengths[:-1] = np.asarray(lengths)[:-1]
                self.lfs[i]._dirty_model = True

    def locus_log_likelihoods(self):
        """{locus index: lnL} of this rank's loci, evaluated together"""
        lfs = [self.lfs[i] for i in self.local]
        for lf in lfs:
            lf._sync()
        if self.on_gpu and len(lfs) > 1:
            from .engine import evaluate_many

            vals = evaluate_many([lf.engine for lf in lfs])
        else:
            vals = [lf.engine.evaluate() for lf in lfs]
        return {i: float(v) for i, v in zip(self.local, vals, strict=True)}

    def get_log_likelihood(self) -> float:
        local = 0.0
        for _i, v in sorted(self.locus_log_likelihoods().items()):
            local += v  # locus order: the same bits on every run
        self._buf[0] = local
        if self.world > 1:
            self._dist.all_reduce(self._buf, group=self.group)
        return float(self._buf.item())
