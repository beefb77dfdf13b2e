"""Multi-GPU plumbing for replicated sketches: one process per GPU, each ingests a disjoint read
shard into its own replica; the replicas are merged by a saturating-add / OR reduction over NVLink peer memory.

torch.distributed is used for what it is good at here — rendezvous, barriers and moving 64-byte IPC handles —
never for the table data itself, which the kernels read directly from the peers' HBM.
"""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous, rank-ordered shard [lo, hi) of n_items (rank order == stream order, so that "first toucher"
    arguments carry over: the lowest rank that touched a bin holds its first toucher)."""
    per, rem = divmod(n_items, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


class ReplicaGroup:
    """The replicas of one sketch across the ranks of a torch.distributed process group."""

    def __init__(self, sketch, dist=None, device=None):
        self.sketch = sketch
        self.dist = dist
        self.device = device
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.attached = False

    def barrier(self):
        if self.dist is not None and self.world > 1:
            self.dist.barrier()

    def attach(self):
        """Exchange the CUDA-IPC handles of every rank's tables and map the peers (collective)."""
        if self.world == 1:
            return
        import torch
        mine = np.ascontiguousarray(self.sketch.ipc_export(), dtype=np.uint8)
        t = torch.from_numpy(mine.copy())
        if self.device is not None:
            t = t.to(self.device)
        gathered = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(gathered, t)
        allh = torch.cat(gathered).cpu().numpy()
        self.sketch.ipc_attach(self.rank, self.world, allh)
        self.attached = True
        self.barrier()

    def merge(self):
        """All replicas become the saturating sum (OR for Bloom filters) of all replicas (collective)."""
        if self.world == 1:
            return
        if not self.attached:
            self.attach()
        self.barrier()                       # every rank has finished ingesting
        self.sketch.reduce_scatter_peers()   # rank r folds slice r of all peers into its own copy
        self.barrier()                       # all slices final before anyone copies them
        self.sketch.all_gather_peers()       # rank r pulls the other slices from their owners
        self.barrier()

    def detach(self):
        if self.attached:
            self.barrier()
            self.sketch.ipc_detach()
            self.attached = False


def split_by_bases(buf, off, max_bases):
    """Cut (buffer, offsets) into runs of whole reads of at most max_bases bases each (a single longer read gets
    its own run: callers create their shards with max_positions >= the longest read)."""
    runs = []
    n = len(off) - 1
    r = 0
    while r < n:
        r1 = r + 1
        while r1 < n and int(off[r1 + 1]) - int(off[r]) <= max_bases:
            r1 += 1
        runs.append((r, r1))
        r = r1
    return runs


class ShardedGroup:
    """An address-sharded sketch across the ranks of a process group: every rank hashes its own reads, routes
    each counter update to the owner of its bin through NVLink peer memory (route), and applies what it
    received (apply).  The barriers between the two phases come from torch.distributed."""

    def __init__(self, shard, dist=None, device=None):
        self.shard = shard
        self.dist = dist
        self.device = device
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1

    def barrier(self):
        if self.dist is not None and self.world > 1:
            self.dist.barrier()

    def attach(self):
        import torch
        mine = np.ascontiguousarray(self.shard.ipc_export(), dtype=np.uint8)
        if self.world == 1:
            self.shard.ipc_attach(mine)
            return
        t = torch.from_numpy(mine.copy())
        if self.device is not None:
            t = t.to(self.device)
        gathered = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(gathered, t)
        self.shard.ipc_attach(torch.cat(gathered).cpu().numpy())
        self.barrier()

    def consume_reads(self, reads, clean=True):
        """Collective: every rank passes its own shard of the reads; returns the k-mers this rank hashed."""
        from . import cabi
        buf, off = cabi.as_reads(reads)
        runs = split_by_bases(buf, off, self.shard.max_positions)
        rounds = len(runs)
        if self.dist is not None and self.world > 1:
            import torch
            t = torch.tensor([rounds], dtype=torch.int64, device=self.device if self.device is not None else "cpu")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            rounds = int(t.item())
        kmers = 0
        for i in range(rounds):
            if i < len(runs):
                r0, r1 = runs[i]
                kmers += self.shard.route((buf, off[r0:r1 + 1]), clean=clean)
            self.barrier()          # every update of this round sits in its owner's store
            self.shard.apply()
            self.barrier()          # every owner has marked the new positions of the round
            self.shard.count_new()
            self.barrier()          # nobody reads a peer's bitmap or store any more: the next round may start
        return kmers
