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
