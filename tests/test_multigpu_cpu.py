"""Host-side logic of the N > 1 path on CPU: world_size-2 gloo process group, stub sketch (no GPU here):
handle exchange order, barrier placement between reduce-scatter and all-gather, shard and slice arithmetic."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class StubSketch:
    """Records what ReplicaGroup asks of a sketch."""

    def __init__(self, rank, log):
        self.rank, self.log, self.attached = rank, log, None

    def _w(self, what):
        with open(self.log, "a") as fh:
            fh.write("%d %s\n" % (self.rank, what))

    def ipc_export(self):
        return np.full(64 * 3, self.rank + 1, dtype=np.uint8)     # 3 tables x 64-byte handle

    def ipc_attach(self, rank, world, handles):
        self.attached = (rank, world, np.array(handles))

    def reduce_scatter_peers(self):
        self._w("rs")

    def all_gather_peers(self):
        self._w("ag")

    def ipc_detach(self):
        self._w("detach")


def _worker(rank, world, port, log, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from khmer_b200.multigpu import ReplicaGroup, shard_range
    sk = StubSketch(rank, log)
    g = ReplicaGroup(sk, dist)
    assert (g.rank, g.world) == (rank, world)
    g.attach()
    r, w, h = sk.attached
    assert (r, w) == (rank, world) and h.shape == (world * 64 * 3,)
    for q in range(world):
        assert (h[q * 192:(q + 1) * 192] == q + 1).all()          # rank-ordered handles
    g.merge()
    g.merge()
    g.detach()
    lo, hi = shard_range(1001, rank, world)
    t = torch.tensor([lo, hi])
    allr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
    if rank == 0:
        spans = [tuple(x.tolist()) for x in allr]
        assert spans[0][0] == 0 and spans[-1][1] == 1001
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        open(out, "w").write("ok")
    dist.destroy_process_group()


def test_replica_group_gloo_world2(tmp_path):
    log, out = str(tmp_path / "log"), str(tmp_path / "out")
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, log, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
    lines = [ln.split() for ln in open(log).read().split("\n") if ln]
    ops = [op for _, op in lines]
    # two merges: in each, both reduce-scatters are logged before either all-gather (barrier in between)
    assert ops[:4].count("rs") == 2 and ops[:2] == ["rs", "rs"] and ops[2:4] == ["ag", "ag"]
    assert ops[4:6] == ["rs", "rs"] and ops[6:8] == ["ag", "ag"] and ops[8:] == ["detach", "detach"]


def test_slice_ranges_cover_every_table_once():
    from khmer_b200 import cabi
    for n_words in (0, 1, 3, 4, 5, 1000, 25000000 - 3, 25000000):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                a, b = cabi.slice_range(n_words, world, r)
                assert a == prev and a <= b <= n_words and (a % 4 == 0 or a == b)
                prev = b
            assert prev == n_words
    with pytest.raises(cabi.KmgpuError):
        cabi.slice_range(10, 2, 2)


def test_shard_range():
    from khmer_b200.multigpu import shard_range
    for n in (0, 1, 7, 8, 1000):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _socket_worker(rank, world, port, log, out):
    from khmer_b200.multigpu import ReplicaGroup, SocketComm
    comm = SocketComm(rank=rank, world=world, addr="127.0.0.1", port=port)
    assert comm.all_gather_bytes(bytes([rank + 7]) * (rank + 1)) == [bytes([7]), bytes([8, 8])][:world]
    assert comm.all_reduce_sum(10 + rank) == sum(10 + r for r in range(world))
    comm.barrier()
    sk = StubSketch(rank, log)
    g = ReplicaGroup(sk, comm)
    g.attach()
    r, w, h = sk.attached
    assert (r, w) == (rank, world) and h.shape == (world * 64 * 3,)
    for q in range(world):
        assert (h[q * 192:(q + 1) * 192] == q + 1).all()
    g.merge()
    g.detach()
    comm.close()
    if rank == 0:
        open(out, "w").write("ok")


def test_replica_group_socket_rendezvous_world2(tmp_path):
    """the dependency-free rendezvous (no torch.distributed): same handle order and barrier placement"""
    log, out = str(tmp_path / "log"), str(tmp_path / "out")
    port = 31500 + (os.getpid() % 500)
    mp.spawn(_socket_worker, args=(2, port, log, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
    ops = [ln.split()[1] for ln in open(log).read().split("\n") if ln]
    assert ops[:2] == ["rs", "rs"] and ops[2:4] == ["ag", "ag"] and ops[4:] == ["detach", "detach"]


class StubShard:
    """Records what ShardedGroup asks of a shard (the protocol of a round: route | offsets | push | apply | count_new)."""

    def __init__(self, rank, log, max_positions):
        self.rank, self.log, self.max_positions = rank, log, max_positions
        self.routed = []

    def _w(self, what):
        with open(self.log, "a") as fh:
            fh.write("%d %s\n" % (self.rank, what))

    def ipc_export(self):
        return np.full(64 * 5, self.rank + 1, dtype=np.uint8)

    def ipc_attach(self, handles):
        self.handles = np.array(handles)

    def route(self, reads, clean=True):
        buf, off = reads
        self.routed.append(len(off) - 1)
        self._w("route")
        return int(off[-1] - off[0])

    def offsets(self):
        self._w("offsets")

    def push(self):
        self._w("push")

    def apply(self):
        self._w("apply")

    def count_new(self):
        self._w("count_new")

    def stats(self):
        return 10 + self.rank, 100 + self.rank, 0


def _sharded_worker(rank, world, port, log, out):
    from khmer_b200.multigpu import ShardedGroup, SocketComm
    comm = SocketComm(rank=rank, world=world, addr="127.0.0.1", port=port)
    sh = StubShard(rank, log, max_positions=1000)
    g = ShardedGroup(sh, comm)
    g.attach()
    assert sh.handles.shape == (world * 64 * 5,)
    # rank 0 has three rounds of reads, rank 1 a single one: it must still take part in rounds 2 and 3 with empty input
    n_reads = 30 if rank == 0 else 8
    reads = ["ACGT" * 25] * n_reads
    got = g.consume_reads(reads)
    assert got == 100 * n_reads
    assert sh.routed == ([10, 10, 10] if rank == 0 else [8, 0, 0])
    assert g.stats() == (21, 201)
    comm.close()
    if rank == 0:
        open(out, "w").write("ok")


def test_sharded_group_protocol_world2(tmp_path):
    """every phase of a round is finished on all ranks before the next one starts anywhere (barriers), and a rank that has run
    out of reads keeps posting empty routes"""
    log, out = str(tmp_path / "log"), str(tmp_path / "out")
    port = 32500 + (os.getpid() % 500)
    mp.spawn(_sharded_worker, args=(2, port, log, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
    ops = [ln.split()[1] for ln in open(log).read().split("\n") if ln]
    want = []
    for _ in range(3):
        for phase in ("route", "offsets", "push", "apply", "count_new"):
            want += [phase, phase]
    assert ops == want


def _sharded_worker_gloo(rank, world, port, log, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from khmer_b200.multigpu import ShardedGroup
    sh = StubShard(rank, log, max_positions=500)
    g = ShardedGroup(sh, dist)
    g.attach()
    n_reads = 4 if rank == 0 else 11                     # 1 round against 3: rank 0 posts two empty routes
    assert g.consume_reads(["ACGT" * 25] * n_reads) == 100 * n_reads
    assert sh.routed == ([4, 0, 0] if rank == 0 else [5, 5, 1])
    assert g.stats() == (21, 201)
    if rank == 0:
        open(out, "w").write("ok")
    dist.destroy_process_group()


def test_sharded_group_protocol_gloo_world2(tmp_path):
    """the same protocol over a torch.distributed (gloo) group"""
    log, out = str(tmp_path / "log"), str(tmp_path / "out")
    port = 29000 + (os.getpid() % 400)
    mp.spawn(_sharded_worker_gloo, args=(2, port, log, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
    ops = [ln.split()[1] for ln in open(log).read().split("\n") if ln]
    want = []
    for _ in range(3):
        for phase in ("route", "offsets", "push", "apply", "count_new"):
            want += [phase, phase]
    assert ops == want
