"""Address-sharded sketches: ranks own bin ranges, k-mers are routed to the owners
through peer memory, the concatenated slices equal the single-sketch tables byte for byte.  The ranks are
emulated in one process on one GPU (peer pointers are then ordinary device pointers); with >= 2 GPUs the
one-process-per-GPU CUDA-IPC form runs as well."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol
from common import synth_reads

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _assemble(shards, n_tables):
    return [np.concatenate([s.slice_bytes(i) for s in shards]) for i in range(n_tables)]


@pytest.mark.parametrize("cls,k", [("Countgraph", 20), ("SmallCountgraph", 17), ("Nodegraph", 31), ("Counttable", 40),
                                   ("SmallCounttable", 33), ("Nodetable", 21)])
@pytest.mark.parametrize("world", [1, 3, 8])
def test_sharded_equals_single_sketch(cls, k, world):
    from khmer_b200 import cabi
    from khmer_b200.multigpu import split_by_bases, shard_range
    kind, hk, _ = ol.CLASSES[cls]
    sizes = ol.primes_near_x(4, 70000) if world < 8 else [1009, 997, 991, 64]     # tiny tables: some ranks own nothing
    reads = synth_reads(world * 7 + k, 900, 130, 6000, err=0.01, with_n=True) + ["A" * 300] * 40 + ["", "ACG"]
    shards = [cabi.Shard(kind, hk, k, sizes, r, world, max_positions=20000) for r in range(world)]
    cabi.attach_local_shards(shards)
    parts = []
    for r in range(world):
        lo, hi = shard_range(len(reads), r, world)
        buf, off = cabi.as_reads(reads[lo:hi])
        parts.append((buf, off, split_by_bases(buf, off, 20000)))
    rounds = max(len(p[2]) for p in parts)
    kmers = okmers = 0
    o = ol.Oracle(cls, k, sizes)
    for i in range(rounds):
        for r in range(world):                       # "all ranks route", then "all ranks apply", then "all ranks count"
            buf, off, runs = parts[r]
            if i < len(runs):
                a, b = runs[i]
                kmers += shards[r].route((buf, off[a:b + 1]))
                # the stream order of a round is rank 0's reads, then rank 1's, ...: what the single sketch is fed
                lo, _ = shard_range(len(reads), r, world)
                okmers += o.consume_reads(reads[lo + a: lo + b])
            else:
                shards[r].route((np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64)))
        for r in range(world):
            shards[r].offsets()
        for r in range(world):
            shards[r].push()
        for r in range(world):
            shards[r].apply()
        for r in range(world):
            shards[r].count_new()
    assert kmers == okmers
    tables = _assemble(shards, 4)
    for i in range(4):
        assert np.array_equal(tables[i], o.table(i)), "table %d" % i
    assert sum(s.local.n_occupied() for s in shards) == o.n_occupied()
    assert sum(s.stats()[1] for s in shards) == o.n_unique_kmers()
    for s in shards:
        s.close()


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
from khmer_b200 import cabi
from khmer_b200.multigpu import ShardedGroup, shard_range, split_by_bases
import oracle_lib as ol
from common import synth_reads
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
sizes = ol.primes_near_x(4, 3000000)
reads = synth_reads(5, 6000, 150, 40000) + ["AC" * 100] * 300
for cls, k in (("Countgraph", 20), ("SmallCountgraph", 31), ("Nodegraph", 32), ("Counttable", 40)):
    kind, hk, _ = ol.CLASSES[cls]
    sh = cabi.Shard(kind, hk, k, sizes, rank, world, device=rank, max_positions=200000)
    g = ShardedGroup(sh, dist, device=dev)
    g.attach()
    lo, hi = shard_range(len(reads), rank, world)
    mine = g.consume_reads(reads[lo:hi])
    # the single sketch is fed the rounds in their stream order: round i = run i of rank 0, of rank 1, ...
    o = ol.Oracle(cls, k, sizes)
    runs = []
    for r in range(world):
        a, b = shard_range(len(reads), r, world)
        buf, off = cabi.as_reads(reads[a:b])
        runs.append([(a + x, a + y) for x, y in split_by_bases(buf, off, 200000)])
    total = 0
    for i in range(max(len(x) for x in runs)):
        for r in range(world):
            if i < len(runs[r]):
                total += o.consume_reads(reads[runs[r][i][0]:runs[r][i][1]])
    t = torch.tensor([mine, sh.stats()[0], sh.stats()[1]], dtype=torch.int64, device=dev)
    dist.all_reduce(t)
    assert int(t[0]) == total and int(t[1]) == o.n_occupied() and int(t[2]) == o.n_unique_kmers(), (cls, t.tolist(), total, o.n_occupied(), o.n_unique_kmers())
    for i in range(4):
        lo_b, hi_b = sh.slice(i)
        want = o.table(i)
        got = sh.slice_bytes(i)
        per = 1 if kind == ol.BYTE else 2 if kind == ol.NIBBLE else 8
        assert np.array_equal(got, want[lo_b // per: lo_b // per + len(got)]), (cls, rank, i)
    dist.barrier()
    sh.close()
if rank == 0:
    print("SHARDED-IPC-OK")
dist.destroy_process_group()
'''


def test_sharded_one_process_per_gpu(tmp_path):
    from khmer_b200 import cabi
    if cabi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    n = min(cabi.device_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
                        "127.0.0.1", "--master-port", "29621", str(script)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "SHARDED-IPC-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


REFUSE = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np
from khmer_b200 import cabi
import oracle_lib as ol
from common import synth_reads
sizes = ol.primes_near_x(4, 3000000)
shards = [cabi.Shard(cabi.BYTE, cabi.TWOBIT, 20, sizes, r, 2, max_positions=200000) for r in range(2)]
cabi.attach_local_shards(shards)
big = synth_reads(1, 1000, 150, 40000)
small = synth_reads(2, 3, 150, 40000)
def round_(reads_by_rank):
    errs = 0
    for r in range(2):
        shards[r].route(cabi.as_reads(reads_by_rank[r]))
    for r in range(2):
        shards[r].offsets()
    for r in range(2):
        shards[r].push()
    for r in range(2):
        try:
            shards[r].apply()
        except cabi.KmgpuError as e:
            assert "arena" in str(e)
            errs += 1
    for r in range(2):
        shards[r].count_new()
    return errs
assert round_([big, big]) == 2            # 2 x 131 k k-mers x 4 tables against arenas of 20000 records: both owners refuse
assert sum(s.stats()[0] for s in shards) == 0
assert round_([small, small]) == 0        # the next round is taken as if nothing had happened
o = ol.Oracle("Countgraph", 20, sizes)
o.consume_reads(small); o.consume_reads(small)
assert sum(s.stats()[0] for s in shards) == o.n_occupied() and sum(s.stats()[1] for s in shards) == o.n_unique_kmers()
print("REFUSE-OK")
'''


def test_sharded_round_larger_than_the_arena_is_refused(tmp_path):
    script = tmp_path / "refuse.py"
    script.write_text(REFUSE % {"root": ROOT})
    r = subprocess.run([sys.executable, str(script)], env=dict(os.environ, KMGPU_SHARD_ARENA="20000"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "REFUSE-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
