"""Replicated sketches merged over NVLink peer memory == one sketch fed all reads (needs >= 2 GPUs; skipped on a
1-GPU box).  Single-process form (kmgpu_reduce_replicas) and the one-process-per-GPU CUDA-IPC form."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol
from common import synth_reads

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    from khmer_b200 import cabi
    return cabi.device_count()


@pytest.mark.parametrize("cls", ["Countgraph", "SmallCountgraph", "Nodegraph"])
def test_reduce_replicas_single_process(cls):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    from khmer_b200 import cabi
    n = min(_ngpu(), 4)
    kind, hk, _ = ol.CLASSES[cls]
    sizes = ol.primes_near_x(3, 50000)
    reads = synth_reads(3, 1200, 100, 4000) + ["A" * 140] * 400
    reps = [cabi.Sketch(kind, hk, 17, sizes, device=d) for d in range(n)]
    for d, sk in enumerate(reps):
        sk.consume_reads(reads[d::n])
    cabi.reduce_replicas(reps)
    o = ol.Oracle(cls, 17, sizes)
    o.consume_reads(reads)
    for sk in reps:
        for i in range(3):
            assert np.array_equal(sk.table(i), o.table(i))
        assert sk.n_occupied() == o.n_occupied()


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
from khmer_b200 import cabi
from khmer_b200.multigpu import ReplicaGroup, shard_range
import oracle_lib as ol
from common import synth_reads
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
sizes = ol.primes_near_x(4, 300000)
reads = synth_reads(9, 4000, 120, 20000) + ["AC" * 70] * 500
for cls in ("Countgraph", "SmallCountgraph", "Nodegraph"):
    kind, hk, _ = ol.CLASSES[cls]
    sk = cabi.Sketch(kind, hk, 21, sizes, device=rank)
    lo, hi = shard_range(len(reads), rank, world)
    sk.consume_reads(reads[lo:hi])
    g = ReplicaGroup(sk, dist, device=torch.device("cuda", rank))
    g.merge()
    o = ol.Oracle(cls, 21, sizes)
    o.consume_reads(reads)
    for i in range(4):
        assert np.array_equal(sk.table(i), o.table(i)), (cls, rank, i)
    assert sk.n_occupied() == o.n_occupied()
    g.detach()
    sk.close()
dist.barrier()
if rank == 0:
    print("IPC-MERGE-OK")
dist.destroy_process_group()
'''


def test_ipc_merge_one_process_per_gpu(tmp_path):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    n = min(_ngpu(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
                        "127.0.0.1", "--master-port", "29611", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "IPC-MERGE-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
