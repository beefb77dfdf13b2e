#!/bin/bash
# round-2 GPU session J (2 GPUs): sharded tests after the arena fix, multi-process tests, feed batch size of the file path
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2j
mkdir -p $OUT
echo "== sharded + multi-GPU tests" | tee $OUT/progress.txt
timeout 1500 python -m pytest -q -x -m gpu tests/test_sharded_gpu.py tests/test_multigpu_gpu.py > $OUT/tests_sharded.log 2>&1; echo "tests rc=$?" | tee -a $OUT/progress.txt
tail -25 $OUT/tests_sharded.log | cut -c1-300 | tee -a $OUT/progress.txt
echo "== file path: bases per parsed batch" | tee -a $OUT/progress.txt
for fb in 33554432 50331648 75497472 150994944; do
  KMGPU_FEED_BASES=$fb timeout 300 python - <<'PY' 2>&1 | tail -1 | tee -a $OUT/progress.txt
import os, sys, time, tempfile
sys.path.insert(0, os.getcwd())
import numpy as np
import bench, khmer_b200
n = 2_000_000
buf, off, _ = bench.synth_batch(777, n)
td = tempfile.mkdtemp(dir="/dev/shm")
path = os.path.join(td, "reads.fa")
block = np.empty((n, 154), dtype=np.uint8); block[:, :3] = np.frombuffer(b">r\n", dtype=np.uint8); block[:, 3:153] = buf.reshape(n, 150); block[:, -1] = 10
open(path, "wb").write(block.tobytes())
t = khmer_b200.Countgraph(20, 1e8, 4)
best = 1e9
for i in range(5):
    t0 = time.perf_counter(); r, k = t.consume_seqfile(path); dt = time.perf_counter() - t0
    if i: best = min(best, dt)
print("KMGPU_FEED_BASES=%s: %.2f G k-mers/s (%.1f ms)" % (os.environ["KMGPU_FEED_BASES"], k / best / 1e9, best * 1e3))
os.unlink(path); os.rmdir(td)
PY
done
find gpurun_out -size +20M -delete
