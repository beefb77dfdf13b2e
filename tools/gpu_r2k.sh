#!/bin/bash
# round-2 GPU session K (8 GPUs of one box): config C5 address-sharded at 128 GB of tables per GPU, replicated scaling line
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2k
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | tee $OUT/progress.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
P=29811
for N in 8 4; do
  echo "== address-sharded bench, N=$N (128 GB of tables per GPU)" | tee -a $OUT/progress.txt
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+N)) bench.py --gpus $N --mode sharded --steps 6 --warmup 2 > $OUT/bench_sharded_n$N.json 2> $OUT/bench_sharded_n$N.err; echo "sharded rc=$?" | tee -a $OUT/progress.txt
  tail -1 $OUT/bench_sharded_n$N.json | cut -c1-1500 | tee -a $OUT/progress.txt
  tail -2 $OUT/bench_sharded_n$N.err | cut -c1-300 | tee -a $OUT/progress.txt
done
echo "== replicated bench, N=8" | tee -a $OUT/progress.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((P+20)) bench.py --gpus 8 --steps 8 --warmup 3 > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "bench rc=$?" | tee -a $OUT/progress.txt
tail -1 $OUT/bench_n8.json | cut -c1-1800 | tee -a $OUT/progress.txt
tail -2 $OUT/bench_n8.err | cut -c1-300 | tee -a $OUT/progress.txt
echo "== one process per GPU tests (4 ranks)" | tee -a $OUT/progress.txt
timeout 600 python -m pytest -q -x -m gpu tests/test_sharded_gpu.py tests/test_multigpu_gpu.py -k "one_process" > $OUT/tests.log 2>&1; echo "tests rc=$?" | tee -a $OUT/progress.txt
tail -3 $OUT/tests.log | cut -c1-300 | tee -a $OUT/progress.txt
find gpurun_out -size +20M -delete
