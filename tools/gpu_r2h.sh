#!/bin/bash
# round-2 GPU session H (N GPUs of one box): multi-process tests, replicated and address-sharded bench lines
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
OUT=gpurun_out/r2h_n$N
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | tee $OUT/progress.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== multi-GPU tests" | tee -a $OUT/progress.txt
timeout 1200 python -m pytest -q -x -m gpu tests/test_multigpu_gpu.py tests/test_sharded_gpu.py tests/test_gpu_round2.py -k "ipc or one_process or first_touch or reduce_replicas" > $OUT/tests_multi.log 2>&1; echo "multi tests rc=$?" | tee -a $OUT/progress.txt
tail -4 $OUT/tests_multi.log | tee -a $OUT/progress.txt
P=29611
echo "== replicated bench, N=$N" | tee -a $OUT/progress.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 8 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench rc=$?" | tee -a $OUT/progress.txt
tail -1 $OUT/bench_n$N.json | cut -c1-1800 | tee -a $OUT/progress.txt
tail -3 $OUT/bench_n$N.err | tee -a $OUT/progress.txt
echo "== address-sharded bench, N=$N (128 GB of tables per GPU)" | tee -a $OUT/progress.txt
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus $N --mode sharded --steps 6 --warmup 2 > $OUT/bench_sharded_n$N.json 2> $OUT/bench_sharded_n$N.err; echo "sharded rc=$?" | tee -a $OUT/progress.txt
tail -1 $OUT/bench_sharded_n$N.json | cut -c1-1500 | tee -a $OUT/progress.txt
tail -3 $OUT/bench_sharded_n$N.err | tee -a $OUT/progress.txt
if [ "$N" -ge 4 ]; then
  echo "== reference arm contract check (rank 0 only)" | tee -a $OUT/progress.txt
fi
find gpurun_out -size +20M -delete
du -sm gpurun_out | tee -a $OUT/progress.txt
