#!/bin/bash
# round-2 GPU session I (2 GPUs): address-sharded sketches, exchange v3 (local grouping, exact receive layout, bulk push)
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
OUT=gpurun_out/r2i_n$N
mkdir -p $OUT
echo "== sharded tests" | tee $OUT/progress.txt
timeout 1500 python -m pytest -q -x -m gpu tests/test_sharded_gpu.py > $OUT/tests_sharded.log 2>&1; echo "sharded tests rc=$?" | tee -a $OUT/progress.txt
tail -30 $OUT/tests_sharded.log | cut -c1-300 | tee -a $OUT/progress.txt
P=29711
echo "== address-sharded bench, N=$N" | tee -a $OUT/progress.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --mode sharded --steps 6 --warmup 2 > $OUT/bench_sharded_n$N.json 2> $OUT/bench_sharded_n$N.err; echo "sharded rc=$?" | tee -a $OUT/progress.txt
tail -1 $OUT/bench_sharded_n$N.json | cut -c1-1600 | tee -a $OUT/progress.txt
tail -3 $OUT/bench_sharded_n$N.err | cut -c1-400 | tee -a $OUT/progress.txt
echo "== address-sharded bench, N=1" | tee -a $OUT/progress.txt
timeout 600 python bench.py --mode sharded --steps 6 --warmup 2 > $OUT/bench_sharded_n1.json 2> $OUT/bench_sharded_n1.err; echo "sharded n1 rc=$?" | tee -a $OUT/progress.txt
tail -1 $OUT/bench_sharded_n1.json | cut -c1-1600 | tee -a $OUT/progress.txt
echo "== bench default (file legs with the SIMD packer)" | tee -a $OUT/progress.txt
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?" | tee -a $OUT/progress.txt
python -c "
import json; d=json.load(open('$OUT/bench_default.json'))
print('value %.2f e2e %.2f file %.2f' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e_file']['value']/1e9)); print(json.dumps(d['e2e_file'].get('compressed'))); print(d['parity_check']['ok'])" | tee -a $OUT/progress.txt
find gpurun_out -size +20M -delete
