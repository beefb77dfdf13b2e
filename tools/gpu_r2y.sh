#!/bin/bash
# round-2: replicated bench on 2 GPUs, final tree
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2y
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 8 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench rc=$?" | tee $OUT/progress.txt
python -c "
import json; d=json.loads(open('$OUT/bench_n2.json').read().strip().split('\n')[-1])
print('N=2 value %.2f e2e %.2f packed %.2f' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e_packed']['value']/1e9), d['merge_check'])" | tee -a $OUT/progress.txt
tail -3 $OUT/bench_n2.err | cut -c1-300 | tee -a $OUT/progress.txt
