#!/bin/bash
# round-2 GPU session N: k_apply2 with fewer instructions per lightly loaded bucket
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2n
mkdir -p $OUT
echo "== tests of the grouped path" | tee $OUT/progress.txt
timeout 1500 python -m pytest -q -x -m gpu tests/test_gpu_parity.py tests/test_gpu_benchscale.py -k "group or many_buckets or benchscale" > $OUT/tests_group.log 2>&1; echo "group tests rc=$?" | tee -a $OUT/progress.txt
tail -4 $OUT/tests_group.log | cut -c1-300 | tee -a $OUT/progress.txt
echo "== large tables" | tee -a $OUT/progress.txt
timeout 900 python tools/bench_configs.py --no-queries C3 C4 C4S C4L > $OUT/configs.jsonl 2> $OUT/configs.err; echo "configs rc=$?" | tee -a $OUT/progress.txt
cut -c1-330 $OUT/configs.jsonl | tee -a $OUT/progress.txt
echo "== headline shape through the grouped path" | tee -a $OUT/progress.txt
KMGPU_PREFER_BINS=0 timeout 600 python bench.py --no-cpu --no-check --no-file --steps 8 --warmup 4 2> $OUT/bench_grouped.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('grouped path on C1: value %.2f e2e %.2f' % (d['value'] / 1e9, d['e2e']['value'] / 1e9))" | tee -a $OUT/progress.txt
find gpurun_out -size +20M -delete
