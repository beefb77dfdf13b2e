#!/bin/bash
# round-2 GPU session D
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2d
mkdir -p $OUT
echo "== smoke" | tee $OUT/progress.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/progress.txt
echo "== changed tests first" | tee -a $OUT/progress.txt
timeout 2400 python -m pytest -q -x -s -m gpu tests/test_sharded_gpu.py tests/test_gpu_round2.py tests/test_gpu_normalize.py tests/test_host_gpu.py > $OUT/tests_new.log 2>&1; echo "new tests rc=$?" | tee -a $OUT/progress.txt
tail -6 $OUT/tests_new.log | tee -a $OUT/progress.txt
grep "normalize 1M" $OUT/tests_new.log | tee -a $OUT/progress.txt
echo "== bench default (full line)" | tee -a $OUT/progress.txt
timeout 1200 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?" | tee -a $OUT/progress.txt
python -c "
import json; d=json.load(open('$OUT/bench_default.json'))
print('value %.2f e2e %.2f file %.2f' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e_file']['value']/1e9)); print(json.dumps(d['secondary'])); print(d['parity_check']['ok'])" | tee -a $OUT/progress.txt
echo "== BASELINE table sizes" | tee -a $OUT/progress.txt
timeout 1500 python tools/bench_configs.py C2 C3 C4 C4S NORM C5 > $OUT/configs.jsonl 2> $OUT/configs.err; echo "configs rc=$?" | tee -a $OUT/progress.txt
cat $OUT/configs.jsonl | cut -c1-500 | tee -a $OUT/progress.txt
KMGPU_PREFER_DELTA=0 timeout 600 python tools/bench_configs.py C2 C4S > $OUT/configs_grouped.jsonl 2>> $OUT/configs.err
cat $OUT/configs_grouped.jsonl | cut -c1-400 | tee -a $OUT/progress.txt
echo "== sharded bench, one GPU (whole tables on it)" | tee -a $OUT/progress.txt
timeout 600 python bench.py --mode sharded --steps 4 --warmup 2 > $OUT/bench_sharded_n1.json 2> $OUT/bench_sharded_n1.err; echo "sharded rc=$?" | tee -a $OUT/progress.txt
cut -c1-900 $OUT/bench_sharded_n1.json | tee -a $OUT/progress.txt
tail -3 $OUT/bench_sharded_n1.err | tee -a $OUT/progress.txt
echo "== full gpu test suite" | tee -a $OUT/progress.txt
timeout 3000 python -m pytest tests -q -x -m gpu > $OUT/tests_all.log 2>&1; echo "all tests rc=$?" | tee -a $OUT/progress.txt
tail -5 $OUT/tests_all.log | tee -a $OUT/progress.txt
echo "== ncu: fallback paths (atomic / red sector counters)" | tee -a $OUT/progress.txt
KMGPU_GROUP=0 KMGPU_BUCKETS=0 timeout 600 ncu --set full --clock-control none -k regex:"k_scatter|k_fold" --launch-skip 8 -c 4 -o $OUT/delta_full python bench.py --no-cpu --no-check --no-file --steps 1 --warmup 1 > $OUT/ncu_delta.log 2>&1; echo "ncu delta rc=$?" | tee -a $OUT/progress.txt
KMGPU_GROUP=0 KMGPU_DELTA=0 timeout 600 ncu --set full --clock-control none -k regex:"k_ingest" --launch-skip 2 -c 3 -o $OUT/cas_full python bench.py --no-cpu --no-check --no-file --steps 1 --warmup 1 --reads 500000 > $OUT/ncu_cas.log 2>&1; echo "ncu cas rc=$?" | tee -a $OUT/progress.txt
echo "== ncu: default path (bins + bucketize + apply), full" | tee -a $OUT/progress.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_hashbins|k_bucketize|k_apply" --launch-skip 12 -c 6 -o $OUT/default_full python bench.py --no-cpu --no-check --no-file --steps 2 --warmup 1 > $OUT/ncu_default.log 2>&1; echo "ncu default rc=$?" | tee -a $OUT/progress.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --no-cpu --no-check --no-file --steps 2 --warmup 1 > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?" | tee -a $OUT/progress.txt
ls -la $OUT | tee -a $OUT/progress.txt
