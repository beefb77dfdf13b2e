#!/bin/bash
# round-2 GPU session B: persistent grouping + new features
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2b
mkdir -p $OUT
echo "== smoke" | tee $OUT/progress.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/progress.txt
echo "== quick grouped tests" | tee -a $OUT/progress.txt
timeout 900 python -m pytest -q -x -m gpu tests/test_gpu_parity.py -k "golden_C1 or (golden and 25k) or many_buckets" > $OUT/tests_quick.log 2>&1; echo "quick rc=$?" | tee -a $OUT/progress.txt
tail -4 $OUT/tests_quick.log | tee -a $OUT/progress.txt
for v in "KMGPU_X=0" "KMGPU_PART_T=16384" "KMGPU_PERSIST=0" "KMGPU_PERSIST=0 KMGPU_PART_T=16384" "KMGPU_GROUP=0"; do
  tag=$(echo $v | tr '= ' '__')
  env $v timeout 600 python bench.py --no-cpu --no-check --steps 8 --warmup 4 > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err; echo "bench $v rc=$?" | tee -a $OUT/progress.txt
  python -c "import json,sys; d=json.load(open('$OUT/bench_$tag.json')); print('  value %.2f G  e2e %.2f G  ms/step %.2f  nobig %.2f' % (d['value']/1e9, d['e2e']['value']/1e9, d['ms_per_step'], d.get('value_bigcount_off',0)/1e9))" 2>&1 | tee -a $OUT/progress.txt
done
echo "== new feature tests" | tee -a $OUT/progress.txt
timeout 1500 python -m pytest -q -x -s -m gpu tests/test_gpu_round2.py tests/test_gpu_normalize.py tests/test_sharded_gpu.py > $OUT/tests_new.log 2>&1; echo "new tests rc=$?" | tee -a $OUT/progress.txt
tail -6 $OUT/tests_new.log | tee -a $OUT/progress.txt
echo "== full gpu test suite" | tee -a $OUT/progress.txt
timeout 2400 python -m pytest tests -q -x -m gpu > $OUT/tests_all.log 2>&1; echo "all tests rc=$?" | tee -a $OUT/progress.txt
tail -5 $OUT/tests_all.log | tee -a $OUT/progress.txt
echo "== bench full default" | tee -a $OUT/progress.txt
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?" | tee -a $OUT/progress.txt
echo "== ncu launch list" | tee -a $OUT/progress.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv python bench.py --no-cpu --no-check --steps 2 --warmup 1 > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?" | tee -a $OUT/progress.txt
echo "== ncu full" | tee -a $OUT/progress.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_part|k_apply2" --launch-skip 8 -c 4 -o $OUT/ingest_full python bench.py --no-cpu --no-check --steps 2 --warmup 1 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $OUT/progress.txt
ls -la $OUT | tee -a $OUT/progress.txt
