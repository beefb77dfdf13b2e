#!/bin/bash
# round-2 GPU session C: test suite, bench (bins path default), grouped-path variants, BASELINE table sizes
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2c
mkdir -p $OUT
nproc > $OUT/nproc.txt; free -g >> $OUT/nproc.txt
echo "== smoke" | tee $OUT/progress.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/progress.txt
echo "== new/changed tests first" | tee -a $OUT/progress.txt
timeout 2400 python -m pytest -q -x -s -m gpu tests/test_gpu_round2.py tests/test_gpu_normalize.py tests/test_sharded_gpu.py tests/test_host_gpu.py > $OUT/tests_new.log 2>&1; echo "new tests rc=$?" | tee -a $OUT/progress.txt
tail -6 $OUT/tests_new.log | tee -a $OUT/progress.txt
grep "normalize 1M" $OUT/tests_new.log | tee -a $OUT/progress.txt
echo "== bench default (full line)" | tee -a $OUT/progress.txt
timeout 1200 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?" | tee -a $OUT/progress.txt
cut -c1-1500 $OUT/bench_default.json | tee -a $OUT/progress.txt
for v in "KMGPU_PREFER_BINS=0 KMGPU_APPLY_GATE=1" "KMGPU_PREFER_BINS=0 KMGPU_APPLY_GATE=0" "KMGPU_PREFER_BINS=0 KMGPU_APPLY_GATE=0 KMGPU_PART_T=16384"; do
  tag=$(echo $v | tr '= ' '__')
  env $v timeout 600 python bench.py --no-cpu --no-check --no-file --steps 8 --warmup 4 > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err; echo "bench $v rc=$?" | tee -a $OUT/progress.txt
  python -c "import json,sys; d=json.load(open('$OUT/bench_$tag.json')); print('  value %.2f G  e2e %.2f G  ms/step %.2f  nobig %.2f' % (d['value']/1e9, d['e2e']['value']/1e9, d['ms_per_step'], d.get('value_bigcount_off',0)/1e9))" 2>&1 | tee -a $OUT/progress.txt
done
echo "== BASELINE table sizes" | tee -a $OUT/progress.txt
timeout 1500 python tools/bench_configs.py C2 C3 C4 C4S C5M NORM C5 > $OUT/configs.jsonl 2> $OUT/configs.err; echo "configs rc=$?" | tee -a $OUT/progress.txt
cat $OUT/configs.jsonl | cut -c1-700 | tee -a $OUT/progress.txt
echo "== full gpu test suite" | tee -a $OUT/progress.txt
timeout 3000 python -m pytest tests -q -x -m gpu > $OUT/tests_all.log 2>&1; echo "all tests rc=$?" | tee -a $OUT/progress.txt
tail -5 $OUT/tests_all.log | tee -a $OUT/progress.txt
echo "== ncu launch list (default bench)" | tee -a $OUT/progress.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --no-cpu --no-check --no-file --steps 2 --warmup 1 > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?" | tee -a $OUT/progress.txt
echo "== ncu full C3 two-level kernels" | tee -a $OUT/progress.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_part|k_apply" --launch-skip 6 -c 6 -o $OUT/c3_full python tools/bench_configs.py C3 > $OUT/ncu_c3.log 2>&1; echo "ncu c3 rc=$?" | tee -a $OUT/progress.txt
ls -la $OUT | tee -a $OUT/progress.txt
