#!/bin/bash
# round-2 GPU session E: device resolver of the normalization, regrouping fixes, diagnostics of the large-table configs,
# ncu summaries made ON the box (the reports themselves are too large to bring back)
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2e
mkdir -p $OUT
echo "== smoke" | tee $OUT/progress.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/progress.txt
echo "== changed tests first" | tee -a $OUT/progress.txt
KMGPU_DEBUG=1 timeout 1500 python -m pytest -q -x -s -m gpu tests/test_gpu_normalize.py tests/test_gpu_round2.py tests/test_sharded_gpu.py > $OUT/tests_new.log 2>&1; echo "new tests rc=$?" | tee -a $OUT/progress.txt
tail -6 $OUT/tests_new.log | tee -a $OUT/progress.txt
grep -E "normalize 1M|normalize_batch" $OUT/tests_new.log | tail -8 | tee -a $OUT/progress.txt
timeout 1500 python -m pytest -q -x -m gpu tests/test_gpu_parity.py -k "group or many_buckets" > $OUT/tests_group.log 2>&1; echo "group tests rc=$?" | tee -a $OUT/progress.txt
tail -4 $OUT/tests_group.log | tee -a $OUT/progress.txt
echo "== BASELINE table sizes (debug notes on stderr: regrouping runs, normalization windows)" | tee -a $OUT/progress.txt
KMGPU_DEBUG=1 timeout 1500 python tools/bench_configs.py C3 C4 C4S C5 NORM > $OUT/configs.jsonl 2> $OUT/configs.err; echo "configs rc=$?" | tee -a $OUT/progress.txt
cut -c1-420 $OUT/configs.jsonl | tee -a $OUT/progress.txt
grep -c "regrouping" $OUT/configs.err | tee -a $OUT/progress.txt
grep -E "regrouping|normalize_batch" $OUT/configs.err | sort | uniq -c | sort -rn | head -12 | tee -a $OUT/progress.txt
echo "== sharded bench, one GPU" | tee -a $OUT/progress.txt
KMGPU_DEBUG=1 timeout 600 python bench.py --mode sharded --steps 4 --warmup 2 > $OUT/bench_sharded_n1.json 2> $OUT/bench_sharded_n1.err; echo "sharded rc=$?" | tee -a $OUT/progress.txt
cut -c1-1000 $OUT/bench_sharded_n1.json | tee -a $OUT/progress.txt
grep -c "regroup" $OUT/bench_sharded_n1.err | tee -a $OUT/progress.txt
echo "== bench default (full line)" | tee -a $OUT/progress.txt
timeout 1200 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?" | tee -a $OUT/progress.txt
python -c "
import json; d=json.load(open('$OUT/bench_default.json'))
print('value %.2f e2e %.2f file %.2f' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e_file']['value']/1e9)); print(json.dumps(d['secondary'])); print(d['parity_check']['ok'])" | tee -a $OUT/progress.txt
echo "== ncu (summaries made here, reports deleted)" | tee -a $OUT/progress.txt
summ() { python tools/ncu_summary.py $OUT/$1.ncu-rep $OUT/$1_summary.json > /dev/null 2>&1; echo "summary $1 rc=$?" | tee -a $OUT/progress.txt; rm -f $OUT/$1.ncu-rep; }
KMGPU_GROUP=0 KMGPU_BUCKETS=0 timeout 600 ncu --set full --clock-control none -k regex:"k_scatter|k_fold" --launch-skip 8 -c 4 -o $OUT/delta_full python bench.py --no-cpu --no-check --no-file --steps 1 --warmup 1 > $OUT/ncu_delta.log 2>&1; echo "ncu delta rc=$?" | tee -a $OUT/progress.txt
summ delta_full
KMGPU_GROUP=0 KMGPU_DELTA=0 KMGPU_BUCKETS=0 timeout 600 ncu --set full --clock-control none -k regex:"k_ingest" --launch-skip 2 -c 3 -o $OUT/cas_full python bench.py --no-cpu --no-check --no-file --steps 1 --warmup 1 --reads 500000 > $OUT/ncu_cas.log 2>&1; echo "ncu cas rc=$?" | tee -a $OUT/progress.txt
summ cas_full
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_hashbins|k_bucketize|k_apply|k_popc" --launch-skip 16 -c 4 -o $OUT/default_full python bench.py --no-cpu --no-check --no-file --steps 2 --warmup 1 > $OUT/ncu_default.log 2>&1; echo "ncu default rc=$?" | tee -a $OUT/progress.txt
summ default_full
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --no-cpu --no-check --no-file --steps 2 --warmup 1 > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?" | tee -a $OUT/progress.txt
find gpurun_out -size +20M -delete
du -sm gpurun_out | tee -a $OUT/progress.txt
ls -la $OUT | tee -a $OUT/progress.txt
