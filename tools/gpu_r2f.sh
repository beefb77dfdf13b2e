#!/bin/bash
# round-2 GPU session F: k_apply2 cut across 2/4 CTAs per bucket, sparse threshold, query kernels
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2f
mkdir -p $OUT
echo "== tests of the changed paths" | tee $OUT/progress.txt
timeout 1500 python -m pytest -q -x -m gpu tests/test_gpu_parity.py -k "group or many_buckets" > $OUT/tests_group.log 2>&1; echo "group tests rc=$?" | tee -a $OUT/progress.txt
tail -4 $OUT/tests_group.log | tee -a $OUT/progress.txt
for sp in 1 2 4; do
  KMGPU_APPLY_SPLIT=$sp KMGPU_SPARSE=0 timeout 900 python -m pytest -q -x -m gpu tests/test_gpu_parity.py -k "test_group_path_many_buckets and (dense or sparse) or golden_cases_grouped" > $OUT/tests_split$sp.log 2>&1; echo "split $sp tests rc=$?" | tee -a $OUT/progress.txt
  tail -2 $OUT/tests_split$sp.log | tee -a $OUT/progress.txt
done
timeout 900 python -m pytest -q -x -m gpu tests/test_gpu_benchscale.py tests/test_gpu_round2.py -k "benchscale or batch_read" > $OUT/tests_misc.log 2>&1; echo "misc tests rc=$?" | tee -a $OUT/progress.txt
tail -2 $OUT/tests_misc.log | tee -a $OUT/progress.txt
echo "== variants on the large tables (ingest only)" | tee -a $OUT/progress.txt
run() { echo "-- $1" | tee -a $OUT/progress.txt; env $1 timeout 600 python tools/bench_configs.py --no-queries $2 2>> $OUT/configs.err | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print(d['config'][:40], '%.2f G k-mers/s, %.1f ms/batch' % (d.get('ingest_gkmers_per_s', 0), d.get('ingest_ms_per_batch', 0)) if 'error' not in d else d['error'])" | tee -a $OUT/progress.txt; }
run "KMGPU_APPLY_SPLIT=0" "C3 C4 C5 C4S"
run "KMGPU_APPLY_SPLIT=1" "C3 C4S"
run "KMGPU_APPLY_SPLIT=2" "C3 C4S"
run "KMGPU_APPLY_SPLIT=4" "C3 C4S"
run "KMGPU_SPARSE=0 KMGPU_APPLY_SPLIT=4" "C4 C5"
run "KMGPU_SPARSE=0 KMGPU_APPLY_SPLIT=2" "C4"
run "KMGPU_SPARSE_MAX_LOAD=256" "C4 C5"
echo "== launch lists" | tee -a $OUT/progress.txt
for c in C3 C4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$c.csv python tools/bench_configs.py --no-queries $c > $OUT/ncu_$c.log 2>&1; echo "ncu $c rc=$?" | tee -a $OUT/progress.txt
done
timeout 300 python tools/medians_probe.py 2>&1 | tee -a $OUT/progress.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_medians.csv python tools/medians_probe.py > $OUT/ncu_medians.log 2>&1; echo "ncu medians rc=$?" | tee -a $OUT/progress.txt
python - <<'PY' | tee -a $OUT/progress.txt
import csv, collections, glob
for f in sorted(glob.glob('gpurun_out/r2f/launches_*.csv')):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        k = r[4].split('(')[0][:60]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[-1].replace(',', ''))
    print(f)
    for k, (n, t) in agg.items(): print('   %-60s x%-3d %10.1f us total' % (k, n, t / 1e3 if t > 1e5 else t))
PY
echo "== bench default" | tee -a $OUT/progress.txt
timeout 1200 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?" | tee -a $OUT/progress.txt
python -c "
import json; d=json.load(open('$OUT/bench_default.json'))
print('value %.2f e2e %.2f file %.2f' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e_file']['value']/1e9)); print(json.dumps(d['secondary'])); print(d['parity_check']['ok'])" | tee -a $OUT/progress.txt
find gpurun_out -size +20M -delete
du -sm gpurun_out | tee -a $OUT/progress.txt
