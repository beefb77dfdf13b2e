#!/bin/bash
# round-2 GPU session L: sparse apply with batched loads / swaps and one counter update per CTA
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2l
mkdir -p $OUT
echo "== tests of the grouped path" | tee $OUT/progress.txt
timeout 1500 python -m pytest -q -x -m gpu tests/test_gpu_parity.py tests/test_sharded_gpu.py tests/test_gpu_round2.py -k "group or many_buckets or sharded or first_touch or more_than_ten" > $OUT/tests_group.log 2>&1; echo "group tests rc=$?" | tee -a $OUT/progress.txt
tail -4 $OUT/tests_group.log | cut -c1-300 | tee -a $OUT/progress.txt
echo "== large tables" | tee -a $OUT/progress.txt
timeout 900 python tools/bench_configs.py --no-queries C5 C4 C3 > $OUT/configs.jsonl 2> $OUT/configs.err; echo "configs rc=$?" | tee -a $OUT/progress.txt
cut -c1-330 $OUT/configs.jsonl | tee -a $OUT/progress.txt
timeout 600 python bench.py --mode sharded --steps 6 --warmup 2 > $OUT/bench_sharded_n1.json 2> $OUT/bench_sharded_n1.err; echo "sharded n1 rc=$?" | tee -a $OUT/progress.txt
tail -1 $OUT/bench_sharded_n1.json | cut -c1-1400 | tee -a $OUT/progress.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_C5.csv python tools/bench_configs.py --no-queries C5 > $OUT/ncu_C5.log 2>&1; echo "ncu C5 rc=$?" | tee -a $OUT/progress.txt
python - <<'PY' | tee -a $OUT/progress.txt
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2l/launches_C5.csv')) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r[4].split('(')[0][:60], [0, 0.0]); a[0] += 1; a[1] += float(r[-1].replace(',', ''))
for k, (n, t) in agg.items(): print('   %-60s x%-3d %10.3f ms total' % (k, n, t / 1e6))
PY
find gpurun_out -size +20M -delete
