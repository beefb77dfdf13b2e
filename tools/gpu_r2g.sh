#!/bin/bash
# round-2 GPU session G: HLL registers, count passes, compressed input, larger chunks on the multi-GB tables, full suite
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2g
mkdir -p $OUT
echo "== new tests" | tee $OUT/progress.txt
timeout 1200 python -m pytest -q -x -m gpu tests/test_gpu_round2.py -k "hll or batch_read" > $OUT/tests_new.log 2>&1; echo "hll tests rc=$?" | tee -a $OUT/progress.txt
tail -3 $OUT/tests_new.log | tee -a $OUT/progress.txt
timeout 1200 python -m pytest -q -x -m gpu tests/test_gpu_parity.py -k "count-passes or turns" > $OUT/tests_cp.log 2>&1; echo "count-passes tests rc=$?" | tee -a $OUT/progress.txt
tail -3 $OUT/tests_cp.log | tee -a $OUT/progress.txt
echo "== bench default" | tee -a $OUT/progress.txt
timeout 1200 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?" | tee -a $OUT/progress.txt
python -c "
import json; d=json.load(open('$OUT/bench_default.json'))
print('value %.2f e2e %.2f file %.2f' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e_file']['value']/1e9)); print(json.dumps(d['e2e_file'].get('compressed'))); print(json.dumps(d['secondary'])); print(d['parity_check']['ok'])" | tee -a $OUT/progress.txt
tail -3 $OUT/bench_default.err | tee -a $OUT/progress.txt
timeout 300 python tools/medians_probe.py 2>&1 | tee -a $OUT/progress.txt
echo "== larger chunks" | tee -a $OUT/progress.txt
timeout 900 python tools/bench_configs.py --no-queries C3L C4L > $OUT/configs_large.jsonl 2> $OUT/configs_large.err; echo "configs rc=$?" | tee -a $OUT/progress.txt
cut -c1-330 $OUT/configs_large.jsonl | tee -a $OUT/progress.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_C5.csv python tools/bench_configs.py --no-queries C5 > $OUT/ncu_C5.log 2>&1; echo "ncu C5 rc=$?" | tee -a $OUT/progress.txt
python - <<'PY' | tee -a $OUT/progress.txt
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2g/launches_C5.csv')) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r[4].split('(')[0][:60], [0, 0.0]); a[0] += 1; a[1] += float(r[-1].replace(',', ''))
for k, (n, t) in agg.items(): print('   %-60s x%-3d %10.3f ms total' % (k, n, t / 1e6))
PY
echo "== ncu: atomic / reduction sectors of the fallback kernels" | tee -a $OUT/progress.txt
M=lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
KMGPU_GROUP=0 KMGPU_BUCKETS=0 timeout 600 ncu --metrics $M --clock-control none -k regex:"k_scatter|k_fold" --launch-skip 8 -c 4 --csv --log-file $OUT/delta_atomics.csv python bench.py --no-cpu --no-check --no-file --steps 1 --warmup 1 > $OUT/ncu_delta.log 2>&1; echo "ncu delta rc=$?" | tee -a $OUT/progress.txt
KMGPU_GROUP=0 KMGPU_DELTA=0 KMGPU_BUCKETS=0 timeout 600 ncu --metrics $M --clock-control none -k regex:"k_ingest" --launch-skip 2 -c 3 --csv --log-file $OUT/cas_atomics.csv python bench.py --no-cpu --no-check --no-file --steps 1 --warmup 1 --reads 500000 > $OUT/ncu_cas.log 2>&1; echo "ncu cas rc=$?" | tee -a $OUT/progress.txt
KMGPU_PREFER_BINS=0 timeout 600 ncu --metrics $M --clock-control none -k regex:"k_part|k_apply2" --launch-skip 8 -c 4 --csv --log-file $OUT/grouped_atomics.csv python bench.py --no-cpu --no-check --no-file --steps 1 --warmup 1 > $OUT/ncu_grouped.log 2>&1; echo "ncu grouped rc=$?" | tee -a $OUT/progress.txt
echo "== full gpu test suite" | tee -a $OUT/progress.txt
timeout 3000 python -m pytest tests -q -x -m gpu > $OUT/tests_all.log 2>&1; echo "all tests rc=$?" | tee -a $OUT/progress.txt
tail -5 $OUT/tests_all.log | tee -a $OUT/progress.txt
find gpurun_out -size +20M -delete
du -sm gpurun_out | tee -a $OUT/progress.txt
