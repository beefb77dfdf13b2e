#!/bin/bash
# round-2 final GPU session: the whole suite, smoke, both bench arms, every configuration, launch list of the default run
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2z
mkdir -p $OUT
echo "== smoke" | tee $OUT/progress.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/progress.txt
echo "== full gpu test suite" | tee -a $OUT/progress.txt
timeout 3000 python -m pytest tests -q -x -m gpu > $OUT/tests_all.log 2>&1; echo "all tests rc=$?" | tee -a $OUT/progress.txt
tail -5 $OUT/tests_all.log | cut -c1-300 | tee -a $OUT/progress.txt
echo "== bench (ours)" | tee -a $OUT/progress.txt
timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/progress.txt
python -c "
import json; d=json.load(open('$OUT/bench.json'))
print('value %.2f e2e %.2f packed %.2f file %.2f nobig %.2f' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e_packed']['value']/1e9, (d['e2e_file']['value'] or 0)/1e9, d['value_bigcount_off']/1e9)); print(json.dumps(d['e2e_file'].get('compressed'))); print(json.dumps(d['secondary'])); print(d['parity_check']['ok'], d['roofline']['frac'], d['roofline']['dram_frac'], d['cpu_baseline']['value'])" | tee -a $OUT/progress.txt
tail -4 $OUT/bench.err | cut -c1-300 | tee -a $OUT/progress.txt
echo "== bench (reference arm)" | tee -a $OUT/progress.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference rc=$?" | tee -a $OUT/progress.txt
cut -c1-600 $OUT/bench_reference.json | tee -a $OUT/progress.txt
echo "== every configuration" | tee -a $OUT/progress.txt
timeout 1500 python tools/bench_configs.py C2 C3 C4 C4S C5M C5 NORM > $OUT/configs.jsonl 2> $OUT/configs.err; echo "configs rc=$?" | tee -a $OUT/progress.txt
cut -c1-420 $OUT/configs.jsonl | tee -a $OUT/progress.txt
echo "== launch list of the default run" | tee -a $OUT/progress.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --no-cpu --no-check --no-file --steps 2 --warmup 1 > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?" | tee -a $OUT/progress.txt
find gpurun_out -size +20M -delete
du -sm gpurun_out | tee -a $OUT/progress.txt
