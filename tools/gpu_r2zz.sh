#!/bin/bash
# round-2 closing GPU session: the whole suite and the bench line on the final tree
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2zz
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
print('value %.2f e2e %.2f packed %.2f file %.2f nobig %.2f' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e_packed']['value']/1e9, (d['e2e_file']['value'] or 0)/1e9, d['value_bigcount_off']/1e9)); print(json.dumps(d['secondary'])); print(d['parity_check']['ok'], d['roofline']['frac'], d['roofline']['dram_frac'], d['cpu_baseline']['value'])" | tee -a $OUT/progress.txt
find gpurun_out -size +20M -delete
