#!/bin/bash
# round-2: ncu --set full of the grouped path on the final tree (C3: two levels + k_apply2 split 2; C5: k_apply_sparse)
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2x
mkdir -p $OUT
summ() { python tools/ncu_summary.py $OUT/$1.ncu-rep $OUT/$1_summary.json > /dev/null 2>&1; echo "summary $1 rc=$?" | tee -a $OUT/progress.txt; rm -f $OUT/$1.ncu-rep; }
echo "== C3" | tee $OUT/progress.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_part|k_apply" --launch-skip 4 -c 4 -o $OUT/c3_full python tools/bench_configs.py --no-queries C3 > $OUT/ncu_c3.log 2>&1; echo "ncu c3 rc=$?" | tee -a $OUT/progress.txt
summ c3_full
echo "== C5" | tee -a $OUT/progress.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_hash64|k_part|k_apply" --launch-skip 5 -c 5 -o $OUT/c5_full python tools/bench_configs.py --no-queries C5 > $OUT/ncu_c5.log 2>&1; echo "ncu c5 rc=$?" | tee -a $OUT/progress.txt
summ c5_full
find gpurun_out -size +20M -delete
