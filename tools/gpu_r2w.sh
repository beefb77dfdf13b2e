#!/bin/bash
# round-2: k_part shared memory sized by the partitions an instantiation sorts into (two CTAs per SM for the super-bucket level)
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2w
mkdir -p $OUT
timeout 200 python -m pytest -q -x -m gpu tests/test_gpu_parity.py -k "test_group_path_many_buckets or golden_cases_grouped" > $OUT/tests.log 2>&1; echo "tests rc=$?" | tee $OUT/progress.txt
tail -2 $OUT/tests.log | cut -c1-200 | tee -a $OUT/progress.txt
timeout 150 python tools/bench_configs.py --no-queries C3 C5 2> $OUT/configs.err | cut -c1-300 | tee -a $OUT/progress.txt
