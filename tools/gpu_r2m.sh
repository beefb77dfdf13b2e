#!/bin/bash
# round-2 GPU session M: sparse apply threshold on C4
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2m
mkdir -p $OUT
echo "== C4 with the sparse apply taken up to 2048 expected records per bucket" | tee $OUT/progress.txt
for v in 2048 1024; do
KMGPU_SPARSE_MAX_LOAD=$v timeout 600 python tools/bench_configs.py --no-queries C4 2>> $OUT/configs.err | cut -c1-330 | tee -a $OUT/progress.txt
done
