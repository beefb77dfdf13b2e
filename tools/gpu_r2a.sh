#!/bin/bash
# round-2 GPU session A: sanity (sanitizer on a small case), grouped-path parity tests, bench variants, ncu.
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
OUT=gpurun_out/r2a
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke" | tee $OUT/progress.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/progress.txt
echo "== small grouped case" | tee -a $OUT/progress.txt
KMGPU_TEST_CLS=Countgraph KMGPU_GROUP_MIN_BUCKETS=0 KMGPU_CHUNK_BASES=16384 KMGPU_PART_BASES=1000 timeout 300 \
   python -m pytest -q -x -m gpu tests/test_gpu_parity.py -k test_gpu_vs_oracle_random_inner > $OUT/sanitizer.log 2>&1; echo "sanitizer rc=$?" | tee -a $OUT/progress.txt
echo "== grouped-path tests" | tee -a $OUT/progress.txt
timeout 1500 python -m pytest -q -x -m gpu tests/test_gpu_parity.py -k "group or golden_C1 or many_buckets or (golden and 25k)" > $OUT/tests_group.log 2>&1; echo "group tests rc=$?" | tee -a $OUT/progress.txt
tail -5 $OUT/tests_group.log | tee -a $OUT/progress.txt
echo "== normalize tests" | tee -a $OUT/progress.txt
timeout 1200 python -m pytest -q -x -s -m gpu tests/test_gpu_normalize.py > $OUT/tests_norm.log 2>&1; echo "normalize tests rc=$?" | tee -a $OUT/progress.txt
tail -5 $OUT/tests_norm.log | tee -a $OUT/progress.txt
echo "== bench default" | tee -a $OUT/progress.txt
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?" | tee -a $OUT/progress.txt
cat $OUT/bench_default.json | cut -c1-600 | tee -a $OUT/progress.txt
for v in "KMGPU_PART_T=16384" "KMGPU_GROUP=0"; do
  tag=$(echo $v | tr '=' '_')
  env $v timeout 600 python bench.py --no-cpu --no-check --steps 8 --warmup 4 > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err; echo "bench $v rc=$?" | tee -a $OUT/progress.txt
  cut -c1-300 $OUT/bench_$tag.json | tee -a $OUT/progress.txt
done
echo "== full gpu test suite" | tee -a $OUT/progress.txt
timeout 2400 python -m pytest tests -q -x -m gpu > $OUT/tests_all.log 2>&1; echo "all tests rc=$?" | tee -a $OUT/progress.txt
tail -5 $OUT/tests_all.log | tee -a $OUT/progress.txt
echo "== ncu launch list" | tee -a $OUT/progress.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv python bench.py --no-cpu --no-check --steps 2 --warmup 1 > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?" | tee -a $OUT/progress.txt
echo "== ncu full (k_part, k_apply2)" | tee -a $OUT/progress.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_part|k_apply2" --launch-skip 16 -c 6 -o $OUT/ingest_full python bench.py --no-cpu --no-check --steps 2 --warmup 1 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $OUT/progress.txt
ls -la $OUT | tee -a $OUT/progress.txt
