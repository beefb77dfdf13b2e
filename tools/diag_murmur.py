#!/usr/bin/env python
"""Diagnostic: ingest rate of a Murmur-hashed Counttable in the bigcount-heavy regime of tools/bench_configs.py."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from khmer_b200 import cabi

def run(hk, k, reps):
    sizes = bench.primes_near_x(4, int(1e8))
    sk = cabi.Sketch(cabi.BYTE, hk, k, sizes)
    sk.set_use_bigcount(True)
    batches = [cabi.Batch(bench.synth_batch(77 + b, 2_000_000)[:2], k) for b in range(2)]
    sk.consume_batch(batches[0]); sk.reset()
    for rep in range(reps):
        for bt in batches:
            sk.timer_start(); n = sk.consume_batch(bt); ms = sk.timer_stop()
            print(json.dumps({"hash": hk, "k": k, "rep": rep, "kmers": n, "ms": ms, "gkmers_per_s": n / ms / 1e6, "profile": sk.profile_get(),
                              "bigcounts": len(sk.bigcounts()[0])}), flush=True)
    sk.close()

run(cabi.MURMUR, 40, int(sys.argv[1]) if len(sys.argv) > 1 else 3)
