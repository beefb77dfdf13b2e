#!/usr/bin/env python
"""Secondary measurements (not the bench line): the other BASELINE configs at their table sizes on ONE GPU, through the C ABI
with device-resident batches (ingest) and host buffers (queries, normalization).  One JSON object per configuration.

    python tools/bench_configs.py [--no-queries] [C1 C2 C3 C4 C5 NORM ...]      (default: all)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from khmer_b200 import cabi  # noqa: E402


NO_QUERIES = "--no-queries" in sys.argv


def run(name, storage, hashkind, k, x, n_reads=2_000_000, reps=2, bigcount=False, queries=True):
    queries = queries and not NO_QUERIES
    sizes = bench.primes_near_x(4, int(x))
    t0 = time.perf_counter()
    sk = cabi.Sketch(storage, hashkind, k, sizes)
    if bigcount:
        sk.set_use_bigcount(True)
    batches = []
    for b in range(2):
        buf, off, _ = bench.synth_batch(77 + b, n_reads)
        batches.append((cabi.Batch((buf, off), k), (buf, off)))
    sk.consume_batch(batches[0][0])       # warm-up (allocations), then reset
    sk.reset()
    sk.sync()
    out = {"config": name, "k": k, "tables": "4 x %.3g bins" % x, "reads_per_batch": n_reads, "setup_s": round(time.perf_counter() - t0, 2)}
    sk.profile_reset()
    sk.timer_start()
    kmers = 0
    for r in range(reps):
        for bt, _ in batches:
            kmers += sk.consume_batch(bt)
    ms = sk.timer_stop()
    kern_ms, _, launches = sk.profile_get()
    out["ingest_gkmers_per_s"] = kmers / ms / 1e6
    out["ingest_ms_per_batch"] = ms / (reps * len(batches))
    out["kernel_group_ms_per_batch"] = kern_ms / (reps * len(batches))
    out["launches_per_batch"] = launches / (reps * len(batches))
    out["frac_of_random_rmw_roofline"] = kmers * 4 * 64 / (kern_ms * 1e-3) / 1e9 / bench.measured_peak()[0]
    if queries:
        buf, off = batches[0][1]
        nq = 200_000
        q = (buf[: nq * 150], off[: nq + 1])
        sk.read_medians(q)
        t0 = time.perf_counter()
        med, avg, sd, nk = sk.read_medians(q)
        dt = time.perf_counter() - t0
        out["median_reads_per_s"] = nq / dt
        out["median_gkmers_per_s"] = float(nk.sum()) / dt / 1e9
        if storage != cabi.BIT:
            tracking = cabi.Sketch(cabi.BIT, hashkind, k, sizes)
            t0 = time.perf_counter()
            hist = sk.abundance_distribution((buf, off), tracking)
            dt = time.perf_counter() - t0
            out["abundance_dist_gkmers_per_s"] = n_reads * (150 - k + 1) / dt / 1e9
            out["abundance_dist_distinct"] = int(hist.sum())
            tracking.close()
    print(json.dumps(out), flush=True)
    sk.close()


def run_normalize(n_reads=4_000_000, x=2e9, genome=20_000_000):
    """config C3's loop (normalize-by-median -k 20 -C 20) at C3's table size: 30x of a 20 Mbp genome in stream order."""
    sizes = bench.primes_near_x(4, int(x))
    sk = cabi.Sketch(cabi.BYTE, cabi.TWOBIT, 20, sizes)
    sk.set_use_bigcount(True)
    rng = np.random.default_rng(5)
    g = rng.integers(0, 4, genome, dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    starts = rng.integers(0, genome - 150, n_reads)
    buf = lut[g[starts[:, None] + np.arange(150)[None, :]]].reshape(-1)
    off = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(150)
    sk.normalize_batch((buf[: 150 * 20000], off[: 20001]), 20)     # warm-up
    sk.reset()
    t0 = time.perf_counter()
    keep, kmers = sk.normalize_batch((buf, off), 20)
    dt = time.perf_counter() - t0
    print(json.dumps({"config": "C3 normalize-by-median k=20 C=20, 4 x %.3g bytes" % x, "reads": n_reads, "coverage": n_reads * 150 / genome,
                      "kept": int(keep.sum()), "reads_per_s": n_reads / dt, "kmers_looked_up_per_s": n_reads * 131 / dt, "seconds": dt}), flush=True)
    sk.close()


CONFIGS = {
    "C1": lambda: run("C1 Countgraph k=20 x=1e8 bigcount", cabi.BYTE, cabi.TWOBIT, 20, 1e8, bigcount=True),
    "C1G": lambda: run("C1 through the fused grouping kernel (KMGPU_PREFER_BINS=0 must be set)", cabi.BYTE, cabi.TWOBIT, 20, 1e8, bigcount=True, queries=False),
    "C2": lambda: run("C2 Nodegraph k=32 x=1e9 bits", cabi.BIT, cabi.TWOBIT, 32, 1e9),
    "C3": lambda: run("C3 Countgraph k=20 x=2e9 bytes (8 GB)", cabi.BYTE, cabi.TWOBIT, 20, 2e9, bigcount=True),
    "C4": lambda: run("C4 SmallCountgraph k=31 x=8e9 nibbles (16 GB)", cabi.NIBBLE, cabi.TWOBIT, 31, 8e9),
    "C3L": lambda: run("C3 Countgraph k=20 x=2e9 bytes (8 GB), 5 M reads per batch (one chunk)", cabi.BYTE, cabi.TWOBIT, 20, 2e9, n_reads=5_000_000, bigcount=True, queries=False),
    "C4L": lambda: run("C4 SmallCountgraph k=31 x=8e9 nibbles (16 GB), 5 M reads per batch (one chunk)", cabi.NIBBLE, cabi.TWOBIT, 31, 8e9, n_reads=5_000_000, queries=False),
    "C4S": lambda: run("C4-shape SmallCountgraph k=31 x=4e8 nibbles", cabi.NIBBLE, cabi.TWOBIT, 31, 4e8),
    "C5M": lambda: run("C5 hash path, Counttable k=40 (Murmur) x=1e8", cabi.BYTE, cabi.MURMUR, 40, 1e8, bigcount=False),
    "C5": lambda: run("C5 single-GPU slice, Counttable k=40 (Murmur) 4 x 3.2e10 bytes (128 GB)", cabi.BYTE, cabi.MURMUR, 40, 3.2e10, queries=False),
    "NORM": run_normalize,
}

if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["C1", "C2", "C3", "C4", "C4S", "C5M", "C5", "NORM"]
    for n in names:
        try:
            CONFIGS[n]()
        except Exception as e:   # a configuration that does not fit must not hide the others
            print(json.dumps({"config": n, "error": str(e)[:300]}), flush=True)
