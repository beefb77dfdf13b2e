#!/usr/bin/env python
"""Secondary measurements (not the bench line): the other BASELINE configs at single-GPU scale, through the C ABI
with device-resident batches.  Prints one JSON object per configuration."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from khmer_b200 import cabi  # noqa: E402


def run(name, storage, hashkind, k, x, n_reads=2_000_000, reps=3, bigcount=False):
    sizes = bench.primes_near_x(4, int(x))
    sk = cabi.Sketch(storage, hashkind, k, sizes)
    if bigcount:
        sk.set_use_bigcount(True)
    batches = []
    for b in range(2):
        buf, off, _ = bench.synth_batch(77 + b, n_reads)
        batches.append((cabi.Batch((buf, off), k), (buf, off)))
    sk.consume_batch(batches[0][0])       # warm-up (allocations), then reset
    sk.reset()
    out = {"config": name, "k": k, "tables": "4 x %.3g bins" % x, "reads_per_batch": n_reads}
    sk.timer_start()
    kmers = 0
    for r in range(reps):
        for bt, _ in batches:
            kmers += sk.consume_batch(bt)
    ms = sk.timer_stop()
    out["ingest_gkmers_per_s"] = kmers / ms / 1e6
    # queries on the loaded sketch: medians of 200k reads, per-k-mer counts
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
    print(json.dumps(out), flush=True)
    sk.close()


if __name__ == "__main__":
    run("C1 Countgraph k=20 x=1e8 bigcount", cabi.BYTE, cabi.TWOBIT, 20, 1e8, bigcount=True)
    run("C2 Nodegraph k=32 x=1e9 bits", cabi.BIT, cabi.TWOBIT, 32, 1e9)
    run("C4-shape SmallCountgraph k=31 x=4e8 nibbles", cabi.NIBBLE, cabi.TWOBIT, 31, 4e8)
    run("C5-shape Counttable k=40 (Murmur) x=1e8", cabi.BYTE, cabi.MURMUR, 40, 1e8, bigcount=True)
    run("C3-shape Countgraph k=20 x=2e9 (8 GB, HBM-resident path)", cabi.BYTE, cabi.TWOBIT, 20, 2e9, bigcount=True)
