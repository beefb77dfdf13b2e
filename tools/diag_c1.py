import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from khmer_b200 import cabi
sizes = bench.primes_near_x(4, int(1e8))
for big in (True, False):
    sk = cabi.Sketch(cabi.BYTE, cabi.TWOBIT, 20, sizes)
    sk.set_use_bigcount(big)
    bts = []
    for b in range(2):
        buf, off, _ = bench.synth_batch(77 + b, 2_000_000)
        bts.append(cabi.Batch((buf, off), 20))
    sk.consume_batch(bts[0]); sk.reset()
    for r in range(4):
        for i, bt in enumerate(bts):
            sk.profile_reset()
            t0 = time.perf_counter()
            n = sk.consume_batch(bt)
            dt = time.perf_counter() - t0
            ms, nl, al = sk.profile_get()
            t = sk.table(0)
            print("big=%s rep %d batch %d: %.1f ms (ingest kernels %.1f ms, %d launches) -> %.2f G/s  max=%d n255=%d bigmap=%d" % (
                big, r, i, dt * 1e3, ms, al, n / dt / 1e9, int(t.max()), int((t == 255).sum()), len(sk.bigcounts()[0])), flush=True)
    sk.close()
