#!/usr/bin/env python
"""Per-read medians over a device-resident batch on the headline table (used under ncu for the launch list of the query kernels)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from khmer_b200 import cabi  # noqa: E402

sizes = bench.primes_near_x(4, int(1e8))
sk = cabi.Sketch(cabi.BYTE, cabi.TWOBIT, 20, sizes)
buf, off, _ = bench.synth_batch(77, 1_000_000)
b = cabi.Batch((buf, off), 20)
sk.consume_batch(b)
for stats in (False, True):
    sk.batch_read_medians(b, stats=stats)
    sk.timer_start()
    med = sk.batch_read_medians(b, stats=stats)[0]
    ms = sk.timer_stop()
    print("stats=%s: %.2f ms, %.2f G k-mers/s, median of medians %d" % (stats, ms, len(med) * 131 / ms / 1e6, int(sorted(med)[len(med) // 2])))
