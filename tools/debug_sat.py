import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from khmer_b200 import cabi
import bench
sizes = bench.primes_near_x(4, int(1e8))
sk = cabi.Sketch(cabi.BYTE, cabi.TWOBIT, 20, sizes)
sk.set_use_bigcount(True)
for b in range(4):
    buf, off, _ = bench.synth_batch(1000 + b, 2_000_000)
    n = sk.consume_reads((buf, off))
    t0 = sk.table(0)
    print("batch", b, "kmers", n, "max", int(t0.max()), "n>=200", int((t0 >= 200).sum()), "argmax", int(t0.argmax()), "stats", sk.stats(), "big", len(sk.bigcounts()[0]))
    if t0.max() >= 200:
        h = np.bincount(t0, minlength=256)
        print("hist tail", {i: int(h[i]) for i in range(100, 256) if h[i]})
keys, vals = sk.bigcounts()
print("bigcounts", list(zip(keys.tolist(), vals.tolist()))[:10])
from tests import oracle_lib as ol
for k in keys.tolist()[:5]:
    print(k, ol.revhash(k, 20))
