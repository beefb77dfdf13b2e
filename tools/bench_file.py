#!/usr/bin/env python
"""End-to-end file ingestion through the host layer: khmer_b200.Countgraph(20, 1e8, 4).consume_seqfile(path) on a FASTA
of synthetic reads (plain and gzip), next to the reference's own consume_seqfile (oracle/_ref, all cores)."""
import json, os, sys, time, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import khmer_b200 as kh

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
buf, off, _ = bench.synth_batch(5, n_reads)
rows = buf.reshape(n_reads, 150)
block = np.empty((n_reads, 154), dtype=np.uint8)
block[:, :3] = np.frombuffer(b">r\n", dtype=np.uint8); block[:, 3:153] = rows; block[:, -1] = 10
path = "/tmp/bench_reads.fa"
open(path, "wb").write(block.tobytes())
subprocess.check_call(["gzip", "-1", "-k", "-f", path])
out = {"reads": n_reads, "file_mb": os.path.getsize(path) / 1e6, "host_cores": os.cpu_count()}
for label, p in (("plain", path), ("gzip", path + ".gz")):
    for rep in range(2):
        cg = kh.Countgraph(20, 1e8, 4)
        cg.set_use_bigcount(True)
        t0 = time.perf_counter()
        reads, kmers = cg.consume_seqfile(p)
        dt = time.perf_counter() - t0
        del cg
    out[label + "_gkmers_per_s"] = kmers / dt / 1e9
    out[label + "_seconds"] = dt
import oracle_lib as ol
if ol.have_ref():
    r = ol.Ref("Countgraph", 20, bench.primes_near_x(4, int(1e8)))
    r.set_use_bigcount(True)
    t0 = time.perf_counter()
    _, kmers = r.consume_seqfile(path, threads=os.cpu_count())
    out["reference_plain_gkmers_per_s"] = kmers / (time.perf_counter() - t0) / 1e9
print(json.dumps(out))
