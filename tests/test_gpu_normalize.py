"""Digital normalization on the device (kmgpu_normalize_batch) against the serial loop: the reference script's own md5s
(tests/test_script_output.py:51-70), the oracle's keep flags and tables on synthetic reads for every sketch class, pairs,
tiny windows (every bundle in between), and 1 M reads.  -m gpu."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol
from common import synth_reads, synth_buffer

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def make_gpu(cls, k, sizes):
    from khmer_b200 import cabi
    kind, hk, _ = ol.CLASSES[cls]
    return cabi.Sketch(kind, hk, k, sizes)


def _same(g, o, n_tables):
    assert g.stats() == (o.n_occupied(), o.n_unique_kmers())
    for i in range(n_tables):
        assert np.array_equal(g.table(i), o.table(i)), "table %d" % i


def test_reference_script_md5s(datadir):
    recs, name = [], None
    for ln in open(os.path.join(datadir, "simple-genome-reads.fa")):
        ln = ln.rstrip("\n")
        if ln.startswith(">"):
            name = ln[1:]
        else:
            recs.append((name, ln))
    cleaned = [s.upper().replace("N", "A") for _, s in recs]
    sizes = ol.primes_near_x(4, int(1e7 / 4))
    for cutoff, want in ((20, "942e9024c25a8d85033d755d86aba4a3"), (15, "0d1b4b9d4c76cb8cdeee5a98f6e70163")):
        g = make_gpu("Countgraph", 21, sizes)
        o = ol.Oracle("Countgraph", 21, sizes)
        keep, kmers = g.normalize_batch(cleaned, cutoff)
        okeep, okmers = o.normalize_reads(cleaned, cutoff)
        out = "".join(">%s\n%s\n" % (n, s) for (n, s), k in zip(recs, keep) if k)
        assert hashlib.md5(out.encode()).hexdigest() == want
        assert np.array_equal(keep, okeep) and kmers == okmers
        _same(g, o, 4)


def _body(cls):
    kind, hk, _ = ol.CLASSES[cls]
    rng = np.random.default_rng(11)
    k = 21 if hk == ol.TWOBIT else 35
    sizes = ol.primes_near_x(3, 60000)
    g, o = make_gpu(cls, k, sizes), ol.Oracle(cls, k, sizes)
    if kind == ol.BYTE:
        g.set_use_bigcount(True)
        o.set_use_bigcount(True)
    cutoff = 12 if kind == ol.BYTE else 5 if kind == ol.NIBBLE else 1
    for part in range(3):
        # deep coverage of a small genome: most reads end up discarded, decisions depend on the order inside a window
        reads = synth_reads(100 + part, 4000, 100, 3000, err=0.01) + ["ACGT", "", "A" * 60, "ACGTTGCA" * 20]
        paired = None
        if part == 1:   # every other pair of neighbours is a bundle
            paired = np.zeros(len(reads), dtype=np.uint8)
            paired[0:len(reads) - 6:4] = 1
        keep, kmers = g.normalize_batch(reads, cutoff, paired=paired)
        okeep, okmers = o.normalize_reads(reads, cutoff, paired=paired)
        assert np.array_equal(keep, okeep), "keep flags differ (%d vs %d kept)" % (keep.sum(), okeep.sum())
        assert kmers == okmers
        assert 0 < keep.sum() < len(reads)
        _same(g, o, 3)
    if kind == ol.BYTE:
        gk, gv = g.bigcounts()
        assert dict(zip(gk.tolist(), gv.tolist())) == o.bigcounts()


@pytest.mark.parametrize("cls", list(ol.CLASSES))
@pytest.mark.parametrize("window", [None, 7, 300])
def test_normalize_vs_oracle(cls, window):
    if window is None:
        _body(cls)
        return
    env = dict(os.environ, KMGPU_NORM_WINDOW=str(window), KMGPU_NORM_CLS=cls)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.abspath(__file__), "-k", "test_normalize_inner"],
                       env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_normalize_inner():
    cls = os.environ.get("KMGPU_NORM_CLS")
    if not cls:
        pytest.skip("driver-invoked only")
    _body(cls)


def test_normalize_one_million_reads():
    """config C3's loop at 1 M reads: Countgraph k=20, C=20, 30x of a 5 Mbp genome, 4 x 4e7-byte tables (grouped path for the
    commits), default 64 Ki-read windows."""
    import time
    buf, off = synth_buffer(2024, 1_000_000, 150, 5_000_000)
    sizes = ol.primes_near_x(4, 4e7)
    g, o = make_gpu("Countgraph", 20, sizes), ol.Oracle("Countgraph", 20, sizes)
    g.normalize_batch((buf[:150 * 300_000], off[:300_001]), 20)      # warm-up: device workspaces are allocated on first use
    g.reset()
    t0 = time.perf_counter()
    keep, kmers = g.normalize_batch((buf, off), 20)
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    okeep, okmers = o.normalize_reads((buf, off), 20)
    t_cpu = time.perf_counter() - t0
    print("normalize 1M reads: device %.2f s (%.2f M reads/s), serial oracle %.1f s; kept %d" % (t_gpu, 1.0 / t_gpu, t_cpu, keep.sum()))
    assert np.array_equal(keep, okeep)
    assert kmers == okmers
    _same(g, o, 4)
