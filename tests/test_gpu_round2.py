"""Round-2 additions through the C ABI, against the oracle: per-k-mer "was new" flags (the tagging consumer), sketches with more
than ten tables, the first-touch log that makes n_unique_kmers and abundance_distribution exact across replicas, a bigcount
mask sketch, the packed host path.  -m gpu."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol
from common import synth_reads

pytestmark = pytest.mark.gpu


def make_gpu(cls, k, sizes, device=0):
    from khmer_b200 import cabi
    kind, hk, _ = ol.CLASSES[cls]
    return cabi.Sketch(kind, hk, k, sizes, device=device)


def _same(g, o, n_tables):
    assert g.stats() == (o.n_occupied(), o.n_unique_kmers())
    for i in range(n_tables):
        assert np.array_equal(g.table(i), o.table(i)), "table %d" % i


def _newflags_body(cls):
    kind, hk, _ = ol.CLASSES[cls]
    k = 19 if hk == ol.TWOBIT else 37
    sizes = ol.primes_near_x(3, 9000)      # small: plenty of collisions, "new" differs from "first occurrence"
    g, o = make_gpu(cls, k, sizes), ol.Oracle(cls, k, sizes)
    for part in range(3):
        reads = synth_reads(300 + part, 500, 120, 2500, err=0.02) + ["ACGT", "", "T" * 90, "".join("ACGT"[i % 4] for i in range(9000))]
        n, n_new, flags = g.consume_reads_new(reads, clean=True)
        want = np.zeros(len(flags), dtype=np.uint8)
        at = total = 0
        for r in reads:
            hs = o.kmer_hashes(ol.clean(r)) if len(r) >= k else []
            for i, h in enumerate(hs):
                want[at + i] = o.add(int(h))
            total += len(hs)
            at += len(r)
        assert n == total and n_new == int(want.sum())
        assert np.array_equal(flags, want)
        _same(g, o, 3)


@pytest.mark.parametrize("cls", list(ol.CLASSES))
@pytest.mark.parametrize("variant", ["default", "grouped-small-chunks"])
def test_consume_reads_new_flags(cls, variant):
    if variant == "default":
        _newflags_body(cls)
        return
    env = dict(os.environ, KMGPU_R2_CLS=cls, KMGPU_GROUP_MIN_BUCKETS="0", KMGPU_CHUNK_BASES="16384", KMGPU_PART_BASES="1000")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.abspath(__file__), "-k", "test_newflags_inner"],
                       env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_newflags_inner():
    cls = os.environ.get("KMGPU_R2_CLS")
    if not cls:
        pytest.skip("driver-invoked only")
    _newflags_body(cls)


@pytest.mark.parametrize("cls,nt", [("Countgraph", 11), ("SmallCountgraph", 32), ("Nodegraph", 17), ("Counttable", 12)])
def test_more_than_ten_tables(cls, nt):
    """NibbleStorage allows 32 tables (storage.hh:278-280), the others any number: these shapes take the grouped path."""
    kind, hk, _ = ol.CLASSES[cls]
    k = 21 if hk == ol.TWOBIT else 33
    sizes = ol.primes_near_x(nt, 50000)
    g, o = make_gpu(cls, k, sizes), ol.Oracle(cls, k, sizes)
    if kind == ol.BYTE:
        g.set_use_bigcount(True)
        o.set_use_bigcount(True)
    for part in range(2):
        reads = synth_reads(40 + part, 1500, 140, 8000, err=0.01, with_n=True) + ["ACGTTGCA" * 40] * 300
        assert g.consume_reads(reads) == o.consume_reads(reads)
        _same(g, o, nt)
    probe = [p for p in synth_reads(7, 30, 120, 8000)]
    hs = np.concatenate([o.kmer_hashes(p) for p in probe])
    assert np.array_equal(g.get_counts(hs), np.array([o.get(int(h)) for h in hs], dtype=np.uint16))
    if kind == ol.BYTE:
        gk, gv = g.bigcounts()
        assert len(gk) > 0 and dict(zip(gk.tolist(), gv.tolist())) == o.bigcounts()


@pytest.mark.parametrize("cls", ["Countgraph", "SmallCountgraph", "Nodegraph", "Nodetable"])
@pytest.mark.parametrize("world", [2, 5])
def test_first_touch_log_replicas_exact_unique(cls, world):
    """Replicas (emulated on one GPU, one process) fed contiguous shards in rank order, logs on: after the merge every replica
    reports the n_unique_kmers of ONE sketch fed all reads — and the per-rank local counts do NOT add up to it."""
    from khmer_b200 import cabi
    kind, hk, _ = ol.CLASSES[cls]
    k = 21 if hk == ol.TWOBIT else 35
    sizes = ol.primes_near_x(4, 30000)      # collision-heavy: local and global "new" disagree often
    reads = synth_reads(world * 3 + k, 3000, 110, 9000, err=0.02, with_n=True) + ["GATTACA" * 20] * 40
    from khmer_b200.multigpu import shard_range
    reps = [make_gpu(cls, k, sizes) for _ in range(world)]
    for r, sk in enumerate(reps):
        sk.first_touch_log(True)
    # a second epoch starts from the merged state, common to all replicas: meaningful for Bloom filters only (merging counting
    # replicas that share a base would add the base `world` times — counting replicas start a merge epoch empty)
    for epoch in range(2 if kind == ol.BIT else 1):
        o = ol.Oracle(cls, k, sizes) if epoch == 0 else o
        local_sum = 0
        for r, sk in enumerate(reps):
            lo, hi = shard_range(len(reads), r, world)
            part = reads[lo:hi] if epoch == 0 else [x[::-1] for x in reads[lo:hi]]
            before = sk.n_unique_kmers()
            sk.consume_reads(part)
            local_sum += sk.n_unique_kmers() - before
            o.consume_reads(part)
        cabi.reduce_replicas(reps)
        for sk in reps:
            _same(sk, o, 4)
        if epoch == 0:
            assert local_sum > o.n_unique_kmers()   # the log made a difference


def test_first_touch_log_abundance_across_replicas():
    """config C4's query: abundance_distribution over shards with one tracking filter per rank == one tracking filter, one pass."""
    from khmer_b200 import cabi
    from khmer_b200.multigpu import shard_range
    world, k = 3, 21
    sizes = ol.primes_near_x(4, 40000)
    reads = synth_reads(91, 4000, 100, 6000, err=0.01)
    counts, oc = make_gpu("SmallCountgraph", k, sizes), ol.Oracle("SmallCountgraph", k, sizes)
    counts.consume_reads(reads)
    oc.consume_reads(reads)
    tracking = [make_gpu("Nodegraph", k, sizes) for _ in range(world)]
    for t in tracking:
        t.first_touch_log(True)
    for r, t in enumerate(tracking):
        lo, hi = shard_range(len(reads), r, world)
        counts.abundance_distribution(reads[lo:hi], t)
    cabi.attach_replicas(tracking)
    hist = np.zeros(65536, dtype=np.uint64)
    for t in tracking:
        t.first_touch_resolve(hist=hist)
    ot = ol.Oracle("Nodegraph", k, sizes)
    want = oc.abundance_distribution(reads, ot)
    assert np.array_equal(hist, want)
    for t in tracking:
        t.ipc_detach()


def test_mask_sketch_with_bigcount_counts_above_255():
    """consume_seqfile_with_mask compares mask->get_count(kmer), which for a bigcount ByteStorage exceeds 255 (hashtable.cc:177-178,
    storage.hh:640-647)."""
    k, sizes = 17, ol.primes_near_x(3, 20000)
    hot = synth_reads(5, 4, 80, 400)
    reads = synth_reads(6, 600, 80, 3000)
    mask, omask = make_gpu("Countgraph", k, sizes), ol.Oracle("Countgraph", k, sizes)
    mask.set_use_bigcount(True)
    omask.set_use_bigcount(True)
    mask.consume_reads(hot * 400 + reads)
    omask.consume_reads(hot * 400 + reads)
    assert max(omask.bigcounts().values()) >= 400
    for thr, ge in ((300, True), (300, False), (255, True), (1000, False)):
        g, o = make_gpu("Countgraph", k, sizes), ol.Oracle("Countgraph", k, sizes)
        n = g.consume_reads(hot + reads, mask=(mask, thr, ge))
        # the oracle has no mask loop: apply the predicate k-mer by k-mer (Hashtable::consume_seqfile_with_mask, hashtable.cc:152-190)
        want = 0
        for r in hot + reads:
            for h in o.kmer_hashes(ol.clean(r)):
                c = omask.get(int(h))
                if (c >= thr) if ge else (c <= thr):
                    o.add(int(h))
                    want += 1
        assert n == want
        _same(g, o, 3)


def test_trim_batch_matches_reference_semantics():
    """Hashtable::trim_on_abundance / trim_below_abundance (src/oxli/hashtable.cc:504-560) for a batch, against the oracle's counts"""
    k, sizes = 17, ol.primes_near_x(3, 50000)
    g, o = make_gpu("Countgraph", k, sizes), ol.Oracle("Countgraph", k, sizes)
    reads = synth_reads(3, 800, 90, 1500, err=0.03)
    g.consume_reads(reads)
    o.consume_reads(reads)
    queries = synth_reads(4, 300, 90, 1500, err=0.05) + ["ACGT", "A" * 17, "A" * 18, ""]
    for abund, below in ((3, False), (8, False), (5, True), (40, True)):
        got = g.trim_batch(queries, abund, below=below)
        for q, pos in zip(queries, got):
            c = o.kmer_counts(q).tolist() if len(q) >= k else []
            bad = (lambda v: v > abund) if below else (lambda v: v < abund)
            if len(c) <= 1 or bad(c[0]):
                want = 0
            else:
                want = next((k - 1 + i for i in range(1, len(c)) if bad(c[i])), len(q))
            assert pos == want, (q, abund, below)


@pytest.mark.parametrize("cls", ["Countgraph", "SmallCounttable", "Nodegraph"])
def test_batch_read_medians_equals_host_call(cls):
    """kmgpu_batch_read_medians (device-resident batch) against kmgpu_read_medians and the oracle, incl. reads without k-mers and
    a sketch of five tables (the count kernel takes the tables four at a time)"""
    from khmer_b200 import cabi
    kind, hk, _ = ol.CLASSES[cls]
    k = 21 if hk == ol.TWOBIT else 33
    sizes = ol.primes_near_x(5, 40000)
    g, o = make_gpu(cls, k, sizes), ol.Oracle(cls, k, sizes)
    reads = synth_reads(11, 1500, 140, 6000, err=0.01)
    assert g.consume_reads(reads) == o.consume_reads(reads)
    q = reads[:400] + ["ACGT", "", "A" * (k - 1), "C" * 300]
    b = cabi.Batch(q, k, clean=True)
    med, avg, sd, nk = g.batch_read_medians(b)
    med2, avg2, sd2, nk2 = g.read_medians(q, clean=True)
    assert np.array_equal(med, med2) and np.array_equal(nk, nk2)
    assert avg.tobytes() == avg2.tobytes() and sd.tobytes() == sd2.tobytes()
    for i in (0, 7, 399, 403):
        if len(q[i]) >= k:
            m, a, s = o.median(q[i])
            assert (int(med[i]), np.float32(avg[i]).tobytes(), np.float32(sd[i]).tobytes()) == (m, np.float32(a).tobytes(), np.float32(s).tobytes())
    assert g.batch_read_medians(b, stats=False)[0].tolist() == med.tolist()
    b.close()


@pytest.mark.parametrize("k,p", [(20, 14), (32, 10), (41, 16), (5, 4), (75, 12)])
def test_hll_registers_equal_reference(k, p):
    """kmgpu_hll_consume: the registers of the reference's HLLCounter (oracle restatement pinned to it in test_oracle.py; the
    compiled reference itself where it travelled), over several calls, with dirty reads cleaned, merged with another counter"""
    from khmer_b200 import cabi
    reads_a = synth_reads(21, 3000, 150, 40000, err=0.01, with_n=True) + ["ACGT", "", "N" * 200, "acgtn" * 40]
    reads_b = synth_reads(22, 2000, 90, 40000, err=0.0)
    g = cabi.HLL(k, p)
    n = g.consume_reads(reads_a, clean=True)
    want, n_want = ol.hll_consume(reads_a, k, p, clean=True)
    assert n == n_want and np.array_equal(g.registers(), want)
    n2 = g.consume_reads(reads_b, clean=True)
    want2, n_want2 = ol.hll_consume(reads_b, k, p, counters=want, clean=True)
    assert n2 == n_want2 and np.array_equal(g.registers(), want2)
    if ol.have_ref():
        try:
            ref = ol.RefHLL(p, k)
        except ol.RefError:
            ref = None
        if ref is not None:
            for r in reads_a + reads_b:
                ref.consume_string(ol.clean(r))
            assert np.array_equal(g.registers(), ref.counters())
    other = cabi.HLL(k, p)
    other.consume_reads(reads_b, clean=True)
    other.merge_registers(want)                       # HLLCounter::merge
    assert np.array_equal(other.registers(), want2)
    other.merge_registers(want, replace=True)         # set_counters
    assert np.array_equal(other.registers(), want)
    if k > 32:
        with pytest.raises(cabi.KmgpuError):
            g.consume_reads(["ACGTNACGT" * 20], clean=False)   # Murmur hashes the letters themselves: non-ACGT needs cleaning
