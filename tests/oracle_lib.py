"""ctypes bindings for the two parity checkers (TEST INFRASTRUCTURE ONLY).

* ``Oracle``  — oracle/_build/libkhmer_oracle.so, our plain-C restatement (oracle/khmer_oracle.c).
* ``Ref``     — oracle/_ref/libkhmer_ref.so, the unmodified reference liboxli compiled by oracle/Makefile
                (present in the build container and shipped prebuilt to the GPU box; optional).

Nothing under khmer_b200/ imports this module.
"""
import ctypes as C
import gzip
import bz2
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libkhmer_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libkhmer_ref.so")

BYTE, NIBBLE, BIT = 0, 1, 2
TWOBIT, MURMUR = 0, 1

# reference class name -> (storage kind, hash kind, ref_new kind id)
CLASSES = {
    "Countgraph": (BYTE, TWOBIT, 0),
    "SmallCountgraph": (NIBBLE, TWOBIT, 1),
    "Nodegraph": (BIT, TWOBIT, 2),
    "Counttable": (BYTE, MURMUR, 3),
    "SmallCounttable": (NIBBLE, MURMUR, 4),
    "Nodetable": (BIT, MURMUR, 5),
}

u64p = C.POINTER(C.c_uint64)
u16p = C.POINTER(C.c_uint16)


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
            os.path.join(ROOT, "oracle", "khmer_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return ORACLE_SO


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        L = C.CDLL(build_oracle())
        L.ko_new.restype = C.c_void_p
        L.ko_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, u64p]
        L.ko_free.argtypes = [C.c_void_p]
        L.ko_set_use_bigcount.argtypes = [C.c_void_p, C.c_int]
        for f in ("ko_n_unique", "ko_n_occupied", "ko_n_bigcounts"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        L.ko_table.restype = C.POINTER(C.c_uint8)
        L.ko_table.argtypes = [C.c_void_p, C.c_int]
        L.ko_table_nbytes.restype = C.c_uint64
        L.ko_table_nbytes.argtypes = [C.c_void_p, C.c_int]
        L.ko_bigcounts.argtypes = [C.c_void_p, u64p, u16p]
        L.ko_add.argtypes = [C.c_void_p, C.c_uint64]
        L.ko_get_count.argtypes = [C.c_void_p, C.c_uint64]
        L.ko_get_count.restype = C.c_uint
        L.ko_hash.argtypes = [C.c_void_p, C.c_char_p, u64p]
        L.ko_kmer_hashes.restype = C.c_int64
        L.ko_kmer_hashes.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, u64p, C.c_int64]
        L.ko_consume_string.restype = C.c_int64
        L.ko_consume_string.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        L.ko_consume_reads.restype = C.c_int64
        L.ko_consume_reads.argtypes = [C.c_void_p, C.c_char_p, u64p, C.c_uint64, C.c_int, C.c_int, C.c_uint64,
                                       C.c_uint64]
        L.ko_get_kmer_counts.restype = C.c_int64
        L.ko_get_kmer_counts.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, u16p, C.c_int64]
        L.ko_median.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, u16p, C.POINTER(C.c_float),
                                C.POINTER(C.c_float)]
        L.ko_median_at_least.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint]
        L.ko_abundance_distribution.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, u64p, C.c_uint64, C.c_int, u64p]
        L.ko_normalize_reads.restype = C.c_int64
        L.ko_normalize_reads.argtypes = [C.c_void_p, C.c_char_p, u64p, C.c_uint64, C.c_void_p, C.c_uint, C.c_void_p]
        L.ko_update_from.argtypes = [C.c_void_p, C.c_void_p]
        L.ko_save.argtypes = [C.c_void_p, C.c_char_p]
        L.ko_hash_twobit.argtypes = [C.c_char_p, C.c_int, u64p, u64p, u64p]
        L.ko_hash_murmur.argtypes = [C.c_char_p, C.c_int, u64p, u64p, u64p]
        L.ko_revhash.argtypes = [C.c_uint64, C.c_int, C.c_char_p]
        L.ko_primes_near_x.argtypes = [C.c_uint32, C.c_uint64, u64p]
        L.ko_band_interval.argtypes = [C.c_uint, C.c_uint, u64p, u64p]
        L.ko_clean.argtypes = [C.c_char_p, C.c_size_t]
        _oracle = L
    return _oracle


def _b(s):
    return s if isinstance(s, (bytes, bytearray)) else s.encode()


def pack_reads(reads):
    """list of str/bytes -> (concatenated bytes, offsets uint64[n+1])"""
    bs = [_b(r) for r in reads]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    return b"".join(bs), off


def hash_twobit(kmer, k=None):
    L = oracle_lib()
    kmer = _b(kmer)
    k = len(kmer) if k is None else k
    f, r, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    rc = L.ko_hash_twobit(kmer, k, C.byref(f), C.byref(r), C.byref(c))
    if rc:
        raise ValueError("ko_hash_twobit rc=%d" % rc)
    return c.value, f.value, r.value


def hash_murmur(kmer, k=None):
    L = oracle_lib()
    kmer = _b(kmer)
    k = len(kmer) if k is None else k
    f, r, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    L.ko_hash_murmur(kmer, k, C.byref(f), C.byref(r), C.byref(c))
    return c.value, f.value, r.value


def revhash(h, k):
    buf = C.create_string_buffer(k + 1)
    oracle_lib().ko_revhash(h, k, buf)
    return buf.value.decode()


def primes_near_x(n, x):
    out = (C.c_uint64 * max(n, 1))()
    got = oracle_lib().ko_primes_near_x(int(n), int(x), out)
    return [out[i] for i in range(got)]


def band_interval(num_bands, band):
    lo, hi = C.c_uint64(), C.c_uint64()
    if oracle_lib().ko_band_interval(num_bands, band, C.byref(lo), C.byref(hi)):
        raise ValueError("bad band")
    return lo.value, hi.value


def clean(seq):
    b = C.create_string_buffer(_b(seq), len(seq))
    oracle_lib().ko_clean(b, len(seq))
    return b.raw.decode()


class Oracle:
    """Single-threaded CPU sketch (oracle port)."""

    def __init__(self, cls, k, sizes):
        self.L = oracle_lib()
        self.kind, self.hashkind, _ = CLASSES[cls]
        self.cls, self.k = cls, k
        self.sizes = [int(s) for s in sizes]
        arr = (C.c_uint64 * len(self.sizes))(*self.sizes)
        self.h = self.L.ko_new(self.kind, self.hashkind, k, len(self.sizes), arr)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ko_free(self.h)
            self.h = None

    def set_use_bigcount(self, on):
        if self.L.ko_set_use_bigcount(self.h, int(on)):
            raise ValueError("bigcount is not supported for this storage.")

    def hash(self, kmer):
        o = C.c_uint64()
        self.L.ko_hash(self.h, _b(kmer), C.byref(o))
        return o.value

    def add(self, h):
        if not isinstance(h, int):
            h = self.hash(h)
        return bool(self.L.ko_add(self.h, h))

    def get(self, h):
        if not isinstance(h, int):
            h = self.hash(h)
        return self.L.ko_get_count(self.h, h)

    def consume(self, seq):
        seq = _b(seq)
        return self.L.ko_consume_string(self.h, seq, len(seq))

    def consume_reads(self, reads, clean=True, band=None):
        seqs, off = pack_reads(reads)
        lo, hi = band if band else (0, 0)
        return self.L.ko_consume_reads(self.h, seqs, off.ctypes.data_as(u64p), len(off) - 1, int(clean),
                                       int(band is not None), lo, hi)

    def kmer_hashes(self, seq):
        seq = _b(seq)
        n = max(0, len(seq) - self.k + 1)
        out = np.zeros(max(n, 1), dtype=np.uint64)
        got = self.L.ko_kmer_hashes(self.h, seq, len(seq), out.ctypes.data_as(u64p), n)
        return out[:got]

    def kmer_counts(self, seq):
        seq = _b(seq)
        n = max(0, len(seq) - self.k + 1)
        out = np.zeros(max(n, 1), dtype=np.uint16)
        got = self.L.ko_get_kmer_counts(self.h, seq, len(seq), out.ctypes.data_as(u16p), n)
        return out[:got]

    def median(self, seq):
        seq = _b(seq)
        m, a, s = C.c_uint16(), C.c_float(), C.c_float()
        if self.L.ko_median(self.h, seq, len(seq), C.byref(m), C.byref(a), C.byref(s)):
            raise ValueError("no k-mer counts for this string; too short?")
        return m.value, a.value, s.value

    def median_at_least(self, seq, cutoff):
        seq = _b(seq)
        return bool(self.L.ko_median_at_least(self.h, seq, len(seq), cutoff))

    def abundance_distribution(self, reads, tracking, clean=True):
        seqs, off = pack_reads(reads)
        dist = np.zeros(65536, dtype=np.uint64)
        self.L.ko_abundance_distribution(self.h, tracking.h, seqs, off.ctypes.data_as(u64p), len(off) - 1,
                                         int(clean), dist.ctypes.data_as(u64p))
        return dist

    def normalize_reads(self, reads, cutoff, paired=None):
        """serial digital normalization (scripts/normalize-by-median.py:155-179): (keep flags, k-mers consumed)"""
        if isinstance(reads, tuple):
            buf, off = reads
            seqs = buf.tobytes() if hasattr(buf, "tobytes") else bytes(buf)
            off = np.ascontiguousarray(off, dtype=np.uint64)
        else:
            seqs, off = pack_reads(reads)
        n = len(off) - 1
        keep = np.zeros(max(n, 1), dtype=np.uint8)
        pw = np.ascontiguousarray(paired, dtype=np.uint8) if paired is not None else None
        kmers = self.L.ko_normalize_reads(self.h, seqs, off.ctypes.data_as(u64p), n,
                                          pw.ctypes.data if pw is not None else None, int(cutoff), keep.ctypes.data)
        return keep[:n], kmers

    def update(self, other):
        if self.L.ko_update_from(self.h, other.h):
            raise ValueError("both nodegraphs must have same table sizes")

    def n_unique_kmers(self):
        return self.L.ko_n_unique(self.h)

    def n_occupied(self):
        return self.L.ko_n_occupied(self.h)

    def table(self, i):
        n = self.L.ko_table_nbytes(self.h, i)
        return np.ctypeslib.as_array(self.L.ko_table(self.h, i), shape=(n,)).copy()

    def bigcounts(self):
        n = self.L.ko_n_bigcounts(self.h)
        keys = np.zeros(max(n, 1), dtype=np.uint64)
        vals = np.zeros(max(n, 1), dtype=np.uint16)
        self.L.ko_bigcounts(self.h, keys.ctypes.data_as(u64p), vals.ctypes.data_as(u16p))
        return dict(zip(keys[:n].tolist(), vals[:n].tolist()))

    def save(self, path):
        if self.L.ko_save(self.h, _b(path)):
            raise OSError("ko_save failed")


# --------------------------------------------------------------------------------------------
# compiled reference (optional)
# --------------------------------------------------------------------------------------------
_ref = None


def have_ref():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        L.ref_last_error.restype = C.c_char_p
        L.ref_new.restype = C.c_void_p
        L.ref_new.argtypes = [C.c_int, C.c_int, u64p, C.c_int]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_primes.argtypes = [C.c_uint32, C.c_uint64, u64p]
        L.ref_set_use_bigcount.argtypes = [C.c_void_p, C.c_int]
        L.ref_consume_seqfile.argtypes = [C.c_void_p, C.c_char_p, C.c_int, u64p, u64p]
        L.ref_consume_seqfile_banding.argtypes = [C.c_void_p, C.c_char_p, C.c_uint, C.c_uint, u64p, u64p]
        L.ref_consume_seqfile_with_mask.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint, C.c_int, u64p,
                                                    u64p]
        L.ref_consume_string.restype = C.c_int64
        L.ref_consume_string.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_hash_dna.argtypes = [C.c_void_p, C.c_char_p, u64p]
        L.ref_add_hash.argtypes = [C.c_void_p, C.c_uint64]
        L.ref_get_count_hash.argtypes = [C.c_void_p, C.c_uint64]
        L.ref_get_count_kmer.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_get_kmer_hashes.restype = C.c_int64
        L.ref_get_kmer_hashes.argtypes = [C.c_void_p, C.c_char_p, u64p, C.c_int64]
        L.ref_get_kmer_counts.restype = C.c_int64
        L.ref_get_kmer_counts.argtypes = [C.c_void_p, C.c_char_p, u16p, C.c_int64]
        L.ref_get_median_count.argtypes = [C.c_void_p, C.c_char_p, u16p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ref_median_at_least.argtypes = [C.c_void_p, C.c_char_p, C.c_uint]
        L.ref_abundance_distribution.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, u64p]
        L.ref_save.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_load.argtypes = [C.c_void_p, C.c_char_p]
        for f in ("ref_n_unique_kmers", "ref_n_occupied"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_n_tables.argtypes = [C.c_void_p]
        L.ref_ksize.argtypes = [C.c_void_p]
        L.ref_tablesizes.argtypes = [C.c_void_p, u64p]
        L.ref_raw_table.restype = C.POINTER(C.c_uint8)
        L.ref_raw_table.argtypes = [C.c_void_p, C.c_int]
        L.ref_nodegraph_update.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_hash_twobit.argtypes = [C.c_char_p, C.c_int, u64p, u64p, u64p]
        L.ref_hash_murmur.argtypes = [C.c_char_p, C.c_int, u64p, u64p, u64p]
        L.ref_revhash.argtypes = [C.c_uint64, C.c_int, C.c_char_p]
        if hasattr(L, "ref_n_tags"):   # a prebuilt library from before the tagging wrappers lacks them
            L.ref_consume_seqfile_and_tag.argtypes = [C.c_void_p, C.c_char_p, u64p, u64p]
            L.ref_n_tags.restype = C.c_int64
            L.ref_n_tags.argtypes = [C.c_void_p]
            L.ref_save_tagset.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_parse_clean.restype = C.c_int64
        L.ref_parse_clean.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, u64p, C.c_uint64, u64p]
        _ref = L
    return _ref


class RefError(Exception):
    pass


class Ref:
    """The compiled, unmodified reference behind a flat C wrapper (oracle/ref_driver.cc)."""

    def __init__(self, cls, k, sizes):
        self.L = ref_lib()
        self.kind, self.hashkind, kid = CLASSES[cls]
        self.cls, self.k = cls, k
        self.sizes = [int(s) for s in sizes]
        arr = (C.c_uint64 * len(self.sizes))(*self.sizes)
        self.h = self.L.ref_new(kid, k, arr, len(self.sizes))
        if not self.h:
            raise RefError(self.L.ref_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_free(self.h)
            self.h = None

    def _chk(self, rc):
        if rc < 0:
            raise RefError(self.L.ref_last_error().decode())
        return rc

    def set_use_bigcount(self, on):
        self._chk(self.L.ref_set_use_bigcount(self.h, int(on)))

    def consume_seqfile(self, path, threads=1):
        r, k = C.c_uint64(), C.c_uint64()
        self._chk(self.L.ref_consume_seqfile(self.h, _b(path), threads, C.byref(r), C.byref(k)))
        return r.value, k.value

    def consume_seqfile_and_tag(self, path):
        r, k = C.c_uint64(), C.c_uint64()
        self._chk(self.L.ref_consume_seqfile_and_tag(self.h, _b(path), C.byref(r), C.byref(k)))
        return r.value, k.value

    def n_tags(self):
        return self._chk(self.L.ref_n_tags(self.h))

    def save_tagset(self, path):
        self._chk(self.L.ref_save_tagset(self.h, _b(path)))

    def consume_seqfile_banding(self, path, num_bands, band):
        r, k = C.c_uint64(), C.c_uint64()
        self._chk(self.L.ref_consume_seqfile_banding(self.h, _b(path), num_bands, band, C.byref(r), C.byref(k)))
        return r.value, k.value

    def consume_seqfile_with_mask(self, path, mask, threshold=0, consume_masked=False):
        r, k = C.c_uint64(), C.c_uint64()
        self._chk(self.L.ref_consume_seqfile_with_mask(self.h, _b(path), mask.h, threshold, int(consume_masked),
                                                       C.byref(r), C.byref(k)))
        return r.value, k.value

    def consume(self, seq):
        return self._chk(self.L.ref_consume_string(self.h, _b(seq)))

    def hash(self, kmer):
        o = C.c_uint64()
        self._chk(self.L.ref_hash_dna(self.h, _b(kmer), C.byref(o)))
        return o.value

    def add(self, h):
        if not isinstance(h, int):
            h = self.hash(h)
        return bool(self._chk(self.L.ref_add_hash(self.h, h)))

    def get(self, h):
        if isinstance(h, int):
            return self._chk(self.L.ref_get_count_hash(self.h, h))
        return self._chk(self.L.ref_get_count_kmer(self.h, _b(h)))

    def kmer_hashes(self, seq):
        n = max(0, len(seq) - self.k + 1)
        out = np.zeros(max(n, 1), dtype=np.uint64)
        got = self._chk(self.L.ref_get_kmer_hashes(self.h, _b(seq), out.ctypes.data_as(u64p), n))
        return out[:got]

    def kmer_counts(self, seq):
        n = max(0, len(seq) - self.k + 1)
        out = np.zeros(max(n, 1), dtype=np.uint16)
        got = self._chk(self.L.ref_get_kmer_counts(self.h, _b(seq), out.ctypes.data_as(u16p), n))
        return out[:got]

    def median(self, seq):
        m, a, s = C.c_uint16(), C.c_float(), C.c_float()
        self._chk(self.L.ref_get_median_count(self.h, _b(seq), C.byref(m), C.byref(a), C.byref(s)))
        return m.value, a.value, s.value

    def median_at_least(self, seq, cutoff):
        return bool(self._chk(self.L.ref_median_at_least(self.h, _b(seq), cutoff)))

    def abundance_distribution(self, path, tracking):
        dist = np.zeros(65536, dtype=np.uint64)
        self._chk(self.L.ref_abundance_distribution(self.h, _b(path), tracking.h, dist.ctypes.data_as(u64p)))
        return dist

    def update(self, other):
        self._chk(self.L.ref_nodegraph_update(self.h, other.h))

    def save(self, path):
        self._chk(self.L.ref_save(self.h, _b(path)))

    def load(self, path):
        self._chk(self.L.ref_load(self.h, _b(path)))
        self.k = self.L.ref_ksize(self.h)

    def n_unique_kmers(self):
        return self.L.ref_n_unique_kmers(self.h)

    def n_occupied(self):
        return self.L.ref_n_occupied(self.h)

    def tablesizes(self):
        n = self.L.ref_n_tables(self.h)
        out = (C.c_uint64 * n)()
        self.L.ref_tablesizes(self.h, out)
        return list(out)

    def table(self, i):
        size = self.tablesizes()[i]
        n = size if self.kind == BYTE else size // 2 + 1 if self.kind == NIBBLE else size // 8 + 1
        return np.ctypeslib.as_array(self.L.ref_raw_table(self.h, i), shape=(n,)).copy()


def ref_parse_clean(path):
    """(list of cleaned sequences) as the reference's bulk loaders see the file."""
    L = ref_lib()
    total = C.c_uint64()
    n = L.ref_parse_clean(_b(path), None, 0, None, 0, C.byref(total))
    if n < 0:
        raise RefError(L.ref_last_error().decode())
    buf = C.create_string_buffer(total.value + 1)
    off = np.zeros(n + 1, dtype=np.uint64)
    L.ref_parse_clean(_b(path), buf, total.value, off.ctypes.data_as(u64p), n + 1, C.byref(total))
    raw = buf.raw
    return [raw[int(off[i]):int(off[i + 1])].decode() for i in range(n)]


# --------------------------------------------------------------------------------------------
# tiny FASTA/FASTQ reader for tests (host-side product parser is tested against it and against Ref)
# --------------------------------------------------------------------------------------------
def read_fastx(path):
    """Raw (uncleaned) sequences of a FASTA/FASTQ file, plain / .gz / .bz2."""
    with open(path, "rb") as fh:
        magic = fh.read(3)
    if magic[:2] == b"\x1f\x8b":
        data = gzip.open(path, "rb").read()
    elif magic == b"BZh":
        data = bz2.open(path, "rb").read()
    else:
        data = open(path, "rb").read()
    lines = data.split(b"\n")
    seqs = []
    i = 0
    while i < len(lines):
        ln = lines[i].rstrip(b"\r")
        if ln.startswith(b">"):
            i += 1
            parts = []
            while i < len(lines) and not lines[i].startswith(b">"):
                parts.append(lines[i].strip())
                i += 1
            seqs.append(b"".join(parts).decode())
        elif ln.startswith(b"@"):
            seqs.append(lines[i + 1].strip().decode())
            i += 4
        else:
            i += 1
    return seqs


# ---- HyperLogLog registers (HLLCounter, src/oxli/hllcounter.cc) --------------------------------------------
def hll_consume(reads, k, p, counters=None, clean=True):
    """plain-C restatement: (registers, k-mers consumed)"""
    L = oracle_lib()
    L.ko_hll_consume.restype = C.c_int64
    L.ko_hll_consume.argtypes = [C.c_char_p, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p]
    seqs, off = pack_reads(reads)
    regs = np.zeros(1 << p, dtype=np.uint8) if counters is None else np.ascontiguousarray(counters, dtype=np.uint8).copy()
    n = L.ko_hll_consume(seqs, off.ctypes.data_as(u64p), len(off) - 1, k, p, int(clean), regs.ctypes.data)
    return regs, n


class RefHLL:
    """the reference's HLLCounter (oracle/_ref)"""

    def __init__(self, p, k):
        self.L = ref_lib()
        if not hasattr(self.L, "ref_hll_new"):
            raise RefError("oracle/_ref was built before the HLL wrappers were added")
        self.L.ref_hll_new.restype = C.c_void_p
        self.L.ref_hll_new.argtypes = [C.c_int, C.c_int]
        self.L.ref_hll_free.argtypes = [C.c_void_p]
        self.L.ref_hll_consume_string.restype = C.c_int64
        self.L.ref_hll_consume_string.argtypes = [C.c_void_p, C.c_char_p]
        self.L.ref_hll_consume_seqfile.argtypes = [C.c_void_p, C.c_char_p, u64p, u64p]
        self.L.ref_hll_counters.argtypes = [C.c_void_p, C.c_void_p]
        self.L.ref_hll_set_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self.L.ref_hll_estimate.restype = C.c_int64
        self.L.ref_hll_estimate.argtypes = [C.c_void_p]
        self.p = p
        self.h = self.L.ref_hll_new(1 << p, k)
        if not self.h:
            raise RefError(self.L.ref_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_hll_free(self.h)
            self.h = None

    def consume_string(self, s):
        n = self.L.ref_hll_consume_string(self.h, _b(s))
        if n < 0:
            raise RefError(self.L.ref_last_error().decode())
        return n

    def consume_seqfile(self, path):
        r, k = C.c_uint64(), C.c_uint64()
        if self.L.ref_hll_consume_seqfile(self.h, _b(path), C.byref(r), C.byref(k)) < 0:
            raise RefError(self.L.ref_last_error().decode())
        return r.value, k.value

    def counters(self):
        out = np.zeros(1 << self.p, dtype=np.uint8)
        assert self.L.ref_hll_counters(self.h, out.ctypes.data) == len(out)
        return out

    def set_counters(self, regs):
        regs = np.ascontiguousarray(regs, dtype=np.uint8)
        assert self.L.ref_hll_set_counters(self.h, regs.ctypes.data, len(regs)) == 0

    def estimate(self):
        return self.L.ref_hll_estimate(self.h)
