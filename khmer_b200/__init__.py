"""khmer_b200 — B200-native k-mer ingestion backend for khmer's sketches.

Layers (bottom up):
  csrc/           sm_100a kernels + the C ABI of include/kmgpu.h           -> libkmgpu.so
  host/           liboxli-compatible C++ host layer over that ABI          -> liboxli_b200.so
  _oxli           its CPython binding (same class/method names as khmer._oxli / khmer._khmer)
  cabi            literal ctypes view of the C ABI (used by the C-ABI tests and bench.py)

`from khmer_b200 import Countgraph, ReadParser, ...` mirrors `from khmer import ...` for the ingestion path:
Countgraph / SmallCountgraph / Nodegraph / Counttable / SmallCounttable / Nodetable with consume_seqfile,
consume, get, get_median_count, median_at_least, abundance_distribution, save/load, n_unique_kmers ...
Tables live in GPU memory; without the built extensions or without a CUDA device everything raises.
"""
__version__ = "0.1.0"

_EXPORTS = ("Countgraph", "SmallCountgraph", "Nodegraph", "Counttable", "SmallCounttable", "Nodetable", "Hashtable",
            "ReadParser", "FastxParser", "Read", "forward_hash", "forward_hash_no_rc", "reverse_hash", "hash_murmur3",
            "hash_no_rc_murmur3", "reverse_complement", "get_n_primes_near_x", "compute_band_interval", "MAX_KCOUNT",
            "MAX_BIGCOUNT")


def __getattr__(name):
    if name in _EXPORTS:
        from . import _oxli   # built by __graft_entry__.build(); ImportError if missing — no fallback
        return getattr(_oxli, name)
    raise AttributeError(name)


def calc_expected_collisions(graph, force=False, max_false_pos=.2):
    """False-positive rate of a loaded table (khmer/__init__.py:181-215): (occupancy / min table size) ** n_tables."""
    sizes = graph.hashsizes()
    n_ht = float(len(sizes))
    occupancy = float(graph.n_occupied())
    min_size = min(sizes)
    fp_one = occupancy / min_size
    fp_all = fp_one ** n_ht
    if fp_all > max_false_pos:
        import sys
        print("**", file=sys.stderr)
        print("** ERROR: the graph structure is too small for ", file=sys.stderr)
        print("** this data set.  Increase data structure size", file=sys.stderr)
        print("** with --max_memory_usage/-M.", file=sys.stderr)
        print("**", file=sys.stderr)
        print("** Do not use these results!!", file=sys.stderr)
        print("**", file=sys.stderr)
        print("** (estimated false positive rate of %.3f;" % fp_all, file=sys.stderr, end=' ')
        print("max recommended %.3f)" % max_false_pos, file=sys.stderr)
        print("**", file=sys.stderr)
        if not force:
            sys.exit(1)
    return fp_all
