"""TEST INFRASTRUCTURE — runs the REFERENCE's own scripts, read in place from a khmer source tree (nothing copied), against
khmer_b200._oxli: the proof behind "load-into-counting.py / abundance-dist.py / load-graph.py work unchanged".

The reference's Python layer (khmer/*.py, oxli/*.py, scripts/*.py) is pure Python on top of two compiled modules,
khmer._khmer (CPython) and khmer._oxli.* (Cython).  Here those module names are bound to khmer_b200._oxli — same class and
function names — plus stubs for what the three scripts import but never call on this path (labels, partitions, assembler, HLL),
and minimal `screed` / `bz2file` stand-ins (neither is installable here; the scripts use them to sniff file types and to read
FASTA/FASTQ records in pure Python).  `khmer/__init__.py` itself is then executed from the reference tree.

    run_script(ref_root, "load-into-counting.py", ["-k", "20", "-x", "1e3", "-N", "2", out, infile]) -> (rc, stdout, stderr)
"""
import contextlib
import gzip
import io
import os
import runpy
import sys
import types


class _Stub:
    """Placeholder for out-of-scope classes: importable, not constructible."""

    def __init__(self, *a, **k):
        raise NotImplementedError("outside the k-mer ingestion path this backend replaces")


class _Record:
    def __init__(self, name, sequence, quality=None):
        self.name, self.sequence, self.quality = name, sequence, quality
        if quality is not None:
            self.annotations = ""

    def __len__(self):
        return len(self.sequence)


def _screed_open(filename, *a, **k):
    """FASTA/FASTQ records (plain or gzip) the way screed.open yields them."""
    with open(filename, "rb") as fh:
        magic = fh.read(2)
    data = (gzip.open(filename, "rb") if magic == b"\x1f\x8b" else open(filename, "rb")).read().decode()
    lines = data.split("\n")
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln.startswith(">"):
            name, seq = ln[1:], []
            i += 1
            while i < len(lines) and not lines[i].startswith(">"):
                seq.append(lines[i].strip())
                i += 1
            yield _Record(name, "".join(seq))
        elif ln.startswith("@"):
            yield _Record(ln[1:], lines[i + 1].strip(), lines[i + 3].strip())
            i += 4
        else:
            i += 1


def install(ref_root):
    """Bind the module names the reference's Python layer imports.  Idempotent per process."""
    if "khmer" in sys.modules and getattr(sys.modules["khmer"], "_b200_harness", False):
        return sys.modules["khmer"]
    import khmer_b200
    from khmer_b200 import _oxli as ox

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    # third-party modules that are not installable here
    screed = mod("screed", open=_screed_open, Record=_Record, __version__="shim")
    mod("screed.screedRecord", Record=_Record)
    mod("screed.utils", to_str=lambda s: s.decode() if isinstance(s, bytes) else s)
    screed.screedRecord = sys.modules["screed.screedRecord"]
    import bz2
    mod("bz2file", BZ2File=bz2.BZ2File, open=bz2.open)

    # the compiled modules of the reference -> khmer_b200._oxli
    khmer = types.ModuleType("khmer")
    khmer.__path__ = [os.path.join(ref_root, "khmer")]
    khmer.__file__ = os.path.join(ref_root, "khmer", "__init__.py")
    khmer._b200_harness = True
    sys.modules["khmer"] = khmer
    mod("khmer._khmer", Read=ox.Read, forward_hash=ox.forward_hash, forward_hash_no_rc=ox.forward_hash_no_rc,
        reverse_hash=ox.reverse_hash, hash_murmur3=ox.hash_murmur3, hash_no_rc_murmur3=ox.hash_no_rc_murmur3,
        reverse_complement=ox.reverse_complement, get_version_cpp=lambda: "khmer_b200 " + khmer_b200.__version__,
        ReadParser=ox.ReadParser, FILETYPES={"COUNTING_HT": 1, "HASHBITS": 2, "TAGS": 3, "STOPTAGS": 4, "SUBSET": 5,
                                             "LABELSET": 6, "SMALLCOUNT": 7, "QFCOUNT": 8})
    pkg = mod("khmer._oxli")
    pkg.__path__ = []
    mod("khmer._oxli.graphs", Counttable=ox.Counttable, QFCounttable=_Stub, CyclicCounttable=_Stub, Nodetable=ox.Nodetable,
        SmallCounttable=ox.SmallCounttable, Countgraph=ox.Countgraph, SmallCountgraph=ox.SmallCountgraph, Nodegraph=ox.Nodegraph,
        Hashtable=ox.Hashtable, Hashgraph=ox.Hashtable)
    mod("khmer._oxli.labeling", GraphLabels=_Stub)
    mod("khmer._oxli.legacy_partitioning", SubsetPartition=_Stub, PrePartitionInfo=_Stub)
    mod("khmer._oxli.readaligner", ReadAligner=_Stub)
    mod("khmer._oxli.assembly", LinearAssembler=_Stub, SimpleLabeledAssembler=_Stub, JunctionCountAssembler=_Stub)
    mod("khmer._oxli.hashset", HashSet=_Stub)
    mod("khmer._oxli.hllcounter", HLLCounter=_Stub)

    def is_prime(n):
        return n > 1 and (n == 2 or (n % 2 and all(n % i for i in range(3, int(n ** 0.5) + 1, 2))))

    mod("khmer._oxli.utils", get_n_primes_near_x=ox.get_n_primes_near_x, is_prime=is_prime)

    class UnpairedReadsError(ValueError):
        def __init__(self, msg, r1, r2):
            super().__init__(msg)
            self.read1, self.read2 = r1, r2

    def _split(name):
        return name.split(None, 1)[0] if name else name

    def check_is_left(name):
        n = _split(name)
        return n.endswith("/1") or (len(name.split()) > 1 and name.split()[1].startswith("1:"))

    def check_is_right(name):
        n = _split(name)
        return n.endswith("/2") or (len(name.split()) > 1 and name.split()[1].startswith("2:"))

    def check_is_pair(r1, r2):
        a, b = _split(r1.name), _split(r2.name)
        if a.endswith("/1") and b.endswith("/2"):
            return a[:-2] == b[:-2]
        return a == b and check_is_left(r1.name) and check_is_right(r2.name)

    mod("khmer._oxli.parsing", FastxParser=ox.FastxParser, check_is_left=check_is_left, check_is_right=check_is_right,
        check_is_pair=check_is_pair, UnpairedReadsError=UnpairedReadsError, Sequence=_Record, BrokenPairedReader=_Stub,
        SplitPairedReader=_Stub, _split_left_right=lambda n: (n, ""))
    # the reference's own pure-Python files, executed where they lie
    with open(khmer.__file__) as fh:
        exec(compile(fh.read(), khmer.__file__, "exec"), khmer.__dict__)
    oxli = types.ModuleType("oxli")
    oxli.__path__ = [os.path.join(ref_root, "oxli")]
    oxli.__file__ = os.path.join(ref_root, "oxli", "__init__.py")
    sys.modules["oxli"] = oxli
    with open(oxli.__file__) as fh:
        exec(compile(fh.read(), oxli.__file__, "exec"), oxli.__dict__)
    return khmer


def run_script(ref_root, script, argv, cwd=None):
    """Run <ref_root>/scripts/<script> with argv in this process; returns (exit status, stdout, stderr)."""
    install(ref_root)
    path = os.path.join(ref_root, "scripts", script)
    out, err = io.StringIO(), io.StringIO()
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = [path] + [str(a) for a in argv]
    rc = 0
    try:
        if cwd:
            os.chdir(cwd)
        with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
            try:
                runpy.run_path(path, run_name="__main__")
            except SystemExit as e:
                rc = e.code if isinstance(e.code, int) else (0 if e.code is None else 1)
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
    return rc, out.getvalue(), err.getvalue()
