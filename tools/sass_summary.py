#!/usr/bin/env python
"""Per-kernel SASS summary of khmer_b200/libkmgpu.so: instruction counts and the mnemonics that prove which hardware paths a
kernel uses (UBLKCP = cp.async.bulk, SYNCS = mbarrier, ATOMS/ATOMG/REDG = shared / global atomics), plus the lines around every
bulk copy.

    python tools/sass_summary.py [regex of kernel names] > profiles/rN_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else r"k_part|k_apply2|k_apply_sparse|k_bigscan2|k_hash64|k_ft_|k_norm_")
sass = subprocess.run(["cuobjdump", "-sass", "-arch", "sm_100a", os.path.join(ROOT, "khmer_b200", "libkmgpu.so")],
                      capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
print("# cuobjdump -sass -arch sm_100a khmer_b200/libkmgpu.so, kernels matching /%s/" % pat.pattern)
for part in re.split(r"\n\s*Function : ", sass)[1:]:
    name = part.split("\n", 1)[0].strip()
    if not pat.search(name):
        continue
    lines = part.split("\n")
    ins = [(i, m.group(1)) for i, ln in enumerate(lines) for m in [re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)] if m]
    tally = collections.Counter(op for _, op in ins if op.startswith(("UBLKCP", "SYNCS", "ATOMS", "ATOMG", "REDG", "RED.", "UTMA", "LDGSTS")))
    print("\n%s\n    %d instructions;  %s" % (demangle(name)[:240], len(ins), "  ".join("%s x%d" % kv for kv in sorted(tally.items()))))
    for i, op in ins:
        if op.startswith(("UBLKCP", "SYNCS")):
            print("        " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", lines[i]).strip())
