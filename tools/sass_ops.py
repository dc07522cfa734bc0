#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show what a kernel is built from (cuobjdump -sass of the built library):
UBLKCP / SYNCS = TMA bulk copies completing on mbarriers, FFMA2 / FMUL2 / FADD2 = packed FP32, RED / REDG = reductions
to global memory (scalar and vector), ATOMS = shared-memory atomics, LDGMC = multimem (NVLS) load-reduce, ACQBULK /
PREEXIT = griddepcontrol.wait / launch_dependents (programmatic dependent launch), MUFU, LDS/LDG.
usage: python tools/sass_ops.py > profiles/sass_ops.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "gi-gs_b200", "lib", "libgigs_b200.so")
OPS = ["UBLKCP", "SYNCS", "ACQBULK", "PREEXIT", "FFMA2", "FMUL2", "FADD2", "FFMA", "MUFU", "REDG", "RED", "ATOMS", "ATOMG", "LDGMC", "STG", "LDG",
       "LDS", "STS", "SHFL", "VOTE", "BAR", "LDL", "STL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    kern, rows = None, collections.OrderedDict()
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", name).replace("gigs::", "").replace("(anonymous namespace)::", "")
            rows[kern] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", ln)
        if m and kern:
            op = m.group(1)
            rows[kern]["TOTAL"] += 1
            for o in OPS:
                if op == o or (o in ("RED", "REDG") and op.startswith(o) and not (o == "RED" and op.startswith("REDG"))):
                    rows[kern][o] += 1
                    break
    cols = ["TOTAL"] + OPS
    w = max(len(k) for k in rows) + 2
    print("SASS mnemonic counts per kernel of gi-gs_b200/lib/libgigs_b200.so (static instruction counts, sm_100a)")
    print("kernel".ljust(w) + "".join(c.rjust(8) for c in cols))
    for k, c in rows.items():
        print(k.ljust(w) + "".join(str(c.get(x, 0)).rjust(8) for x in cols))


if __name__ == "__main__":
    main()
