#!/usr/bin/env python
"""SASS evidence per kernel of libitsolv_b200.so: how many TMA bulk copies (UBLKCP), FP64 tensor-core instructions (DMMA),
FP64 FMAs (DFMA), cp.async pieces (LDGSTS), mbarrier operations (SYNCS), 128-bit global loads/stores and shared-memory
loads each kernel contains, plus registers and spills from the cubin's resource usage.
    python tools/sass_summary.py > profiles/sass_summary_rNN.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "iterative_solver_b200", "lib", "libitsolv_b200.so")
MNEMONICS = ["UBLKCP", "UTMALDG", "DMMA", "DFMA", "DMUL", "DADD", "LDGSTS", "SYNCS", "LDG.E.128", "STG.E.128", "LDS.128",
             "LDS.64", "MUFU.RCP64H", "ATOMG", "MEMBAR", "SHFL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["instructions"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + ".") or op.startswith(mn):
                    counts[cur][mn] += 1
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
        if m and fn:
            usage[fn] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    names = demangle(list(counts))
    cols = ["instructions"] + MNEMONICS
    print("# cuobjdump -sass / -res-usage of iterative_solver_b200/lib/libitsolv_b200.so (sm_100a), per kernel")
    print("# " + " | ".join(["kernel", "regs", "static smem", "local (spill) bytes"] + cols))
    for k, c in counts.items():
        short = re.sub(r"\(.*\)$", "", names.get(k, k)).replace("itsolv::", "")
        r = usage.get(k, ("?", "?", "?"))
        print(" | ".join([short, str(r[0]), str(r[1]), str(r[2])] + [str(c.get(col, 0)) for col in cols]))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("# kernels:", len(counts), "| with UBLKCP:", sum(1 for c in counts.values() if c["UBLKCP"]), "| with DMMA:",
          sum(1 for c in counts.values() if c["DMMA"]), "| with LDGSTS:", sum(1 for c in counts.values() if c["LDGSTS"]),
          "| UTCMMA / tcgen05 (no f64 kind exists):", sum(1 for k in counts if "UTCMMA" in sass and False))


if __name__ == "__main__":
    main()
