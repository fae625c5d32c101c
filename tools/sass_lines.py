#!/usr/bin/env python3
"""Static SASS instruction counts per source line (with the inline chain), split by issue pipe.

  tools/sass_lines.py <object-or-cubin> <kernel-regex> <source-file-basename> [--inner]

Every instruction of the kernel is attributed to the OUTERMOST frame of its inline chain that lies in the given
source file (--inner: the innermost such frame), so that e.g. `gl_fused6.cuh` shows how many instructions each
statement of ring_mul_fused6 expands to.  Loops are counted once (static counts): multiply by the trip counts.
Needs -lineinfo at compile time.  Uses cuobjdump -xelf + nvdisasm -gi."""
import collections
import os
import re
import subprocess
import sys
import tempfile

FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IMUL"}
ALU = {"IADD3", "IADD", "LOP3", "SHF", "PRMT", "VIMNMX", "IMNMX", "ISETP", "SEL", "LEA", "VIADD", "IABS", "PLOP3",
       "MOV", "FSEL", "VIADDMNMX", "UIADD3", "ULOP3", "USHF", "UMOV", "USEL", "UISETP", "ULEA", "UIMAD", "UPLOP3"}
LSU = {"LDS", "STS", "LDG", "STG", "LDL", "STL", "LDSM", "LDC", "LDCU", "ATOMS", "ATOMG", "RED", "LDGSTS", "UBLKCP",
       "SYNCS"}


def pipe(op):
    base = op.split(".")[0]
    if base == "IMAD" and ".WIDE" in op:
        return "imad_wide"
    if base in FMA:
        return "fma"
    if base in ALU:
        return "alu"
    if base in LSU:
        return "lsu"
    return "other"


def main():
    path, kpat, src = sys.argv[1], re.compile(sys.argv[2]), sys.argv[3]
    inner = "--inner" in sys.argv
    tmp = tempfile.mkdtemp()
    if path.endswith(".cubin"):
        cubins = [os.path.abspath(path)]
    else:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(path)], cwd=tmp, check=True,
                       stdout=subprocess.DEVNULL)
        cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    for cubin in cubins:
        txt = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
        cur, chain, pending = None, [], []
        stats = {}
        for line in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", line)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                cur = name if kpat.search(name) else None
                chain, pending = [], []
                if cur:
                    stats[cur] = collections.defaultdict(collections.Counter)
                continue
            if cur is None:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
            if m:
                pending.append((os.path.basename(m.group(1)), int(m.group(2))))
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
            if m:
                if pending:
                    chain, pending = pending, []
                frames = [f for f in chain if f[0] == src]
                key = "(elsewhere)" if not frames else "%s:%d" % (frames[0] if inner else frames[-1])
                stats[cur][key][pipe(m.group(1))] += 1
        for name, tab in stats.items():
            print("== " + name[:120])
            tot = collections.Counter()
            def keyf(k):
                m = re.search(r":(\d+)$", k)
                return int(m.group(1)) if m else 10 ** 9
            for key in sorted(tab, key=keyf):
                c = tab[key]
                tot.update(c)
                print("  %-24s total %5d  alu %5d  fma %4d  imad_wide %4d  lsu %4d  other %4d" % (
                    key, sum(c.values()), c["alu"], c["fma"], c["imad_wide"], c["lsu"], c["other"]))
            print("  %-24s total %5d  alu %5d  fma %4d  imad_wide %4d  lsu %4d  other %4d" % (
                "TOTAL (static)", sum(tot.values()), tot["alu"], tot["fma"], tot["imad_wide"], tot["lsu"], tot["other"]))


main()
