#!/usr/bin/env python3
"""Instruction mix per kernel from cuobjdump -sass (static counts; loops counted once).
Usage: tools/sass_mix.py <obj-or-so> [kernel-name-regex]"""
import collections, re, subprocess, sys
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IMUL"}
def pipe(op):
    base = op.split(".")[0]
    if base in FMA: return "fma"
    if base in ("IADD3", "IADD", "LOP3", "SHF", "PRMT", "VIMNMX", "IMNMX", "ISETP", "SEL", "LEA", "VIADD", "IABS", "PLOP3", "UIADD3", "ULOP3", "USHF", "MOV", "FSEL", "VIADDMNMX", "IADD3.X"): return "alu"
    if base in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "LDSM", "LDC", "ATOMS", "RED", "LDGSTS"): return "lsu"
    return "other"
def main():
    path = sys.argv[1]; pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur = None; mixes = collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); mixes[cur] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur: mixes[cur][m.group(1)] += 1
    for k, c in mixes.items():
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        if pat and not pat.search(name): continue
        tot = sum(c.values()); byp = collections.Counter()
        for op, n in c.items(): byp[pipe(op)] += n
        print("== %s\n   total %d  " % (name[:110], tot) + "  ".join("%s=%d" % kv for kv in byp.most_common()))
        print("   " + "  ".join("%s:%d" % kv for kv in c.most_common(22)))
main()
