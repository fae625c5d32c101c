#!/usr/bin/env python3
"""Summarise build/csrc/*.ptxas.log: registers / spills / smem per kernel."""
import glob, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for f in sorted(glob.glob(os.path.join(ROOT, "build/csrc/*.ptxas.log"))):
    txt = open(f).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", txt, re.S):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void sr::", "")
        print("%-70s regs=%-4s stack=%-5s spill_st=%-5s spill_ld=%-5s" % (name[:70], m.group(5), m.group(2), m.group(3), m.group(4)))
