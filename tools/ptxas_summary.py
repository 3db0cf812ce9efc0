#!/usr/bin/env python
"""Registers / spills / shared memory per kernel from `nvcc -Xptxas -v` (the in-tree build in verbose mode).

    python tools/ptxas_summary.py [filter-regex] [-DNAME ...]
"""
import re
import subprocess
import sys

args = [a for a in sys.argv[1:] if not a.startswith("-D")]
defs = [a for a in sys.argv[1:] if a.startswith("-D")]
out = subprocess.run([sys.executable, "-m", "fast_speech_enhancement_metrics_b200.build", "--verbose"] + defs,
                     capture_output=True, text=True).stderr
pat = re.compile(args[0]) if args else None
name = None
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip()
        spill = ""
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and name:
        spill = "stack %s spill st/ld %s/%s" % m.groups()
        continue
    m = re.search(r"Used (\d+) registers.*?(?:, (\d+) bytes smem)?", line)
    if m and name:
        smem = re.search(r"(\d+) bytes smem", line)
        if pat is None or pat.search(name):
            print("%-60s regs %3s  smem %6s  %s" % (name[:60], m.group(1), smem.group(1) if smem else "0", spill))
        name = None
