"""Registers / spills per kernel of one .cu file (ptxas -v), for checking a change before spending GPU time.

  python tools/ptxas_report.py mmhand_b200/csrc/elementwise.cu [filter]
"""
import re
import subprocess
import sys

src = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xptxas", "-v", "-c",
                      src, "-o", "/dev/null"], capture_output=True, text=True).stderr
name = None
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        stack = None
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        stack = m.groups()
    m = re.search(r"Used (\d+) registers", line)
    if m and name and flt in name:
        print("%-90s regs %3s  stack/spill st/ld %s" % (name[-90:], m.group(1), "/".join(stack or ())))
