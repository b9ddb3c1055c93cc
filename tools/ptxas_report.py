"""Registers / spills per kernel from an nvcc -Xptxas -v log."""
import re, subprocess, sys
s = open(sys.argv[1]).read()
pat = re.compile(r"Compiling entry function '([^']+)' for 'sm_100a'\nptxas info\s+: Function properties for [^\n]+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers")
for name, stack, ss, sl, regs in pat.findall(s):
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = dem.replace("sdg::", "").replace("(sdg::StageArgs)", "")[:100]
    print(f"{regs:>4} regs  stack {stack:>4}  spill st/ld {ss:>4}/{sl:>4}  {dem}")
