"""SASS instruction count / code bytes per kernel of an object file: python tools/sass_sizes.py file.o"""
import re, subprocess, sys
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
name, n, out = None, 0, []
for l in txt.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        if name: out.append((n, name))
        name, n = m.group(1), 0
    elif re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", l):
        n += 1
out.append((n, name))
for n, name in sorted(out):
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().replace("sdg::", "")[:90]
    print(f"{n:6d} instr {n * 16 // 1024:4d} KB  {dem}")
