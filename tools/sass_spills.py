#!/usr/bin/env python
"""Where do the register spills of a kernel sit?  tools/sass_spills.py file.cubin 'substring of the mangled kernel name'
Prints, per source line (nvcc -lineinfo), the number of LDL / STL instructions in the kernel's SASS (nvdisasm -g)."""
import re
import subprocess
import sys
from collections import Counter

cubin, key = sys.argv[1], sys.argv[2]
txt = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(txt) if l.startswith(".text.") and key in l)
end = next((i for i in range(start + 1, len(txt)) if txt[i].startswith(".text.") or txt[i].startswith(".section")), len(txt))
cur, c, n = None, Counter(), 0
for l in txt[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    if re.search(r"^\s+/\*[0-9a-f]{4,}\*/", l):
        n += 1
    m2 = re.search(r"\b(LDL|STL)\b", l)
    if m2:
        c[(cur, m2.group(1))] += 1
print("instructions:", n)
for k, v in sorted(c.items(), key=lambda kv: (kv[0][0] or ("", 0))):
    print(k, v)
