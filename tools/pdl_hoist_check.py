#!/usr/bin/env python
"""List, per kernel in libia2c_b200.so, the global loads that SASS places BEFORE the first ACQBULK (griddepcontrol.wait).

A load there reads memory the programmatic-dependent-launch primary may not have written yet.  ptxas is free to hoist
LDG.E.CONSTANT (ld.global.nc: __ldg, or a const __restrict__ kernel parameter) above the wait, so every entry printed
here has to be a buffer that no kernel of the same stream writes (filter tables, ...).  Usage: tools/pdl_hoist_check.py [lib]"""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "ia2c_b200/libia2c_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
name, body = None, []
funcs = []
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if name:
            funcs.append((name, body))
        name, body = m.group(1), []
    elif name and re.search(r"/\*[0-9a-f]{4}\*/", line):
        body.append(line.strip())
if name:
    funcs.append((name, body))
for name, body in funcs:
    acq = [i for i, l in enumerate(body) if "ACQBULK" in l]
    if not acq:
        continue
    early = [l for l in body[:acq[0]] if re.search(r"\bLDG", l)]
    demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"\(anonymous namespace\)::|ia2c::", "", demangled).split("(")[0]
    print(f"{short}: ACQBULK at instr {acq[0]}/{len(body)}, {len(early)} global loads before it")
    for l in early:
        print("    ", re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l))
