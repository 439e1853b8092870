#!/usr/bin/env python
"""Per-kernel SASS summary of libia2c_b200.so: architecture, registers / spills (from a verbose rebuild), and counts of
the instructions that characterise the design (FFMA2 packed fp32, DFMA fp64, bulk async copies, mbarrier, cp.async,
PDL, tensor-core / TMEM opcodes).   usage: python tools/sass_summary.py > profiles/r02_sass_summary.md"""
import os
import re
import subprocess
import sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ia2c_b200 import build as B  # noqa: E402

WATCH = ["FFMA2", "FMUL2", "FADD2", "FFMA", "DFMA", "DADD", "DMUL", "MUFU", "LDG", "STG", "LDS", "STS", "LDGSTS", "UBLKCP", "SYNCS",
         "SHFL", "REDUX", "MATCH", "ATOMS", "ACQBULK", "UTMALDG", "UTCHMMA", "UTCQMMA", "LDTM", "HMMA", "IMMA", "BAR"]


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except Exception:
        return n


def main():
    regs = {}
    for src in B.SOURCES:
        r = subprocess.run([B.NVCC, *B.FLAGS, "-Xptxas=-v", "-c", os.path.join(B.CSRC, src), "-o", "/dev/null"], capture_output=True, text=True)
        cur = None
        for line in r.stderr.splitlines():
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
            if m:
                cur = m.group(1)
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur:
                regs.setdefault(cur, {})["spill"] = (int(m.group(2)), int(m.group(3)))
            m = re.search(r"Used (\d+) registers", line)
            if m and cur:
                regs.setdefault(cur, {})["regs"] = int(m.group(1))
    sass = subprocess.run(["cuobjdump", "-sass", B.LIB], capture_output=True, text=True).stdout
    arch = Counter(re.findall(r"arch = (sm_\w+)", sass))
    counts, cur = defaultdict(Counter), None
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    counts[cur][w] += 1
    print("# SASS summary of libia2c_b200.so (round 2)\n")
    print(f"Built with `{' '.join(B.FLAGS)}`; cubins: {dict(arch)}; {len(counts)} kernels.\n")
    print("No tensor-core (`UTC*MMA`, `HMMA`), TMEM (`LDTM`) or tensor-map TMA (`UTMALDG`) opcode is expected: hidden_size = 6 leaves no dense")
    print("contraction (SURVEY.md §7.3).  The Blackwell-specific instructions in use are the packed fp32 pipe (`FFMA2`/`FMUL2`/`FADD2`),")
    print("bulk async copies global->shared on an mbarrier (`UBLKCP` + `SYNCS`, the single-pass update kernel), `LDGSTS` (cp.async ring of")
    print("the belief kernel), `REDUX`/`MATCH` warp reductions and programmatic dependent launch.\n")
    cols = ["FFMA2", "FMUL2", "FADD2", "FFMA", "DFMA", "MUFU", "LDGSTS", "UBLKCP", "SYNCS", "SHFL", "REDUX", "ATOMS", "BAR"]
    print("| kernel | regs | spill st/ld (B) | SASS instrs | " + " | ".join(cols) + " |")
    print("|---|---|---|---|" + "---|" * len(cols))
    for fn in sorted(counts, key=lambda f: -counts[f]["_total"]):
        name = demangle(fn)
        name = re.sub(r"ia2c::\(anonymous namespace\)::", "", name)
        name = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name)[:70]
        rg = regs.get(fn, {})
        sp = rg.get("spill", ("?", "?"))
        print(f"| `{name}` | {rg.get('regs', '?')} | {sp[0]}/{sp[1]} | {counts[fn]['_total']} | " + " | ".join(str(counts[fn][c]) for c in cols) + " |")
    tot = Counter()
    for fn in counts:
        tot.update(counts[fn])
    print("\nTotals over all kernels: " + ", ".join(f"{w} {tot[w]}" for w in WATCH if tot[w]))


if __name__ == "__main__":
    main()
