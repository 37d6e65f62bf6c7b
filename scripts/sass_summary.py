#!/usr/bin/env python
"""Opcode histogram per kernel of libsib_b200.so (cuobjdump -sass), written to profiles/<round>_sass_summary.txt:
the committed proof of which Blackwell instructions the hand-written kernels use (UTCHMMA = tcgen05.mma, UTMALDG /
UTMASTG = TMA loads / stores, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier, ELECT, ...).
Usage: python scripts/sass_summary.py [round-tag]      (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "speech-inpainting_b200", "libsib_b200.so")
KEY = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "ELECT", "UCGABAR_ARV",
       "UCGABAR_WAIT", "HMMA", "FFMA", "FADD", "FMUL", "HFMA2", "HMNMX2", "MUFU", "LDS", "STS", "LDG", "STG", "LDSM", "SHFL", "BAR",
       "FENCE", "MEMBAR", "ACQBULK", "UTMACMDFLUSH", "UTMACCTL", "PREEXIT", "ATOMS", "RED", "F2FP"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["__total__"] += 1
    demangled = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines() \
        if kernels else []
    if len(demangled) != len(kernels):
        demangled = list(kernels)
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt")
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True).stdout.strip()
    tot = collections.Counter()
    with open(path, "w") as f:
        f.write(f"# cuobjdump -sass speech-inpainting_b200/libsib_b200.so (sm_100a), tree at git {head}: instruction counts per kernel\n")
        f.write("# columns: total instructions, then the opcodes that matter for the Blackwell claim (absent = 0)\n\n")
        for (mangled, c), name in zip(kernels.items(), demangled):
            short = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
            short = re.sub(r"\(CUtensorMap_st.*", "(...)", short)[:110]
            hits = "  ".join(f"{k}={c[k]}" for k in KEY if c.get(k))
            f.write(f"{short}\n    total={c['__total__']}  {hits}\n")
            tot.update(c)
        f.write("\n# whole library\n    " + "  ".join(f"{k}={tot[k]}" for k in ["__total__"] + KEY if tot.get(k)) + "\n")
    print(open(path).read()[-1200:])


if __name__ == "__main__":
    main()
