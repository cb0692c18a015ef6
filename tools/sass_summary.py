"""Per-kernel SASS evidence from the built library (no GPU needed): `cuobjdump -sass dupl_b200/libdupl.so` reduced to the
mnemonics that prove which hardware path a kernel takes (B200_PROFILING.md):

  UTCHMMA / UTCQMMA   tcgen05.mma            LDTM / STTM   tcgen05.ld / tcgen05.st (tensor memory)
  UTMALDG / UTMASTG   TMA bulk tensor copies  UTCBAR        tcgen05.commit -> mbarrier      SYNCS   mbarrier ops
  LDS.128 / LDG.E.128 vector width of the shared / global accesses of the HBM-bound kernels

usage: python tools/sass_summary.py [path/to/libdupl.so] > profiles/rNN_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "LDS.128", "LDS", "LDG.E.128", "LDG",
         "STG.E.128", "STG", "FFMA", "MUFU", "HMMA", "BAR"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "dupl_b200", "libdupl.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = collections.Counter()
            kernels[m.group(1)] = cur
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["_n"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                cur[w] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("| kernel | SASS instr | " + " | ".join(WATCH) + " |")
    print("|---|---|" + "---|" * len(WATCH))
    rows = []
    for (mangled, c), name in zip(kernels.items(), demangle):
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("dupl::", "")
        rows.append((name, c))
    for name, c in sorted(rows):
        print(f"| `{name}` | {c['_n']} | " + " | ".join(str(c[w]) if c[w] else "" for w in WATCH) + " |")


if __name__ == "__main__":
    main()
