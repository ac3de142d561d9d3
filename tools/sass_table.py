"""Per-kernel SASS evidence table (runs anywhere: only needs cuobjdump on the built library).

Counts the mnemonics that prove which hardware path a kernel uses (B200_PROFILING.md, "What proves a Blackwell-native kernel"):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor copies, UBLKCP = 1-D bulk copy,
SYNCS = mbarrier, HMMA = legacy mma.sync, LDGSTS = cp.async, REDG/ATOMG = global reductions/atomics.
Usage: python tools/sass_table.py [out.md]"""
import re, subprocess, sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "detikzify_b200" / "csrc" / "libdtk_b200.so"
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "LDGSTS", "LDSM", "REDG", "ATOMG", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            kernels[cur]["_n"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_n"] += 1
        base = op.split(".")[0]
        for k in KEYS:
            if base == k or (k in ("UTCHMMA", "UTCQMMA") and base.startswith(k)):
                kernels[cur][k] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for (name, c), dm in zip(kernels.items(), demangled):
        short = dm.replace("(anonymous namespace)::", "").replace("dtk::", "").replace("void ", "")
        short = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", short)   # drop the argument list, keep template arguments
        rows.append((short, c))
    out = ["# SASS evidence per kernel (`tools/sass_table.py`, cuobjdump -sass of libdtk_b200.so)", "",
           "| kernel | instr | " + " | ".join(KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
    for short, c in rows:
        out.append(f"| `{short}` | {c['_n']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 1:
        Path(sys.argv[1]).write_text(text)
    print(text)


if __name__ == "__main__":
    main()
