"""Per-kernel counts of the SASS mnemonics that prove which hardware paths the library uses
(B200_PROFILING.md): DMMA (FP64 tensor), UTCIMMA/UTCHMMA... (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UBLKCP
(cp.async.bulk), UTMALDG (TMA tensor loads), SYNCS (mbarrier), plus registers / shared memory from -res-usage.

    python tools/sass_summary.py > profiles/sass_summary.txt      (no GPU needed: reads the built .so)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bayesopt_smart_b200", "csrc", "libbo_b200.so")
MNEMONICS = ["DMMA", "DFMA", "DADD", "DMUL", "UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCCP",
             "UBLKCP", "UTMALDG", "SYNCS", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "MUFU", "ATOM", "RED", "VOTE",
             "IMAD", "PRMT"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["total"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn):
                    counts[cur][mn] += 1
                    break
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage, cur = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
    names = demangle(order)
    arch = re.search(r"arch = (sm_\w+)", sass)
    print(f"# {os.path.relpath(LIB, ROOT)}  ({arch.group(1) if arch else '?'}); cuobjdump -sass / -res-usage; "
          "instruction counts per kernel (static, all template instances listed)")
    cols = [mn for mn in MNEMONICS if any(c[mn] for c in counts.values())]
    print("kernel".ljust(64) + "".join(c.rjust(9) for c in ["instr"] + cols + ["regs", "smem_B"]))
    for fn in sorted(order, key=lambda f: names[f]):
        short = names[fn].replace("(anonymous namespace)::", "").replace("void ", "").replace("bo::", "")
        short = re.sub(r"\((?!.*>).*$", "", short)  # drop the argument list, keep template arguments
        c = counts[fn]
        reg, smem = usage.get(fn, (0, 0))
        print(short[:63].ljust(64) + str(c["total"]).rjust(9) + "".join(str(c[mn]).rjust(9) for mn in cols) +
              str(reg).rjust(9) + str(smem).rjust(9))


if __name__ == "__main__":
    sys.exit(main())
