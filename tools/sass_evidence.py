"""Per-kernel counts of the SASS mnemonics that prove which hardware paths the built library uses
(cuobjdump -sass of gr-ltetrigger_b200/lib/libltetrigger_b200.so; B200_PROFILING.md names the mnemonics).
  python tools/sass_evidence.py > profiles/sass_evidence_rNN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gr-ltetrigger_b200", "lib", "libltetrigger_b200.so")
MNEMONICS = ["UTCIMMA", "LDTM", "UTMALDG", "UTCBAR", "UBLKCP", "SYNCS", "FFMA2", "LDGSTS"]


def per_kernel(lib=LIB):
    """{demangled kernel name without its parameter list: Counter(mnemonic -> instructions)}"""
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    rows, cur = [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = (m.group(1), collections.Counter())
            rows.append(cur)
        elif cur:
            for k in MNEMONICS:
                if re.search(r"\b" + k + r"\b", line):
                    cur[1][k] += 1
    names = subprocess.run(["cu++filt"] + [r[0] for r in rows], capture_output=True, text=True, check=True).stdout.splitlines()
    out = {}
    for (_, c), d in zip(rows, names):
        d = d.replace("void ", "").replace("(int)", "")
        d = re.sub(r">\(.*$", ">", d) if ">(" in d else re.sub(r"\(.*$", "", d)
        out[d] = c
    return out


if __name__ == "__main__":
    print("# SASS evidence (cuobjdump -sass gr-ltetrigger_b200/lib/libltetrigger_b200.so, sm_100a): instructions per kernel")
    print("# UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit,")
    print("# UBLKCP = cp.async.bulk, SYNCS = mbarrier, FFMA2 = packed FP32 FMA, LDGSTS = cp.async")
    print("# kernel | " + " | ".join(MNEMONICS))
    for name, c in sorted(per_kernel(sys.argv[1] if len(sys.argv) > 1 else LIB).items()):
        print(name + " | " + " | ".join(str(c[k]) for k in MNEMONICS))
