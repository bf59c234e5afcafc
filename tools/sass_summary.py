#!/usr/bin/env python
"""Per-kernel count of the SASS mnemonics that prove the Blackwell-native path (B200_PROFILING.md "What proves a
Blackwell-native kernel"): UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA), MUFU.TANH, and the
legacy HMMA (must be 0).  Runs `cuobjdump -sass` on the built libnppc_b200.so (no GPU needed):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "generative-audio_b200", "libnppc_b200.so")
PAT = collections.OrderedDict([("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
                               ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("MUFU.TANH", r"MUFU\.TANH"),
                               ("HMMA(legacy)", r"\bHMMA\b")])


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur:
            for k, rx in PAT.items():
                if re.search(rx, line):
                    counts[cur][k] += 1
    total = collections.Counter()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a) — instruction counts per kernel; kernels without any are omitted")
    print("kernel | " + " | ".join(PAT))
    for k, c in counts.items():
        if sum(c.values()):
            print(f"{k[:110]} | " + " | ".join(str(c[p]) for p in PAT))
            total.update(c)
    print("TOTAL | " + " | ".join(str(total[p]) for p in PAT))


if __name__ == "__main__":
    sys.exit(main())
