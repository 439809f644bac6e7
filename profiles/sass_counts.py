#!/usr/bin/env python
"""Instruction-class counts per kernel of the shipped library (cuobjdump -sass), so the Blackwell-native claim can be
audited from the repository: tcgen05 MMAs (UTCHMMA / .2CTA), TMEM loads/stores (LDTM / STTM), TMA loads
(UTMALDG), TMA reductions (UTMAREDG), tcgen05 commits (UTCBAR), legacy tensor-core ops (HMMA — must be 0), MUFU.

    python profiles/sass_counts.py [path/to/libflexam_b200.so] > profiles/sass_r2.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASSES = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAREDG", "UTMASTG", "SYNCS", "HMMA",
           "MUFU.EX2", "MUFU", "ATOM", "RED", "LDG", "STG", "BAR.SYNC"]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "flexam_b200", "csrc", "libflexam_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    name = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        c = counts[name]
        c["total"] += 1
        for k in CLASSES:
            if k == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    c[k] += 1
            elif op == k or op.startswith(k + "."):
                c[k] += 1
    cols = ["total"] + CLASSES
    print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)} — instruction-class counts per kernel (static SASS, sm_100a)")
    print("| kernel | " + " | ".join(cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for n, c in counts.items():
        print(f"| `{n}` | " + " | ".join(str(c.get(k, 0)) for k in cols) + " |")
    tc = sum(c["UTCHMMA"] for c in counts.values())
    print(f"\nTotals: UTCHMMA {tc} (of which .2CTA {sum(c['UTCHMMA.2CTA'] for c in counts.values())}), "
          f"HMMA {sum(c['HMMA'] for c in counts.values())}, LDTM {sum(c['LDTM'] for c in counts.values())}, "
          f"UTMALDG {sum(c['UTMALDG'] for c in counts.values())}, UTMAREDG {sum(c['UTMAREDG'] for c in counts.values())}")


if __name__ == "__main__":
    main()
