#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove which hardware paths the library's kernels use (cuobjdump -sass on
the in-tree libb200rag.so; no GPU needed).  Writes profiles/<tag>_sass_summary.txt.

    python tools/sass_summary.py [tag]

UTCHMMA = tcgen05.mma (.2CTA: cta_group::2), UTMALDG = cp.async.bulk.tensor (TMA tile loads), UBLKCP = cp.async.bulk
(1-D bulk copies), LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit, SYNCS = mbarrier operations,
ATOMS.ADD = native shared-memory integer atomics, IDP.2A = dp2a (16-bit x 8-bit integer dot-product steps of the 8-bit
candidate scan), LDG.E.256 = 256-bit global loads."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "audio-rag_b200", "b200rag", "libb200rag.so")
WATCH = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "ATOMS.ADD", "ATOMS.CAST", "ATOMS.MAX",
         "ATOMG", "RED", "IDP.2A", "LDG.E.256", "LDG.E.128", "LDS.128", "STS.128", "FFMA", "DFMA", "DADD", "DMUL", "SHFL", "BAR.SYNC",
         "MEMBAR", "CCTL", "ST.E.STRONG.SYS", "LD.E.STRONG.SYS"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kern = None
    counts = collections.OrderedDict()
    total = collections.Counter()
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*$", "", kern)
            counts[kern] = collections.Counter()
            continue
        if kern is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m:
            continue
        op = m.group(1)
        total[kern] += 1
        toks = op.split(".")
        for w in WATCH:
            wt = w.split(".")
            if toks[0] == wt[0] and all(t in toks[1:] for t in wt[1:]):
                counts[kern][w] += 1
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt")
    with open(path, "w") as f:
        f.write(f"# cuobjdump -sass {os.path.relpath(SO, ROOT)} -- per-kernel counts of selected mnemonics (tools/sass_summary.py)\n")
        f.write("# a mnemonic matches when its base and every listed modifier appear (UTCHMMA counts include the .2CTA forms; "
                "LDG.E.256 matches LDG.E.ENL2.256.CONSTANT)\n")
        for k, c in counts.items():
            if not total[k]:
                continue
            keep = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
            f.write(f"{k}\n    instructions={total[k]} {keep}\n")
    print(path)


if __name__ == "__main__":
    main()
