#!/usr/bin/env python
"""SASS evidence per hot kernel (no GPU needed): counts of the Blackwell-native mnemonics in the built library.
    python tools/sass_evidence.py > profiles/r2_sass_evidence.txt
UBLKCP = cp.async.bulk (1-D TMA), UTMALDG = cp.async.bulk.tensor (tiled TMA), UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld,
DMMA = mma.sync f64, REDG = red.global, ATOMG = atom.global."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "symtensor_b200", "lib", "libsymtensor_b200.so")
MNEMONICS = ["UBLKCP", "UTMALDG", "UTMASTG", "UTCHMMA", "LDTM", "STTM", "UTCBAR", "DMMA", "HMMA", "REDG", "ATOMG", "SYNCS", "LDGSTS", "ACQBULK"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for mn in MNEMONICS:
            if re.search(r"\b" + mn + r"\b|\b" + mn + r"\.", line):
                counts[cur][mn] += 1
        counts[cur]["_instructions"] += 1 if re.search(r"/\*[0-9a-f]{4,6}\*/\s+[A-Z@]", line) else 0
    print("SASS mnemonic counts per kernel of", os.path.relpath(LIB, ROOT), "(cuobjdump -sass; sm_100a)")
    print("%-64s %8s  %s" % ("kernel", "instrs", "Blackwell-native mnemonics"))
    for k, c in counts.items():
        tags = "  ".join(f"{mn}:{c[mn]}" for mn in MNEMONICS if c[mn])
        print("%-64s %8d  %s" % (k[:64], c["_instructions"], tags))


if __name__ == "__main__":
    sys.exit(main())
