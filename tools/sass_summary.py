#!/usr/bin/env python
"""Per-kernel count of the SASS mnemonics that show a Blackwell-native kernel (B200_PROFILING.md):
tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM / STTM, TMA -> UTMALDG / UTMASTG, tcgen05.commit -> UTCBAR,
legacy tensor path -> HMMA.   python tools/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "repurpose_b200/librepurpose_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "MUFU.EX2", "HMMA", "UCGABAR", "USETMAXREG"]
cur, tab = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name.replace("rp::(anonymous namespace)::", "").replace("void ", ""))
        cur = tab.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        cur["instructions"] += 1
        for k in keys:
            if op.startswith(k):
                cur[k] += 1
print(f"SASS mnemonic counts per kernel of {lib} (cuobjdump -sass, sm_100a)\n")
print(f"{'kernel':58s} {'instr':>7s} " + " ".join(f"{k:>9s}" for k in keys))
tot = collections.Counter()
for name, c in tab.items():
    print(f"{name[:58]:58s} {c['instructions']:7d} " + " ".join(f"{c[k]:9d}" for k in keys))
    tot.update(c)
print(f"{'total':58s} {tot['instructions']:7d} " + " ".join(f"{tot[k]:9d}" for k in keys))
