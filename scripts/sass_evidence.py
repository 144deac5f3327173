#!/usr/bin/env python
"""Counts the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md, "What proves a Blackwell-native
kernel") per kernel of libwisb200.so -> profiles/r01_sass_evidence.csv.  Runs without a GPU (cuobjdump only)."""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "willow_inference_server_b200", "libwisb200.so")
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r01_sass_evidence.csv")
GROUPS = collections.OrderedDict([
    ("tcgen05_mma(UTC*MMA)", re.compile(r"\bUTC[A-Z]*MMA")), ("tcgen05_ld_st(LDTM/STTM)", re.compile(r"\b(LDTM|STTM)")),
    ("tma_tensor(UTMALDG/UTMASTG)", re.compile(r"\b(UTMALDG|UTMASTG)")), ("tma_bulk(UBLKCP)", re.compile(r"\bUBLKCP")),
    ("mbarrier(SYNCS)", re.compile(r"\bSYNCS")), ("cp_async(LDGSTS)", re.compile(r"\bLDGSTS")),
    ("warp_mma(HMMA)", re.compile(r"\bHMMA")), ("ldmatrix(LDSM)", re.compile(r"\bLDSM")), ("fp32_fma(FFMA)", re.compile(r"\bFFMA")),
])
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
name, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "").replace("wisb::", "")
        name = re.sub(r"\((?!.*>).*$", "", name) if ">" in name else re.sub(r"\(.*", "", name)
        name = re.sub(r">\(.*$", ">", name)
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and name:
        total[name] += 1
        for g, rx in GROUPS.items():
            if rx.match(m.group(1)):
                counts[name][g] += 1
with open(OUT, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "sass_instructions"] + list(GROUPS))
    for k, c in counts.items():
        w.writerow([k, total[k]] + [c[g] for g in GROUPS])
print(open(OUT).read())
