#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass).  usage: tools/sass_evidence.py [libdvo.so] > profiles/xxx.md"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "droplet_visual_odometry_b200/libdvo.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["UTCIMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "REDUX", "VABSDIFF4", "IDP", "VIMNMX3", "POPC", "DFMA", "DMUL",
        "DADD", "ATOMS", "UCGABAR", "LDS", "LDG", "STG", "SHFL"]
pretty = {"ILb1EEE": "<1>", "ILb0EEE": "<0>", "ILi1EEE": "<1>", "ILi3EEE": "<3>"}
rows, cur = collections.OrderedDict(), None
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        short = re.search(r"(k_[a-z0-9_]+)", name)
        tag = next((v for k, v in pretty.items() if k in name), "")
        cur = (short.group(1) if short else name) + tag
        rows[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        op = m.group(1)
        rows[cur]["total"] += 1
        for c in cols:
            if op.startswith(c):
                rows[cur][c] += 1
print("# SASS evidence (cuobjdump -sass libdvo.so, sm_100a): instruction mnemonics per kernel\n")
print("| kernel | total | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for k, c in rows.items():
    print("| %s | %d | " % (k, c["total"]) + " | ".join(str(c[x]) for x in cols) + " |")
print("""
UTCIMMA = tcgen05.mma kind::i8 (5th-generation tensor core, int8), LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit,
UBLKCP = cp.async.bulk (bulk async copy global -> shared), UTMALDG = TMA tile load (cp.async.bulk.tensor), SYNCS = mbarrier,
REDUX = redux.sync, VABSDIFF4 = packed byte |a-b|, IDP = dp2a, VIMNMX3 = three-input min/max, UCGABAR = cluster barrier.
Template arguments: k_fast_nms<1> = TMA staging, k_nn_tensor<1> = ratio matcher (runner-up kept), k_nn<1> = split cross-check,
k_ransac<1> = cluster mode, k_ingest<3> = BGR.""")
