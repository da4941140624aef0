"""SASS mnemonic counts per kernel of the built library (the evidence that the hot kernels use tcgen05 / TMA / st.async):
    python profiles/sass_mnemonics.py > profiles/r2_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "genvox_b200", "lib", "libgenvox_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WANT = ("UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "HMMA", "STAS", "SYNCS", "LDGSTS", "LDSM", "MEMBAR", "REDG", "ATOMG", "UTCATOMSWS")
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    if kern:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and m.group(1).split(".")[0] in WANT:
            counts[kern][m.group(1).split(".")[0]] += 1
print("SASS mnemonic counts per kernel (cuobjdump -sass genvox_b200/lib/libgenvox_b200.so, round 2, final).")
print("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA 1-D bulk copy), UTMALDG = TMA tensor-map load,")
print("HMMA = mma.sync, LDSM = ldmatrix, LDGSTS = cp.async, STAS = st.async (DSMEM push with mbarrier complete_tx), SYNCS = mbarrier ops.\n")
for k, c in counts.items():
    if any(c[w] for w in ("UTCHMMA", "HMMA", "UBLKCP", "UTMALDG", "STAS", "LDGSTS")):
        print(f"{k[:70]:70s} " + "  ".join(f"{w}:{c[w]}" for w in WANT if c[w]))
