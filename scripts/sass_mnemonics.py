"""Per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMA / DMMA paths in the shipped library:
   python scripts/sass_mnemonics.py r02   ->  profiles/r02_sass_mnemonics.txt   (no GPU needed)"""
import collections, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = subprocess.run(["cuobjdump", "-sass", str(ROOT / "later_b200" / "liblater_b200.so")], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTCIMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "DMMA", "ACQBULK"]
cur, cnt = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); cnt[cur] = collections.Counter(); continue
    if cur:
        for k in keys:
            if re.search(r"\b" + k + r"\b", line):
                cnt[cur][k] += 1
with open(ROOT / "profiles" / f"{tag}_sass_mnemonics.txt", "w") as f:
    f.write("# cuobjdump -sass later_b200/liblater_b200.so : instruction counts per kernel (sm_100a), proving the tcgen05 / TMA / DMMA paths\n")
    f.write("# UTCHMMA/UTCIMMA = tcgen05.mma kind::f16 / kind::i8, UTMALDG/UTMASTG = TMA load/store, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld,\n")
    f.write("# DMMA = fp64 tensor-core mma, ACQBULK = griddepcontrol.wait (programmatic dependent launch)\n")
    for fn, c in cnt.items():
        if sum(c.values()) == 0:
            continue
        dem = subprocess.run(["c++filt", "-p", fn], capture_output=True, text=True).stdout.strip()
        dem = dem.replace("(anonymous namespace)::", "")
        f.write(f"{dem[:70]:70s} " + " ".join(f"{k} {c[k]:3d}" for k in keys if c[k]) + "\n")
print(open(ROOT / "profiles" / f"{tag}_sass_mnemonics.txt").read())
