"""SASS mnemonic counts per kernel of the built library (run in the build container: python profiles/sass_counts.py).
Evidence that the cubins are sm_100a-only and which Blackwell-era instructions the kernels use: UBLKCP (1-D TMA bulk
copy), SYNCS (mbarrier), FFMA2 / FADD2 / FMUL2 (packed FP32), FMNMX3 (3-input min/max), REDUX / CREDUX (warp reductions)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dusty-gan_b200", "lib", "libdustyb200.so")
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
elfs = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout.split()
funcs = re.split(r"\n\s*Function : ", txt)[1:]
KEYS = ["UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FMNMX3", "FMNMX", "REDUX", "CREDUX", "LDS.128", "LDG.E.128", "STG.E.128",
        "SHFL", "VOTE", "MATCH", "ATOMG", "ATOMS", "REDG", "BAR.SYNC", "UTMALDG", "UTCHMMA", "HMMA"]
rows, tot = [], collections.Counter()
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f)
    cnt = {k: sum(1 for o in ops if (o == k or o.startswith(k + ".")) or (k in ("LDS.128", "LDG.E.128", "STG.E.128") and o.startswith(k.split(".")[0]) and k.split(".", 1)[1] in o))
           for k in KEYS}
    rows.append((name, len(ops), cnt))
    tot.update(cnt)
names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
out = ["SASS mnemonic counts per kernel of dusty-gan_b200/lib/libdustyb200.so (cuobjdump -sass; python profiles/sass_counts.py)",
       "cubins: " + " ".join(e for e in elfs if e.endswith(".cubin")), "",
       "%-72s %6s " % ("kernel", "instrs") + " ".join("%9s" % k for k in KEYS)]
for (name, n, cnt), d in sorted(zip(rows, names), key=lambda x: -x[0][1]):
    d = re.sub(r"\(.*", "", d).replace("dusty::", "").replace("void ", "")
    out.append("%-72s %6d " % (d[:72], n) + " ".join("%9d" % cnt[k] for k in KEYS))
out.append("%-72s %6d " % ("TOTAL", sum(r[1] for r in rows)) + " ".join("%9d" % tot[k] for k in KEYS))
open(os.path.join(ROOT, "profiles", "sass_r2.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
