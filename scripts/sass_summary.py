"""SASS mnemonic summary per kernel of libltxb.so (cuobjdump -sass): the instructions that prove the tcgen05 / TMEM / TMA path
(UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor load,
SYNCS = mbarrier) next to the legacy tensor-core forms that must NOT appear (HMMA = mma.sync).  Writes markdown to stdout.
    python scripts/sass_summary.py [path/to/libltxb.so] > profiles/r2/sass_summary.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "mlx-video_b200/csrc/libltxb.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
want = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP", "SYNCS", "UCGABAR", "MUFU", "FFMA2", "HMMA", "ACQBULK", "USETMAXREG", "ELECT"]
kernels, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur["_total"] += 1
        op = m.group(1)
        for w in want:
            if op.startswith(w):
                cur[w] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
print(f"SASS mnemonic counts per kernel, `cuobjdump -sass {lib}` (sm_100a).  UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = "
      "tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce-add, SYNCS = mbarrier ops, UCGABAR = cluster barrier, USETMAXREG = setmaxnreg; HMMA (mma.sync) must be 0.\n")
print("| kernel | instructions | " + " | ".join(want) + " |")
print("|---|---|" + "---|" * len(want))
tot = collections.Counter()
for (name, c), dn in zip(kernels.items(), demangled):
    short = re.sub(r"\(.*", "", dn).replace("void ltxb::", "").replace("ltxb::", "")
    print(f"| `{short}` | {c['_total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in want) + " |")
    tot.update(c)
print(f"| **all {len(kernels)} kernels** | {tot['_total']} | " + " | ".join(str(tot[w]) for w in want) + " |")
