"""SASS opcode histogram per kernel of libkccot.so -> profiles/<tag>_sass_opcodes.md (cuobjdump + cu++filt)."""
import collections
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
sass = subprocess.run(["cuobjdump", "-sass", "kccotgan_b200/libkccot.so"], capture_output=True, text=True).stdout
cur, hist = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        hist[cur][m.group(1).split(".")[0]] += 1
names = list(hist)
dem = subprocess.run(["cu++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
rows = []
for n, d in zip(names, dem):
    d = d.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("kccot::", "").replace("void ", "")
    cut = d.find(">(")
    short = d[:cut + 1] if cut >= 0 else re.sub(r"\(.*", "", d)
    short = short.replace("(int)", "").replace("(bool)", "")
    rows.append((short, hist[n]))
rows.sort(key=lambda r: (-(r[1].get("UTCHMMA", 0) + r[1].get("UTMALDG", 0)), r[0]))
cols = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "FFMA", "MUFU", "SHFL", "LDS", "STS",
        "LDG", "STG", "BAR"]
with open(f"profiles/{tag}_sass_opcodes.md", "w") as fh:
    fh.write(f"# SASS opcode histogram of libkccot.so (cuobjdump -sass, sm_100a), {tag}\n\n")
    fh.write("UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add, LDTM / STTM = tcgen05.ld / st,\n"
             "UTCBAR = tcgen05.commit, SYNCS = mbarrier operations.  Static instruction counts per kernel (loops count once).\n"
             "Regenerate with `python scripts/sass_histogram.py <tag>` after `python -m kccotgan_b200.build`.\n\n")
    fh.write("| kernel | " + " | ".join(cols) + " | total |\n|---|" + "---|" * (len(cols) + 1) + "\n")
    for s, c in rows:
        fh.write(f"| `{s}` | " + " | ".join(str(c.get(k, 0)) for k in cols) + f" | {sum(c.values())} |\n")
print(open(f"profiles/{tag}_sass_opcodes.md").read()[:2500])
