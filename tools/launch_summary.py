"""ncu launch list (csv with gpu__time_duration.sum per launch) -> per-kernel summary text.
usage: python tools/launch_summary.py <launches.csv> <out.txt> [header line]"""
import csv, collections, os, re, sys
src, out = sys.argv[1], sys.argv[2]
hdr = sys.argv[3] if len(sys.argv) > 3 else ""
rows = [r for r in csv.reader(open(src, errors="ignore")) if len(r) > 5]
head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
col = {n: i for i, n in enumerate(rows[head])}
agg = collections.defaultdict(lambda: [0.0, 0])
tot = 0.0
n = 0
body = rows[head + 1:]
marks = [i for i, r in enumerate(body) if "triu_tril" in r[col["Kernel Name"]]]
if marks:
    body = body[marks[-1] + 1:]
for r in body:
    try:
        v = float(r[col["Metric Value"]].replace(",", ""))
    except (ValueError, IndexError):
        continue
    unit = r[col["Metric Unit"]]
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])[:90]
    agg[name][0] += us; agg[name][1] += 1
    tot += us; n += 1
build = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "monocular-depth-estimation-cil_b200", "build_id.txt")).read().strip()
with open(out, "w") as f:
    f.write(hdr + "\n")
    f.write(f"{n} launches of one train step (eager issue of the captured step body), total {tot / 1e3:.2f} ms (cold-cache, serialised: compare shares); build {build}\n")
    for name, (us, k) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        f.write(f"{us / 1e3:10.3f} ms {100 * us / tot:5.1f}%  n={k:4d}  {name}\n")
print(open(out).read()[:4000])
