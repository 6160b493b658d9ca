"""Per-kernel share of one rate-path step from an `ncu --metrics gpu__time_duration.sum` launch list."""
import collections, csv, io, re, sys
txt = open(sys.argv[1]).read()
txt = txt[txt.index('"ID"'):]
rows = list(csv.DictReader(io.StringIO(txt)))
marks = [i for i, r in enumerate(rows) if "rd_epilogue" in r["Kernel Name"]]
a, b = marks[-2] + 1, marks[-1] + 1          # one full step's worth of launches (step boundary shifted, same multiset)
step = rows[a:b]
def short(n):
    n = re.sub(r"\(.*", "", n); n = n.replace("void ", "").replace("mmnc::", "")
    return n[:70]
tot = sum(float(r["Metric Value"]) for r in step)
by = collections.defaultdict(lambda: [0, 0.0])
for r in step:
    k = short(r["Kernel Name"]); by[k][0] += 1; by[k][1] += float(r["Metric Value"])
print(f"one step = {len(step)} launches, sum of kernel durations {tot/1e6:.3f} ms (serialised, cold-cache: compare SHARES)")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t/1e6:.3f} | {100*t/tot:.1f} % |")
mine = sum(t for k, (n, t) in by.items() if not k.startswith("at::") and "cub" not in k and "elementwise" not in k and "Memset" not in k)
print(f"\nmmnc kernels: {100*mine/tot:.1f} % of the step's GPU time")
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as f:
        w = csv.writer(f); w.writerow(["kernel", "grid", "block", "duration_ns"])
        for r in step: w.writerow([short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], r["Metric Value"]])
