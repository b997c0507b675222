"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
tot = collections.defaultdict(float); cnt = collections.Counter()
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
seen = 0
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    seen += 1
    if seen <= skip:
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void |fm::\(anonymous namespace\)::|fm::", "", name)
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    tot[name] += v; cnt[name] += 1
total = sum(tot.values())
print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:70]:70s} {cnt[k]:8d} {v:12.1f} {v/cnt[k]:10.2f} {100*v/total:6.1f}%")
print(f"{'TOTAL':70s} {sum(cnt.values()):8d} {total:12.1f}")
