"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = None, []
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
agg = collections.OrderedDict()
for d in data:
    name = re.sub(r"\(.*", "", d["Kernel Name"])
    name = re.sub(r".*::", "", name)[:50]
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
    a = agg.setdefault(name + " grid=" + d["Grid Size"], [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':75s} launches  avg ms   share")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:75s} n={n:3d} avg={ms / n:8.3f} ms  {100 * ms / tot:5.1f}%")
print(f"total {tot:.3f} ms over {len(data)} launches")
