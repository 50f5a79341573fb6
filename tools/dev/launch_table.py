"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch log into a markdown table (shares per kernel)."""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows:
    if r is hdr or len(r) <= iv or r[ik] == "Kernel Name":
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    if r[iu] in ("nsecond", "ns"):
        v /= 1e3
    elif r[iu] in ("msecond", "ms"):
        v *= 1e3
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"^void ", "", name).replace("(anonymous namespace)::", "")
    name = name[:70]
    a = tot.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
all_us = sum(v[1] for v in tot.values())
print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (k, n, us, 100 * us / all_us, us / n))
print("\ntotal: %d launches, %.1f us" % (sum(v[0] for v in tot.values()), all_us))
