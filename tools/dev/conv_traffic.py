"""Sum DRAM bytes of the convolution launches of an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` log."""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
per = collections.OrderedDict()
mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
n_launch = collections.Counter()
for r in rows:
    if r is hdr or r[ik] == "Kernel Name" or "dram__bytes" not in r[im]:
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")[:60]
    v = float(r[iv].replace(",", "")) * mult.get(r[iu], 1.0)
    per[name] = per.get(name, 0.0) + v
    if "read" in r[im]:
        n_launch[name] += 1
tot = sum(per.values())
for k, v in sorted(per.items(), key=lambda kv: -kv[1]):
    print("| `%s` | %d | %.1f MB | %.1f MB/launch |" % (k, n_launch[k], v / 1e6, v / 1e6 / max(n_launch[k], 1)))
print("total %.1f MB over %d launches" % (tot / 1e6, sum(n_launch.values())))
