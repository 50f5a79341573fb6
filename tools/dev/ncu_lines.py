"""Dynamic warp-instructions per CUDA source line for one kernel: ncu_lines.py rep so_path kernel_mangled_substr divisor [min]"""
import csv, io, re, subprocess, sys, os, tempfile
from collections import defaultdict
rep, so, ksub, div = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
mn = float(sys.argv[5]) if len(sys.argv) > 5 else 4
tmp = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s >/dev/null 2>&1" % (tmp, os.path.abspath(so)), shell=True)
ins = []
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    txt = subprocess.run("nvdisasm -g -c %s/%s" % (tmp, f), shell=True, capture_output=True, text=True).stdout
    on = False; cur = None
    for l in txt.split("\n"):
        if l.startswith(".text."):
            on = ksub in l
            continue
        if not on: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m: ins.append((m.group(2), cur))
    if ins: break
out = subprocess.run("ncu -i %s --page source --csv --print-source sass" % rep, shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; data = rows[2:]
ia = h.index("Instructions Executed")
assert len(ins) == len(data), (len(ins), len(data))
per = defaultdict(int)
for (txt, cur), r in zip(ins, data): per[cur] += int(r[ia])
srcs = {}
tot = 0
for (f, ln), n in sorted(per.items()):
    tot += n
    if n / div >= mn:
        line = ""
        for root in ("v2e_b200/csrc",):
            pth = os.path.join(root, f)
            if os.path.exists(pth):
                srcs.setdefault(pth, open(pth).read().split("\n"))
                line = srcs[pth][ln - 1].strip()[:100]
        print("%7.1f %s:%d  %s" % (n / div, f, ln, line))
print("total %.1f" % (tot / div))
