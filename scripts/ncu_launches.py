"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into profiles/ (share of each kernel).
usage: python scripts/ncu_launches.py gpurun_out/launches_seg.csv profiles/r01_ncu_launches_seg_v2.txt "<command line captured>" """
import csv, sys, re

src, dst, what = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
with open(src, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
agg = {}
n = 0
for r in rd:
    if len(r) <= iv:
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("<unnamed>::", "").replace("void ", "").strip()
    name = re.sub(r"at::native::|\(anonymous namespace\)::", "", name)[:90]
    ms = float(r[iv].replace(",", "")) * scale.get(r[iu], 1e-6)
    a = agg.setdefault(name, [0.0, 0])
    a[0] += ms
    a[1] += 1
    n += 1
tot = sum(a[0] for a in agg.values())
with open(dst, "w") as f:
    f.write(f"# ncu launch list summary (gpu__time_duration.sum, --clock-control none): {what}\n")
    f.write(f"# {n} launches captured; total {tot:.2f} ms; per-launch times are cold-cache/serialised: compare SHARES\n")
    f.write("share%  total_ms  launches  kernel\n")
    for name, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
        f.write(f"{100 * ms / tot:6.2f}  {ms:8.3f}  {c:8d}  {name}\n")
print(open(dst).read()[:3000])
