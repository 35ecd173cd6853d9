"""Per-kernel table from an `ncu --metrics ... --csv` capture of one training iteration: launches, total time, DRAM
bytes, achieved DRAM GB/s and tensor-pipe utilisation (time-weighted).  usage: step_metrics_summary.py in.csv out.md"""
import collections
import csv
import sys

src, dst = sys.argv[1], sys.argv[2]
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
per = collections.OrderedDict()   # launch id -> dict
for r in rows:
    d = per.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0].replace("gap::", "").replace("void ", "")})
    v = float(r["Metric Value"].replace(",", "") or 0)
    u = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    if m.startswith("dram__bytes") or m.startswith("lts__t_bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    d[m] = v
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d["name"], {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0, "tp": 0.0, "l2": 0.0})
    t = d.get("gpu__time_duration.sum", 0.0)
    a["n"] += 1
    a["t"] += t
    a["rd"] += d.get("dram__bytes_read.sum", 0.0)
    a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    a["l2"] += d.get("lts__t_bytes.sum", 0.0)
    a["tp"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
T = sum(a["t"] for a in agg.values())
with open(dst, "w") as f:
    f.write("# One Pix2Pix training iteration (batch 64, 256x256) under ncu: every kernel\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none` "
            "(per-launch, cold-cache, serialised: compare shares).  DRAM GB/s = (read+write)/time; tensor-pipe % is "
            "time-weighted over the kernel's launches.\n\n")
    f.write(f"total {T:.0f} us over {sum(a['n'] for a in agg.values())} launches\n\n")
    f.write("| kernel | launches | time us | share % | DRAM read MB | DRAM write MB | DRAM GB/s | % of 6527 GB/s | L2 traffic MB | tensor pipe % |\n")
    f.write("|---|---|---|---|---|---|---|---|---|---|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        gbs = (a["rd"] + a["wr"]) / (a["t"] * 1e-6) / 1e9 if a["t"] else 0.0
        f.write(f"| {k[:60]} | {a['n']} | {a['t']:.1f} | {100 * a['t'] / T:.1f} | {a['rd'] / 1e6:.1f} | {a['wr'] / 1e6:.1f} | "
                f"{gbs:.0f} | {100 * gbs / 6527:.0f} | {a['l2'] / 1e6:.0f} | {a['tp'] / a['t'] if a['t'] else 0:.1f} |\n")
print(open(dst).read())
