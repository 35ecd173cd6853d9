"""Per kernel of an `ncu --set full --import-source on` report: duration, DRAM traffic, issue rate, instruction mix and the
SASS lines that hold most warp-stall samples (the view that found the instruction-bound / prologue-bound non-GEMM kernels
of round 2).  usage: ncu_hotspots.py out.md report.ncu-rep [kernel-regex]   (runs here, no GPU needed)"""
import collections
import csv
import io
import re
import subprocess
import sys


def norm(name):
    """`void gap::k<(bool)1, (int)4>(args)` and `void k<1, 4>` -> `k<1, 4>`"""
    name = re.sub(r"\(.*?\)(?=[0-9])", "", name.split("(const")[0])      # template value casts
    name = re.sub(r"\((?!bool|int).*", "", name) if "<" not in name else re.sub(r">\(.*", ">", name)
    return name.replace("void ", "").replace("gap::", "").strip()


def page(rep, which, extra=()):
    return subprocess.run(["ncu", "-i", rep, "--page", which, "--csv", *extra], capture_output=True, text=True).stdout


def main():
    if len(sys.argv) < 3 or sys.argv[1].endswith(".ncu-rep"):
        sys.exit(__doc__)
    out_path, rep = sys.argv[1], sys.argv[2]
    rx = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = list(csv.reader(io.StringIO(page(rep, "raw"))))
    hdr = raw[0]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, k, d=""):
        return r[col[k]] if k in col and col[k] < len(r) else d

    seen = {}
    for r in raw[2:]:
        name = get(r, "Kernel Name")
        base = norm(name)
        if rx and not rx.search(base):
            continue
        t = float(get(r, "gpu__time_duration.sum", "0") or 0)
        if base not in seen or t > seen[base][0]:          # the longest launch of each kernel
            seen[base] = (t, r)
    src = list(csv.reader(io.StringIO(page(rep, "source"))))
    blocks, cur = {}, None
    for r in src:
        if r and r[0] == "Kernel Name":
            base = norm(r[1])
            cur = blocks.setdefault(base, {"hdr": None, "rows": []}) if base not in blocks else None
            continue
        if cur is None:
            continue
        if cur["hdr"] is None:
            cur["hdr"] = r
        else:
            cur["rows"].append(r)
    lines = ["# Stall-sample hot spots per kernel (`tools/ncu_hotspots.py`; the longest launch of each kernel; first "
             "profiled instance for the SASS view)", ""]
    for base, (t, r) in sorted(seen.items(), key=lambda kv: -kv[1][0]):
        rd = float(get(r, "dram__bytes_read.sum", "0") or 0)
        wr = float(get(r, "dram__bytes_write.sum", "0") or 0)
        unit = raw[1][col["dram__bytes_read.sum"]] if "dram__bytes_read.sum" in col else ""
        lines.append(f"## `{base}`  {t:.1f} us, DRAM {rd:.1f} + {wr:.1f} {unit}, issue/cycle/SMSP "
                     f"{get(r, 'smsp__issue_active.avg.per_cycle_active')}, registers {get(r, 'launch__registers_per_thread')}, "
                     f"grid {get(r, 'launch__grid_size')}")
        b = blocks.get(base)
        if not b or not b["hdr"]:
            lines += ["(no source page)", ""]
            continue
        h = b["hdr"]
        i_s, i_n, i_m = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        body = [x for x in b["rows"] if len(x) > i_n and x[i_n].isdigit()]
        tot = sum(int(x[i_m]) for x in body) or 1
        ops = collections.Counter()
        for x in body:
            tk = x[i_s].split()
            ops[(tk[1] if tk[0].startswith("@") else tk[0]).split(".")[0]] += int(x[i_n])
        n = sum(ops.values()) or 1
        lines.append("instruction mix: " + ", ".join(f"{k} {100 * v / n:.0f} %" for k, v in ops.most_common(8)))
        lines.append("")
        lines.append("| stall samples | share | SASS |")
        lines.append("|---|---|---|")
        for x in sorted(body, key=lambda x: -int(x[i_m]))[:6]:
            lines.append(f"| {x[i_m]} | {100 * int(x[i_m]) / tot:.1f} % | `{x[i_s].strip()[:80]}` |")
        lines.append("")
    open(out_path, "w").write("\n".join(lines) + "\n")
    print("wrote", out_path)


if __name__ == "__main__":
    main()
