"""Summarise `.ncu-rep` files (ncu --set full) into one compact CSV/markdown table per report.
usage: ncu_summary.py out.md rep1.ncu-rep [rep2 ...]   (runs here, no GPU needed)"""
import csv
import io
import subprocess
import sys

WANT = [
    ("Kernel Name", "kernel"),
    ("Grid Size", "grid"),
    ("Block Size", "block"),
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_MB"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem_B"),
    ("sm__cycles_elapsed.avg.per_second", "sm_GHz"),
]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    idx = {}
    for name, short in WANT:
        for i, h in enumerate(hdr):
            if h == name:
                idx[short] = i
                break
    res = []
    for r in rd[2:]:
        d = {}
        for name, short in WANT:
            if short in idx:
                v = r[idx[short]]
                u = units[idx[short]]
                if short == "kernel":
                    v = v.split("(")[0].replace("gap::", "")[:48]
                if short in ("dram_rd_MB", "dram_wr_MB", "l2_to_sm_MB"):
                    f = float(v.replace(",", "")) if v else 0.0
                    f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                    v = f"{f:.1f}"
                if short == "time_us":
                    f = float(v.replace(",", "")) if v else 0.0
                    f *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
                    v = f"{f:.1f}"
                d[short] = v
        res.append(d)
    return res


def main():
    out_path, reps = sys.argv[1], sys.argv[2:]
    if out_path.endswith(".ncu-rep") or not reps:      # a report given as the output path would be overwritten
        sys.exit(__doc__)
    cols = [s for _, s in WANT]
    with open(out_path, "w") as f:
        f.write("# ncu --set full summaries (B200, --clock-control none; per-launch, cold-cache, serialised)\n\n")
        for rep in reps:
            rows = rows_of(rep)
            f.write(f"## {rep.split('/')[-1]}\n\n")
            f.write("| " + " | ".join(cols) + " |\n|" + "---|" * len(cols) + "\n")
            for d in rows:
                f.write("| " + " | ".join(str(d.get(c, "")) for c in cols) + " |\n")
            f.write("\n")
    print("wrote", out_path)


if __name__ == "__main__":
    main()
