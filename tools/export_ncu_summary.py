"""Compact per-kernel table from an `ncu --set full` report exported with `ncu -i X.ncu-rep --page raw --csv > raw.csv`.
usage: python tools/export_ncu_summary.py raw.csv out.md "title" — the raw CSV is committed next to the table so every figure can be recomputed."""
import csv
import sys

raw, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor %"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma %"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue %"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "XU(MUFU) %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem")]
cols = [(k, n) for k, n in COLS if k in ix]


def fmt(k, v, u):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    if k.startswith("gpu__time"):
        x = x / 1e3 if u in ("ns", "nsecond") else (x if u in ("us", "usecond") else x * 1e3)
        return f"{x:.1f} us"
    if "bytes" in k or "shared_mem" in k:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return f"{x * scale / 1e6:.2f} MB" if x * scale >= 1e5 else f"{x * scale / 1e3:.1f} KB"
    return f"{x:.1f}"


with open(out, "w") as f:
    f.write(f"# {title}\n\nsource: `{raw}` (raw `ncu --page raw --csv` export of the `--set full --clock-control none` capture, one launch per kernel after two warm-up "
            "launches; `tools/ncu_kernels_r02.py` lists the shapes).  Per-launch times under ncu are serialised and cold in the instruction cache: use them for "
            "the shares and the pipe / memory percentages, not as bench values.\n\n")
    f.write("| # | kernel | grid | " + " | ".join(n for _, n in cols) + " |\n|---|---|---|" + "---|" * len(cols) + "\n")
    for r in data:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        f.write(f"| {r[ix['ID']]} | `{name[:48]}` | {r[ix['Grid Size']]} | " + " | ".join(fmt(k, r[ix[k]], units[ix[k]]) for k, _ in cols) + " |\n")
print(open(out).read())
