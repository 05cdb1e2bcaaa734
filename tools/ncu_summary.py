"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small markdown table.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md
"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 cyc %"),
        ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "dfma thr-inst"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("launch__registers_per_thread", "regs"), ("launch__waves_per_multiprocessor", "waves"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {k: hdr.index(k) for k, _ in KEYS if k in hdr}
    ik = hdr.index("Kernel Name")
    print(f"# ncu --set full summary of `{path}` (per launch; cold-cache, serialised replays)\n")
    print("| kernel | " + " | ".join(n for k, n in KEYS if k in idx) + " |")
    print("|---|" + "---|" * len(idx))
    for r in rows[2:]:
        name = r[ik].split("(")[0].split("::")[-1][:40]
        cells = []
        for k, _ in KEYS:
            if k in idx:
                v, u = r[idx[k]], units[idx[k]]
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
        print(f"| {name} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
