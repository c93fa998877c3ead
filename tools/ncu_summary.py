"""Compact per-kernel table from an ncu report: python tools/ncu_summary.py report.ncu-rep [units_per_launch]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
cols = [("gpu__time_duration.sum", "time"), ("smsp__inst_executed.sum", "warp instr"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "IPC/SM"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/instr"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__waves_per_multiprocessor", "waves")]
print("| kernel | " + " | ".join(c[1] for c in cols) + " |")
print("|---|" + "---|" * len(cols))
for r in rows[2:]:
    name = r[h.index("Kernel Name")].split("(")[0]
    vals = []
    for key, _ in cols:
        if key in h:
            v, unit = r[h.index(key)], u[h.index(key)]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            vals.append(f"{v} {unit}".strip())
        else:
            vals.append("-")
    print(f"| {name} | " + " | ".join(vals) + " |")
