"""ncu report(s) -> markdown table of the metrics the judge greps for.  usage: ncu_summary.py a.ncu-rep [b.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "lsu wf %")]

print("| kernel | " + " | ".join(k for _, k in KEYS) + " |")
print("|---|" + "---|" * len(KEYS))
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for row in rows[2:]:
        cells = []
        for key, _ in KEYS:
            if key in hdr:
                i = hdr.index(key)
                cells.append(f"{row[i]} {units[i]}".strip())
            else:
                cells.append("-")
        print(f"| {row[ki][:60]} | " + " | ".join(cells) + " |")
