"""Summarise an .ncu-rep into a small markdown table (committed under profiles/).
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_summary.md "title" """
import csv, io, subprocess, sys

rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("gpu__time_duration.sum", "time us"), ("smsp__cycles_active.avg", "cycles"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %")]
idx = [(hdr.index(c), n) for c, n in cols if c in hdr]
with open(out, "w") as f:
    f.write(f"# {title}\n\nSource: `{rep}` (ncu --set full --clock-control none; per-launch numbers are cold-cache, serialised replays)\n\n")
    f.write("| # | " + " | ".join(n for _, n in idx) + " |\n|" + "---|" * (len(idx) + 1) + "\n")
    for i, r in enumerate(data):
        cells = []
        for j, n in idx:
            v = r[j]
            if n == "kernel":
                v = v.split("(")[0].replace("void ", "")[:40]
            else:
                try:
                    v = f"{float(v.replace(',', '')):.1f}" if "." in v else v
                except ValueError:
                    pass
            cells.append(v)
        f.write(f"| {i} | " + " | ".join(cells) + " |\n")
print("wrote", out)
