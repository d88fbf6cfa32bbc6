"""Summarise an .ncu-rep into a small markdown table (committed under profiles/).
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_summary.md "title" [traffic.json precision]
With the two optional arguments the average DRAM bytes (read + write) per tc_conv_kernel launch of the capture are
stored in traffic.json under the precision key (bench.py's roofline.traffic)."""
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
if len(sys.argv) > 5:
    import json, os
    tj_path, prec = sys.argv[4], sys.argv[5]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n = 0.0, 0
    for r in data:
        if "tc_conv_kernel" in r[ki]:
            tot += float(r[ri].replace(",", "")) * scale[units[ri]] + float(r[wi].replace(",", "")) * scale[units[wi]]
            n += 1
    tj = json.load(open(tj_path)) if os.path.exists(tj_path) else {}
    import hashlib
    ksrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sfd2_b200", "csrc", "tc_conv.cu")
    tj[prec] = {"tc_conv_bytes_per_launch": tot / max(n, 1), "launches": n, "source": os.path.basename(rep),
                "kernel_sha": hashlib.sha1(open(ksrc, "rb").read()).hexdigest()[:12],
                "metric": "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full"}
    json.dump(tj, open(tj_path, "w"), indent=1)
    print("traffic", prec, tot / max(n, 1))
