"""Who waits for whom inside tc_conv_kernel: execution counts of the mbarrier.try_wait instructions per barrier,
read from the SASS source page of an ncu capture (every retry of a spin loop is one execution).

    python tools/ncu_waits.py gpurun_out/prof.ncu-rep first_launch n_launches out.md name1,name2,...

The barriers sit at fixed offsets behind the epilogue staging tiles (tc_conv.cu): full / empty = operand ring
(MMA issuer waits on `full` = starved by loads; TMA producer waits on `empty` = ring full, consumer-bound),
tfull / tempty = accumulator hand-over (epilogue waits on `tfull` = idle; MMA issuer waits on `tempty` = no free
accumulator, epilogue-bound), fullA / emptyA = the A ring of the split-ring (halo) mode.
"""
import collections, csv, io, re, subprocess, sys

rep, first, count, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
names = sys.argv[5].split(",") if len(sys.argv) > 5 else []
# offsets behind the staging tiles (1, 2 or 3 tiles of 4 KB per epilogue warp, TcConvArgs::stiles) + 1152 bytes of bias
REL = [(0x0, "full"), (0x60, "empty"), (0xc0, "tfull"), (0xd0, "tempty"), (0x120, "fullA"), (0x140, "emptyA")]
BASES = [0x8480, 0x480]      # captures before / after the staging size became a runtime value (the constant part is the 1152 bias bytes)
OFFS = [(hex(BASES[0] + r), n) for r, n in REL]
ALIAS = {hex(b + r): hex(BASES[0] + r) for b in BASES for r, _ in REL}
lines = ["| # | layer | time us | " + " | ".join(n for _, n in OFFS) + " | reading |", "|---|---|---|" + "---|" * (len(OFFS) + 1)]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
ti = rows[0].index("gpu__time_duration.sum")
for k in range(count):
    i = first + k
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(i), "--launch-count", "1"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    r = list(csv.reader(io.StringIO(src)))
    hdr, data = r[1], r[2:]
    ai, ei = hdr.index("Source"), hdr.index("Instructions Executed")
    cnt = collections.Counter()
    for d in data:
        m = re.search(r"TRYWAIT.*\+(0x[0-9a-f]{3,5})\]", d[ai])
        if m and m.group(1) in ALIAS:
            cnt[ALIAS[m.group(1)]] += int(d[ei] or 0)
    c = {n: cnt[o] for o, n in OFFS}
    if c["tempty"] > 20 * 950:
        reading = "MMA issuer waits for a free accumulator: epilogue-bound"
    elif c["full"] > c["empty"]:
        reading = "MMA issuer waits for operands: load-bound"
    else:
        reading = "producer waits for ring slots, epilogue for accumulators: MMA-bound"
    t = float(rows[2 + i][ti].replace(",", ""))
    lines.append(f"| {i} | {names[k] if k < len(names) else ''} | {t:.1f} | " + " | ".join(str(c[n]) for _, n in OFFS) + f" | {reading} |")
open(out, "w").write("# mbarrier.try_wait executions per barrier (all CTAs), one 1600x1200 extraction\n\n"
                     f"Source: `{rep}`, `ncu --page source`; see tools/ncu_waits.py for how to read the columns.\n\n" + "\n".join(lines) + "\n")
print("wrote", out)
