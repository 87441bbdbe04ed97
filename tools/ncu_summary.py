"""Summarise an .ncu-rep (raw page) into a compact per-kernel table."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "us"), ("sm__cycles_elapsed.max", "cyc"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("lts__t_bytes.sum", "ltsMB"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("smsp__inst_executed.sum", "inst")]
print("%-44s" % "kernel" + "".join("%10s" % c[1] for c in cols))
for r in data:
    name = r[idx["Kernel Name"]]
    vals = []
    for c, _ in cols:
        if c in idx:
            v = r[idx[c]].replace(",", "")
            u = units[idx[c]]
            try:
                f = float(v)
                if u == "byte": f /= 1e6
                elif u == "Kbyte": f /= 1e3
                elif u == "Gbyte": f *= 1e3
                elif u == "ns": f /= 1e3
                elif u == "ms": f *= 1e3
                vals.append("%10.2f" % f)
            except ValueError:
                vals.append("%10s" % v[:9])
        else:
            vals.append("%10s" % "-")
    print("%-44s" % name[:43] + "".join(vals))
