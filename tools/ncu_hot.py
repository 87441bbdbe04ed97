"""Hottest SASS lines (by warp-stall samples) of one kernel of an .ncu-rep: ncu_hot.py REP KERNEL_REGEX INDEX [N]"""
import csv, subprocess, sys
rep, rx, ix = sys.argv[1], sys.argv[2], sys.argv[3]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", "::regex:%s:%s" % (rx, ix)],
                     capture_output=True, text=True).stdout
rows = [r for r in csv.reader(raw.splitlines())]
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
print(rows[0][1][:100] if rows[0] else "")
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
def num(x):
    try: return int(float(x))
    except ValueError: return 0
tot = sum(num(r[i_s]) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in sorted(data, key=lambda r: -num(r[i_s]))[:n]:
    st = sorted(((h, num(r[hdr.index(h)])) for h in stalls), key=lambda kv: -kv[1])[:2]
    print("%6d %5.1f%% %9d  %-72s %s" % (num(r[i_s]), 100.0 * num(r[i_s]) / max(tot, 1), num(r[i_ex]), r[i_src][:72],
                                          [(k.replace("stall_", ""), v) for k, v in st if v]))
