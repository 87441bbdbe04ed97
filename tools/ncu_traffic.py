"""profiles/r<N>_ncu_traffic.json from an `ncu --set full` capture of one 256-pair pass (tools/prof_step.py 128 3 with
-k regex:conv_tc -s 14 -c 7): DRAM bytes read / written and tensor-pipe activity of the seven conv launches, in layer
order.  bench.py reports `roofline.traffic` from the newest such file.   python tools/ncu_traffic.py REP OUT.json"""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
def val(r, name):
    v, u = float(r[idx[name]].replace(",", "")), units[idx[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
layers = {}
for name, r in zip(["cnv1", "cnv2", "cnv3", "cnv4", "cnv5", "cnv6", "cnv7"], data):
    layers[name] = {"dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
                    "pairs_per_launch": 256,
                    "tensor_pipe_pct": float(r[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
                    "kernel": r[idx["Kernel Name"]][:48]}
json.dump({"source": "ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 14 -c 7, tools/prof_step.py 128 3 (%s)" % rep,
           "layers": layers}, open(out, "w"), indent=1)
print(json.dumps({k: (round(v["dram_bytes_read"] / 1e6), round(v["dram_bytes_write"] / 1e6), round(v["tensor_pipe_pct"], 1)) for k, v in layers.items()}))
