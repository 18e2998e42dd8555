"""profiles/ncu_traffic.json from an `ncu --set full` report: DRAM bytes (read + write) per launch
of every kernel, keyed "<scene>:<stage>" the way bench.py looks them up.
Usage: python tools/ncu_traffic.py gpurun_out/r01d.ncu-rep fluid_million"""
import csv, io, json, subprocess, sys
from pathlib import Path
rep, scene = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in rows[2:]:
    name = r[ki].split("(")[0].split("::")[-1].split("<")[0].replace("k_", "")
    b = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
    acc.setdefault(name, []).append(b)
out_path = Path(__file__).resolve().parent.parent / "profiles" / "ncu_traffic.json"
out = json.loads(out_path.read_text()) if out_path.exists() else {}
for name, vals in acc.items():
    out[f"{scene}:{name}"] = sum(vals) / len(vals)
out["_source"] = f"{Path(rep).name}: ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean per launch"
out_path.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
print(json.dumps(out, indent=1, sort_keys=True))
