"""profiles/ncu_traffic.json from the COMMITTED ncu summaries (profiles/ncu_r02*.txt, written by
tools/ncu_summary.py from `ncu --set full` captures of the kernels that ship): DRAM bytes (read +
write) per launch, keyed "<scene>:<regime>:<stage>[_brick]" the way bench.py looks them up
(regime t0 = capture at substep 5, post_impact = substep 110, settled = substep 200).  Reproducible from tracked
files:  python tools/ncu_traffic_from_summaries.py"""
import json
import re
from pathlib import Path

PROF = Path(__file__).resolve().parent.parent / "profiles"
SOURCES = [  # (file, regime)
    ("ncu_r02a_shipped_r01_kernels_t0.txt", "t0"),
    ("ncu_r02a_shipped_r01_kernels_step200.txt", "settled"),
    ("ncu_r02f_lambda_delta_global_vs_persistent_brick_t0.txt", "t0"),
    ("ncu_r02v_shipped_kernels_step110.txt", "post_impact"),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def stage_of(kernel: str):
    brick = "brick" in kernel
    for key, stage in (("LambdaOp", "lambda"), ("DeltaOp", "delta"), ("XsphOp", "xsph"), ("k_lambda", "lambda"),
                       ("k_delta", "delta"), ("k_xsph", "xsph"), ("k_neighbors", "neighbors")):
        if key in kernel:
            return stage + ("_brick" if brick else "")
    return None


out, src = {}, {}
for fname, regime in SOURCES:
    kernel, acc = None, {}
    for line in (PROF / fname).read_text().splitlines():
        if line.startswith("====="):
            kernel = line[5:].strip()
            continue
        m = re.match(r"\s+dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
        if m and kernel:
            acc.setdefault(kernel, 0.0)
            acc[kernel] += float(m.group(2)) * SCALE[m.group(3)]
    per_stage = {}
    for kernel, b in acc.items():
        st = stage_of(kernel)
        if st:
            per_stage.setdefault(st, []).append(b)
    for st, vals in per_stage.items():
        key = f"fluid_million:{regime}:{st}"
        out[key] = sum(vals) / len(vals)   # delta: mean of the <.,0,.> and <.,1,.> (last iteration) launches
        src[key] = fname
out["_source"] = "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, parsed from " + ", ".join(
    sorted(set(src.values())))
(PROF / "ncu_traffic.json").write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
print(json.dumps(out, indent=1, sort_keys=True))
