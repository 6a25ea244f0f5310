#!/usr/bin/env python
"""Build profiles/ncu_traffic.json -- the DRAM traffic / tensor-pipe figures bench.py's `roofline` object quotes -- from
`ncu --set full` captures (read here with `ncu -i`, no GPU needed).

    python tools/ncu_traffic.py "conv2 fwd (64->64 @192^2)=gpurun_out/r2h_conv2fwd.ncu-rep" ... [--headline "conv2 fwd (64->64 @192^2)"]

Every entry records dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, its duration under the profiler and
sm__pipe_tensor_cycles_active (%), plus the name of the capture it came from."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M = {"read": "dram__bytes_read.sum", "write": "dram__bytes_write.sum", "time": "gpu__time_duration.sum",
     "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "%": 1.0, "usecond": 1.0,
         "nsecond": 1e-3, "msecond": 1e3}


def read_rep(path: str) -> dict:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(raw)) if len(r) > 8]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, row = rows[hi], rows[hi + 1], rows[-1]    # last captured launch
    out = {"kernel": row[hdr.index("Kernel Name")].split("(")[0]}
    for k, name in M.items():
        i = hdr.index(name)
        out[k] = float(row[i].replace(",", "")) * SCALE.get(units[i], 1.0)
    return out


def main() -> None:
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    headline = sys.argv[sys.argv.index("--headline") + 1] if "--headline" in sys.argv else None
    if headline in args:
        args.remove(headline)
    detail = {"unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), ncu --set full, batch 64"}
    tensor, us, src = {}, {}, {}
    for a in args:
        name, path = a.rsplit("=", 1)
        r = read_rep(path)
        detail[name] = r["read"] + r["write"]
        tensor[name], us[name] = round(r["tensor"], 1), round(r["time"], 1)
        src[name] = os.path.basename(path)
    detail["tensor_pipe_active_pct"] = tensor
    detail["us_under_ncu"] = us
    detail["source"] = src
    headline = headline or next(iter(tensor))
    out = {"traffic": detail[headline], "traffic_kernel": headline, "detail": detail}
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
