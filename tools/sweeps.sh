#!/bin/bash
# usage (on the GPU box): tools/sweeps.sh TAG -> batch-size and workload sweeps of bench.py into gpurun_out/TAG_sweeps.json
TAG=${1:-sweeps}
python - <<P
import json, subprocess
out = {}
runs = [("b1", ["--batch", "1"]), ("b2", ["--batch", "2"]), ("b4", ["--batch", "4"]), ("b8", ["--batch", "8"]), ("b16", ["--batch", "16"]),
        ("b64", ["--batch", "64"]), ("c2", ["--workload", "c2_micro"]), ("c5", ["--workload", "c5_stress"])]
for name, extra in runs:
    r = subprocess.run(["timeout", "-s", "KILL", "120", "python", "bench.py", "--no-cpu-baseline", "--no-latency"] + extra, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        out[name] = {k: d[k] for k in ("value", "ms_per_step", "config", "e2e", "gpu_launches", "kernels", "stages", "roofline")}
        print(name, round(d["value"], 1), round(d["ms_per_step"], 4), round(d["e2e"]["value"], 1))
    except Exception as e:
        print(name, "failed", e, r.stderr[-300:])
json.dump(out, open("gpurun_out/${TAG}_sweeps.json", "w"), indent=1)
P
