"""Developer helper: bench.py throughput as a function of the per-GPU batch size (tail / wave quantisation)."""
import json
import subprocess
import sys

for lb in (sys.argv[1:] or ["19", "20", "21", "22", "23"]):
    out = subprocess.run([sys.executable, "bench.py", "--log2-batch", lb, "--steps", "10", "--no-cpu-baseline"],
                         capture_output=True, text=True).stdout
    d = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])
    print("log2_batch %s  value %.4e perm/s  %.3f ms/step  e2e %.4e" % (lb, d["value"], d["ms_per_step"], d["e2e"]["value"]), flush=True)
