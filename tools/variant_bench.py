"""Developer tool: time kernel build variants (mjhmc_b200/_variants/lib_*.so) on a few workloads.
Usage (GPU box): python tools/variant_bench.py v0 v1 ... [--workloads a,b] [--parity]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = [a for a in sys.argv[1:] if not a.startswith("--")]
wl = "roughwell2d_mjhmc,roughwell2d_control,testgauss2d_control_L1,roughwell2d_control_L1,funnel10d_cthmc"
parity = "--parity" in sys.argv
for a in sys.argv[1:]:
    if a.startswith("--workloads="):
        wl = a.split("=", 1)[1]
for v in args:
    env = dict(os.environ, MJHMC_B200_LIB=os.path.join(ROOT, "mjhmc_b200", "_variants", "lib_%s.so" % v))
    for w in wl.split(","):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", w, "--steps", "10", "--warmup", "3",
                            "--no-cpu-baseline"], env=env, capture_output=True, text=True)
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            print("%-4s %-26s value %.4g  ms/step %.4f  hbm_frac %.3f  e2e %.3g" % (
                v, w, j["value"], j["ms_per_step"], j["roofline"]["frac"], j["e2e"]["value"]), flush=True)
        except Exception:   # noqa: BLE001
            print(v, w, "FAILED", r.stdout[-300:], r.stderr[-600:], flush=True)
    if parity:
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "-q", "-x",
                            "-k", "golden or philox"], env=env, capture_output=True, text=True, cwd=ROOT)
        print(v, "parity:", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
