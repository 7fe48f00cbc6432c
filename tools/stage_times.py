"""Serialised per-stage times of the C2 batch (the bench's roofline leg) for the library named by HYORB_LIB: quick A/B of kernel variants."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in sys.argv[1:] or [""]:
    env = dict(os.environ)
    if lib:
        env["HYORB_LIB"] = os.path.join(ROOT, "hyslam_b200", "lib", lib)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu"], capture_output=True, text=True, env=env)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(lib or "default", round(d["value"]), round(d["e2e"]["value"]), {k: round(v["ms_per_step"], 3) for k, v in d["stages"].items()}, flush=True)
    except Exception as e:
        print(lib, "FAILED", r.stderr[-500:], e)
