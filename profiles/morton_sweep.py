"""Resident and end-to-end Mq/s of cfg2 knn=1 against the number of Morton bits the batch is sorted on."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for bits in (30, 27, 24, 21, 18, 15):
    env = dict(os.environ, PICO_B200_MORTON_BITS=str(bits))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "50", "--no-cpu-baseline"],
                       capture_output=True, text=True, env=env)
    import json
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print("bits %2d : resident %.0f Mq/s (%.3f ms/step)  e2e %.0f Mq/s (%.3f ms)" % (
            bits, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), flush=True)
    except Exception as ex:
        print(bits, "failed", r.stderr[-300:])
