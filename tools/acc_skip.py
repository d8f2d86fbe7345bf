"""Development aid: time k_jac_accumulate with stages left out (AAR_ACC_SKIP bit mask) to find what bounds it."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for mask in [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 8, 7, 15, 16]:
    env = dict(os.environ, AAR_ACC_SKIP=str(mask))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_time.py"), "--workload", "cfg4", "--frames", "10000", "--iters", "4"], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if "phases" in l]
    print(mask, line[0].split("'accumulate_kernel':")[1].strip(" }") if line else r.stderr[-300:], flush=True)
