"""Development aid: time LM iterations of one workload on cuda:0 with per-phase device times.
--libs a.so,b.so times several builds of the library (kernel variants) on the same rig, one child process each."""
import argparse, os, pickle, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "automatic-ar_b200", "python"))

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3"); ap.add_argument("--frames", type=int, default=None); ap.add_argument("--iters", type=int, default=8)
ap.add_argument("--libs", default=None); ap.add_argument("--rig", default=None)
a = ap.parse_args()
if a.libs:
    from aar_b200 import synth
    rig = synth.make_config(a.workload, frames=a.frames)
    path = f"/tmp/quick_time_rig_{os.getpid()}.pkl"
    with open(path, "wb") as fh: pickle.dump(rig, fh, protocol=4)
    for spec in a.libs.split(","):          # lib[:ENV=value[:ENV=value]]
        lib, *sets = spec.split(":")
        env = dict(os.environ)
        if lib != "default": env["AAR_LIB"] = os.path.abspath(lib)
        for kv in sets: env[kv.split("=")[0]] = kv.split("=")[1]
        print(f"== {spec}", flush=True)
        subprocess.run([sys.executable, __file__, "--workload", a.workload, "--iters", str(a.iters), "--rig", path], env=env)
    os.remove(path)
    sys.exit(0)
from aar_b200 import binding, synth
t = time.time()
if a.rig:
    with open(a.rig, "rb") as fh: rig = pickle.load(fh)
else:
    rig = synth.make_config(a.workload, frames=a.frames)
t_gen = time.time() - t
t = time.time(); p = binding.Problem(rig); t_create = time.time() - t
z0 = p.mats2evec()
print(f"{a.workload}: C={rig.C} M={rig.M} F={rig.F} N={p.num_obs} n_r={p.n_r} gen {t_gen:.1f}s create {t_create:.2f}s", flush=True)
prm = binding.Problem.default_params(ignore_stop_rules=1)
p.lm_begin(z0, prm); p.lm_iterate(2)
p.set_profiling(True)
t = time.time()
rep, tr = p.lm_iterate(a.iters, trace_capacity=a.iters)      # returns after the last try's state has been read back
dt = time.time() - t
print("trace cost:", tr[:, 0], "tries", tr[:, 3])
ph = p.phase_ms()
print(f"{dt / a.iters * 1e3:.3f} ms/iter  {4 * p.num_obs * a.iters / dt / 1e9:.3f} G corner-obs/s; phases (ms/iter):", {k: round(v / a.iters, 3) for k, v in ph.items()})
print(f"jacobian kernel: {p.num_obs * 9.0e3 / (ph['jacobian'] / a.iters * 1e-3) / 1e12:.2f} TFLOP/s algorithmic (9.0 kflop/obs)")
