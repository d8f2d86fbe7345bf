"""CPU column of SURVEY 8(d): the reference's own sparselevmarq.h driven by the restated MultiCamMapper residual / Jacobian
(oracle/, test infrastructure) on the host cores — full solves of BASELINE configs 1 and 2, a fixed 3 iterations of config 3.
Prints one JSON object per config.  Usage: cpu_reference_table.py [cfg1 cfg2 cfg3]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "automatic-ar_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py
from aar_b200 import synth

for name in (sys.argv[1:] or ["cfg1", "cfg2", "cfg3"]):
    rig = synth.make_config(name)
    o = oracle_py.Oracle(rig); z0 = o.mats2evec()
    n_obs = o.num_rows // 8
    out = {"config": name, "cameras": rig.C, "markers": rig.M, "frames": rig.F, "marker_observations": n_obs, "num_vars": o.num_vars,
           "cores": int(o.L.aar_oracle_omp_threads()), "kind": "reference" if o.is_ref else "port", "host": os.uname().nodename}
    if name == "cfg3":
        secs = o.time_ref_steps(z0, 3)
        out.update(mode="3 SparseLevMarq::step calls from the initial estimate", seconds=secs, iterations=3)
    else:
        t = time.time(); z, cost, iters, trace = o.solve(z0); secs = time.time() - t
        out.update(mode="full solve (SparseLevMarq::solve, reference stop rules)", seconds=secs, iterations=int(iters), initial_cost=float(trace[0][0]) if len(trace) else None, final_cost=float(cost))
    out["lm_iters_per_s"] = out["iterations"] / out["seconds"]; out["corner_obs_per_s"] = 4.0 * n_obs * out["iterations"] / out["seconds"]
    print(json.dumps(out), flush=True)
