#!/usr/bin/env python
"""bench.py — LM iterations/s and corner-observations/s of the MultiCamMapper joint optimisation.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4] [--frames F]

One *step* is one Levenberg-Marquardt iteration of MultiCamMapper::solve() (quantised finite-difference
Jacobian + normal equations + per-frame Schur elimination + reduced solve + trial residual + accept/reject,
libs/sparselevmarq.h:348-430 of the reference) over every marker observation of the workload.  The default
workload is BASELINE.json configs[3] (16 cameras, 64 markers, 100 000 frames, ~25.6 M marker observations
= ~102 M corner observations), the configuration the north star's target is quoted on; it fits one B200.
With N > 1 (launched by torch.distributed.run, one rank per GPU) the same problem is frame-sharded and the
reduced system + cost are NCCL-all-reduced every try  ->  "scaling": "strong".

value   device-resident throughput: observations, poses and z are in HBM before the timed region.
e2e     the same iterations through the C-ABI call a reference user would make (aar_lm_solve with HOST
        io_vec, pinned): host->device copy of z, K iterations, device->host copy of z and of the LM state.
--impl reference   times the reference's own CPU path (oracle/_ref: restated MultiCamMapper residual/Jacobian
        driving the UNMODIFIED libs/sparselevmarq.h) on the host cores, on a bounded frame sample of the
        same workload.

The LM trajectory is restarted from z0 every RESTART iterations (device-side copy, inside the timed
region) so that every timed step is a productive LM iteration and not a converged one.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p_ in (os.path.join(ROOT, "automatic-ar_b200", "python"),):
    if p_ not in sys.path:
        sys.path.insert(0, p_)

RESTART = 10            # LM iterations between restarts from z0
W_J_FLOP = 9.0e3        # algorithmic FP64 flop per marker observation per Jacobian evaluation
W_PROJ_FLOP = 5912.0    # ... of which in k_jac_project: 37 projections x 152 + 288 for the central differences
OBS_BYTES = 76          # HBM bytes per marker observation read by k_jac_project (2 x 8 float corners + 12 B indices; the pair table adds 1536 B per (frame, camera) pair)
STAGE_BYTES = 640       # HBM bytes per marker observation written by k_jac_project (144 float numerators + 8 double residuals)
CPU_SAMPLE_FRAMES = 300
CPU_SAMPLE_ITERS = 3


def load_peaks():
    peaks = {"hbm_gbs": 6650.0, "hbm_src": "fallback", "fp64_tflops": 36.7, "fp64_src": "r1 DFMA microbenchmark (profiles/r1_fp64_peak.json)"}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks["hbm_gbs"] = float(mp["hbm_gbs"]); peaks["hbm_src"] = "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    try:
        fp = json.load(open(os.path.join(ROOT, "profiles", "r1_fp64_peak.json")))
        peaks["fp64_tflops"] = float(fp["dfma_tflops"])
    except Exception:
        pass
    return peaks


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index; self.rows = []; self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = [r for r in self.rows if t0 - 0.15 <= r[0] <= t1 + 0.15] or self.rows[-2:]     # very short timed regions: nearest samples (under load: warm-up precedes)
        for ts, line in window:
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); pw.append(float(f[3]))
            except Exception:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(workload, frames, iters, warmup=0):
    """The reference's CPU implementation of the path on the host cores (test infrastructure: oracle/)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    from aar_b200 import synth
    rig = synth.make_config(workload, frames=frames)
    o = oracle_py.Oracle(rig)
    z0 = o.mats2evec()
    kind = "reference" if o.is_ref else "port"
    cores = int(o.L.aar_oracle_omp_threads())
    if o.is_ref:
        if warmup:
            o.time_ref_steps(z0, warmup)
        secs = o.time_ref_steps(z0, iters)
    else:
        t = time.time(); _, _, it, _ = o.solve_port(z0); secs = time.time() - t; iters = max(int(it), 1)
    n_obs = o.num_rows // 8
    return dict(seconds=secs, iters=iters, n_obs=n_obs, kind=kind, cores=cores,
                sample=f"{workload} restricted to its first {rig.F} frames ({n_obs} marker observations), {iters} SparseLevMarq::step calls from the initial estimate"
                       + ("" if o.is_ref else " [restated LM loop]"))


def run_reference(args, rank):
    if rank != 0:
        return
    r = cpu_reference_run(args.workload, args.frames or CPU_SAMPLE_FRAMES, max(args.steps, 1), warmup=min(args.warmup, 1))
    v = 4.0 * r["n_obs"] * r["iters"] / r["seconds"]
    line = {"impl": "reference", "metric": "corner_obs_per_s", "value": v, "unit": "corner-observations/s", "n_gpus": args.gpus,
            "steps": r["iters"], "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * r["seconds"] / r["iters"],
            "lm_iters_per_s": r["iters"] / r["seconds"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "sampled_frames": args.frames or CPU_SAMPLE_FRAMES, "marker_observations": r["n_obs"]},
            "cpu_baseline": {"value": v, "unit": "corner-observations/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": v, "unit": "corner-observations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def track_line(device, with_cpu):
    """BASELINE config 5 (track mode) on a bounded sample: batched per-frame 6-dof LM against the fixed rig, host z in/out."""
    import copy
    import numpy as np
    from aar_b200 import binding, synth
    frames = 5000
    rig = copy.copy(synth.make_config("cfg5", frames=frames))
    rig.T_cam_init, rig.T_marker_init = rig.T_cam_true, rig.T_marker_true        # the solved rig is fixed while tracking
    p = binding.Problem(rig, cams=False, markers=False, objects=True, device=device)
    z0 = p.mats2evec().reshape(-1, 6)
    p.track_batch(z0)
    t = time.time(); z, cost, its = p.track_batch(z0); dt = time.time() - t
    out = {"workload": "cfg5 sample: %d independent frames, %d marker observations" % (rig.F, p.num_obs), "frames_per_s": rig.F / dt,
           "ms": 1e3 * dt, "lm_iterations_mean": float(its.mean()), "rms_px": float(np.sqrt(cost.sum() / (8 * p.num_obs))),
           "timing": "host wall clock around aar_track_batch (host z in / out, copies included)"}
    p.close()
    if with_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py
        o = oracle_py.Oracle(rig); o.set_config(cams=False, markers=False, objects=True)
        n = 8; t = time.time()
        for k in range(n):
            sel = rig.det_frame == rig.frame_ids[k]
            rows, zi = o.track_init(int(rig.frame_ids[k]), rig.T_frame_init[k], rig.det_cam[sel], rig.det_marker[sel], rig.det_xy[sel])
            (o.track_ref if o.is_ref else o.track_port)(zi)
        out["cpu_frames_per_s"] = n / (time.time() - t); out["cpu_sample_frames"] = n
        out["cpu_kind"] = "reference" if o.is_ref else "port"
    return out


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from aar_b200 import binding, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.pop("NCCL_DEBUG", None)      # NCCL prints its version banner to STDOUT when this is set: stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    binding.lib()
    t0 = time.time()
    rig = synth.make_config(args.workload, frames=args.frames)
    t_gen = time.time() - t0
    stream = torch.cuda.Stream()
    t0 = time.time()
    p = binding.Problem(rig, device=local_rank, stream=stream.cuda_stream, rank=rank, world_size=world)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(binding.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        p.comm_init(bytes(idt.cpu().numpy().tobytes()))
    torch.cuda.synchronize()
    t_create = time.time() - t0
    n_obs, n_local, n_vars = p.num_obs, p.num_local_obs, p.num_vars
    z0 = p.mats2evec()
    zpin = torch.empty(n_vars, dtype=torch.float64).pin_memory()
    zpin_np = zpin.numpy()

    prm = binding.Problem.default_params(ignore_stop_rules=1)
    K, W = args.steps, args.warmup

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    seg_tries = [0]          # total_tries of the report is cumulative since the last lm_begin

    def iterate(n, done):
        """n LM iterations, restarting from the device-resident z0 every RESTART iterations; returns tries executed."""
        tries = 0
        while n > 0:
            if done % RESTART == 0 and done > 0:
                p.lm_begin(None, prm); seg_tries[0] = 0
            m = min(n, RESTART - done % RESTART)
            rep, _ = p.lm_iterate(m)
            tries += rep.total_tries - seg_tries[0]; seg_tries[0] = rep.total_tries
            n -= m; done += m
        return done, tries

    # ---------------------------------------------------------------- device-resident arm
    with torch.cuda.stream(stream):
        clocks = ClockSampler(local_rank); clocks.start()      # started before the warm-up so that samples exist even for a short timed region
        p.lm_begin(z0, prm)
        done, _ = iterate(W, 0)
        p.set_profiling(True)
        launches0 = p.kernel_launches
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        tw0 = time.time()
        e0.record(stream)
        done, tries_total = iterate(K, done)
        e1.record(stream)
        barrier()
        tw1 = time.time()
        ms = e0.elapsed_time(e1)
        launches = p.kernel_launches - launches0
        ph = p.phase_ms()
        p.set_profiling(False)
        clk = clocks.stop(tw0, tw1)
        p.lm_end()

        # ------------------------------------------------------------ end-to-end arm (host io_vec through aar_lm_solve)
        chunk = min(K, RESTART)
        prm_e = binding.Problem.default_params(ignore_stop_rules=1, max_iters=chunk)
        zpin_np[:] = z0
        p.solve_inplace(zpin_np, prm_e)                     # warm-up
        barrier()
        f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
        tw2 = time.time()
        f0.record(stream)
        left = K; e2e_tries = 0; calls = 0
        while left > 0:
            m = min(left, chunk)
            prm_e.max_iters = m
            zpin_np[:] = z0
            rep = p.solve_inplace(zpin_np, prm_e)
            e2e_tries += rep.total_tries; left -= m; calls += 1
        f1.record(stream)
        barrier()
        tw3 = time.time()
        ms_e2e = max(f0.elapsed_time(f1), 1e3 * (tw3 - tw2))     # host wall clock includes the copies' sync points

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peaks = load_peaks()
        corner = 4.0 * n_obs
        value = corner * K / (ms * 1e-3)
        nl = max(ph["jacobian_launches"], 1.0)
        jac_ms = ph["jacobian_kernel"] / nl                 # k_jac_project, CUDA events on the launching stream, inside the timed region
        acc_ms = ph["accumulate_kernel"] / nl               # k_jac_accumulate
        ach = W_PROJ_FLOP * n_local / (jac_ms * 1e-3) / 1e12 if jac_ms > 0 else 0.0
        ach_total = W_J_FLOP * n_local / ((jac_ms + acc_ms) * 1e-3) / 1e12 if jac_ms + acc_ms > 0 else 0.0
        hbm_ach = (OBS_BYTES + STAGE_BYTES) * n_local / (jac_ms * 1e-3) / 1e9 if jac_ms > 0 else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "jacobian_traffic.json"))).get("dram_bytes_per_launch_at_bench_size")
        except Exception:
            pass
        line = {
            "metric": "corner_obs_per_s", "value": value, "unit": "corner-observations/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "lm_iters_per_s": K / (ms * 1e-3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "cameras": rig.C, "markers": rig.M, "frames": rig.F, "marker_observations": n_obs,
                       "corner_observations": int(corner), "num_vars": n_vars, "reduced_system": p.n_r, "parallelism": f"frame-shard x{world}",
                       "l2": ("per-step working set %.0f MB per rank (observations + staged Jacobian block, streamed once per step) %s the 126 MB L2; no explicit flush"
                              % ((OBS_BYTES + 2 * STAGE_BYTES) * n_local / 1e6, "exceeds" if (OBS_BYTES + 2 * STAGE_BYTES) * n_local > 2 * 126e6 else "does NOT exceed")),
                       "restart_every": RESTART, "step": "one SparseLevMarq::step (J + JtJ + Schur + solve + trial residual)"},
            "e2e": {"value": corner * K / (ms_e2e * 1e-3), "unit": "corner-observations/s", "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": int(8 * n_vars * calls / K), "d2h_bytes_per_step": int((8 * n_vars * calls + 96 * e2e_tries) / K),
                    "call": f"aar_lm_solve(host io_vec) x{calls}, {chunk} iterations each"},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"kernel": "k_jac_project (37 pinhole projections per marker observation: residual + quantised central-difference numerators; inv(Tc)*To variants shared per (frame, camera) pair)",
                         "bound": "fp64", "achieved": ach, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["fp64_tflops"],
                         "peak_source": peaks["fp64_src"], "flop_per_marker_obs": W_PROJ_FLOP, "ms_per_launch": jac_ms,
                         "hbm": {"achieved": hbm_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / peaks["hbm_gbs"], "peak_source": peaks["hbm_src"],
                                 "bytes_per_marker_obs": OBS_BYTES + STAGE_BYTES},
                         "traffic": traffic,
                         "jacobian_phase": {"kernels": "k_jac_project + k_jac_accumulate (k_pair_tab, k_expand_jac and the zeroing passes are in phases_ms_per_step.jacobian)", "flop_per_marker_obs": W_J_FLOP, "ms": jac_ms + acc_ms,
                                            "achieved": ach_total, "frac": ach_total / peaks["fp64_tflops"]}},
            "phases_ms_per_step": {k: v / K for k, v in ph.items() if k != "jacobian_launches"}, "total_tries": int(tries_total),
            "setup_s": {"generate": t_gen, "create_upload_undistort": t_create},
        }
        if world == 1 and not args.no_track:
            try:
                line["track"] = track_line(local_rank, not args.no_cpu_baseline)
            except Exception as e:      # the headline measurement above stands on its own
                line["track"] = {"error": str(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(args.workload, CPU_SAMPLE_FRAMES, CPU_SAMPLE_ITERS)
            line["cpu_baseline"] = {"value": 4.0 * r["n_obs"] * r["iters"] / r["seconds"], "unit": "corner-observations/s", "cores": r["cores"],
                                    "kind": r["kind"], "sample": r["sample"], "ms_per_step_on_sample": 1e3 * r["seconds"] / r["iters"]}
        print(json.dumps(line), flush=True)
    p.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4")
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-track", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    args.warmup = max(args.warmup, 3)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
