"""Multi-rank host logic on CPU (gloo, world_size 2): the frame shards of aar_shard_plan partition the problem, and the
frame-eliminated reduced systems of the shards sum (the data-path all-reduce of SURVEY 8e) to the reduced system of the
whole problem.  The per-shard reduced systems come from the CPU oracle — the CUDA path is exercised under -m gpu."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

WORLD = 2


def _worker(rank, port, out_dir):
    for p in (os.path.join(ROOT, "automatic-ar_b200", "python"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import torch
    from aar_b200 import binding, synth
    import oracle_py
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    rig = synth.make_rig(C=3, M=6, F=60, obs_per_frame=6.0, seed=11)
    fb, fe, ob, oe, n = binding.Problem.shard_plan(rig, rank, WORLD)
    # every rank learns every shard
    mine = torch.tensor([fb, fe, ob, oe, n], dtype=torch.int64)
    allp = [torch.zeros(5, dtype=torch.int64) for _ in range(WORLD)]
    dist.all_gather(allp, mine)
    allp = torch.stack(allp).numpy()
    assert allp[0, 0] == 0 and allp[-1, 1] == rig.F and np.all(allp[1:, 0] == allp[:-1, 1])          # frames partitioned, in order
    assert allp[0, 2] == 0 and allp[-1, 3] == n and np.all(allp[1:, 2] == allp[:-1, 3])              # observations partitioned
    assert np.all(allp[:, 4] == n)
    cnt = allp[:, 3] - allp[:, 2]
    per_frame = np.bincount(np.searchsorted(rig.frame_ids, rig.det_frame), minlength=rig.F)
    assert cnt.max() - cnt.min() <= 2 * per_frame.max()                                                # balanced by observation count
    # the shard's observation range is exactly the observations of its frames
    o = oracle_py.Oracle(rig)
    obs = o.observations()
    fidx = np.searchsorted(rig.frame_ids, obs["frame_id"])
    assert np.array_equal(np.nonzero((fidx >= fb) & (fidx < fe))[0], np.arange(ob, oe))
    # reduced system of the shard -> all-reduce(sum) == reduced system of the whole problem
    z = o.mats2evec(); mu = 321.0
    S, b, c = o.reduced_system(z, mu, fb, fe)
    S_full, b_full, c_full = o.reduced_system(z, mu)
    buf = torch.from_numpy(np.concatenate([S.ravel(), b, [c]]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    buf = buf.numpy(); nr = len(b)
    # the oracle adds Hrr (the camera/marker blocks of the shard's observations) per shard, like a rank does
    assert np.abs(buf[:nr * nr].reshape(nr, nr) - S_full).max() <= 1e-12 * np.abs(S_full).max()
    assert np.abs(buf[nr * nr:nr * nr + nr] - b_full).max() <= 1e-12 * np.abs(b_full).max()
    assert abs(buf[-1] - c_full) <= 1e-12 * c_full
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_frame_shards_partition_and_reduce(tmp_path, oracle_mod):
    from aar_b200 import binding
    binding.build()
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    assert all(os.path.exists(os.path.join(tmp_path, f"ok{r}")) for r in range(WORLD))


def test_shard_plan_edge_cases():
    from aar_b200 import binding, synth
    binding.build()
    rig = synth.make_rig(C=2, M=3, F=5, obs_per_frame=4.0, seed=2)
    # more ranks than frames: trailing ranks may own nothing, the union is still the whole problem
    plans = [binding.Problem.shard_plan(rig, r, 8) for r in range(8)]
    assert plans[0][0] == 0 and plans[-1][1] == rig.F
    assert all(plans[i][1] == plans[i + 1][0] for i in range(7))
    assert sum(p[3] - p[2] for p in plans) == plans[0][4]
    # single rank owns everything
    assert binding.Problem.shard_plan(rig, 0, 1)[:2] == (0, rig.F)
