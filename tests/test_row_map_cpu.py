"""Host logic of the drop-in without a GPU: the observation -> residual-row map (fill_iteration_arrays,
multicam_mapper.cpp:345-377) of the C ABI against the CPU oracle, bit-exact, on randomised inputs
(hypothesis): shuffled detection order inside (frame, camera) groups, arbitrary file order, unknown camera /
marker / frame ids, duplicated detections, sparse id sets."""
import copy

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from aar_b200 import synth


@pytest.fixture(scope="module")
def binding():
    from aar_b200 import binding as b
    b.build()
    return b


def _check(binding, oracle_mod, rig, rig_oracle=None):
    m = binding.Problem.row_map(rig)
    obs = oracle_mod.Oracle(rig_oracle or rig).observations()
    assert len(m["frame_idx"]) == len(obs["frame_id"])
    assert np.array_equal(rig.frame_ids[m["frame_idx"]], obs["frame_id"])
    assert np.array_equal(rig.cam_ids[m["cam_idx"]], obs["cam_id"])
    assert np.array_equal(rig.marker_ids[m["marker_idx"]], obs["marker_id"])
    assert np.array_equal(m["has_jac"], obs["has_jac"])


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), C=st.integers(1, 4), M=st.integers(1, 7), F=st.integers(1, 12), n_extra=st.integers(0, 6), shuffle_all=st.booleans())
def test_row_map_bit_exact_randomised(binding, oracle_mod, seed, C, M, F, n_extra, shuffle_all):
    rig = synth.make_rig(C=C, M=M, F=F, obs_per_frame=3.0, seed=seed)
    if rig.N == 0:
        return
    rng = np.random.default_rng(seed)
    rig = copy.copy(rig)
    if n_extra:   # unknown camera / marker ids and duplicates of existing detections
        pick = rng.integers(0, rig.N, n_extra)
        ef, ec, em = rig.det_frame[pick].copy(), rig.det_cam[pick].copy(), rig.det_marker[pick].copy()
        kind = rng.integers(0, 3, n_extra)
        ec[kind == 0] = 77; em[kind == 1] = 4242          # kind 2: exact duplicate (frame, cam, marker)
        rig.det_frame = np.concatenate([rig.det_frame, ef]); rig.det_cam = np.concatenate([rig.det_cam, ec])
        rig.det_marker = np.concatenate([rig.det_marker, em]); rig.det_xy = np.concatenate([rig.det_xy, rig.det_xy[pick] + 0.5])
    if shuffle_all:  # the ABI accepts detections in any order; rows are (frame, cam, order of appearance)
        perm = rng.permutation(len(rig.det_frame))
        rig.det_frame, rig.det_cam, rig.det_marker, rig.det_xy = rig.det_frame[perm], rig.det_cam[perm], rig.det_marker[perm], rig.det_xy[perm]
    _check(binding, oracle_mod, rig)


def test_row_map_drops_frames_without_pose(binding, oracle_mod):
    rig = synth.make_rig(C=2, M=4, F=6, obs_per_frame=3.0, seed=3)
    rig2 = copy.copy(rig)
    rig2.det_frame = np.concatenate([rig.det_frame, [10 ** 6]]).astype(np.int32); rig2.det_cam = np.concatenate([rig.det_cam, rig.det_cam[:1]])
    rig2.det_marker = np.concatenate([rig.det_marker, rig.det_marker[:1]]); rig2.det_xy = np.concatenate([rig.det_xy, rig.det_xy[:1]])
    _check(binding, oracle_mod, rig2, rig_oracle=rig)     # the reference would throw on a detection of a frame without a pose; the ABI drops it


def test_empty_problem(binding):
    rig = synth.make_rig(C=2, M=3, F=4, obs_per_frame=3.0, seed=1)
    rig = copy.copy(rig)
    rig.det_frame = rig.det_frame[:0]; rig.det_cam = rig.det_cam[:0]; rig.det_marker = rig.det_marker[:0]; rig.det_xy = rig.det_xy[:0]
    m = binding.Problem.row_map(rig)
    assert len(m["frame_idx"]) == 0
    assert binding.Problem.shard_plan(rig, 0, 2)[2:] == (0, 0, 0)
