import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "automatic-ar_b200", "python"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


PARITY_LOG = os.path.join(ROOT, "gpurun_out", "parity_achieved.jsonl")


def parity_record(check, **values):
    """Keeps the deviation a parity test ACHIEVED next to the bar it asserts: one JSON line per check
    in gpurun_out/parity_achieved.jsonl (merged back from the GPU box; the round's copy lives in profiles/r2_parity.jsonl)."""
    import json
    import numpy as np
    def plain(v):
        if isinstance(v, (np.floating, np.integer)): return v.item()
        if isinstance(v, np.ndarray): return v.tolist()
        return v
    rec = {"check": check}; rec.update({k: plain(v) for k, v in values.items()})
    os.makedirs(os.path.dirname(PARITY_LOG), exist_ok=True)
    with open(PARITY_LOG, "a") as fh:
        fh.write(json.dumps(rec) + "\n")
    print("PARITY", json.dumps(rec))
    return rec


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "cv2_golden.npz"))


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU oracle (test infrastructure).  Built on demand when missing."""
    import oracle_py
    if not os.path.exists(oracle_py.lib_path(False)):
        oracle_py.build()
    return oracle_py


def rig_from_golden(g):
    import numpy as np
    from aar_b200 import synth
    kw = {}
    for f in synth.Rig.__dataclass_fields__:
        v = g["rig_" + f]
        if f == "image_size":
            v = tuple(int(x) for x in v)
        elif f == "marker_size":
            v = np.float32(v)
        elif f in ("root_cam", "root_marker"):
            v = int(v)
        kw[f] = v
    return synth.Rig(**kw)
