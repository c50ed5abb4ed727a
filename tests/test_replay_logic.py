"""CPU check of the index logic of the rotation-replay kernels (slot ring of the round-robin schedule, column
orientation, the lean kernel's per-lane sign mask) against the plain product of the recorded rotations:
`tools/emulate_replay.py` mirrors the kernels lane by lane in NumPy."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location('emulate_replay', os.path.join(ROOT, 'tools', 'emulate_replay.py'))
emu = importlib.util.module_from_spec(spec)
spec.loader.exec_module(emu)


@pytest.mark.parametrize("d", [1, 2, 3, 5, 10, 32, 33, 64])
def test_lane_level_replay_equals_product_of_rotations(d):
    rng = np.random.RandomState(d)
    dd = d + (d & 1); npairs = dd // 2; per = dd - 1
    log = []
    for g in range(2 * per):
        row = []
        for k in range(npairs):
            a0, b0 = emu.schedule(dd, g % per, k)
            if max(a0, b0) >= d:
                row.append((1.0, 0.0))
            else:
                t = rng.uniform(-1, 1); c = 1 / np.sqrt(1 + t * t)
                row.append((c, t * c))
        log.append(row)
    V = emu.lean(d, log)
    assert np.max(np.abs(V - emu.reference(d, log))) < 1e-14
    assert np.max(np.abs(V.T.dot(V) - np.eye(d))) < 1e-13
