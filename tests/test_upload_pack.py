"""Host side of the upload pipeline (camlasercalibratool_b200/csrc/clc_upload.inl): what the pack threads write for any
point range of a gathered problem (ragged and empty frames, ranges that start and end inside frames), and the planarity
verdict they reach on the way.  No GPU: clc_debug_pack touches no CUDA call."""
import numpy as np
import pytest

from camlasercalibratool_b200 import debug_pack


def _frames(rng, counts, planar=True):
    frames = []
    for c in counts:
        f = rng.normal(size=(c, 3))
        if planar:
            f[:, 2] = 0.0
        frames.append(f)
    return frames


@pytest.mark.parametrize("counts", [[5, 0, 0, 7, 1, 0, 12], [1], [0, 0, 3], [100, 2, 2, 2, 50]])
def test_pack_ranges_match_concatenation(counts):
    rng = np.random.default_rng(3)
    frames = _frames(rng, counts, planar=False)
    flat = np.concatenate(frames, axis=0)
    P = flat.shape[0]
    for a in range(0, P + 1, max(1, P // 7)):
        for b in range(a, P + 1, max(1, P // 5)):
            xyz, _ = debug_pack(frames, a, b, xy=False)
            np.testing.assert_array_equal(xyz, flat[a:b])
            xy, nonplanar = debug_pack(frames, a, b, xy=True)
            np.testing.assert_array_equal(xy, flat[a:b, :2])
            assert nonplanar == int(b > a)  # random z: never exactly zero


def test_planarity_verdict():
    rng = np.random.default_rng(4)
    frames = _frames(rng, [40, 0, 9, 31], planar=True)
    frames[2][3, 2] = -0.0  # negative zero is zero
    _, nonplanar = debug_pack(frames, 0, 80, xy=True)
    assert nonplanar == 0
    frames[3][30, 2] = 1e-300  # the very last point
    assert debug_pack(frames, 0, 80, xy=True)[1] == 1
    assert debug_pack(frames, 0, 79, xy=True)[1] == 0  # a range that stops before it
    frames[3][30, 2] = 0.0
    frames[0][0, 2] = np.nan  # NaN is not zero: the general (three-stream) path must see it
    assert debug_pack(frames, 0, 1, xy=True)[1] == 1
