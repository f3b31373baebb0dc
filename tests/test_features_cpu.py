"""is_feature (utils/cv.rs:22-212) as restated in the oracle against OpenCV's FAST 9_16, of which the
reference says it is "a direct port/adaptation" (threshold 30, no non-max suppression).  The reference has
no test of its own for this function; OpenCV is the independent pin.  CPU only."""
import numpy as np
import pytest

from oracle import oracle_py as O

cv2 = pytest.importorskip("cv2")


def _images():
    rng = np.random.default_rng(5)
    yield rng.integers(0, 256, (50, 70), dtype=np.uint8)
    yield rng.integers(0, 256, (41, 33), dtype=np.uint8)
    img = cv2.GaussianBlur(rng.integers(0, 256, (60, 80), dtype=np.uint8), (0, 0), 1.5)
    img[15:35, 20:50] = np.clip(img[15:35, 20:50].astype(int) + 90, 0, 255).astype(np.uint8)
    img[40:52, 5:18] = np.clip(img[40:52, 5:18].astype(int) - 90, 0, 255).astype(np.uint8)
    yield img
    yield np.full((20, 20), 7, dtype=np.uint8)
    chk = (np.indices((48, 48)).sum(0) // 6 % 2 * 200 + 20).astype(np.uint8)
    yield chk


def test_is_feature_equals_opencv_fast_9_16():
    fast = cv2.FastFeatureDetector_create(threshold=30, nonmaxSuppression=False, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    total = 0
    for img in _images():
        want = np.zeros(img.shape, bool)
        for k in fast.detect(img, None):
            want[int(k.pt[1]), int(k.pt[0])] = True
        got = O.is_feature_map(img[..., None])
        assert np.array_equal(got, want)
        total += int(want.sum())
    assert total > 500


def test_is_feature_reads_channel_zero_of_a_colour_plane():
    rng = np.random.default_rng(6)
    img = rng.integers(0, 256, (30, 40, 3), dtype=np.uint8)
    a = O.is_feature_map(img)
    b = O.is_feature_map(np.ascontiguousarray(img[..., :1]))
    assert np.array_equal(a, b) and a.any()
