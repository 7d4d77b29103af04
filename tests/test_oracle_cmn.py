"""The CMN oracle (oracle/cmn_oracle.py) against hand-computed vectors of Kaldi's SlidingWindowCmn window rule [ext]
(reference pipe: tensorflow/tf_extract.py:63) — both the centred mode the reference uses and the causal mode with
--min-cmn-window."""
import numpy as np

from oracle import cmn_oracle


def test_windows_center_hand_computed():
    # T = 5, window 4: start = t - 2, end = start + 4, shifted into [0, 5)
    assert [cmn_oracle.window(t, 5, 4, True) for t in range(5)] == [(0, 4), (0, 4), (0, 4), (1, 5), (1, 5)]
    x = np.array([[1.0], [2.0], [3.0], [4.0], [5.0]], np.float32)
    want = np.array([[-1.5], [-0.5], [0.5], [0.5], [1.5]], np.float32)
    np.testing.assert_array_equal(cmn_oracle.apply_cmvn_sliding_naive(x, 4, True), want)
    np.testing.assert_array_equal(cmn_oracle.apply_cmvn_sliding(x, 4, True), want)


def test_windows_causal_min_window_hand_computed():
    # T = 5, window 3, min_window 2: [t-3, t+1) shifted right, then end = max(t+1, 2) while the window would look past t
    assert [cmn_oracle.window(t, 5, 3, False, 2) for t in range(5)] == [(0, 2), (0, 2), (0, 3), (0, 4), (1, 5)]
    x = np.array([[1.0], [2.0], [3.0], [4.0], [5.0]], np.float32)
    want = np.array([[-0.5], [0.5], [1.0], [1.5], [1.5]], np.float32)
    np.testing.assert_array_equal(cmn_oracle.apply_cmvn_sliding_naive(x, 3, False, 2), want)
    # utterance shorter than min_window: every frame is normalised by the whole utterance
    assert [cmn_oracle.window(t, 3, 300, False, 100) for t in range(3)] == [(0, 3)] * 3
    # frame 0 is NOT normalised by itself alone (that would zero it): it looks ahead to min_window frames
    assert cmn_oracle.window(0, 1000, 300, False, 100) == (0, 100)
    assert cmn_oracle.window(99, 1000, 300, False, 100) == (0, 100)
    assert cmn_oracle.window(100, 1000, 300, False, 100) == (0, 101)
    assert cmn_oracle.window(500, 1000, 300, False, 100) == (200, 501)


def test_reference_mode_windows():
    # --center=true --cmn-window=300 (tf_extract.py:63)
    assert cmn_oracle.window(0, 1000) == (0, 300)
    assert cmn_oracle.window(149, 1000) == (0, 300)
    assert cmn_oracle.window(151, 1000) == (1, 301)
    assert cmn_oracle.window(999, 1000) == (700, 1000)
    assert cmn_oracle.window(10, 200) == (0, 200)          # utterance shorter than the window: the whole utterance


def test_running_sums_match_naive_loop():
    rng = np.random.default_rng(2)
    for T in (25, 99, 100, 101, 299, 300, 301, 750):
        x = rng.standard_normal((T, 5)).astype(np.float32) + 3
        for center in (True, False):
            np.testing.assert_allclose(cmn_oracle.apply_cmvn_sliding(x, 300, center), cmn_oracle.apply_cmvn_sliding_naive(x, 300, center),
                                       atol=1e-5)
