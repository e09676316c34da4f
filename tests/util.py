"""Shared helpers for the parity tests."""
import numpy as np


def to_np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def assert_hough_rows_equal(got, want, is_train, rtol=1e-4, atol=1e-4, as_set=False):
    """Compare the five Houghvotinggpu outputs.  Integer-valued fields (batch, class, votes,
    domain, weight) exactly; geometry within the stated tolerance."""
    g = [to_np(x) for x in got]
    w = [to_np(x) for x in want]
    assert g[0].shape == w[0].shape, (g[0].shape, w[0].shape)
    if as_set:
        group = 9 if is_train else 1
        def order(a):
            keys = [tuple(np.round(a[0][i * group][[0, 1]]).astype(int)) +
                    tuple(np.round(a[1][i * group][4:7] * 1e3).astype(int)) for i in range(a[0].shape[0] // group)]
            idx = sorted(range(len(keys)), key=lambda i: keys[i])
            rows = np.concatenate([np.arange(i * group, (i + 1) * group) for i in idx]) if idx else np.arange(0)
            return [x[rows] for x in a]
        if g[0].shape[0] % group == 0:
            g, w = order(g), order(w)
    np.testing.assert_array_equal(g[0][:, [0, 1, 6]], w[0][:, [0, 1, 6]])       # batch, class, votes
    np.testing.assert_allclose(g[0][:, 2:6], w[0][:, 2:6], rtol=rtol, atol=atol)  # box
    np.testing.assert_allclose(g[1], w[1], rtol=rtol, atol=atol)                  # pose
    np.testing.assert_allclose(g[2], w[2], rtol=1e-6, atol=1e-6)                  # target (copied quaternions)
    np.testing.assert_array_equal(g[3], w[3])                                     # weight
    np.testing.assert_array_equal(g[4], w[4])                                     # domain
