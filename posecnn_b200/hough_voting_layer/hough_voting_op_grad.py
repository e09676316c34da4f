"""Gradient registration stub for lib/hough_voting_layer/hough_voting_op_grad.py (`ops.RegisterGradient("Houghvoting")`,
:20-37 — zeros for both inputs through `hough_voting_grad`).  `lib/networks/network.py:18` imports it unconditionally.
The CPU Houghvoting op is baseline-only here (`oracle/cpu_hough_ransac.cpp`), so there is nothing
to register; `hough_voting_op.hough_voting_grad` raises when called."""
try:
    from . import hough_voting_op  # noqa: F401
except ImportError:  # posecnn_b200/ itself on sys.path (reference-style imports)
    import hough_voting_layer.hough_voting_op as hough_voting_op  # noqa: F401
