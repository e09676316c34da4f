"""Gradient registration stub for lib/triplet_loss/triplet_loss_op_grad.py (`ops.RegisterGradient("Triplet")` in the
reference): importing it must succeed (`lib/networks/network.py:6-26`); the op itself is out of scope, so there is
nothing to register — `triplet_loss_op.triplet_loss_grad` raises when called."""
try:
    from . import triplet_loss_op  # noqa: F401
except ImportError:  # posecnn_b200/ itself on sys.path (reference-style imports)
    import triplet_loss.triplet_loss_op as triplet_loss_op  # noqa: F401
