"""Gradient registration stub for lib/computing_flow_layer/computing_flow_op_grad.py (`ops.RegisterGradient("Computeflow")` in the
reference): importing it must succeed (`lib/networks/network.py:6-26`); the op itself is out of scope, so there is
nothing to register — `computing_flow_op.compute_flow_grad` raises when called."""
try:
    from . import computing_flow_op  # noqa: F401
except ImportError:  # posecnn_b200/ itself on sys.path (reference-style imports)
    import computing_flow_layer.computing_flow_op as computing_flow_op  # noqa: F401
