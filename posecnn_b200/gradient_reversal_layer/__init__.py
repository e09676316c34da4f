"""Import stub: out of the hot path; see the op module."""
