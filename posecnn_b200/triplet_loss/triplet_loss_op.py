"""Triplet — import stub for lib/triplet_loss/triplet_loss_op.py:1-7.

`lib/networks/network.py:6-26` imports this module unconditionally, so it must exist for the reference's network code
to import against this package ("stub modules for the other ops").  The op is OUTSIDE the hot path
this package implements (DESIGN.md §1 "out of scope"); the symbols exist, calling them fails loudly.
There is no CPU or library fallback.
"""
from __future__ import annotations


def _out_of_scope(name):
    def op(*args, **kwargs):
        raise NotImplementedError("%s (%s) is outside the PoseCNN inference hot path posecnn_b200 implements "
                                  "(SURVEY.md §8); the vgg16_convs network never calls it" % (name, "Triplet"))
    op.__name__ = name
    return op


triplet_loss = _out_of_scope("triplet_loss")
triplet_loss_grad = _out_of_scope("triplet_loss_grad")
