"""The fp32 PyTorch restatement of the vgg16_convs graph lives in oracle/ref_network.py (it is also the network half of
`bench.py --impl reference`, the CPU arm of the full path); the tests import it under its old name."""
from oracle.ref_network import *  # noqa: F401,F403
from oracle.ref_network import conv, deconv, deconv_filter, heads, heads_from_scores, trunk  # noqa: F401
