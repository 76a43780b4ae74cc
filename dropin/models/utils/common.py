"""Shim: the reference's `models/utils/common.py` module path, served by the B200 path (pcnbr_b200.common over
libpcnbr.so).  Same names / signatures / state_dict keys as /root/reference/models/utils/common.py:6-301."""
from pcnbr_b200.common import *                                  # noqa: F401,F403
from pcnbr_b200.common import (FeaturePropagation, InvResMLP, MiniPointNet, SetAbstraction, UnitPointNet,   # noqa: F401
                               group, interpolate, reduce, sample)
