"""B200-native point-cloud neighbourhood hot path (drop-in for the reference's models/utils/common.py
and models/dgcnn/dgcnn.py).  The directory name is not a Python identifier; load it with
`__graft_entry__.load_package()` (registers it as `pcnbr_b200`)."""
from . import _lib, ops, common, dgcnn, train, synthetic, metrics, block_datasets, dgcnn_utils   # noqa: F401
from .pointnetpp import PointNetpp, PointNetppMSG   # noqa: F401
from .pointnext import PointNeXt                # noqa: F401
from .dgcnn import DGCNN, DGCNNWithColor        # noqa: F401

__version__ = "0.1.0"
