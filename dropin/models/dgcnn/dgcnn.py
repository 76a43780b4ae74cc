"""Shim: the reference's `models/dgcnn/dgcnn.py` module path, served by the B200 path (pcnbr_b200.dgcnn over
libpcnbr.so).  Same names / signatures / state_dict keys as /root/reference/models/dgcnn/dgcnn.py:7-280."""
from pcnbr_b200.dgcnn import *                                   # noqa: F401,F403
from pcnbr_b200.dgcnn import DGCNN, DGCNNWithColor, EdgeConv, get_graph_feature, get_loss, get_model, knn   # noqa: F401
