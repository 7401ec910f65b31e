"""``pygda.models`` estimators on the accelerated path."""
from .base import BaseGDA
from .a2gnn import A2GNN
from .udagcn import UDAGCN
from .grade import GRADE
from .adagcn import AdaGCN
from .gnn import GNN
from .tdss import TDSS
from .dgsda import DGSDA
from .strurw import StruRW

__all__ = ["BaseGDA", "A2GNN", "UDAGCN", "GRADE", "AdaGCN", "GNN", "TDSS", "DGSDA", "StruRW"]   # DistA2GNN: import pygda_b200.models.dist_a2gnn
