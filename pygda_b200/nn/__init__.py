"""``pygda.nn`` modules on the accelerated path."""
from .reverse_layer import GradReverse
from .prop_gcn_conv import PropGCNConv, GCNConv, gcn_norm
from .a2gnn_base import A2GNNBase
from .cached_gcn_conv import CachedGCNConv
from .ppmi_conv import PPMIConv
from .attention import Attention
from .udagcn_base import UDAGCNBase
from .grade_base import GRADEBase
from .adagcn_base import AdaGCNBase
from .gnn_base import GNNBase
from .gat_conv import GATConv
from .dgsda_base import BernProp, DGSDABase
from .reweight_gnn import GCN_reweight, GS_reweight, ReweightGNN
from .mixup_gcnconv import MixUpGCNConv
from .mixup_base import MixupBase

__all__ = ["GradReverse", "PropGCNConv", "GCNConv", "gcn_norm", "A2GNNBase", "CachedGCNConv", "PPMIConv", "Attention",
           "UDAGCNBase", "GRADEBase", "AdaGCNBase", "GNNBase", "GATConv", "BernProp", "DGSDABase",
           "GCN_reweight", "GS_reweight", "ReweightGNN", "MixUpGCNConv", "MixupBase"]
