"""TDSS -- drop-in for pygda/models/tdss.py:93-700 on the B200 path.

A2GNN's network and objective (``A2GNNBase`` + CE + MMD) plus a Laplacian smoothness term on a
smoothing graph ``target_data.edge_index_smooth`` that ``fit`` builds once (tdss.py:503):
K-hop (TwoHopNeighbor, :21-90) or random-walk (:367-373) neighbourhoods.  Same constructor (:168-214),
``forward_model(source_data, target_data, alpha)`` (:241-312), ``smoothness`` (:314-383),
``compute_laplacian_loss`` (:385-449), ``fit`` / ``predict``.
"""
import torch
import torch.nn.functional as F

from .. import smooth
from .a2gnn import A2GNN


class TDSS(A2GNN):
    def __init__(self, in_dim, hid_dim, num_classes, mode='node', smooth_mode='RW', num_layers=2, dropout=0.,
                 act=F.relu, s_pnums=0, t_pnums=30, k=2, rw_len=4, alpha=0.001, beta=1e-4, weight_decay=0.005,
                 adv=False, lr=0.01, epoch=200, device='cuda:0', batch_size=0, num_neigh=-1, verbose=2, **kwargs):
        assert mode == 'node', 'TDSS only supports node-level tasks'                       # :205
        assert adv == False, 'TDSS does not support adversarial training'  # noqa: E712    # :206
        super().__init__(in_dim=in_dim, hid_dim=hid_dim, num_classes=num_classes, mode=mode, num_layers=num_layers,
                         dropout=dropout, act=act, s_pnums=s_pnums, t_pnums=t_pnums, adv=adv, weight=alpha,
                         weight_decay=weight_decay, lr=lr, epoch=epoch, device=device, batch_size=batch_size,
                         num_neigh=num_neigh, verbose=verbose, **kwargs)
        self.smooth_mode = smooth_mode
        self.k = k
        self.rw_len = rw_len
        self.alpha = alpha
        self.beta = beta

    def _extra_loss_terms(self, target_features, target_data):                             # :303-304
        lap = self.compute_laplacian_loss(target_features, target_data.edge_index_smooth)
        return [(lap, float(self.beta))]

    def forward_model(self, source_data, target_data, alpha, mmd_indices=None):
        self.weight = self.alpha                                                           # :300 (MMD trade-off)
        return super().forward_model(source_data, target_data, alpha, mmd_indices=mmd_indices)

    def smoothness(self, edge_index, edge_attr, num_nodes):                                # :314-383
        """Returns ``(edge_index_smooth, edge_attr)`` like the reference; built on the GPU.  Only
        ``edge_attr=None`` is supported (no reference script passes attributes)."""
        if edge_attr is not None:
            raise NotImplementedError("TDSS.smoothness with edge attributes is outside the accelerated path")
        ei = edge_index if edge_index.is_cuda else edge_index.to(self.device)
        if self.smooth_mode == 'RW':
            out = smooth.random_walk_edge_index(ei, num_nodes, self.rw_len)
            return out, torch.ones(out.size(1), dtype=torch.float32, device=out.device)    # dense_to_sparse values
        return smooth.khop_edge_index(ei, num_nodes, self.k), None

    @staticmethod
    def compute_laplacian_loss(features, edge_index):                                      # :385-449
        return smooth.laplacian_loss(features, edge_index)

    def fit(self, source_data, target_data):
        if self.mode == 'node':                                                            # :497-506
            num_target_nodes = target_data.x.shape[0]
            print('before smoothness')
            print(target_data.edge_index.shape)
            target_data.edge_index_smooth, target_data.edge_attr_smooth = self.smoothness(
                target_data.edge_index, getattr(target_data, 'edge_attr', None), num_target_nodes)
            print('after smoothness')
            print(target_data.edge_index_smooth.shape)
        super().fit(source_data, target_data)
