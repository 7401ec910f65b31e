"""MixupBase -- drop-in for pygda/nn/mixup_base.py:10-200 (StruRW's 'mixup' mode).

The convolutions are ``MixUpGCNConv`` (libgda GEMM + aggregation), the interpolations ``a * lam + b * (1 - lam)``
``ops.lerp2`` (gda_scale_f32 + gda_axpy_f32); only the node permutation ``x[id_new_value_old]`` is a torch gather (data
movement).  This mode is API surface -- no benchmark script selects it (benchmark/node/strurw.py:39 defaults to 'erm')."""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .layers import Linear
from .mixup_gcnconv import MixUpGCNConv


class MixupBase(nn.Module):
    def __init__(self, in_dim, hid_dim, num_classes, num_layers=1, dropout=0.1, act=F.relu, rw_lmda=0.8, **kwargs):
        super().__init__()
        self.in_dim, self.hid_dim, self.num_classes, self.num_layers = in_dim, hid_dim, num_classes, num_layers
        self.dropout, self.act, self.rw_lmda = dropout, act, rw_lmda
        self.convs = nn.ModuleList()
        self.convs.append(MixUpGCNConv(self.in_dim, self.hid_dim))
        for _ in range(self.num_layers - 1):
            self.convs.append(MixUpGCNConv(self.hid_dim, self.hid_dim))
        self.cls = Linear(self.hid_dim, self.num_classes)

    def forward(self, x, edge_index, edge_index_b, lam, id_new_value_old, edge_weight):
        x = self.feat_bottleneck(x, edge_index, edge_index_b, lam, id_new_value_old, edge_weight)
        return self.feat_classifier(x)

    def feat_classifier(self, x):
        return self.cls(x)

    def feat_bottleneck(self, x, edge_index, edge_index_b, lam, id_new_value_old, edge_weight):     # :99-178
        def conv(i, a, cen, ei):
            return self.convs[i](a, cen, ei, edge_weight, self.rw_lmda)

        def act(t):
            return ops.act_dropout(t, self.act, 0.0, False)

        def act_drop(t):
            return ops.act_dropout(t, self.act, self.dropout, self.training)

        def drop(t):
            return ops.act_dropout(t, None, self.dropout, self.training)

        if isinstance(id_new_value_old, np.ndarray):
            id_new_value_old = torch.from_numpy(id_new_value_old)
        perm = torch.as_tensor(id_new_value_old, dtype=torch.long).to(x.device)
        lam = float(lam)
        x1 = act_drop(conv(0, x, x, edge_index))                              # :130-132
        x2 = act_drop(conv(1, x1, x1, edge_index))                            # :134-136
        x0_b, x1_b = x[perm], x1[perm]                                        # :138-139
        x_mix = ops.lerp2(x, x0_b, lam)                                       # :141
        new_x1 = act(conv(0, x, x_mix, edge_index))                           # :143-146
        new_x1_b = act(conv(0, x0_b, x_mix, edge_index_b))
        x1_mix = drop(ops.lerp2(new_x1, new_x1_b, lam))                       # :148-149
        new_x2 = act(conv(1, x1, x1_mix, edge_index))                         # :151-154
        new_x2_b = act(conv(1, x1_b, x1_mix, edge_index_b))
        x_mix = drop(ops.lerp2(new_x2, new_x2_b, lam))                        # :156-157
        x = x2
        for i in range(2, len(self.convs)):                                   # :162-176
            x_t = act_drop(conv(i, x, x, edge_index))
            x_b = x[perm]
            new_x = act(conv(i, x, x_mix, edge_index))
            new_x_b = act(conv(i, x_b, x_mix, edge_index_b))
            x_mix = drop(ops.lerp2(new_x, new_x_b, lam))
            x = x_t
        return x_mix
