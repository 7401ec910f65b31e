"""UDAGCNBase -- drop-in for pygda/nn/udagcn_base.py:9-267.

``GNN`` (:9-89): L shared-weight ``CachedGCNConv`` layers, act + Dropout(0.1) between layers
(not after the last).  The reference keeps the dropout layers in a plain Python list (:47), so
they are never registered and stay in training mode even during ``predict``; reproduced
(``ops.act_dropout(..., training=True)``).  The ctor's ``dropout`` argument is ignored by the
reference (:136-151) and here.

``ppmi=True`` (the reference's default) adds a second, weight-sharing encoder over ``PPMIConv`` layers
(path_len 10, :153) fused with the first by ``Attention`` (:262-264); its PPMI graphs are built on the GPU
(pygda_b200/nn/ppmi_conv.py)."""
import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .attention import Attention
from .cached_gcn_conv import CachedGCNConv
from .ppmi_conv import PPMIConv
from .layers import Linear, ReLUDropout


class GNN(nn.Module):
    def __init__(self, in_dim, hid_dim, gnn_type='gcn', num_layers=3, base_model=None, act=F.relu, **kwargs):
        super().__init__()
        model_cls = PPMIConv if gnn_type == 'ppmi' else CachedGCNConv                   # :51
        if base_model is None:
            weights, biases = [None] * num_layers, [None] * num_layers
        else:
            weights = [c.weight for c in base_model.conv_layers]
            biases = [c.bias for c in base_model.conv_layers]
        self.gnn_type, self.act = gnn_type, act
        self.dropout_p = [0.1 for _ in weights]              # the unregistered nn.Dropout(0.1) list
        self.conv_layers = nn.ModuleList()
        self.conv_layers.append(model_cls(in_dim, hid_dim, weight=weights[0], bias=biases[0], **kwargs))
        for idx in range(1, num_layers):
            self.conv_layers.append(model_cls(hid_dim, hid_dim, weight=weights[idx], bias=biases[idx], **kwargs))

    def forward(self, x, edge_index, cache_name, first_product=None):
        """``first_product``: ``x @ conv_layers[0].weight`` when the caller already has it (UDAGCNBase.encode)."""
        for i, conv_layer in enumerate(self.conv_layers):
            if i == 0 and first_product is not None:
                x = conv_layer.propagate_product(first_product, edge_index, cache_name)
            else:
                x = conv_layer(x, edge_index, cache_name)
            if i < len(self.conv_layers) - 1:
                x = ops.act_dropout(x, self.act, self.dropout_p[i], True)   # always "training"
        return x


class UDAGCNBase(nn.Module):
    def __init__(self, in_dim, hid_dim, num_classes, num_layers=3, dropout=0.1, act=F.relu, ppmi=True,
                 adv_dim=40, feature_dtype=None, **kwargs):
        super().__init__()
        if feature_dtype not in (None, torch.float32, torch.bfloat16):
            raise ValueError("feature_dtype must be torch.float32 or torch.bfloat16")
        self.bf16 = feature_dtype == torch.bfloat16        # bf16 rows through the encoders (BASELINE config 3)
        self.ppmi = ppmi
        self.share_first_product = False                   # opt-in (unmeasured in round 1), see encode()
        self.encoder = GNN(in_dim=in_dim, hid_dim=hid_dim, gnn_type='gcn', act=act, num_layers=num_layers)
        if self.ppmi:
            self.ppmi_encoder = GNN(in_dim=in_dim, hid_dim=hid_dim, base_model=self.encoder,
                                    num_layers=num_layers, gnn_type='ppmi', path_len=10)
        self.cls_model = nn.Sequential(Linear(hid_dim, num_classes))
        # indices 0..3 as in the reference's Sequential(Linear, ReLU, Dropout(0.1), Linear) so that
        # state_dict keys match (domain_model.0.*, domain_model.3.*); ReLU+Dropout run as one kernel
        self.domain_model = nn.Sequential(Linear(hid_dim, adv_dim), ReLUDropout(0.1), nn.Identity(),
                                          Linear(adv_dim, 2))
        self.att_model = Attention(hid_dim)
        self.models = [self.encoder, self.cls_model, self.domain_model]
        if self.ppmi:
            self.models.extend([self.ppmi_encoder, self.att_model])
        self.loss_func = ops.softmax_cross_entropy

    def _x(self, data):
        # the input features are constant: their bf16 copy is made once per tensor (ops.bf16_cache)
        return ops.bf16_cache.get(data.x) if (self.bf16 and data.x.dtype != torch.bfloat16) else data.x

    def _out(self, enc):
        # heads, attention and losses run in fp32: one cast at the encoder boundary (its backward casts back)
        return ops.CastFn.apply(enc, False) if enc.dtype == torch.bfloat16 else enc

    def gcn_encode(self, data, cache_name, mask=None):
        encoded_output = self._out(self.encoder(self._x(data), data.edge_index, cache_name))
        return encoded_output if mask is None else encoded_output[mask]

    def ppmi_encode(self, data, cache_name, mask=None):
        encoded_output = self._out(self.ppmi_encoder(self._x(data), data.edge_index, cache_name))
        return encoded_output if mask is None else encoded_output[mask]

    def encode(self, data, cache_name, mask=None):
        if self.ppmi and self.share_first_product and not self.bf16:
            # The two views apply the SAME first-layer weight to the SAME features (the PPMI encoder is built on the
            # adjacency encoder's parameters, :168-169): x @ W0 -- the product that streams the [N, F] features --
            # is evaluated once, and so is its weight gradient (autograd sums the two views' gradients first).
            c0, p0 = self.encoder.conv_layers[0], self.ppmi_encoder.conv_layers[0]
            if c0.weight is p0.weight:
                xw = ops.graph_conv(data.x, c0.weight, None, None, 0, w_in_out=True)
                outs = [self._out(enc(data.x, data.edge_index, cache_name, first_product=xw))
                        for enc in (self.encoder, self.ppmi_encoder)]
                if mask is not None:
                    outs = [o[mask] for o in outs]
                return self.att_model(outs)
        gcn_output = self.gcn_encode(data, cache_name, mask)
        if self.ppmi:
            return self.att_model([gcn_output, self.ppmi_encode(data, cache_name, mask)])
        return gcn_output
