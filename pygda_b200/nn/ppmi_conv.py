"""PPMIConv -- drop-in for pygda/nn/ppmi_conv.py:8-184: a ``CachedGCNConv`` whose cached graph is the
PPMI graph of random walks over the input graph.  Same constructor (``path_len=5``) and ``norm`` contract.

The reference builds that graph in pure Python (dict adjacency, 40 rounds x every node, Counter objects:
minutes to hours at benchmark scale, once per (layer, cache_name)); here it is built on the GPU
(pygda_b200/ppmi.py, csrc/ppmi.cu) in milliseconds.  NumPy's unseeded random stream cannot be reproduced,
so -- exactly like two runs of the reference -- two builds differ in their walks; everything downstream of
the visit counts is the reference's arithmetic (tests/test_gpu_ppmi.py)."""
from ..graph import IMPROVED, NORM_SYM_ROW, SELF_LOOPS, Graph
from ..ppmi import ppmi_edges
from .cached_gcn_conv import CachedGCNConv


class PPMIConv(CachedGCNConv):
    def __init__(self, in_channels, out_channels, weight=None, bias=None, improved=False, use_bias=True,
                 path_len=5, **kwargs):
        super().__init__(in_channels, out_channels, weight, bias, improved, use_bias, **kwargs)
        self.path_len = path_len

    def _ppmi_graph(self, edge_index, num_nodes, improved=False, seed=None):
        ei, w = ppmi_edges(edge_index, num_nodes, self.path_len, seed=seed)               # :98-172
        flags = SELF_LOOPS | NORM_SYM_ROW | (IMPROVED if improved else 0)                 # :174-184
        return Graph(ei, num_nodes, w, flags)

    def norm(self, edge_index, num_nodes, edge_weight=None, improved=False, dtype=None):
        """``(edge_index, normalised weights)`` like the reference's ``norm`` (:56-184); ``edge_weight`` is
        ignored there too (it is overwritten at :171)."""
        return self._ppmi_graph(edge_index, num_nodes, improved).coo()

    def _graph(self, edge_index, num_nodes, cache_name, edge_weight=None):
        if cache_name not in self.cache_dict:                                             # cached_gcn_conv.py:132-136
            self.cache_dict[cache_name] = self._ppmi_graph(edge_index, num_nodes, self.improved)
        return self.cache_dict[cache_name]
