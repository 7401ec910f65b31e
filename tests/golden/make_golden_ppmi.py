"""Golden vectors for the PPMI path, made by EXECUTING THE REFERENCE'S OWN FILES
(pygda/nn/ppmi_conv.py, pygda/nn/udagcn_base.py, pygda/models/udagcn.py) with np.random seeded:

    python tests/golden/make_golden_ppmi.py        # build container only (needs /root/reference)

Writes tests/golden/ppmi.pt: (1) PPMIConv.norm on a small graph (path_len 5 and 10), (2)
UDAGCN.forward_model with ppmi=True (the reference's default), loss / logits / gradients, together with the
PPMI graphs its four PPMIConv caches ended up holding -- so that the GPU path, whose walks come from another
random stream, can be checked with those graphs injected."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference, REPO  # noqa: E402

sys.path.insert(0, REPO)
from make_golden import small_graph  # noqa: E402


def main():
    ref = load_reference()
    out = {}
    g = small_graph(60, 200, 8, 3, seed=11)
    cases = {}
    for path_len in (5, 10):
        conv = ref.ppmi_conv.PPMIConv(8, 4, path_len=path_len)
        np.random.seed(100 + path_len)
        ei, w = conv.norm(g.edge_index, 60)
        cases[path_len] = {"np_seed": 100 + path_len, "edge_index_out": ei, "weight_out": w}
    out["norm"] = {"edge_index": g.edge_index, "num_nodes": 60, "cases": cases}

    src, tgt = small_graph(50, 160, 24, 4, seed=3), small_graph(44, 150, 24, 4, seed=4, with_loops=False)
    torch.manual_seed(21)
    est = ref.udagcn.UDAGCN(in_dim=24, hid_dim=16, num_classes=4, mode='node', num_layers=2, ppmi=True,
                            adv_dim=10, epoch=300, device='cpu')
    est.udagcn = est.init_model()
    net = est.udagcn
    # dropout RNG streams cannot match: identities in the always-on encoder lists, eval for the domain model
    net.encoder.dropout_layers = [torch.nn.Identity() for _ in net.encoder.dropout_layers]
    net.ppmi_encoder.dropout_layers = [torch.nn.Identity() for _ in net.ppmi_encoder.dropout_layers]
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.uniform_(-0.1, 0.1)
    for m in net.models:
        m.eval()
    state = {k: v.clone() for k, v in net.state_dict().items()}
    np.random.seed(77)
    loss, s_logits, t_logits = est.forward_model(src, tgt, 0.04, 120)
    net.zero_grad()
    loss.backward()
    caches = {}
    for li, conv in enumerate(net.ppmi_encoder.conv_layers):
        for name, (ei, w) in conv.cache_dict.items():
            caches[(li, name)] = (ei.clone(), w.clone())
    out["udagcn_ppmi"] = {
        "source": {"x": src.x, "edge_index": src.edge_index, "y": src.y},
        "target": {"x": tgt.x, "edge_index": tgt.edge_index, "y": tgt.y},
        "hparams": dict(in_dim=24, hid_dim=16, num_classes=4, num_layers=2, ppmi=True, adv_dim=10, epoch=300),
        "alpha": 0.04, "epoch": 120, "np_seed": 77, "state": state, "ppmi_caches": caches,
        "loss": loss.detach().clone(), "source_logits": s_logits.detach().clone(),
        "target_logits": t_logits.detach().clone(),
        "grads": {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}}
    torch.save(out, os.path.join(HERE, "ppmi.pt"))
    print("wrote ppmi.pt", os.path.getsize(os.path.join(HERE, "ppmi.pt")), "bytes")


if __name__ == "__main__":
    main()
