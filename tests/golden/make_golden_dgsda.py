"""Golden vectors for DGSDA's Bernstein propagation, made by EXECUTING THE REFERENCE'S OWN FILES
(pygda/nn/dgsda_base.py, pygda/models/dgsda.py):

    python tests/golden/make_golden_dgsda.py        # build container only (needs /root/reference)

Writes tests/golden/dgsda.pt: (1) BernProp forward + backward (input and temp gradients) on a small graph,
(2) DGSDA.forward_model (CE + temp L1 + MMD + weighted target entropy), loss / logits / all gradients."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference, REPO  # noqa: E402

sys.path.insert(0, REPO)
from make_golden import small_graph  # noqa: E402


def main():
    ref = load_reference()
    out = {}
    torch.manual_seed(5)
    g = small_graph(60, 200, 8, 3, seed=11)                  # has self loops and a duplicate edge
    cases = {}
    for K in (1, 4, 8):
        prop = ref.dgsda_base.BernProp(K)
        with torch.no_grad():
            prop.temp.copy_(torch.linspace(1.3, -0.2, K + 1))          # a negative entry: relu(temp)
        x = torch.randn(60, 6, requires_grad=True)
        y = prop(x, g.edge_index)
        go = torch.randn_like(y)
        y.backward(go)
        cases[K] = {"temp": prop.temp.detach().clone(), "x": x.detach().clone(), "y": y.detach().clone(), "gout": go,
                    "gx": x.grad.clone(), "gtemp": prop.temp.grad.clone()}
    out["bernprop"] = {"edge_index": g.edge_index, "num_nodes": 60, "cases": cases}

    src, tgt = small_graph(50, 160, 24, 4, seed=3), small_graph(44, 150, 24, 4, seed=4, with_loops=False)
    hp = dict(in_dim=24, hid_dim=16, num_classes=4, K=6, alpha=0.05, beta=0.5, gamma=0.05, dropout=0.0)
    torch.manual_seed(41)
    est = ref.dgsda.DGSDA(device='cpu', **hp)
    est.dgsda = est.init_model()
    with torch.no_grad():
        est.dgsda.prop2.temp.copy_(torch.linspace(1, 0.1, hp["K"] + 1))
        est.dgsda.prop3.temp.copy_(torch.linspace(0.9, 0.3, hp["K"] + 1))
    est.dgsda.train()
    state = {k: v.clone() for k, v in est.dgsda.state_dict().items()}
    torch.manual_seed(43)                                    # MMD sample indices (pygda/utils/mmd.py:148-149)
    loss, s_logits = est.forward_model(src, tgt)
    est.dgsda.zero_grad()
    loss.backward()
    out["dgsda"] = {"source": {"x": src.x, "edge_index": src.edge_index, "y": src.y},
                    "target": {"x": tgt.x, "edge_index": tgt.edge_index, "y": tgt.y},
                    "hparams": hp, "seed": 43, "state": state, "loss": loss.detach().clone(),
                    "source_logits": s_logits.detach().clone(),
                    "grads": {k: p.grad.clone() for k, p in est.dgsda.named_parameters() if p.grad is not None}}
    torch.save(out, os.path.join(HERE, "dgsda.pt"))
    print("wrote dgsda.pt", os.path.getsize(os.path.join(HERE, "dgsda.pt")), "bytes")


if __name__ == "__main__":
    main()
