"""Host-side profile (cProfile) of the AdaGCN graph-level step (BASELINE config 5 shape, one GPU): where do the
~80 ms per step of host time go?"""
import cProfile, io, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygda_b200.data import DataLoader, DeviceGraphDataset
from pygda_b200.models import AdaGCN
from pygda_b200.optim import Adam
from pygda_b200.synthetic import graph_dataset
from pygda_b200._lib import load
dev = torch.device("cuda", 0)
batch, graphs = 512, int(os.environ.get("GRAPHS", 6000))
hp = dict(in_dim=14, hid_dim=128, num_classes=2, mode="graph", num_layers=2, dropout=0.4, gnn_type="gcn", adv_dim=40,
          gp_weight=5.0, domain_weight=0.1, weight_decay=0.01, lr=0.01, epoch=400, device=str(dev), batch_size=batch, verbose=0)
ds_s = DeviceGraphDataset(graph_dataset(graphs, 30, 2.05, 14, 2, seed=50), dev)
ds_t = DeviceGraphDataset(graph_dataset(graphs, 39, 3.7, 14, 2, seed=51), dev)
torch.manual_seed(0)
model = AdaGCN(**hp)
model.adagcn = model.init_model()
opt = Adam(model.adagcn.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
model.init_critic()
loaders = (DataLoader(ds_s, batch_size=batch, shuffle=True, device=dev), DataLoader(ds_t, batch_size=batch, shuffle=True, device=dev))
def batches():
    while True:
        for sb, tb in zip(*loaders):
            yield sb, tb
it = batches()
def step():
    sb, tb = next(it)
    return model.train_step(sb, tb, opt)[0]
for _ in range(4): step()
torch.cuda.synchronize()
if os.environ.get("NCU_ONE_STEP") == "1":            # ncu --profile-from-start off: the launch list of ONE step
    torch.cuda.profiler.start(); step(); torch.cuda.synchronize(); torch.cuda.profiler.stop(); sys.exit(0)
lib = load(); n0 = lib.gda_launch_count()
t0 = time.perf_counter()
for _ in range(10): loss = step()
torch.cuda.synchronize()
print("ms per step %.2f, launches per step %.0f" % ((time.perf_counter() - t0) * 100, (lib.gda_launch_count() - n0) / 10))
pr = cProfile.Profile(); pr.enable()
for _ in range(10): loss = step()
torch.cuda.synchronize()
pr.disable()
for key in ("cumulative", "tottime"):
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats(key).print_stats(35); print(s.getvalue()[:6000])
