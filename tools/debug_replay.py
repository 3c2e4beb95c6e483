import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dualvgr_oracle as orc
import dualvgr_videoqa_b200.model.models as M
from dualvgr_videoqa_b200.engine import TrainEngine
cfg = (6, 20, 8, 32, 60, 2)
B, N, L, A, V, U = cfg
model = M.DualVGR(vocab=orc.make_vocab(V, A), num_of_nodes=N, graph_module="GAT", graph_layers=1, unit_layers=U)
model.load_state_dict(orc.make_state_dict(U, A, V), strict=True)
for m in model.modules():
    if isinstance(m, torch.nn.Dropout): m.p = 0.0
    if hasattr(m, "dropout") and isinstance(getattr(m, "dropout"), float): m.dropout = 0.0
model = model.cuda().train()
batch = [t.cuda() for t in orc.make_inputs(B, N, L, A, V)]
eng = TrainEngine(model, lr=1e-5)
for i in range(2):
    eng.train_step(*batch); torch.cuda.synchronize()
    print("eager", eng.last_stats.tolist(), "gflat finite", bool(torch.isfinite(eng.gflat).all()), "flat finite", bool(torch.isfinite(eng.flat).all()))
eng.capture(*batch, warmup=2)
for i in range(3):
    eng.replay(); torch.cuda.synchronize()
    print("replay", eng.last_stats.tolist(), "gflat finite", bool(torch.isfinite(eng.gflat).all()), "flat finite", bool(torch.isfinite(eng.flat).all()),
          "logits finite", bool(torch.isfinite(eng.last_logits).all()))
    if not torch.isfinite(eng.gflat).all():
        off = 0
        names = {id(p): n for n, p in model.named_parameters()}
        for p, n in zip(eng.params, eng.sizes):
            g = eng.gflat[off:off + p.numel()]
            if not torch.isfinite(g).all():
                print("   non-finite grad:", names[id(p)])
            off += n
        break
