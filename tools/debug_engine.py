import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import dualvgr_oracle as orc
from test_engine_gpu import make
from dualvgr_videoqa_b200.engine import TrainEngine
import dualvgr_videoqa_b200.utils as U
cfg = (4, 8, 6, 10, 30, 1)
model, batch = make(cfg)
ref_model = copy.deepcopy(model)
eng = TrainEngine(model, lr=1e-3, max_norm=0.05)
opt = torch.optim.Adam(ref_model.parameters(), lr=1e-3)
N = cfg[1]
p0 = {n: p.detach().clone() for n, p in ref_model.named_parameters()}
eng.model.train(); eng.gflat.zero_()
out = eng.model(*batch[:4]); total, ce, com, dep, _ = eng.loss(out, batch[4]); total.backward()
g_eng = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
print("eng loss", float(total), float(ce), float(com), float(dep))
opt.zero_grad()
out = ref_model(*batch[:4]); logits, _, _, ca, cm, aq, mq = out
loss = torch.nn.functional.cross_entropy(logits, batch[4]); n = len(aq)
c2 = sum(U.common_loss(ca[i], cm[i]) for i in range(n)); d2 = sum(U.loss_dependence(aq[i], ca[i], N) + U.loss_dependence(mq[i], cm[i], N) for i in range(n))
loss = loss + c2 / n + 1e-8 * d2 / n
loss.backward()
print("ref loss", float(loss), float(c2), float(d2))
worst = []
for nme, p in ref_model.named_parameters():
    a, b = g_eng[nme].double(), p.grad.double()
    worst.append((float((a - b).norm() / b.norm().clamp_min(1e-30)), float(b.norm()), nme))
for w in sorted(worst)[-8:]: print("grad diff", w)
gn_e = float(eng.gflat.double().norm()); gn_r = float(torch.cat([p.grad.reshape(-1) for p in ref_model.parameters()]).double().norm())
print("grad norms", gn_e, gn_r)
eng.optimizer_step()
torch.nn.utils.clip_grad_norm_(ref_model.parameters(), max_norm=0.05); opt.step()
worst = []
for (n1, p1), (n2, p2) in zip(model.named_parameters(), ref_model.named_parameters()):
    d1, d2_ = (p1 - p0[n1]).double(), (p2 - p0[n1]).double()
    worst.append((float((d1 - d2_).norm() / d2_.norm().clamp_min(1e-30)), float(d2_.norm()), n1))
for w in sorted(worst)[-8:]: print("update diff", w)
