"""Debug aid: gradients arriving at the three encoders, fused unit stack vs module-by-module path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dualvgr_oracle as orc
import dualvgr_videoqa_b200.model.models as M
from dualvgr_videoqa_b200 import autograd as ag, fused_stack as fs

def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

cfg = (5, 20, 9, 16, 50, int(os.environ.get('U', '1')))
B, N, L, A, V, U = cfg
model = M.DualVGR(vocab=orc.make_vocab(V, A), num_of_nodes=N, graph_module="GAT", graph_layers=1, unit_layers=U)
model.load_state_dict(orc.make_state_dict(U, A, V), strict=True)
for m in model.modules():
    if isinstance(m, torch.nn.Dropout): m.p = 0.0
    if hasattr(m, "dropout") and isinstance(getattr(m, "dropout"), float): m.dropout = 0.0
model = model.cuda().train()
app, mot, q, qlen, ans = [t.cuda() for t in orc.make_inputs(B, N, L, A, V)]
g = torch.Generator().manual_seed(0)
D = 768
a0 = (torch.randn((B, N, D), generator=g) * 0.3).to(torch.bfloat16).cuda()
m0 = (torch.randn((B, N, D), generator=g) * 0.3).to(torch.bfloat16).cuda()
dq0 = (torch.randn((B, L, D), generator=g) * 0.3).to(torch.bfloat16).cuda()
w0 = torch.tanh(torch.randn((B, L, 300), generator=g)).cuda()
unit = model.visual_input_unit

def run(fused):
    ag.begin_forward()
    a, m, dq, w = [t.clone().requires_grad_(True) for t in (a0, m0, dq0, w0)]
    if fused:
        wp = torch.nn.functional.pad(w, (0, 4)).to(torch.bfloat16)
        x0 = torch.empty((2, B * N, D), dtype=torch.bfloat16, device="cuda")
        outs = unit.fused(a, m, dq.view(B * L, D), wp, qlen.to(torch.int32))
    else:
        outs = unit(a, m, dq, w, qlen)
    visual, aq, mq, ca, cm, aqf, mqf = outs
    loss = visual.float().pow(2).mean() + aq.float().pow(2).mean() + mq.float().pow(2).mean() + sum(t.float().pow(2).mean() for l in (ca, cm, aqf, mqf) for t in l)
    loss.backward()
    torch.cuda.synchronize()
    return [visual, aq, mq] + ca + cm + aqf + mqf, [a.grad, m.grad, dq.grad, w.grad], {n: p.grad.clone() for n, p in unit.named_parameters() if p.grad is not None}

o1, g1, p1 = run(True)
for p in unit.parameters(): p.grad = None
o2, g2, p2 = run(False)
print("outputs", [round(rel(x, y), 5) for x, y in zip(o1, o2)])
print("input grads (app, mot, dq, words)", [round(rel(x, y), 5) for x, y in zip(g1, g2)], [float(x.float().norm()) for x in g1], [float(x.float().norm()) for x in g2])
bad = [(n, round(rel(p1[n], p2[n]), 4)) for n in p2 if float(p2[n].norm()) > 1e-4 and rel(p1[n], p2[n]) > 3e-2]
print("unit param grads off:", len(bad), "of", len(p2))
for n in p2:
    if float(p2[n].norm()) > 1e-6:
        print(f"  {n:60s} {rel(p1[n], p2[n]):.4f}  {float(p1[n].norm()):.3e} {float(p2[n].norm()):.3e}")
