"""Diagnostic: per-parameter CE-gradient norms of the CUDA model against a golden fixture (got / ref ratio), and the
full-gradient error against the live oracle at a chosen batch size. Usage: diag_grad_norms.py [golden name] [B N U]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import dualvgr_oracle as orc
import test_model_gpu as tm

name = sys.argv[1] if len(sys.argv) > 1 else "g2_B3_N20_U3"
g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=False)
cfg = [int(x) for x in g["cfg"]]
model, inputs, ans = tm.build(cfg, training=True)
out = model(*inputs)
ce = torch.nn.functional.cross_entropy(out[0], ans)
names = [str(n) for n in g["grad_names"]]
params = dict(model.named_parameters())
grads = torch.autograd.grad(ce, [params[n] for n in names], allow_unused=True)
got = np.array([0.0 if gr is None else float(gr.double().norm()) for gr in grads])
ref = g["f64_grad_ce"][:, 0]
print(name, "logits rel", tm.rel(out[0], g["f64_logits_train"]), "norm-vector rel err", np.linalg.norm(got - ref) / np.linalg.norm(ref))
order = np.argsort(-ref)[:14]
for i in order:
    print(f"  {names[i]:70s} got {got[i]:10.4f} ref {ref[i]:10.4f} ratio {got[i] / max(ref[i], 1e-30):.4f}")

if len(sys.argv) > 4:
    B, N, U = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    cfg = (B, N, 8, 32, 60, U)
    L, A, V = 8, 32, 60
    model, inputs, ans = tm.build(cfg, training=True)
    out = model(*inputs)
    ce = torch.nn.functional.cross_entropy(out[0], ans)
    pn = [n for n, _ in model.named_parameters()]
    grads = torch.autograd.grad(ce, [p for _, p in model.named_parameters()], allow_unused=True)
    sd = orc.cast_state_dict(orc.make_state_dict(U, A, V), torch.float64)
    for v in sd.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    app, mot, q, qlen, ans_c = orc.make_inputs(B, N, L, A, V)
    ref_out = orc.dualvgr_forward(sd, U, app.double(), mot.double(), q, qlen, training=True)
    ref_ce = torch.nn.functional.cross_entropy(ref_out[0], ans_c)
    ref_grads = torch.autograd.grad(ref_ce, [sd[n] for n in pn], allow_unused=True)
    num = den = 0.0
    rows = []
    for n, gr, rg in zip(pn, grads, ref_grads):
        if rg is None:
            continue
        gr = torch.zeros_like(rg) if gr is None else gr.double().cpu()
        e, r = float((gr - rg).pow(2).sum()), float(rg.pow(2).sum())
        num += e; den += r
        rows.append((e, r, float(gr.norm()) / max(float(rg.norm()), 1e-30), n))
    print(f"live oracle B={B} N={N} U={U}: logits rel {tm.rel(out[0], ref_out[0]):.3e}  global grad rel {(num / den) ** 0.5:.3e}")
    for e, r, ratio, n in sorted(rows)[-8:]:
        print(f"  {n:70s} err-share {e / num:.3f} own-rel {(e / max(r, 1e-300)) ** 0.5:.4f} norm-ratio {ratio:.4f}")
