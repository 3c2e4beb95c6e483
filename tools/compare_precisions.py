"""Per-parameter full-loss gradient of bf16 mode against fp32 mode (same weights / inputs): where does bf16 lose the aux gradient?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import dualvgr_oracle as orc
from test_model_gpu import build, total_loss
import dualvgr_videoqa_b200.autograd as ag

cfg = tuple(int(x) for x in os.environ.get("CFG", "16,20,8,32,60,3").split(","))
which = os.environ.get("LOSS", "total")
res = {}
for mode in ("fp32", "bf16"):
    model, inputs, ans = build(cfg, True)
    model.set_precision(mode)
    out = model(*inputs)
    total, ce, com, dep = total_loss(out, ans, cfg[1])
    n = len(out[3])
    loss = {"total": total, "ce": ce, "com": com / n, "dep": 1e-8 * dep / n}[which]
    names = [k for k, _ in model.named_parameters()]
    grads = torch.autograd.grad(loss, [p for _, p in model.named_parameters()], allow_unused=True)
    res[mode] = {k: (torch.zeros(1) if g is None else g.double().cpu()) for k, g in zip(names, grads)}
    print(mode, "losses", float(total), float(ce), float(com), float(dep))
ag.ACT[0] = torch.bfloat16
num = den = 0.0
rows = []
for k in res["fp32"]:
    a, b = res["bf16"][k], res["fp32"][k]
    e, r = float((a - b).pow(2).sum()), float(b.pow(2).sum())
    num += e; den += r
    rows.append((e, r, k))
print(f"global rel-L2 bf16 vs fp32 mode ({which}): {(num / den) ** 0.5:.3e}")
for e, r, k in sorted(rows)[-int(os.environ.get('TOP', '25')):]:
    print(f"  err-share {e / num:6.3f}  own-rel {(e / max(r, 1e-300)) ** 0.5:9.3e}  norm {r ** 0.5:9.3e}  {k}")
