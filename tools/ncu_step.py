"""One EAGER train step (no CUDA graph) inside a cudaProfilerStart/Stop range, for `ncu --profile-from-start off`:
   ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/X python tools/ncu_step.py
CONFIG = svqa | msrvtt | msvd_u1..u5 | clip64 (bench.CONFIGS); BATCH overrides the per-GPU batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dualvgr_oracle as orc
import dualvgr_videoqa_b200.model.models as M
from dualvgr_videoqa_b200.engine import TrainEngine
import bench

c = dict(bench.CONFIGS[os.environ.get("CONFIG", "svqa")]); c["F"], c["Dv"] = bench.F_, bench.DV
if os.environ.get("BATCH"):
    c["B"] = int(os.environ["BATCH"])
dev = torch.device("cuda", 0)
model = M.DualVGR(vocab=orc.make_vocab(c["V"], c["A"]), num_of_nodes=c["N"], graph_module="GAT", graph_layers=1, unit_layers=c["U"])
model.load_state_dict(orc.make_state_dict(c["U"], c["A"], c["V"]), strict=True)
model = model.to(dev).train()
eng = TrainEngine(model)
g = torch.Generator().manual_seed(1)
app = torch.randn((c["B"], c["N"], c["F"], c["Dv"]), generator=g).abs_().to(dev)
mot = torch.randn((c["B"], c["N"], c["Dv"]), generator=g).abs_().to(dev)
qlen = torch.randint(5, c["L"] + 1, (c["B"],), generator=g); qlen[0] = c["L"]
q = (torch.randint(2, c["V"], (c["B"], c["L"]), generator=g) * (torch.arange(c["L"])[None] < qlen[:, None])).to(dev)
ans = torch.randint(0, c["A"], (c["B"],), generator=g).to(dev)
for _ in range(int(os.environ.get("WARMUP", "2"))):
    eng.train_step(app, mot, q, qlen.to(dev), ans)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
loss = eng.train_step(app, mot, q, qlen.to(dev), ans)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(loss), "timeouts", eng.dependency_poll_timeouts())
