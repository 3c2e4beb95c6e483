"""In-situ kernel timing of the captured train step with torch.profiler (CUPTI): warm caches, real overlap."""
import os, sys, collections, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dualvgr_oracle as orc
import dualvgr_videoqa_b200.model.models as M
from dualvgr_videoqa_b200.engine import TrainEngine
import bench

import torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
c = dict(bench.CONFIGS[os.environ.get('CONFIG', 'svqa')]); c['F'], c['Dv'] = bench.F_, bench.DV
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
model = M.DualVGR(vocab=orc.make_vocab(c["V"], c["A"]), num_of_nodes=c["N"], graph_module="GAT", graph_layers=1, unit_layers=c["U"])
model.load_state_dict(orc.make_state_dict(c["U"], c["A"], c["V"]), strict=True)
model = model.to(dev).train()
eng = TrainEngine(model)
g = torch.Generator().manual_seed(1 + int(os.environ.get("RANK", "0")))
app = torch.randn((c["B"], c["N"], c["F"], c["Dv"]), generator=g).abs_().to(dev)
mot = torch.randn((c["B"], c["N"], c["Dv"]), generator=g).abs_().to(dev)
qlen = torch.randint(5, c["L"] + 1, (c["B"],), generator=g); qlen[0] = c["L"]
q = (torch.randint(2, c["V"], (c["B"], c["L"]), generator=g) * (torch.arange(c["L"])[None] < qlen[:, None])).to(dev)
ans = torch.randint(0, c["A"], (c["B"],), generator=g).to(dev)
eng.capture(app, mot, q, qlen.to(dev), ans)
for _ in range(3):
    eng.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        eng.replay()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
tot = 0.0
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"\(.*", "", ev.name)[:100]
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        tot += a[1] * 0
tot = sum(v[1] for v in agg.values())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    eng.replay()
e1.record(); torch.cuda.synchronize()
print(f"wall (CUDA events, 10 replays): {e0.elapsed_time(e1)/10:.3f} ms/step")
print(f"kernel time per step: {tot/3/1e3:.3f} ms over {sum(v[0] for v in agg.values())//3} launches")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{us/3/1e3:8.3f} ms {100*us/tot:5.1f}%  x{n//3:<4d} {name}")

# ---- timeline of ONE replay: start offset, duration, stream of every kernel (to read the critical path / overlap)
with profile(activities=[ProfilerActivity.CUDA]) as prof2:
    eng.replay()
    torch.cuda.synchronize()
evs = [e for e in prof2.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
print("---- timeline (us from first kernel): start  dur  end  name")
for e in evs:
    st = e.time_range.start - t0
    print(f"{st:9.1f} {e.time_range.end - e.time_range.start:8.1f} {e.time_range.end - t0:9.1f}  {re.sub(r'[(<].*', '', e.name)[:60]}")

if world > 1:
    dist.barrier(); torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)
