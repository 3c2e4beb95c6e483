"""Full-size smoke + timing of the other BASELINE.json configs (the bench line itself is configs[1]): one engine per config,
whole step captured as a CUDA graph, a few replays timed with CUDA events. Prints one line per config.
  config 3: MSRVTT-QA  N=16, U=3, A=4002, V=8000, B=256
  config 4: MSVD-QA    N=8,  U=1..5, A=1854, V=4000, B=1024
  config 5: synthetic  N=64, U=3, A=32, V=200, B=512 per GPU"""
import gc, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dualvgr_oracle as orc
import dualvgr_videoqa_b200.model.models as M
from dualvgr_videoqa_b200.engine import TrainEngine
from dualvgr_videoqa_b200 import autograd as ag

CONFIGS = [("cfg3 MSRVTT N=16 U=3 A=4002 B=256", 256, 16, 20, 4002, 8000, 3),
           ("cfg4 MSVD N=8 U=1 A=1854 B=1024", 1024, 8, 20, 1854, 4000, 1),
           ("cfg4 MSVD N=8 U=3 A=1854 B=1024", 1024, 8, 20, 1854, 4000, 3),
           ("cfg4 MSVD N=8 U=5 A=1854 B=1024", 1024, 8, 20, 1854, 4000, 5),
           ("cfg5 synthetic N=64 U=3 A=32 B=512", 512, 64, 20, 32, 200, 3)]
only = os.environ.get("ONLY")
dev = torch.device("cuda", 0)
for name, B, N, L, A, V, U in CONFIGS:
    if only and only not in name:
        continue
    torch.manual_seed(666)
    model = M.DualVGR(vocab=orc.make_vocab(V, A), num_of_nodes=N, graph_module="GAT", graph_layers=1, unit_layers=U)
    model = model.to(dev).train()
    eng = TrainEngine(model, lr=1e-4)
    g = torch.Generator().manual_seed(1)
    app = torch.randn((B, N, 16, 2048), generator=g).abs_().to(dev)
    mot = torch.randn((B, N, 2048), generator=g).abs_().to(dev)
    qlen = torch.randint(5, L + 1, (B,), generator=g); qlen[0] = L
    q = (torch.randint(2, V, (B, L), generator=g) * (torch.arange(L)[None] < qlen[:, None])).to(dev)
    ans = torch.randint(0, A, (B,), generator=g).to(dev)
    eng.capture(app, mot, q, qlen.to(dev), ans, warmup=2)
    for _ in range(2):
        eng.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        loss = eng.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.2f} ms/step = {B / ms * 1e3:.0f} samples/s, loss {float(loss):.4f}, finite {bool(torch.isfinite(loss))}, "
          f"lstm dependency timeouts {ag.lstm_seq_timeouts()}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    eng.close(); del eng, model, app, mot
    ag.SYNC_WORDS.clear()
    gc.collect(); torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
