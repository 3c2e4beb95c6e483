#!/usr/bin/env python
"""bench.py — DualVGR train-step throughput on N B200s of one node (contract: see the task statement / DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (sm_100a kernels through the nn.Module mirror)
  python bench.py --impl reference [...]                          the reference's CPU algorithm (oracle port) on host cores

Workload (BASELINE.json configs[1]): SVQA shapes, N=20 clips x 16 frames x 2048-d appearance + 2048-d motion, GloVe-300
question of L=20 tokens, unit_layers=3, batch 256 per GPU, bf16 activations, full train step
(forward, CE + common + HSIC losses, backward, gradient all-reduce, clip 12, Adam). Synthetic data, seeded weights.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(B=256, N=20, L=20, A=32, V=200, U=3, F=16, Dv=2048)
METRIC, UNIT = "dualvgr_train_samples_per_sec", "samples/s"


def workload_name(n_gpus):
    return (f"SVQA config (svqa_DualVGR_20.yml shapes): N={CFG['N']} clips x {CFG['F']} frames x {CFG['Dv']}-d, L={CFG['L']}, "
            f"A={CFG['A']}, unit_layers={CFG['U']}, batch {CFG['B']}/GPU x {n_gpus} GPU, full train step")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clock / throttle sampling DURING the timed region (recipe line of B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------- reference arm
def cpu_reference_step_rate(sample_B, steps, warmup, threads):
    """The reference's algorithm for the path (oracle port, fp32) on the host cores: full train step on a bounded sample."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dualvgr_oracle as orc
    torch.set_num_threads(threads)
    c = CFG
    sd = orc.make_state_dict(c["U"], c["A"], c["V"])
    params = []
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
            params.append(v)
    opt = torch.optim.Adam(params, lr=1e-4)
    app, mot, q, qlen, ans = orc.make_inputs(sample_B, c["N"], c["L"], c["A"], c["V"])
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        out = orc.dualvgr_forward(sd, c["U"], app, mot, q, qlen, training=True)
        total, ce, com, dep = orc.train_loss(out, ans, c["N"])
        total.backward()
        torch.nn.utils.clip_grad_norm_(params, 12)
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sample_B / statistics.median(times), statistics.median(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_B = 16
    rate, sec = cpu_reference_step_rate(sample_B, max(1, args.steps), max(0, min(args.warmup, 2)), threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus), "note": "CPU arm: each step is a bounded sample of the workload"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"oracle port (oracle/dualvgr_oracle.py, fp32, torch CPU) full train step on {sample_B} of the "
                                   f"{CFG['B']} samples of the workload batch; the Python reference cannot travel to the GPU box"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dualvgr_oracle as orc                      # ONLY make_state_dict / make_inputs / cpu_baseline (checker side)
    import dualvgr_videoqa_b200._lib as L
    import dualvgr_videoqa_b200.model.models as M
    from dualvgr_videoqa_b200.engine import TrainEngine
    from dualvgr_videoqa_b200 import autograd as ag

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to STDOUT when the first communicator is created; rank 0 must print exactly one
        # JSON line there, so file descriptor 1 points at stderr until the engine (and its first collective) exists
        sys.stdout.flush()
        saved_stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    c = CFG
    torch.manual_seed(666)
    model = M.DualVGR(vocab=orc.make_vocab(c["V"], c["A"]), num_of_nodes=c["N"], graph_module="GAT", graph_layers=1,
                      unit_layers=c["U"])
    model.load_state_dict(orc.make_state_dict(c["U"], c["A"], c["V"]), strict=True)
    model = model.to(dev).train()
    eng = TrainEngine(model, lr=1e-4, max_norm=12.0, alpha=1.0, beta=1e-8)
    if world > 1:
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved_stdout_fd, 1)
        os.close(saved_stdout_fd)

    # synthetic shard of this rank, generated on the host, pinned; a device-resident copy for the kernel-side number
    g = torch.Generator().manual_seed(1000 + rank)
    host = {
        "app": torch.randn((c["B"], c["N"], c["F"], c["Dv"]), generator=g).abs_().pin_memory(),
        "mot": torch.randn((c["B"], c["N"], c["Dv"]), generator=g).abs_().pin_memory(),
    }
    qlen = torch.randint(5, c["L"] + 1, (c["B"],), generator=g); qlen[0] = c["L"]
    q = torch.randint(2, c["V"], (c["B"], c["L"]), generator=g) * (torch.arange(c["L"])[None] < qlen[:, None])
    host["q"], host["qlen"] = q.long().pin_memory(), qlen.long().pin_memory()
    host["ans"] = torch.randint(0, c["A"], (c["B"],), generator=g).pin_memory()
    res = {k: v.to(dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    use_graph = not args.no_graph
    if use_graph:
        eng.capture(res["app"], res["mot"], res["q"], res["qlen"], res["ans"], warmup=max(3, args.warmup))

    def step_resident():
        if use_graph:
            return eng.replay()
        return eng.train_step(res["app"], res["mot"], res["q"], res["qlen"], res["ans"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.launch_count() - n0
    if use_graph:
        launches = eng.launches_per_replay * args.steps      # kernels of libdualvgr_b200.so inside each replayed graph
    clocks = sampler.stop() if rank == 0 else None
    # the dominant kernel — lstm_seq_fwd_kernel: the appearance encoder's input projection (K = 2048) and its 16 recurrent
    # steps (K = 384) fused in ONE persistent tcgen05 launch — timed live with CUDA events on its launching stream at the
    # workload's exact shape; operands (335 MB of features, 503 MB of gates out) exceed the 126 MB L2
    import dualvgr_videoqa_b200.ops as ops
    T_, S_, H_ = c["F"], c["B"] * c["N"], 384
    xa = (torch.randn((T_, S_, c["Dv"]), device=dev) * 0.5).to(torch.bfloat16)
    wih = (torch.randn((8 * H_, c["Dv"]), device=dev) * 0.02).to(torch.bfloat16)
    whh = (torch.randn((2, 4 * H_, H_), device=dev) * 0.05).to(torch.bfloat16)
    bih = torch.zeros(8 * H_, device=dev)
    r_ = None
    for _ in range(3):
        r_ = None                                   # one output set live at a time: the caching allocator reuses its blocks,
        r_ = ops.lstm_seq_fwd(xa, wih, whh, bih)     # so no cudaMalloc lands between the two events below
    gemm_ms, seq_timeouts = [], 0
    for _ in range(10):
        r_ = None
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r_ = ops.lstm_seq_fwd(xa, wih, whh, bih); b.record()
        torch.cuda.synchronize()
        gemm_ms.append(a.elapsed_time(b))
        seq_timeouts += int(r_[5][-1])
    seq_timeouts += ag.lstm_seq_timeouts()            # the train steps' own launches (must be 0: no dependency poll gave up)
    del xa, r_
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = c["B"] * world * args.steps / (ms_max / 1e3)

    # ---- e2e: same step through the public API with HOST buffers; H2D of every step's inputs inside the timed region
    #      (double-buffered on a copy stream, as an input pipeline would), D2H read of the loss every step
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    loss_host = torch.zeros(args.steps + 8, dtype=torch.float32).pin_memory()

    def measure_e2e(host_t):
        bufs = [{k: torch.empty_like(v, device=dev) for k, v in host_t.items()} for _ in range(2)]

        def prefetch(i):
            b = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[b])
                for k, v in host_t.items():
                    bufs[b][k].copy_(v, non_blocking=True)
                ready[b].record(copy_stream)

        def run_e2e(n):
            cur = torch.cuda.current_stream()
            for b in range(2):
                consumed[b].record(cur)
            prefetch(0)
            for i in range(n):
                if i + 1 < n:
                    prefetch(i + 1)
                b = i % 2
                cur.wait_event(ready[b])
                if use_graph:   # device-to-device hand-over of the staged batch into the graph's static inputs, then ONE launch
                    eng.load_batch(bufs[b]["app"], bufs[b]["mot"], bufs[b]["q"], bufs[b]["qlen"], bufs[b]["ans"])
                    consumed[b].record(cur)
                    lo = eng.replay()
                else:
                    lo = eng.train_step(bufs[b]["app"], bufs[b]["mot"], bufs[b]["q"], bufs[b]["qlen"], bufs[b]["ans"])
                    consumed[b].record(cur)
                loss_host[i].copy_(lo, non_blocking=True)

        run_e2e(2)
        barrier()
        e0.record()
        run_e2e(args.steps)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        del bufs
        return c["B"] * world * args.steps / (float(t.item()) / 1e3)

    e2e_value = e2e_bf16 = None
    h2d_bf16 = None
    if not args.no_e2e:
        e2e_value = measure_e2e(host)          # the reference's data format: fp32 features (DataLoader.py:61-84)
        if use_graph and not args.no_bf16_e2e:
            # informational second leg: the same step with the features STORED as bf16 on the host, i.e. half
            # the bytes on the PCIe link, which is what bounds the fp32 leg; needs its own captured graph (bf16 prologue)
            host16 = dict(host)
            host16["app"] = host["app"].to(torch.bfloat16).pin_memory()
            host16["mot"] = host["mot"].to(torch.bfloat16).pin_memory()
            h2d_bf16 = sum(v.numel() * v.element_size() for v in host16.values())
            eng.graph = None
            res16 = {k: v.to(dev) for k, v in host16.items()}
            eng.capture(res16["app"], res16["mot"], res16["q"], res16["qlen"], res16["ans"], warmup=1)
            e2e_bf16 = measure_e2e(host16)

    if rank == 0:
        pk = peaks()
        # algorithmic FLOPs: W_ih product of both directions + the recurrent products of steps 1..T-1 (h_0 = 0)
        flops = 2.0 * (c["B"] * c["N"]) * (8 * 384) * (c["F"] * c["Dv"] + (c["F"] - 1) * 384)
        gemm_avg = statistics.mean(gemm_ms) if gemm_ms else float("nan")
        achieved = flops / (gemm_avg * 1e-3) / 1e12 if gemm_ms else None
        traffic = None
        prof = os.path.join(ROOT, "profiles", "dominant_kernel.json")
        if os.path.exists(prof):
            pj = json.load(open(prof))
            if str(pj.get("kernel", "")).startswith("lstm_seq_fwd_kernel"):     # only a capture of THIS kernel counts
                traffic = pj.get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(world), "parallelism": f"dp{world}" if world > 1 else "single",
                       "l2": "no explicit flush: each step streams 713 MB of fp32 features + ~1.5 GB of intermediates, far above the 126 MB L2",
                       "optimizer": "clip 12 + Adam lr 1e-4 (flat fused)", "dropout": "on (train mode, reference rates)",
                       "cuda_graph": bool(use_graph),
                       "final_loss": float(loss)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "note": "host buffers in the reference's format (fp32 features): bound by the host link, not by the kernels"},
            "e2e_bf16_features": {"value": e2e_bf16, "unit": UNIT, "h2d_bytes_per_step": h2d_bf16, "d2h_bytes_per_step": 4,
                                  "note": "informational: same step with the clip features stored as bf16 on the host"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "lstm_seq_fwd_kernel<BN=256> (appearance encoder forward: W_ih product + 16 recurrent steps, one persistent launch)",
                         "bound": "tensor", "achieved": achieved, "peak": pk["tf"], "unit": "TFLOP/s",
                         "frac": (achieved / pk["tf"]) if achieved else None, "traffic": traffic,
                         "peak_source": pk["src"] + ", sustained bf16 figure",
                         "timing": "mean of 10 isolated launches at the workload shape, CUDA events on the launch stream",
                         "launch_ms": gemm_avg, "dependency_poll_timeouts": seq_timeouts},
        }
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            rate, sec = cpu_reference_step_rate(8, 3, 1, threads)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "oracle port, fp32, full train step on 8 of the 256 samples of the batch, "
                                              "median of 3 steps after 1 warm-up"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # the captured graph holds NCCL work objects; tearing the communicator down under it can block forever, so
        # synchronise, drop the graph, and leave without the (optional) destroy_process_group handshake
        dist.barrier()
        torch.cuda.synchronize()
        eng.close()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    if os.environ.get("DVGR_HANG_DUMP"):          # debugging aid: dump all Python stacks if the run stalls
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["DVGR_HANG_DUMP"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true", help="profiling aid: skip the host-buffer e2e leg")
    ap.add_argument("--no-bf16-e2e", action="store_true", help="skip the informational e2e leg with bf16-stored features")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying the captured CUDA graph")
    ap.add_argument("--no-cpu", action="store_true", help="profiling aid: skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
