#!/usr/bin/env python
"""Benchmark of the GNNAE train step (BASELINE.json: jets/sec, fwd+bwd+Adam, Chamfer loss, N=30).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--num-nodes 30|150]
                    [--batch B] [--precision bf16|fp32]

One JSON line on stdout (rank 0).  Workload: reference CLI-default architecture, synthetic JetNet-shaped
jets, 4096 jets per GPU (BASELINE config 2 at N=1; 8 GPUs = config 4's 32768-jet global batch), weak scaling.
"value" times the step with the batch resident in HBM (CUDA events, L2 flushed between steps, max over
ranks); "e2e" goes through the public ``GNNAETrainer.step`` with a pinned host batch per step and a
device->host read of the loss.  ``--impl reference`` times the CPU PyTorch restatement of the reference's
train step (oracle/torch_port.py, "port") on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "jets/sec fwd+bwd GNNAE train step"
UNIT = "jets/s"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(burst=float(p["bf16_tflops"]), sustained=float(p["bf16_tflops_sustained"]), hbm=float(p["hbm_gbs"]),
                    source="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# --------------------------------------------------------------------------------------------------
# CPU baseline: torch-CPU restatement of the reference step on a bounded sample
# --------------------------------------------------------------------------------------------------
def cpu_reference_step_rate(num_nodes, sample_jets, steps, warmup, budget_s=25.0):
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch_port as TP
    from golden_cases import _linear  # noqa: F401  (test-infrastructure initialiser)
    import gnnae_oracle as O
    from gnn_jet_autoencoder_b200.config import DEFAULT_ARCH as A
    from gnn_jet_autoencoder_b200.trainer import synthetic_jets
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.default_rng(0)
    enc_cfg = dict(num_nodes=num_nodes, input_node_size=A["vec_dims"], latent_node_size=A["latent_node_size"],
                   node_sizes=A["node_sizes"], edge_sizes=A["edge_sizes"], num_mps=A["num_mps"], alphas=A["alphas"],
                   latent_map=A["latent_map"])
    dec_cfg = dict(enc_cfg, output_node_size=A["vec_dims"])
    ep = O.init_graphnet_params(rng, A["vec_dims"], A["latent_node_size"], A["node_sizes"], A["edge_sizes"], A["num_mps"],
                                prefix="encoder.", dtype=np.float32)
    h0 = A["node_sizes"][0][0]
    dp = {}
    dp["linear.weight"], dp["linear.bias"] = _linear(rng, num_nodes * h0, A["latent_node_size"])
    dp.update(O.init_graphnet_params(rng, h0, A["vec_dims"], A["node_sizes"], A["edge_sizes"], A["num_mps"],
                                     prefix="decoder.", dtype=np.float32))
    step = TP.TorchTrainStep(TP.make_params(ep), TP.make_params(dp), enc_cfg, dec_cfg)
    x = torch.from_numpy(synthetic_jets(sample_jets, num_nodes, seed=1234))
    for _ in range(max(1, warmup)):
        step.step(x)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        step.step(x)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(value=done * sample_jets / dt, unit=UNIT, cores=cores, kind="port",
                sample=f"{done} steps of {sample_jets} jets (N={num_nodes}) through oracle/torch_port.py "
                       f"(torch {torch.__version__} CPU, fp32, autograd + 2x Adam), {dt:.1f} s"), dt / done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_sample or (256 if args.num_nodes <= 32 else 16)
    cb, sec_per_step = cpu_reference_step_rate(args.num_nodes, sample, args.steps, min(args.warmup, 2), budget_s=120.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args, sample, 1),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, per_gpu_batch, world):
    return {"workload": f"GNNAE full training step (fwd+bwd+Adam, Chamfer loss) N={args.num_nodes} "
                        f"batch {per_gpu_batch} per GPU (BASELINE config {'2' if args.num_nodes == 30 else '3'})",
            "arch": "reference CLI defaults: edge [32,128,64,16], node [[16],[32],[8]], 3 MP steps, latent 20 'mean'",
            "num_nodes": args.num_nodes, "per_gpu_batch": per_gpu_batch, "global_batch": per_gpu_batch * world,
            "parallelism": f"dp{world} (batch shard, one flat-gradient all-reduce)",
            "l2": "flushed between timed steps (256 MiB write, outside the timed events)"}


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [c.strip() for c in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from gnn_jet_autoencoder_b200 import GNNAETrainer, ops, synthetic_jets
    from gnn_jet_autoencoder_b200.config import DEFAULT_ARCH, build_models, edge_macs_per_row, train_flops_per_jet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries ONE JSON line: NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) is dropped, any more verbose
        # debug level the caller asked for goes to stderr
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    N = args.num_nodes
    B = args.batch or (4096 if N <= 32 else 2048)
    enc, dec = build_models(N, DEFAULT_ARCH, device=dev, precision=args.precision, seed=0)
    tr = GNNAETrainer(enc, dec, batch_size=B, use_cuda_graph=not args.no_graph)
    # a few distinct pinned host batches (seed = 1234 + rank), cycled by the e2e loop
    host = [torch.from_numpy(synthetic_jets(B, N, seed=1234 + rank + 101 * i)).pin_memory() for i in range(4)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rank_max(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    # ---- warm-up (also builds the CUDA graph) ----
    tr.load_batch(host[0])
    for _ in range(max(args.warmup, 3)):
        tr.compute_gradients()
        tr.apply_gradients()
    barrier()

    # ---- device-resident timing: K steps, events around each step, L2 flushed in between ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = ops.LAUNCHES["count"]
    barrier()
    for s, e in ev:
        flush.fill_(1)
        s.record()
        tr.compute_gradients()
        tr.apply_gradients()
        e.record()
    barrier()
    launches = ops.LAUNCHES["count"] - launches0
    ms_total = rank_max(sum(s.elapsed_time(e) for s, e in ev))
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- end to end: pinned host batch in, loss out, every step, through the public API ----
    for i in range(2):
        tr.step(host[i % len(host)])
    barrier()
    s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    last = 0.0
    for i in range(args.steps):
        last = tr.step(host[i % len(host)])
    e0.record()
    barrier()
    e2e_ms = rank_max(s0.elapsed_time(e0)) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- dominant kernel alone: backward of the heaviest step (encoder step 1), CUDA events on its stream ----
    peaks = measured_peaks()
    kern = time_dominant_kernel(tr, args) if rank == 0 else None
    if world > 1:
        dist.barrier()

    if rank == 0:
        flops_jet = train_flops_per_jet(N, DEFAULT_ARCH)
        step_tflops = value / world * flops_jet / 1e12          # per GPU
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "fp32", "data": "synthetic",
            "config": workload_config(args, B, world),
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": B * N * 3 * 4, "d2h_bytes_per_step": tr.stats_host.numel() * 4,
                    "api": "GNNAETrainer.step(pinned host batch) -> float loss"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": kern and {
                "bound": "tensor", "achieved": kern["tflops"], "peak": peaks["burst"], "unit": "TFLOP/s",
                "frac": kern["tflops"] / peaks["burst"], "traffic": kernel_traffic(args, B, N), "kernel": kern["name"],
                "us_per_launch": kern["us"], "algorithmic_flop_per_launch": kern["flop"], "peak_source": peaks["source"] + ", burst",
            },
            "step_roofline": {"bound": "tensor", "achieved": step_tflops, "peak": peaks["sustained"], "unit": "TFLOP/s",
                              "frac": step_tflops / peaks["sustained"], "flop_per_jet": flops_jet,
                              "peak_source": peaks["source"] + ", sustained", "note": "whole train step, per GPU, edge-MLP "
                              "dense-formulation FLOPs (3 x forward); padding and recomputation not counted"},
            "last_loss": last,
        }
        if world == 1 and not args.no_cpu_baseline:
            sample = args.cpu_sample or (256 if N <= 32 else 16)
            line["cpu_baseline"], _ = cpu_reference_step_rate(N, sample, steps=8, warmup=1, budget_s=20.0)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def kernel_traffic(args, B, N):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the round's
    `ncu --set full` capture (profiles/r01_edge_bwd2_ncu.txt: 73.85 MB read + 12.46 MB written at B=4096, N=30, bf16);
    null for workloads that were not captured."""
    return 86.3e6 if (args.precision == "bf16" and B == 4096 and N == 30) else None


def time_dominant_kernel(tr, args, reps=20):
    import torch
    from gnn_jet_autoencoder_b200 import ops
    from gnn_jet_autoencoder_b200.config import DEFAULT_ARCH
    # heaviest step: the one with the most edge MACs per row
    steps = tr.enc_steps + tr.dec_steps
    def macs(s):
        d = s["desc"]
        w = [2 * d.node_in + 1] + [d.edge_widths[i] for i in range(d.n_edge_layers)]
        return sum(a * b for a, b in zip(w[:-1], w[1:]))
    idx = max(range(len(tr.enc_steps)), key=lambda i: macs(tr.enc_steps[i]))
    s = tr.enc_steps[idx]
    hin = tr.x if idx == 0 else tr.enc_steps[idx - 1]["out"]
    g = tr.d_enc_out if idx == len(tr.enc_steps) - 1 else tr.enc_steps[idx + 1]["din"]
    din = tr.dx if idx == 0 else s["din"]
    st = torch.cuda.current_stream().cuda_stream
    P, G = tr.flat.data_ptr(), tr.grad.data_ptr()
    def launch():
        ops.raw_mp_bwd(s["desc"], hin.data_ptr(), s["e"].data_ptr(), P + 4 * s["off"], g.data_ptr(), din.data_ptr(),
                       G + 4 * s["off"], tr.ws.data_ptr(), tr.ws_bytes, st)
    from gnn_jet_autoencoder_b200 import _lib
    lib = _lib.load()
    def launch_kernel_only():
        # the fused backward edge kernel ALONE, on the workspace the full call above has populated
        return lib.gj_bench_edge_bwd_only(s["desc"], hin.data_ptr(), P + 4 * s["off"], din.data_ptr(), G + 4 * s["off"],
                                          tr.ws.data_ptr(), tr.ws_bytes, st)
    for _ in range(3):
        launch()
    kernel_only = launch_kernel_only() == 0
    fn = launch_kernel_only if kernel_only else launch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / reps
    flop = 2.0 * 2.0 * macs(s) * tr.N * tr.N * tr.B       # dgrad + wgrad of the dense formulation
    name = (f"edge_bwd2_kernel (fused recompute + dgrad + wgrad of encoder step {idx}, B={tr.B}, N={tr.N}; launched alone "
            f"through gj_bench_edge_bwd_only)") if kernel_only else \
           f"gj_mp_step_bwd (encoder step {idx}, B={tr.B}, N={tr.N}, all of its launches)"
    return dict(name=name, us=us, flop=flop, tflops=flop / (us * 1e-6) / 1e12)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--num-nodes", type=int, default=30)
    ap.add_argument("--batch", type=int, default=0, help="jets per GPU (default 4096 at N=30, 2048 at N=150)")
    ap.add_argument("--precision", default=os.environ.get("GNNJET_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="jets per CPU-baseline step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
