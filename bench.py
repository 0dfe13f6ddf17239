#!/usr/bin/env python
"""GKN training throughput on B200 — the benchmark contract of this repo.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): BBA all-atom-sized frames (504 atoms, 8 A radius graph,
~57k directed edges/frame), KernelNN width 64, kernel_width 1024, depth 6 (12 NNConv
applications), window 10, bf16 tensor-core operands with fp32 accumulation, 16 frames per GPU per
step (the reference's default batch of 128 over 8 GPUs).  One step = one pass of the hot path over
one batch of synthetic frames per GPU: radius-graph build from coordinates -> forward -> LpLoss
-> backward -> (gradient all-reduce) -> Adam.

One JSON line is printed by rank 0:
  value     frames/s over all GPUs with the step's coordinates already resident in HBM
  e2e       same metric through the public API with HOST (pinned) coordinates: H2D copy of the
            step's window, graph build, train step and a D2H read of the loss inside the timing
  roofline  the dominant kernel.  Default (node-first factorisation): the tcgen05 grouped-GEMM
            kernel with the largest share of the step (HBM-bound: algorithmic bytes per launch /
            CUDA-event duration against the measured copy bandwidth in MEASURED_PEAKS.json;
            all three grouped-GEMM kernels are listed under roofline.kernels).  With
            GKN_FACTORIZATION=edge: the fused edge-first contraction (tensor-bound: 2*E*K*w^2
            FLOP per launch against the measured sustained bf16 peak)
  cpu_baseline  the CPU restatement of the reference (oracle/) timed on this box's host cores
`--impl reference` times that CPU restatement alone (all host threads) with the same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

# per-step edge counts vary: grow allocator segments in place instead of re-cudaMalloc-ing
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

import numpy as np  # noqa: E402
import torch  # noqa: E402

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

LR_NOTE = {"lr": 1e-4}
STDOUT_FD = {"fd": None}  # the real stdout while fd 1 is parked on stderr (multi-GPU runs)


def emit_line(line: dict) -> None:
    text = json.dumps(line) + "\n"
    if STDOUT_FD["fd"] is not None:
        sys.stdout.flush()
        os.write(STDOUT_FD["fd"], text.encode())
    else:
        sys.stdout.write(text)
        sys.stdout.flush()
METRIC = "gkn_train_frames_per_s"
UNIT = "frames/s"
WORKLOAD = "bba_allatom"


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d.get("hbm_gbs"), "bf16_tflops": d.get("bf16_tflops"),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained"), "source": "measured"}
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines: list[str] = []
        self.marks: list[tuple[str, int]] = []   # (timed loop, index of its first sample)
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu_index), "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self, name: str):
        """Samples from here on belong to the timed loop `name` (per-loop medians in the JSON line:
        a loop that ran through a power-cap dip shows up as such)."""
        self.marks.append((name, len(self.lines)))

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, watts, idx = [], [], set(), [], []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, line in enumerate(self.lines):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            idx.append(i)
            try:
                watts.append(float(parts[6]))
            except (IndexError, ValueError):
                watts.append(float("nan"))
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        out = {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm), "sm_max_mhz": max(smax),
               "reasons": sorted(reasons), "samples": len(sm)}
        by_loop = {}
        for k, (name, start) in enumerate(self.marks):
            end = self.marks[k + 1][1] if k + 1 < len(self.marks) else len(self.lines)
            sel = [j for j, i in enumerate(idx) if start <= i < end]
            if sel:
                w = [watts[j] for j in sel if watts[j] == watts[j]]
                by_loop[name] = {"sm_mhz": statistics.median(sm[j] for j in sel),
                                 "sm_min_mhz": min(sm[j] for j in sel),
                                 "power_w": statistics.median(w) if w else None, "samples": len(sel)}
        if by_loop:
            out["by_loop"] = by_loop
        return out


def dist_info():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


# ------------------------------------------------------------------------------------------------
# CPU reference (oracle) training step — used by cpu_baseline and --impl reference ONLY
# ------------------------------------------------------------------------------------------------
class CpuReference:
    """fp32 CPU restatement of the reference's training step on this path: SciPy radius graph
    (graph_kernel.py:362-393), KernelNN forward with the materialised [E,w,w] kernel
    (:194-209, :277-309), LpLoss.rel (:105-119), backward, Adam (:541)."""

    def __init__(self, wl, threads: int, seed: int = 0, lr: float = 1e-4):
        from oracle import gkn_oracle as orc
        from molecular_dynamics_neural_operator_b200.graph_kernel import KernelNN

        self.orc = orc
        self.wl = wl
        torch.set_num_threads(threads)
        self.threads = threads
        torch.manual_seed(seed)
        # the module is used here only as a parameter container with the reference's init
        model = KernelNN(wl.width, wl.ker_width, wl.depth, 6, in_width=7, out_width=3,
                         window_size=wl.window, num_residues=wl.num_atoms, precision="fp32")
        self.sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
        for k in list(self.sd):
            if k.startswith("conv2.net."):
                self.sd[k] = self.sd["conv1.net." + k[len("conv2.net."):]]
        uniq = {id(v): v for v in self.sd.values()}
        self.opt = torch.optim.Adam(list(uniq.values()), lr=lr, weight_decay=5e-4)

    def calibrate_gflops(self) -> float:
        a = torch.randn(2048, 1024)
        b = torch.randn(1024, 2048)
        torch.mm(a, b)
        t0 = time.perf_counter()
        for _ in range(3):
            torch.mm(a, b)
        dt = (time.perf_counter() - t0) / 3
        return 2 * 2048 * 1024 * 2048 / dt / 1e9

    def pick_atoms(self, target_seconds: float, edges_per_frame: float) -> int:
        """Number of atoms of a frame (a slab along x, same density) whose training step should
        take about `target_seconds` on this host."""
        wl = self.wl
        gflops = self.calibrate_gflops()
        A = 2 * wl.depth
        flop_frame = A * 3 * 2.0 * edges_per_frame * wl.ker_width * wl.width ** 2 * 1.4
        sec_frame = flop_frame / (gflops * 1e9)
        # cap the sample at a quarter frame: the reference formulation keeps several [E,w,w]
        # fp32 tensors per application alive for backward (~35 GB for a full frame)
        frac = min(0.25, target_seconds / sec_frame)
        return max(32, int(wl.num_atoms * frac))

    def step(self, x_window: np.ndarray, y: np.ndarray, aa: np.ndarray, n_atoms: int):
        """One training step on the sub-frame made of the n_atoms atoms with the smallest x.
        Returns (seconds, edges)."""
        from types import SimpleNamespace

        wl, orc = self.wl, self.orc
        t0 = time.perf_counter()
        keep = np.sort(np.argsort(x_window[0, 0, :, 0], kind="stable")[:n_atoms])
        xw = x_window[:, :, keep]
        B, W = xw.shape[0], xw.shape[1]
        ei, ea = orc.radius_graph_batched(xw[:, 0], wl.cutoff)
        data = SimpleNamespace(
            x_aminoacid=torch.from_numpy(np.tile(aa[keep], B)),
            x_position=torch.from_numpy(xw.reshape(B * W, n_atoms, 3)),
            edge_index=torch.from_numpy(ei), edge_attr=torch.from_numpy(ea))
        yt = torch.from_numpy(y[:, keep].reshape(B, -1))
        self.opt.zero_grad(set_to_none=True)
        out = orc.kernelnn_forward(self.sd, data, wl.depth, wl.window, n_atoms)
        loss = orc.lp_loss_rel(out.view(B, -1), yt, size_average=False)
        loss.backward()
        self.opt.step()
        return time.perf_counter() - t0, ei.shape[1]


def run_reference_arm(args, wl, edges_per_frame):
    """`--impl reference`: the CPU restatement of the reference, all host threads, bounded sample
    per step.  Rank 0 only."""
    from molecular_dynamics_neural_operator_b200 import synthetic as syn

    world, rank, _ = dist_info()
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ref = CpuReference(wl, threads, lr=args.lr)
    n_atoms = ref.pick_atoms(args.cpu_seconds_per_step, edges_per_frame)
    traj = syn.trajectory(wl, wl.window + 2 + args.steps + args.warmup, seed=0)
    aa = syn.residue_ids(wl)
    times, edges = [], []
    for i in range(args.warmup + args.steps):
        xw, y = syn.window_batch(traj, [i], wl.window)
        dt, e = ref.step(xw, y, aa, n_atoms)
        if i >= args.warmup:
            times.append(dt)
            edges.append(e)
    total = sum(times)
    frames = sum(edges) / edges_per_frame  # frame-equivalents processed (cost is linear in E)
    value = frames / total
    sample = (f"{n_atoms}/{wl.num_atoms} atoms of one frame per step (x-slab, same density), "
              f"{int(np.mean(edges))} edges vs {int(edges_per_frame)} per full frame; "
              f"frames/s = edge-fraction / step time")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": bench_config(wl, args.frames_per_gpu, edges_per_frame),
        "arm_details": arm_details(True),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bench_config(wl, frames_per_gpu, edges_nominal):
    """The workload both arms are quoted on - the SAME dictionary in the B200 line and in the
    `--impl reference` line (`edges_nominal`: edges of frame 0 of the seeded trajectory, which the
    GPU kernel and SciPy agree on bit for bit).  What differs between the arms is under
    `arm_details`."""
    return {
        "workload": (f"{wl.name}: {wl.num_atoms}-atom frames, {wl.cutoff} A radius graph "
                     f"(~{int(edges_nominal)} directed edges/frame incl. self loops), KernelNN "
                     f"width {wl.width} kernel_width {wl.ker_width} depth {wl.depth} "
                     f"({2 * wl.depth} NNConv applications), window {wl.window}; step = graph "
                     f"build + fwd + LpLoss + bwd + Adam(lr={LR_NOTE['lr']}, weight_decay=5e-4)"),
        "frames_per_gpu_per_step": frames_per_gpu,
        "batch_note": "reference default batch_size is 128 (graph_kernel.py:318) = 16 frames per GPU "
                      "on 8 GPUs; per-GPU work is fixed as N grows (weak scaling)",
        "lr_note": "the reference's default lr 0.01 (graph_kernel.py:319) sends THIS random-init "
                   "model on THESE synthetic coordinates to inf after the first update (loss 1.1e8 "
                   "-> inf -> NaN: 12 tied applications amplify every weight change), lr 1e-3 kills "
                   "every ReLU; at 1e-4 the loss falls steadily (1.1e8 -> 29 in 60 steps), so the "
                   "timed steps carry finite activations and gradients.  Throughput does not depend "
                   "on lr",
        "l2_policy": "per-step working set (h_e alone ~0.12 GB per frame) exceeds the 126 MB L2; "
                     "each step uses a different batch of frames",
    }


def arm_details(reference: bool):
    if reference:
        return {
            "formulation": "the reference's own: SciPy distance_matrix graph, per-edge [E,w,w] kernel "
                           "materialised by the edge MLP every application, fp32, PyTorch CPU "
                           "(oracle/gkn_oracle.py restating graph_kernel.py:194-209, :277-309, :362-393)",
            "frames_per_step": "a bounded sub-frame per step (see cpu_baseline.sample); value is "
                               "scaled to whole frames by edge count",
            "parallelism": "one process, all host threads (torch intra-op)",
        }
    from molecular_dynamics_neural_operator_b200 import functional as GF

    return {
        "factorization": GF.FACTORIZATION,
        "node_forward": ("fused producer/consumer kernel" if GF.NODE_FUSED else
                         "own tcgen05 GEMM for U + grouped message kernel" if GF.NODE_U_GEMM else
                         "library GEMM for U + grouped message kernel (gkn_node_msg_bf16)"),
        "node_backward": ("fused dU producer / (dW3, dx) consumer kernel (gkn_node_fused_bwd_bf16)"
                          if GF.NODE_FUSED_BWD else "gkn_node_du_mn_bf16 + two library GEMMs"),
        "untimed_steps_before_timing": "max(8, warmup) resident + max(3, warmup) host-fed + 2 resident; "
                                       "Python cyclic GC disabled inside the timed loops",
        "parallelism": "frames sharded over GPUs, bucketed NCCL gradient all-reduce",
    }


# ------------------------------------------------------------------------------------------------
# extra measurements carried by the N = 1 line: HBM-bound kernels in isolation, the other configs
# ------------------------------------------------------------------------------------------------
def micro_rooflines(peaks):
    """K0 (radius graph) and K2 (segmented node update / mean) on batches of frames, each launch
    timed alone with CUDA events (scripts/bench_kernels.py): algorithmic bytes of SURVEY.md 8(d)
    (B_g = 12 N + 40 E per graph build, B_s = 4 E w + 4 E + 4 N w + 4 N per reduction) / time
    against the measured copy bandwidth."""
    sys.path.insert(0, str(ROOT / "scripts"))
    import bench_kernels as bk

    bk.QUIET = True
    out = {}
    for name, frames in (("bba_allatom", 64), ("synthetic_10k", 16)):
        bk.RESULTS.clear()
        ei, ea, N, E = bk.bench_graph(name, frames)
        bk.bench_node_update(name, frames, ei, N, E)
        del ei, ea
        for r in bk.RESULTS:
            k = r["kernel"]
            if k.startswith("radius_graph device time"):
                key, extra = "radius_graph", {"count_pass_ms": r["count_ms"], "fill_pass_ms": r["fill_ms"]}
            elif k.startswith("segment_matvec fwd"):
                key, extra = "segment_matvec_fwd", {}
            elif k.startswith("segment_matvec bwd"):
                key, extra = "segment_matvec_bwd", {}
            elif k.startswith("scatter_mean_fwd"):
                key, extra = "scatter_mean_fwd", {}
            else:
                continue
            out[f"{key}[{name} x{frames}]"] = {
                "bound": "hbm", "avg_launch_ms": r["ms"], "bytes_per_launch": r["algorithmic_bytes"],
                "achieved_gbs": r["gbs"], "frac": r["gbs"] / peaks["hbm_gbs"], "edges": r["edges"],
                **extra}
    torch.cuda.empty_cache()
    return out


def quick_train(dev, name, frames, steps, warm=3):
    """A few training steps of another BASELINE config on one GPU (coordinates resident, graph
    built every step): frames/s."""
    from molecular_dynamics_neural_operator_b200 import functional as GF
    from molecular_dynamics_neural_operator_b200 import synthetic as syn
    from molecular_dynamics_neural_operator_b200.graph_kernel import KernelNN
    from molecular_dynamics_neural_operator_b200.train import Trainer

    wl = syn.WORKLOADS[name]
    torch.manual_seed(0)
    model = KernelNN(wl.width, wl.ker_width, wl.depth, 6, in_width=7, out_width=3,
                     window_size=wl.window, num_residues=wl.num_atoms, precision="bf16").to(dev)
    trainer = Trainer(model, lr=LR_NOTE["lr"], weight_decay=5e-4)
    traj = syn.trajectory(wl, wl.window + 1 + frames, seed=0)
    aa = torch.from_numpy(syn.residue_ids(wl)).to(dev)
    xw, y = syn.window_batch(traj, list(range(frames)), wl.window)
    x_d, y_d = torch.from_numpy(xw).to(dev), torch.from_numpy(y).to(dev)
    edges = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(warm + steps):
        if i == warm:
            torch.cuda.synchronize()
            ev0.record()
        batch = syn.make_batch(x_d, aa, y_d, wl.cutoff, prepare=True, ker_width=wl.ker_width)
        edges = batch.edge_index.size(1)
        loss = trainer.step(batch, frames)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    node = GF.FACTORIZATION == "node" and GF.node_path_supported(
        wl.width, wl.ker_width, edges, "bf16", True, frames * wl.num_atoms)
    res = {"workload": name, "frames_per_step": frames, "steps": steps, "ms_per_step": ms,
           "frames_per_s": frames / (ms * 1e-3), "edges_per_frame": edges / frames,
           "edges_per_s": edges / (ms * 1e-3), "width": wl.width, "depth": wl.depth,
           "factorization": "node" if node else "edge", "loss_finite": bool(torch.isfinite(loss)),
           "mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
    del trainer, model, batch
    torch.cuda.empty_cache()
    return res


def _tame_rollout_model(model):
    """A random-init KernelNN amplifies its input ~30x per application on raw Angstrom coordinates
    (16 applications at depth 8: inf).  For the rollout measurements the last edge-MLP layer is
    scaled by 0.02 so that every application maps O(1) features to O(1) features: same kernels,
    same work, finite numbers - which the sharded-vs-unsharded agreement check needs."""
    with torch.no_grad():
        model.conv1.net.layers[4].weight.mul_(0.02)
        model.conv1.net.layers[4].bias.mul_(0.02)


def quick_rollout(dev, name, steps, window=2):
    """BASELINE config 5 on ONE GPU: device-resident operator rollout (forward -> new positions ->
    radius graph rebuilt on the GPU), reference graph_kernel.py:396-413: rollout steps/s."""
    from molecular_dynamics_neural_operator_b200 import synthetic as syn
    from molecular_dynamics_neural_operator_b200.graph_kernel import KernelNN, construct_pairdata

    wl = syn.WORKLOADS[name]
    torch.manual_seed(0)
    model = KernelNN(wl.width, wl.ker_width, wl.depth, 6, in_width=7, out_width=3, window_size=window,
                     num_residues=wl.num_atoms, precision="bf16").to(dev).eval()
    _tame_rollout_model(model)
    traj = torch.from_numpy(syn.trajectory(wl, window, seed=0)).to(dev)
    aa = torch.from_numpy(syn.residue_ids(wl)).to(dev)
    data = construct_pairdata(traj, aa, threshold=wl.cutoff, device=dev)
    edges = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for step in range(steps + 1):
            if step == 1:
                torch.cuda.synchronize()
                ev0.record()
            out = model(data, single_example=True)
            # random-init weights: bound the predicted displacement so the rollout stays physical
            pos = data.x_position[-1] + 0.05 * torch.tanh(torch.nan_to_num(out))
            data = construct_pairdata(torch.cat([data.x_position[1:], pos[None]], 0), aa,
                                      threshold=wl.cutoff, device=dev)
            edges.append(int(data.edge_index.size(1)))
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    res = {"workload": name, "mode": "rollout (inference), one GPU", "steps": steps,
           "ms_per_rollout_step": ms, "steps_per_s": 1e3 / ms, "atoms": wl.num_atoms,
           "edges_per_step": edges[-1], "edges_per_s": edges[-1] / (ms * 1e-3), "width": wl.width,
           "depth": wl.depth, "mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
    del model, data
    torch.cuda.empty_cache()
    return res

def sharded_rollout(dev, world, rank, steps=3, window=2, name="synthetic_100k"):
    """BASELINE config 5 with ONE 100k-atom frame sharded over the N ranks of this run
    (sharded.kernelnn_forward_sharded: source-range shards, one NCCL all-reduce of partial node
    sums per application).  The first step's output is checked against the unsharded forward that
    every rank also computes (same kernels, whole frame); then `steps` rollout steps are timed."""
    import torch.distributed as dist

    from molecular_dynamics_neural_operator_b200 import synthetic as syn
    from molecular_dynamics_neural_operator_b200.graph_kernel import KernelNN, construct_pairdata
    from molecular_dynamics_neural_operator_b200.sharded import kernelnn_forward_sharded

    wl = syn.WORKLOADS[name]
    torch.manual_seed(0)
    model = KernelNN(wl.width, wl.ker_width, wl.depth, 6, in_width=7, out_width=3, window_size=window,
                     num_residues=wl.num_atoms, precision="bf16").to(dev).eval()
    _tame_rollout_model(model)
    for p_ in model.parameters():
        dist.broadcast(p_.data, src=0)
    traj = torch.from_numpy(syn.trajectory(wl, window, seed=0)).to(dev)
    aa = torch.from_numpy(syn.residue_ids(wl)).to(dev)
    data = construct_pairdata(traj, aa, threshold=wl.cutoff, device=dev)
    edges = []
    with torch.no_grad():
        ref = model(data, single_example=True)
        out = kernelnn_forward_sharded(model, data)
        diff = float((out - ref).norm() / ref.norm().clamp(min=1e-30))
        del ref
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for step in range(steps + 1):
            if step == 1:
                dist.barrier()
                torch.cuda.synchronize()
                ev0.record()
            out = kernelnn_forward_sharded(model, data)
            pos = data.x_position[-1] + 0.05 * torch.tanh(torch.nan_to_num(out))
            data = construct_pairdata(torch.cat([data.x_position[1:], pos[None]], 0), aa,
                                      threshold=wl.cutoff, device=dev)
            edges.append(int(data.edge_index.size(1)))
        ev1.record()
        dist.barrier()
        torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1) / steps, diff], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, diff = float(t[0]), float(t[1])
    from molecular_dynamics_neural_operator_b200.sharded import exchange_in_use

    res = {"workload": name, "mode": f"rollout, one frame sharded over {world} GPUs by source-node "
                                     f"ranges; exchange per application: {exchange_in_use()}",
           "steps": steps, "ms_per_rollout_step": ms, "steps_per_s": 1e3 / ms, "atoms": wl.num_atoms,
           "edges_per_step": edges[-1], "rel_fro_vs_unsharded_forward": diff,
           "agrees_with_unsharded": diff < 2.0 ** -7}
    del model, data
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200_arm(args, wl):
    import torch.distributed as dist

    from molecular_dynamics_neural_operator_b200 import _lib, ops
    from molecular_dynamics_neural_operator_b200 import synthetic as syn
    from molecular_dynamics_neural_operator_b200.graph_kernel import KernelNN
    from molecular_dynamics_neural_operator_b200.train import Trainer, broadcast_parameters

    world, rank, local = dist_info()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback); "
                           "use --impl reference for the CPU baseline")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries the single JSON line; NCCL's own log (communicator ranks, NVLS, rings)
        # stays ON and goes to stderr.  Level VERSION prints its banner with a bare printf to
        # stdout, so it is raised to INFO (which goes through NCCL's logger and NCCL_DEBUG_FILE);
        # and, belt and braces, file descriptor 1 points at stderr until the JSON line is written.
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,ENV")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        STDOUT_FD["fd"] = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    B = args.frames_per_gpu
    # every distinct batch (distinct edge count => distinct buffer sizes) must have been seen
    # during warm-up, otherwise the caching allocator grows inside the timed region
    pool = max(1, min(args.pool, args.warmup))
    torch.manual_seed(0)
    model = KernelNN(wl.width, wl.ker_width, wl.depth, 6, in_width=7, out_width=3,
                     window_size=wl.window, num_residues=wl.num_atoms, precision="bf16").to(dev)
    broadcast_parameters(model)
    trainer = Trainer(model, lr=args.lr, weight_decay=5e-4)

    # synthetic trajectory; rank r takes frames r, r+world, ... of every global batch
    n_frames = wl.window + 1 + pool * B * world
    traj = syn.trajectory(wl, n_frames, seed=0)
    aa = torch.from_numpy(syn.residue_ids(wl)).to(dev)
    host_x, host_y = [], []
    for b in range(pool):
        starts = [b * B * world + rank + world * j for j in range(B)]
        xw, y = syn.window_batch(traj, starts, wl.window)
        host_x.append(torch.from_numpy(xw).pin_memory())
        host_y.append(torch.from_numpy(y).pin_memory())
    dev_x = [t.to(dev) for t in host_x]
    dev_y = [t.to(dev) for t in host_y]
    h2d_bytes = host_x[0].numel() * 4 + host_y[0].numel() * 4

    edge_counts = []
    prefetch = syn.BatchPrefetcher(dev)
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    loss_events: list = []
    loss_log: list = []  # every step's loss as read back by the host (rank-local sum over frames)

    def build_resident(i):
        return syn.make_batch(dev_x[i % pool], aa, dev_y[i % pool], wl.cutoff, prepare=True,
                              ker_width=wl.ker_width)

    def build_e2e(i):
        # H2D copy of this step's coordinates from pinned host memory, then the graph build
        x = host_x[i % pool].to(dev, non_blocking=True)
        y = host_y[i % pool].to(dev, non_blocking=True)
        return syn.make_batch(x, aa, y, wl.cutoff, prepare=True, ker_width=wl.ker_width)

    def run_loop(build, steps, read_loss):
        """`steps` training steps; the next batch is built on a side stream while the current
        step runs.  read_loss: copy every step's loss to the host (read one step late so the
        host never drains the training stream).  Both timed loops read the loss: it also paces the
        host to one step ahead of the device, which measured steadier than a free-running host
        (intermittent +10 % step times when the host queues two steps ahead)."""
        trace = os.environ.get("GKN_BENCH_TRACE") == "1"
        prefetch.submit(lambda: build(0))
        done: list = []
        for i in range(steps):
            t0 = time.perf_counter()
            if len(done) >= 2:
                done.pop(0).synchronize()  # at most two steps in flight (bounded run-ahead)
            batch = prefetch.get()
            t1 = time.perf_counter()
            if i + 1 < steps:
                prefetch.submit(lambda: build(i + 1))
            t2 = time.perf_counter()
            edge_counts.append(batch.edge_index.size(1))
            loss = trainer.step(batch, B)
            if trace:  # (diagnostic) host time: wait for step i-2, batch build, step enqueue
                print(f"trace step {i}: wait {1e3 * (t1 - t0):.1f} build {1e3 * (t2 - t1):.1f} "
                      f"enqueue {1e3 * (time.perf_counter() - t2):.1f} ms", file=sys.stderr)
            ev_done = torch.cuda.Event()
            ev_done.record()
            done.append(ev_done)
            if read_loss:
                if loss_events:
                    loss_events.pop().synchronize()
                    loss_log.append(float(loss_host[0]))  # D2H read of the previous step's result
                loss_host.copy_(loss.reshape(1), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                loss_events.append(ev)
        if read_loss and loss_events:
            loss_events.pop().synchronize()
            loss_log.append(float(loss_host[0]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(build, steps, read_loss):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        run_loop(build, steps, read_loss)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # the clock sampler (an nvidia-smi process) starts BEFORE warm-up: its start-up talks to the
    # driver and was seen to disturb the first timed steps; in steady state it samples every 200 ms
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    # A cold process on a fresh box needs more than 3 steps to settle (lazy CUDA module loading,
    # cuBLAS handles/heuristics, allocator growth, page cache): measured 107 frames/s with 3
    # warm-up steps vs 124 with 10 in the FIRST process on a box.  Extra untimed settle steps are
    # added in front of the W warm-up steps so that at least 8 steps precede the timed region.
    settle = max(0, 8 - args.warmup)
    # Host pauses are the remaining source of run-to-run jitter (the device queue holds less than
    # one step of work, so a 30 ms pause of the enqueueing thread shows up as idle GPU time):
    # Python's cyclic collector is therefore switched off for the timed loops (nothing here creates
    # reference cycles per step; a full collection runs first).
    import gc
    run_loop(build_resident, settle + args.warmup, True)
    # the host-fed loop allocates its H2D staging tensors on the prefetch stream: it needs its own
    # warm-up for the caching allocator to reach a steady state (one step was not enough: +-15 %
    # between otherwise identical runs)
    run_loop(build_e2e, max(3, args.warmup), True)
    run_loop(build_resident, 2, True)  # back to the resident pattern before it is timed
    gc.collect()
    gc.disable()
    if sampler:
        sampler.lines.clear()  # keep only samples taken during the timed loops
    edge_counts.clear()
    launches0 = _lib.launch_count()
    loss_before = list(loss_log)
    loss_log.clear()
    if sampler:
        sampler.mark("value")
    ms_total = timed(build_resident, args.steps, True)
    launches = _lib.launch_count() - launches0
    loss_timed = list(loss_log)
    if sampler:
        sampler.mark("e2e")
    ms_e2e = timed(build_e2e, args.steps, True)
    # Per-launch CUDA events around the contraction cost ~0.15 ms each on this driver (measured:
    # +10 % step time), so the kernel's launch duration is taken from an instrumented REPEAT of
    # the same timed loop, and `value` from the clean one above.
    timer = {"count": 0, "ms": 0.0}
    ms_instr = None
    if not args.no_kernel_timer:
        ops.KERNEL_TIMER.enable("contract")
        if sampler:
            sampler.mark("instrumented")
        ms_instr = timed(build_resident, args.steps, True)
        timer = ops.KERNEL_TIMER.collect()
        ops.KERNEL_TIMER.disable()
    clocks = sampler.stop() if sampler else None
    gc.enable()
    # The literal north-star kernel (edge-first fused tcgen05 contraction, the path that carries the
    # tensor-core-utilisation target) measured in the same process on the same batches: a few
    # steps with GKN_FACTORIZATION=edge semantics, its launches timed like the ones above.
    edge_first = None
    from molecular_dynamics_neural_operator_b200 import functional as GF
    if world == 1 and GF.FACTORIZATION == "node" and not args.no_edge_first:
        GF.FACTORIZATION = "edge"
        try:
            run_loop(build_resident, 3, True)
            n_edge = max(3, args.steps // 3)
            ms_edge = timed(build_resident, n_edge, True)
            ops.KERNEL_TIMER.enable("contract")
            timed(build_resident, 2, True)
            t_edge = ops.KERNEL_TIMER.collect()
            ops.KERNEL_TIMER.disable()
        finally:
            GF.FACTORIZATION = "node"
        pk = measured_peaks()
        flop_per_launch = 2.0 * (float(np.mean(edge_counts)) ) * wl.ker_width * wl.width ** 2
        avg_ms = t_edge["ms"] / max(t_edge["count"], 1)
        ach = flop_per_launch / (avg_ms * 1e-3) / 1e12
        edge_first = {
            "value": B * n_edge / (ms_edge * 1e-3), "unit": UNIT, "ms_per_step": ms_edge / n_edge,
            "steps": n_edge,
            "roofline": {"bound": "tensor", "kernel": "contract_tc_kernel (fwd+dx)",
                         "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": ach / pk["bf16_tflops_sustained"],
                         "frac_burst": ach / pk["bf16_tflops"], "launches": t_edge["count"],
                         "avg_launch_ms": avg_ms, "flop_per_launch": flop_per_launch,
                         "peak_source": f"{pk['source']} sustained bf16"},
            "note": "same model, batches and step with the edge-first contraction "
                    "(2*E*K*w^2 FLOP per launch); the default node-first path computes the same "
                    "trilinear form with ~40x fewer FLOPs and is HBM-bound",
        }

    with torch.no_grad():
        ei0, _ = ops.radius_graph(torch.from_numpy(syn.trajectory(wl, 1, seed=0)).to(dev), wl.cutoff)
    edges_nominal = int(ei0.size(1))
    del ei0
    frames_total = B * world * args.steps
    value = frames_total / (ms_total * 1e-3)
    e2e_value = frames_total / (ms_e2e * 1e-3)
    edges_per_frame = float(np.mean(edge_counts)) / B
    peaks = measured_peaks()

    # dominant kernel.  Node-first factorisation (default): the grouped-GEMM kernels stream the
    # edge hidden state once per launch and are HBM-bound; the one with the largest share of the
    # step is reported.  Edge-first (GKN_FACTORIZATION=edge): the fused tcgen05 contraction.
    roofline = None
    E_b, N_b, K_, w_, A_ = edges_per_frame * B, wl.num_atoms * B, wl.ker_width, wl.width, 2 * wl.depth
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    node_bytes = {  # algorithmic bytes per launch
        "node_msg": 2.0 * E_b * K_ + 2.0 * N_b * w_ * K_ + 4.0 * E_b * w_,
        "node_du": 2.0 * E_b * K_ + 2.0 * E_b * w_ + 2.0 * N_b * K_ * w_,
        # (+ one bit per element of h when the ReLU backward of layers.2 rides in its epilogue)
        "node_dh": 2.0 * A_ * E_b * max(w_, 64) + 2.0 * A_ * N_b * K_ * w_ + 2.0 * E_b * K_
                   + (E_b * K_ / 8.0 if GF.DH_FUSED_MASK and GF.NODE_MN and K_ <= 1024 else 0.0),
        # fused kernels: U / dU cross the L2 only.  fwd: h in, messages out, U written once for the
        # dh pass; bwd: h and the message gradients in, per-CTA dx partials out
        "node_fused_fwd": 2.0 * E_b * K_ + 4.0 * E_b * w_ + 2.0 * N_b * w_ * K_,
        "node_fused_bwd": 2.0 * E_b * K_ + 2.0 * E_b * w_ + 4.0 * sms * N_b * w_,
    }
    node_t = {k: timer[k] for k in node_bytes if k in timer and timer[k]["count"]}
    if node_t:
        kernels = {}
        for k, t in node_t.items():
            avg_ms = t["ms"] / t["count"]
            kernels[k] = {"launches": t["count"], "avg_launch_ms": avg_ms,
                          "bytes_per_launch": node_bytes[k],
                          "achieved_gbs": node_bytes[k] / (avg_ms * 1e-3) / 1e9,
                          "share_of_step": t["ms"] / ms_total}
        top = max(node_t, key=lambda k: node_t[k]["ms"])
        peak = peaks["hbm_gbs"]
        traffic = None
        tp = ROOT / "profiles" / "node_traffic.json"
        if tp.exists():
            traffic = json.loads(tp.read_text()).get(top, {}).get("dram_bytes_per_launch")
        roofline = {"bound": "hbm", "achieved": kernels[top]["achieved_gbs"], "peak": peak,
                    "unit": "GB/s", "frac": kernels[top]["achieved_gbs"] / peak, "traffic": traffic,
                    "kernel": f"{'node_fused' if 'fused' in top else 'node_gemm'}_kernel ({top})", "launches": kernels[top]["launches"],
                    "avg_launch_ms": kernels[top]["avg_launch_ms"],
                    "bytes_per_launch": node_bytes[top],
                    "share_of_step": kernels[top]["share_of_step"], "kernels": kernels,
                    "measured_in": "instrumented repeat of the timed loop (per-launch events "
                                   f"inflate the step: {ms_instr / args.steps:.1f} ms vs "
                                   f"{ms_total / args.steps:.1f} ms clean)",
                    "peak_source": f"{peaks['source']} HBM copy bandwidth"}
        if top == "node_fused_bwd":
            roofline["note"] = (
                "one launch per application does the work of three kernels of round 1 "
                "(gkn_node_du_mn_bf16 + the library GEMMs dx = dU W3^T and dW3 += x^T dU: 0.94 ms, "
                "5.4 GB of DRAM traffic, node_du alone at 83 % of the copy bandwidth by its own byte "
                "model); bytes_per_launch counts h and the message gradients in and the per-CTA dx "
                "partials out - dU never reaches HBM.  Fewer bytes in less time, at a lower "
                "fraction of the bandwidth: the kernel is latency-bound (ncu: DRAM 52 %, L2 39 %)")
    elif timer["count"]:
        flop_per_launch = 2.0 * (edges_per_frame * B) * wl.ker_width * wl.width ** 2
        avg_ms = timer["ms"] / timer["count"]
        achieved = flop_per_launch / (avg_ms * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        traffic = None
        tp = ROOT / "profiles" / "contract_traffic.json"
        if tp.exists():
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": traffic,
                    # context for frac > 1: the sustained cuBLAS figure was taken at ~1.35 GHz under
                    # the power cap; inside this step the contraction runs at a higher clock
                    "peak_burst": peaks["bf16_tflops"], "frac_burst": achieved / peaks["bf16_tflops"],
                    "kernel": "contract_tc_kernel (fwd+dx)", "launches": timer["count"],
                    "avg_launch_ms": avg_ms, "flop_per_launch": flop_per_launch,
                    "share_of_step": timer["ms"] / ms_total,
                    "measured_in": "instrumented repeat of the timed loop (per-launch events "
                                   f"inflate the step: {ms_instr / args.steps:.1f} ms vs "
                                   f"{ms_total / args.steps:.1f} ms clean)",
                    "peak_source": f"{peaks['source']} sustained bf16 (kernel timed inside a long step)"}

    if roofline and roofline["bound"] == "tensor" and rank == 0:
        # the library GEMM on THIS box, in THIS process, right after the timed loops: what the
        # `peak` denominator (a cuBLAS bf16 matmul) reads here and now
        a_ = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
        b_ = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            torch.matmul(a_, b_)
        e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0_.record()
        for _ in range(40):
            torch.matmul(a_, b_)
        e1_.record()
        torch.cuda.synchronize()
        roofline["cublas_bf16_tflops_same_process"] = 40 * 2.0 * 8192 ** 3 / (e0_.elapsed_time(e1_) * 1e-3) / 1e12
        del a_, b_
    if edge_first and clocks and clocks.get("sm_mhz"):
        sms_ = torch.cuda.get_device_properties(dev).multi_processor_count
        hw_ = sms_ * 8192 * clocks["sm_mhz"] * 1e6 / 1e12  # 4096 bf16 MAC/clk/SM
        edge_first["roofline"]["hw_peak_at_sampled_clock"] = hw_
        edge_first["roofline"]["frac_hw_at_sampled_clock"] = edge_first["roofline"]["achieved"] / hw_
        edge_first["roofline"]["frac_nominal_2250"] = edge_first["roofline"]["achieved"] / 2250.0
    if roofline and roofline["bound"] == "tensor" and clocks and clocks.get("sm_mhz"):
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        hw_peak = sms * 8192 * clocks["sm_mhz"] * 1e6 / 1e12  # 4096 bf16 MAC/clk/SM
        roofline["hw_peak_at_sampled_clock"] = hw_peak
        roofline["frac_hw_at_sampled_clock"] = roofline["achieved"] / hw_peak
    # Strong scaling beside the weak-scaling headline: a fixed GLOBAL batch
    # of 16 frames split evenly over the ranks (16 / N frames per GPU per step).
    strong = None
    if not args.no_strong and 16 % world == 0:
        Bs = 16 // world
        dev_xs = [t[:Bs].contiguous() for t in dev_x]
        dev_ys = [t[:Bs].contiguous() for t in dev_y]

        def build_strong(i):
            return syn.make_batch(dev_xs[i % pool], aa, dev_ys[i % pool], wl.cutoff, prepare=True,
                                  ker_width=wl.ker_width)

        def run_strong(steps):
            prefetch.submit(lambda: build_strong(0))
            for i in range(steps):
                batch = prefetch.get()
                if i + 1 < steps:
                    prefetch.submit(lambda: build_strong(i + 1))
                trainer.step(batch, Bs)

        run_strong(max(4, pool))
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_strong = max(5, args.steps // 2)
        ev0.record()
        run_strong(n_strong)
        ev1.record()
        barrier()
        ms_s = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms_s], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_s = float(t.item())
        strong = {"scaling": "strong", "global_batch": 16, "frames_per_gpu_per_step": Bs,
                  "steps": n_strong, "ms_per_step": ms_s / n_strong, "value": 16 * n_strong / (ms_s * 1e-3),
                  "unit": UNIT}
    sharded = None
    if world > 1 and not args.no_extras:
        # N3 evidence in the driver's own multi-GPU runs: config 5 sharded over these N GPUs
        del trainer, model
        dev_x.clear()
        dev_y.clear()
        torch.cuda.empty_cache()
        try:
            sharded = sharded_rollout(dev, world, rank)
        except Exception as exc:  # noqa: BLE001
            sharded = {"error": f"{type(exc).__name__}: {exc}"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    A = 2 * wl.depth
    unit_flop = 2.0 * edges_per_frame * wl.ker_width * wl.width ** 2
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": bench_config(wl, B, edges_nominal),
        "arm_details": arm_details(False),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "loss": {"first_untimed_step": loss_before[0] if loss_before else None,
                 "timed_first": loss_timed[0] if loss_timed else None,
                 "timed_last": loss_timed[-1] if loss_timed else None,
                 "all_finite": bool(np.isfinite(loss_before + loss_timed).all()),
                 "note": "rank-0 LpLoss.rel summed over its frames, every step read back by the host"},
        "clocks": clocks,
        "roofline": roofline,
        "edge_first": edge_first,
        "strong_scaling": strong,
        "sharded_rollout": sharded,
        "edges_per_s": value * edges_per_frame,
        "flops": {
            "algorithmic_train_flop_per_frame": 4 * A * unit_flop,
            "executed_edge_first_flop_per_frame": (2 * A - 1 + 2) * unit_flop,
            # node-first: per application U, dx, dW3 (2*N*K*w^2 each) + msg, dU (2*E*K*w each); dh once
            "executed_node_first_flop_per_frame":
                A * (3 * 2.0 * wl.num_atoms * wl.ker_width * wl.width ** 2
                     + 2 * 2.0 * edges_per_frame * wl.ker_width * wl.width)
                + 2.0 * edges_per_frame * A * wl.width * wl.ker_width,
            "note": "algorithmic = the reference's formulation (4 contractions of 2*E*K*w^2 per "
                    "application).  The default node-first path contracts W3 with the node features "
                    "first and executes ~40x fewer FLOPs (HBM-bound: see roofline); the edge-first "
                    "path defers dW3/dh to one pass and skips dx of the first application.  Tensor-"
                    "core utilisation must be read from edge_first.roofline / ncu, never from "
                    "algorithmic FLOP / time",
            "algorithmic_tflops": value * 4 * A * unit_flop / 1e12 / world,
        },
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        ref = CpuReference(wl, threads, lr=args.lr)
        n_atoms = ref.pick_atoms(args.cpu_seconds_per_step, edges_per_frame)
        xw, y = syn.window_batch(traj, [0], wl.window)
        aa_np = syn.residue_ids(wl)
        ref.step(xw, y, aa_np, max(32, n_atoms // 4))  # warm-up
        # bounded sample: whole steps on the sub-frame until about 10 s of CPU work are timed
        dt, e, nsteps = 0.0, 0, 0
        while dt < 10.0 and nsteps < 12:
            xw, y = syn.window_batch(traj, [nsteps % B], wl.window)
            dti, ei_ = ref.step(xw, y, aa_np, n_atoms)
            dt += dti
            e += ei_
            nsteps += 1
        line["cpu_baseline"] = {
            "value": (e / edges_per_frame) / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": (f"{nsteps} training steps on {n_atoms}/{wl.num_atoms} atoms of one frame "
                       f"(x-slab, same density): {e // nsteps} edges per step vs "
                       f"{int(edges_per_frame)} per full frame, {dt:.1f} s in total; "
                       f"frames/s = edge-fraction / step time"),
        }
        # the reference's own default is ONE host thread (graph_kernel.py:501): a smaller sample
        ref1 = CpuReference(wl, 1, lr=args.lr)
        n1 = max(32, n_atoms // 3)
        ref1.step(xw, y, aa_np, max(32, n1 // 2))
        dt1, e1_, k1 = 0.0, 0, 0
        while dt1 < 5.0 and k1 < 4:
            dti, ei_ = ref1.step(xw, y, aa_np, n1)
            dt1 += dti
            e1_ += ei_
            k1 += 1
        line["cpu_baseline"]["value_1_thread"] = (e1_ / edges_per_frame) / dt1
        line["cpu_baseline"]["sample_1_thread"] = (f"{k1} steps on {n1}/{wl.num_atoms} atoms, "
                                                   f"{e1_ // k1} edges per step, {dt1:.1f} s")
        torch.set_num_threads(threads)
    if world == 1 and not args.no_extras:
        # the HBM-bound kernels on their own (K0 graph build, K2 segmented reductions) and the
        # other BASELINE configs, measured in this same process after the headline
        del trainer, model
        dev_x.clear()
        dev_y.clear()
        torch.cuda.empty_cache()
        try:
            if line["roofline"] is not None:
                line["roofline"]["kernels"].update(micro_rooflines(peaks))
            line["configs"] = {
                "bba_ca (configs[0] shape: 28 residues, w 32, depth 4, batch 128)":
                    quick_train(dev, "bba_ca", 128, steps=5),
                "synthetic_10k (configs[2]: 10k atoms, 4.2 A, w 64, depth 6, 4 frames/step)":
                    quick_train(dev, "synthetic_10k", 4, steps=3),
                "synthetic_100k (configs[4] on ONE GPU: 100k atoms, w 128, depth 8, rollout)":
                    quick_rollout(dev, "synthetic_100k", steps=3),
            }
        except Exception as exc:  # the headline stands even if an extra fails: say so in the line
            line["extras_error"] = f"{type(exc).__name__}: {exc}"
    emit_line(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-gpu", type=int, default=16,
                    help="frames per GPU per step; 16 = the reference's default batch_size 128 "
                         "(graph_kernel.py:318) spread over the 8 GPUs of one box")
    ap.add_argument("--pool", type=int, default=4, help="distinct batches cycled through")
    ap.add_argument("--lr", type=float, default=1e-4,
                    help="Adam learning rate (see config.optimizer for why not the reference's 0.01)")
    ap.add_argument("--cpu-seconds-per-step", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-timer", action="store_true", help="(diagnostic) no per-launch events")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the isolated K0/K2 rooflines and the other configs (N=1 only)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling steps")
    ap.add_argument("--no-edge-first", action="store_true",
                    help="skip the extra steps that time the edge-first contraction (N=1 only)")
    ap.add_argument("--workload", default=WORKLOAD, help="(diagnostic) another BASELINE.json shape")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: at least 3 warm-up steps

    from molecular_dynamics_neural_operator_b200 import synthetic as syn

    LR_NOTE["lr"] = args.lr
    wl = syn.WORKLOADS[args.workload]
    if args.impl == "reference":
        # edges/frame of the synthetic workload, from the same generator (SciPy on one frame)
        from oracle import gkn_oracle as orc

        ei, _ = orc.radius_graph_batched(syn.trajectory(wl, 1, seed=0), wl.cutoff)
        run_reference_arm(args, wl, float(ei.shape[1]))
    else:
        run_b200_arm(args, wl)


if __name__ == "__main__":
    main()
