"""Operator rollout on BASELINE config 5 shapes (100k-atom frame, width 128, depth 8): model
forward -> new positions -> slide window -> radius graph rebuilt on the GPU -> repeat, all device
resident (reference caller: recursive_propagation, graph_kernel.py:396-413).

    python scripts/rollout_bench.py [--atoms 100000] [--steps 4]
"""
import argparse
import json
import os
import sys
from pathlib import Path

# edge counts drift from step to step: let the caching allocator grow segments in place instead
# of cudaMalloc-ing a slightly larger block every step
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

import torch  # noqa: E402

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from molecular_dynamics_neural_operator_b200 import synthetic as syn  # noqa: E402
from molecular_dynamics_neural_operator_b200.graph_kernel import KernelNN, construct_pairdata  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="synthetic_100k")
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--window", type=int, default=2)
ap.add_argument("--sharded", action="store_true", help="shard the frame over the torchrun ranks")
a = ap.parse_args()
wl = syn.WORKLOADS[a.workload]
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if a.sharded:
    import torch.distributed as dist

    from molecular_dynamics_neural_operator_b200.sharded import kernelnn_forward_sharded

    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = KernelNN(wl.width, wl.ker_width, wl.depth, 6, in_width=7, out_width=3, window_size=a.window,
                 num_residues=wl.num_atoms, precision="bf16").to(dev).eval()
traj = torch.from_numpy(syn.trajectory(wl, a.window, seed=0)).to(dev)
aa = torch.from_numpy(syn.residue_ids(wl)).to(dev)
data = construct_pairdata(traj, aa, threshold=wl.cutoff, device=dev)
edges = []
with torch.no_grad():
    for step in range(a.steps + 1):
        if step == 1:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        out = kernelnn_forward_sharded(model, data) if a.sharded else model(data, single_example=True)
        # random-init weights: bound the predicted displacement so the rollout stays physical
        step_pos = data.x_position[-1] + 0.05 * torch.tanh(torch.nan_to_num(out))
        newpos = torch.cat([data.x_position[1:], step_pos[None]], dim=0)
        data = construct_pairdata(newpos, aa, threshold=wl.cutoff, device=dev)
        edges.append(int(data.edge_index.size(1)))
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
A = 2 * wl.depth
flop = A * 2.0 * (sum(edges[1:]) / max(a.steps, 1)) * wl.ker_width * wl.width ** 2
if a.sharded:
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
if rank == 0:
  print(json.dumps({"workload": wl.name, "sharded_over_gpus": world if a.sharded else 1, "atoms": wl.num_atoms, "width": wl.width, "depth": wl.depth,
                  "edges_per_step": edges, "ms_per_rollout_step": ms, "steps_per_s": 1e3 / ms,
                  "contraction_tflops": flop / ms / 1e9,
                  "mem_gb": torch.cuda.max_memory_allocated() / 1e9}))
if a.sharded:
    dist.destroy_process_group()
