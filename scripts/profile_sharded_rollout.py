"""Per-phase device+host time of one sharded node-first rollout step at config 5 shapes, emulating
rank 0 of `world` ranks on one GPU (the all-reduce is skipped; its cost is measured by the real
multi-GPU run).  python scripts/profile_sharded_rollout.py [world]"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from molecular_dynamics_neural_operator_b200 import sharded as sh, synthetic as syn  # noqa: E402
from molecular_dynamics_neural_operator_b200.graph_kernel import KernelNN, construct_pairdata  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
wl = syn.WORKLOADS["synthetic_100k"]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = KernelNN(wl.width, wl.ker_width, wl.depth, 6, in_width=7, out_width=3, window_size=2,
                 num_residues=wl.num_atoms, precision="bf16").to(dev).eval()
traj = torch.from_numpy(syn.trajectory(wl, 2, seed=0)).to(dev)
aa = torch.from_numpy(syn.residue_ids(wl)).to(dev)


def phase(name, fn, acc):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    return out


with torch.no_grad():
    for it in range(3):
        acc = {}
        data = phase("construct_pairdata", lambda: construct_pairdata(traj, aa, threshold=wl.cutoff, device=dev), acc)
        x = phase("node_encoder", lambda: model.node_encoder(data), acc)
        shard = phase("shard_frame_by_source", lambda: sh.shard_frame_by_source(
            data.edge_index, data.edge_attr, x.size(0), 0, world, getattr(data, "_gkn_row_ptr", None),
            data._gkn_plan.inv_cnt), acc)
        setup = phase("edge_mlp_setup", lambda: sh._node_shard_setup(shard, model.conv1.net, x.size(1)), acc)
        hs, w3r, B3, msg = setup
        sched = (0,) * model.depth + (1,) * model.depth
        xx = x.contiguous()
        for c in sched:
            sums = phase("partial_sums", lambda: sh._node_partial_sums(xx, shard, hs, w3r, msg), acc)
            xx = phase("node_update", lambda: sh._node_update(
                xx, sums, shard.inv_cnt, B3, (model.conv1.root, model.conv2.root)[c],
                (model.conv1.bias, model.conv2.bias)[c], True), acc)
        phase("fc2", lambda: model.fc2(xx), acc)
    print({k: round(v, 2) for k, v in acc.items()}, "total", round(sum(acc.values()), 2))
