"""Device time of the three node-first grouped-GEMM kernels at the bench shape, in isolation
(CUDA events, L2 flushed between launches by the kernels' own working sets of > 1 GB).

    GKN_NODE_HINTS=msg,du,dh python scripts/bench_node_kernels.py [frames]
"""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from molecular_dynamics_neural_operator_b200 import ops, synthetic as syn  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
wl = syn.WORKLOADS["bba_allatom"]
pos = torch.from_numpy(syn.trajectory(wl, B, seed=0)).cuda()
ei, ea, rp = ops.radius_graph(pos, wl.cutoff, return_row_ptr=True)
N, W, K, A = B * wl.num_atoms, wl.width, wl.ker_width, 2 * wl.depth
plan = ops.plan_from_csr(ei, rp, N)
lay = ops.node_layout(plan, dh_chunks=K // 256)
E = ei.size(1)
bf = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
hs = torch.randn(lay.e_pad, K, device="cuda", generator=g).to(bf)
u_all = (torch.randn(A, N, W * K, device="cuda", generator=g) / 8).to(bf)
gs_all = torch.randn(A, lay.e_pad, W, device="cuda", generator=g).to(bf)
gs_all[:, lay.dst_pad < 0] = 0
msg = torch.empty(lay.e_pad, W, device="cuda")
dut = torch.empty(N, K * W, dtype=bf, device="cuda")
dh = torch.empty(lay.e_pad, K, dtype=bf, device="cuda")


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_msg = timed(lambda i=0: ops.node_msg(hs, u_all[i % A].view(N, W, K), lay, 0, lay.num_units, 0, msg), 24)
t_du = timed(lambda i=0: ops.node_du_mn(hs, gs_all[i % A], lay, 0, N, dut), 24)
t_dh = timed(lambda i=0: ops.node_dh_mn(gs_all, u_all.view(A, N, W, K), lay, 0, lay.num_units, 0, dh), 4)
by = {"node_msg": 2.0 * E * K + 2.0 * N * W * K + 4.0 * E * W,
      "node_du": 2.0 * E * K + 2.0 * E * W + 2.0 * N * K * W,
      "node_dh": 2.0 * A * E * W + 2.0 * A * N * W * K + 2.0 * E * K}
out = {"hints": os.environ.get("GKN_NODE_HINTS", "0,0,0"), "frames": B, "edges": E}
for k, t in (("node_msg", t_msg), ("node_du", t_du), ("node_dh", t_dh)):
    out[k] = {"ms": round(t, 4), "algorithmic_gbs": round(by[k] / t / 1e6)}
print(json.dumps(out))
