"""Device time of the dh pass with and without the fused ReLU backward (bench shape).

    GKN_NODE_HINTS=0,0,<dh bits> python scripts/profile_dh_mask.py [frames]
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
bf = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
hs = torch.relu(torch.randn(lay.e_pad, K, device="cuda", generator=g)).to(bf)
u_all = (torch.randn(A, N, W * K, device="cuda", generator=g) / 8).to(bf)
gs_all = torch.randn(A, lay.e_pad, W, device="cuda", generator=g).to(bf)
gs_all[:, lay.dst_pad < 0] = 0
dh = torch.empty(lay.e_pad, K, dtype=bf, device="cuda")
colsum = torch.zeros(K, device="cuda")


def timed(fn, reps=6):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


u4 = u_all.view(A, N, W, K)
t_plain = timed(lambda: ops.node_dh_mn(gs_all, u4, lay, 0, lay.num_units, 0, dh))
bits = ops.relu_bits(hs)
t_bits = timed(lambda: ops.relu_bits(hs))
t_mask = timed(lambda: ops.node_dh_mn(gs_all, u4, lay, 0, lay.num_units, 0, dh, act_bits=bits,
                                      colsum=colsum))
print(json.dumps({"hints": os.environ.get("GKN_NODE_HINTS", "0,0,0"), "plain_ms": round(t_plain, 4),
                  "masked_ms": round(t_mask, 4), "relu_bits_ms": round(t_bits, 4)}))
