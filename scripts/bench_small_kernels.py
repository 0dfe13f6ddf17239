"""Device time of the small HBM-bound kernels around the node-first path at the bench shape
(16 BBA all-atom frames): CUDA events, algorithmic bytes / time against the copy bandwidth.

    python scripts/bench_small_kernels.py [frames]
"""
import json
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
E_pad = lay.e_pad
bf = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
ea_s = torch.randn(E_pad, 6, device="cuda", generator=g)
w1 = torch.randn(K, 6, device="cuda", generator=g)
b1 = torch.randn(K, device="cuda", generator=g)
hs = torch.relu(torch.randn(E_pad, K, device="cuda", generator=g)).to(bf)
dz = torch.randn(E_pad // 4, K, device="cuda", generator=g).to(bf)
gn = torch.randn(N, W, device="cuda", generator=g)
gs = torch.empty(E_pad, W, dtype=bf, device="cuda")
parts = ops.device_sm_count()
dx_part = torch.randn(parts, N, W, device="cuda", generator=g)
rest = torch.zeros(N, W, device="cuda")
X = torch.randn(6 * N, W, device="cuda", generator=g)
G = torch.randn(6 * N, W, device="cuda", generator=g)
cs = torch.zeros(K, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()  # L2 flush between launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


rows = [
    ("edge_layer0 (bf16 out)", lambda: ops.edge_layer0(ea_s, w1, b1, bf), E_pad * (24 + 2 * K)),
    ("relu_bits", lambda: ops.relu_bits(hs), E_pad * K * (2 + 1 / 8)),
    ("node_gather", lambda: ops.node_gather(gn, lay, gs, None), E_pad * (4 + 2 * W) + N * W * 4),
    ("sum_partials", lambda: ops.sum_partials(dx_part, rest, rest), (parts + 2) * N * W * 4),
    ("xty (6 applications)", lambda: ops.xty(X, G, True), 2 * 6 * N * W * 4),
    ("mask_cast_colsum (bf16, in place, quarter of the edges)",
     lambda: ops.mask_cast_colsum(dz, hs[:E_pad // 4], cs, bf), 3 * (E_pad // 4) * K * 2),
]
for name, fn, nbytes in rows:
    ms = timed(fn)
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "algorithmic_bytes": int(nbytes),
                      "gbs": round(nbytes / ms / 1e6)}))
