"""Fused edge-gradient kernels vs the staged pass (edge_outer_sum -> S in HBM -> library GEMM) on the
bench workload's shapes: 8 BBA all-atom frames, A = 12 applications, w = 64, K = 1024."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from molecular_dynamics_neural_operator_b200 import ops, synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
wl = syn.WORKLOADS["bba_allatom"]
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pos = torch.from_numpy(syn.trajectory(wl, frames, seed=0)).to(dev)
ei, ea, rp = ops.radius_graph(pos, wl.cutoff, return_row_ptr=True)
N, E = wl.num_atoms * frames, ei.size(1)
plan = ops.plan_from_csr(ei, rp, N)
A, W, K = 2 * wl.depth, wl.width, wl.ker_width
g = torch.Generator(device=dev).manual_seed(0)
xs = torch.randn(A, N, W, device=dev, generator=g)
gs = torch.randn(A, N, W, device=dev, generator=g)
w3 = (torch.randn(W * W, K, device=dev, generator=g) / W).bfloat16()
w3t = w3.t().contiguous()
h = torch.relu(torch.randn(E, K, device=dev, generator=g)).bfloat16()


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def staged_dh(chunk=16384):
    for e0 in range(0, E, chunk):
        e1 = min(E, e0 + chunk)
        S = ops.edge_outer_sum(xs, gs, plan, e0, e1, torch.bfloat16)
        torch.mm(S, w3)


flop = 2.0 * E * K * W * W
ms = timeit(lambda: ops.edge_grad_dh(xs, gs, plan, w3t))
print(json.dumps({"kernel": "edge_grad_dh (fused, S in smem)", "edges": E, "ms": ms, "tflops": flop / ms / 1e9}))
ms = timeit(staged_dh)
print(json.dumps({"kernel": "staged dh: edge_outer_sum + cuBLAS", "edges": E, "ms": ms, "tflops": flop / ms / 1e9}))
if hasattr(ops, "edge_grad_dw"):
    ht = h.t().contiguous()
    ms = timeit(lambda: ops.edge_grad_dw(xs, gs, plan, ht))
    print(json.dumps({"kernel": "edge_grad_dw (fused, S^T in smem)", "edges": E, "ms": ms, "tflops": flop / ms / 1e9}))

    def staged_dw(chunk=16384):
        dw3 = torch.zeros(W * W, K, device=dev)
        for e0 in range(0, E, chunk):
            e1 = min(E, e0 + chunk)
            S = ops.edge_outer_sum(xs, gs, plan, e0, e1, torch.bfloat16)
            dw3 = torch.addmm(dw3, S.t(), h[e0:e1], out_dtype=torch.float32)
    ms = timeit(staged_dw)
    print(json.dumps({"kernel": "staged dW3: edge_outer_sum + cuBLAS", "edges": E, "ms": ms, "tflops": flop / ms / 1e9}))
