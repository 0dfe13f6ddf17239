"""Micro-benchmarks of the HBM-bound kernels (radius graph, segmented node update) and of the
contraction at the other widths, on batched frames of the BASELINE.json shapes.  Prints one JSON
line per kernel: algorithmic bytes / CUDA-event time vs the measured HBM peak.

    python scripts/bench_kernels.py [--quick]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from molecular_dynamics_neural_operator_b200 import ops, synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {
    "hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


_FLUSH = None


def timeit(fn, iters=10, warm=3, flush_l2=False):
    """Mean device time of fn (CUDA events on the current stream).  flush_l2: every launch is timed
    on its own with a 256 MB write in between, for working sets smaller than the 126 MB L2."""
    global _FLUSH
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if not flush_l2:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    if _FLUSH is None:
        _FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot = 0.0
    for _ in range(iters):
        _FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


RESULTS: list = []   # every record is also kept here (bench.py imports this module)
QUIET = False


def emit(**kw):
    RESULTS.append(kw)
    if not QUIET:
        print(json.dumps(kw), flush=True)


def graph_pass_times(pos, thr, iters=10):
    """Device time of the two passes of the radius-graph build through the C ABI with every buffer
    preallocated (CUDA events on the launching stream; no allocation or host read in between)."""
    from molecular_dynamics_neural_operator_b200 import _lib
    lib = _lib.load()
    F, N = int(pos.size(0)), int(pos.size(1))
    ws_bytes = int(lib.gkn_radius_graph_workspace_bytes(F, N))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    deg = torch.empty(F * N, dtype=torch.int32, device=dev)
    row_ptr = torch.empty(F * N + 1, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    p = lambda t: t.data_ptr()  # noqa: E731

    def count():
        _lib.check(lib.gkn_radius_graph_count(p(pos), F, N, float(thr), p(deg), p(row_ptr), p(ws), ws_bytes, st),
                   "count")
    count()
    E = int(row_ptr[-1].item())
    ei = torch.empty((2, E), dtype=torch.int64, device=dev)
    ea = torch.empty((E, 6), dtype=torch.float32, device=dev)

    def fill():
        _lib.check(lib.gkn_radius_graph_fill(p(pos), F, N, float(thr), p(row_ptr), E, p(ei), p(ea), p(ws),
                                             ws_bytes, st), "fill")
    def both():
        count()
        fill()
    return timeit(count, iters), timeit(fill, iters), timeit(both, iters), E


def bench_graph(name, frames):
    wl = syn.WORKLOADS[name]
    pos = torch.from_numpy(syn.trajectory(wl, frames, seed=0)).to(dev)
    ei, ea = ops.radius_graph(pos, wl.cutoff)
    E, N = ei.size(1), wl.num_atoms * frames
    cms, fms, bms, E2 = graph_pass_times(pos, wl.cutoff)
    assert E2 == E
    nb = 12 * N + 40 * E
    emit(kernel="radius_graph device time (count pass, fill pass, both back to back)", workload=name,
         frames=frames, atoms=N, edges=E, count_ms=cms, fill_ms=fms, ms=bms, algorithmic_bytes=nb,
         gbs=nb / bms / 1e6, frac_hbm=nb / bms / 1e6 / peaks["hbm_gbs"],
         fill_frac_hbm=nb / fms / 1e6 / peaks["hbm_gbs"])
    ms = timeit(lambda: ops.radius_graph(pos, wl.cutoff), iters=5, warm=2)
    nbytes = 12 * N + 40 * E
    emit(kernel="radius_graph (count+fill, incl. host read of E)", workload=name, frames=frames, atoms=N, edges=E,
         ms=ms, algorithmic_bytes=nbytes, gbs=nbytes / ms / 1e6, frac_hbm=nbytes / ms / 1e6 / peaks["hbm_gbs"])
    ms2 = timeit(lambda: ops.prepare_graph(ei, N), iters=5, warm=2)
    emit(kernel="graph_prepare", workload=name, frames=frames, edges=E, ms=ms2)
    return ei, ea, N, E


def bench_node_update(name, frames, ei, N, E):
    wl = syn.WORKLOADS[name]
    W = wl.width
    plan = ops.prepare_graph(ei, N)
    g = torch.Generator(device=dev).manual_seed(0)
    msg = torch.randn(E, W, device=dev, generator=g)
    x = torch.randn(N, W, device=dev, generator=g)
    M = torch.randn(W, W, device=dev, generator=g)
    bias = torch.randn(W, device=dev, generator=g)
    nbytes = 4 * E * W + 4 * E + 4 * N * W + 4 * N
    ms = timeit(lambda: ops.segment_matvec(msg, None, plan.dst_ptr, plan.src, x, plan.inv_cnt, M, x, M, bias, True))
    emit(kernel="segment_matvec fwd (mean + b3 + root + bias + relu)", workload=name, frames=frames, edges=E, W=W,
         ms=ms, algorithmic_bytes=nbytes, gbs=nbytes / ms / 1e6, frac_hbm=nbytes / ms / 1e6 / peaks["hbm_gbs"])
    ms = timeit(lambda: ops.scatter_mean(msg, plan, None, None, bias, True))
    emit(kernel="scatter_mean_fwd (plain segmented mean)", workload=name, frames=frames, edges=E, W=W,
         ms=ms, algorithmic_bytes=nbytes, gbs=nbytes / ms / 1e6, frac_hbm=nbytes / ms / 1e6 / peaks["hbm_gbs"])
    ms = timeit(lambda: ops.segment_matvec(msg, plan.src_perm, plan.src_ptr, plan.dst, x, None, M, x, M, None, False))
    emit(kernel="segment_matvec bwd (scatter to sources)", workload=name, frames=frames, edges=E, W=W,
         ms=ms, algorithmic_bytes=nbytes + 4 * E, gbs=(nbytes + 4 * E) / ms / 1e6,
         frac_hbm=(nbytes + 4 * E) / ms / 1e6 / peaks["hbm_gbs"])


def bench_contract(name, frames, E, N):
    wl = syn.WORKLOADS[name]
    W, K = wl.width, wl.ker_width
    g = torch.Generator(device=dev).manual_seed(0)
    h = torch.relu(torch.randn(E, K, device=dev, generator=g)).bfloat16()
    w3 = (torch.randn(W * W, K, device=dev, generator=g) / 32).bfloat16()
    idx = torch.randint(0, N, (E,), device=dev, generator=g, dtype=torch.int32)
    x = torch.randn(N, W, device=dev, generator=g)
    for kind, fn in (("fwd", ops.contract_fwd), ("dx", ops.contract_dx)):
        ms = timeit(lambda: fn(h, w3, idx, x, "bf16"), iters=5, warm=2)
        flop = 2.0 * E * K * W * W
        emit(kernel=f"contract_{kind}_bf16", workload=name, frames=frames, edges=E, W=W, K=K, ms=ms,
             tflops=flop / ms / 1e9, frac_bf16_burst=flop / ms / 1e9 / peaks["bf16_tflops"])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--graph-only", action="store_true")
    a = ap.parse_args()
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "peaks": peaks}))
    if a.graph_only:
        for name, frames in (("bba_ca", 128), ("bba_allatom", 8), ("bba_allatom", 64), ("bba_allatom", 512),
                             ("synthetic_10k", 16), ("synthetic_10k", 64), ("synthetic_100k", 1),
                             ("synthetic_100k", 4)):
            bench_graph(name, frames)
        sys.exit(0)
    ei, ea, N, E = bench_graph("bba_allatom", 64)
    bench_node_update("bba_allatom", 64, ei, N, E)
    del ei, ea
    ei, ea, N, E = bench_graph("synthetic_10k", 16 if a.quick else 64)
    bench_node_update("synthetic_10k", 64, ei, N, E)
    bench_contract("synthetic_10k", 1, E // (16 if a.quick else 64), 10000)
    del ei, ea
    ei, ea, N, E = bench_graph("synthetic_100k", 1)
    bench_node_update("synthetic_100k", 1, ei, N, E)
    if not a.quick:
        bench_contract("synthetic_100k", 1, E, N)
