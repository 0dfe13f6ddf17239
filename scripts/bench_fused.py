"""Device time of the fused producer/consumer kernels at the bench shape against the two-kernel
forms they replace, with ablations (producer items only, consumer items only, slack, ring).

    python scripts/bench_fused.py [frames] [--only NAME]   (one JSON line per variant)
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from molecular_dynamics_neural_operator_b200 import ops, synthetic as syn  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
B = int(args[0]) if args else 16
wl = syn.WORKLOADS["bba_allatom"]
pos = torch.from_numpy(syn.trajectory(wl, B, seed=0)).cuda()
ei, ea, rp = ops.radius_graph(pos, wl.cutoff, return_row_ptr=True)
N, W, K, A = B * wl.num_atoms, wl.width, wl.ker_width, 4
plan = ops.plan_from_csr(ei, rp, N)
lay = ops.node_layout(plan, dh_chunks=K // 256)
E = ei.size(1)
bf = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
hs = torch.randn(lay.e_pad, K, device="cuda", generator=g).to(bf)
xb = [torch.randn(N, W, device="cuda", generator=g).to(bf) for _ in range(A)]
w3r = (torch.randn(W, W * K, device="cuda", generator=g) / 8).to(bf)
w3k = w3r.view(W, W, K).permute(0, 2, 1).contiguous().view(W, K * W)
gs_all = torch.randn(A, lay.e_pad, W, device="cuda", generator=g).to(bf)
gs_all[:, lay.dst_pad < 0] = 0
u_all = torch.empty(A, N, W * K, dtype=bf, device="cuda")
msg = torch.empty(lay.e_pad, W, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=12):
    fn(0)
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(reps):
        flush.zero_()  # L2 flush between launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def emit(name, ms, **kw):
    print(json.dumps({"variant": name, "frames": B, "edges": E, "ms": round(ms, 4), **kw}), flush=True)


def want(name):
    return only is None or only == name


# ---------------- forward ----------------
def two_kernel(i):
    torch.mm(xb[i % A], w3r, out=u_all[i % A])
    ops.node_msg(hs, u_all[i % A].view(N, W, K), lay, 0, lay.num_units, 0, msg)


if want("fwd_two_kernel"):
    emit("fwd_two_kernel", timed(two_kernel))
if want("fwd_gemm_only"):
    emit("fwd_gemm_only", timed(lambda i: torch.mm(xb[i % A], w3r, out=u_all[i % A])))
if want("fwd_msg_only"):
    emit("fwd_msg_only", timed(lambda i: ops.node_msg(hs, u_all[i % A].view(N, W, K), lay, 0,
                                                     lay.num_units, 0, msg)))


def fused_runner(sched, ubuf_of):
    flags = torch.zeros(64, sched.num_flags, dtype=torch.int32, device="cuda")
    state = {"n": 0}

    def run(i):
        k = state["n"] % 64
        if k == 0:
            flags.zero_()
        state["n"] += 1
        ops.node_fused_fwd(xb[i % A], w3r, hs, ubuf_of(i), sched, flags[k], msg, N, lay.e_pad, K, W)
    return run


for slack in (0.5, 1.0, 1.5):
    name = f"fwd_fused_keep_slack{slack}"
    if want(name):
        s = ops.fused_fwd_schedule(lay, W, K, ring=0, slack=slack)
        emit(name, timed(fused_runner(s, lambda i: u_all[i % A])), items=s.num_items)
for ring in (3, 4):
    name = f"fwd_fused_ring{ring}"
    if want(name):
        s = ops.fused_fwd_schedule(lay, W, K, ring=ring, slack=1.0)
        ub = torch.empty(s.u_rows, W * K, dtype=bf, device="cuda")
        emit(name, timed(fused_runner(s, lambda i: ub)), items=s.num_items)
# ablations: only the producer tiles / only the consumers (U already in memory, no waits)
base = ops.fused_fwd_schedule(lay, W, K, ring=0, slack=1.0)
it = base.items
if want("fwd_fused_P_only"):
    pi = it.clone()
    pi[pi[:, 0] == ops.IT_C, 0] = ops.IT_NOP   # the consumers become no-ops: same dealing
    p_only = ops.FusedSchedule(pi, base.num_items, base.num_flags, base.u_rows, 0, base.grid)
    emit("fwd_fused_P_only", timed(fused_runner(p_only, lambda i: u_all[i % A])), items=p_only.num_items)
if want("fwd_fused_P_only_nosig"):
    pi = it.clone()
    pi[pi[:, 0] == ops.IT_C, 0] = ops.IT_NOP
    pi[:, 6] = -1   # no counter updates: no GPU-scope fences in the epilogue
    p_only = ops.FusedSchedule(pi, base.num_items, base.num_flags, base.u_rows, 0, base.grid)
    emit("fwd_fused_P_only_nosig", timed(fused_runner(p_only, lambda i: u_all[i % A])))
if want("fwd_fused_C_only"):
    c = it.clone()
    c[c[:, 0] == ops.IT_P, 0] = ops.IT_NOP
    c[:, 4] = -1
    c_only = ops.FusedSchedule(c, base.num_items, base.num_flags, base.u_rows, 0, base.grid)
    two_kernel(0)
    emit("fwd_fused_C_only", timed(fused_runner(c_only, lambda i: u_all[0])), items=c_only.num_items)

# ---------------- backward ----------------
if hasattr(ops, "node_fused_bwd"):
    dut = torch.empty(N, K * W, dtype=bf, device="cuda")
    dw3k = torch.zeros(W, K * W, device="cuda")

    def bwd_three_kernel(i):
        ops.node_du_mn(hs, gs_all[i % A], lay, 0, N, dut)
        torch.mm(dut, w3k.t(), out_dtype=torch.float32)
        torch.addmm(dw3k, xb[i % A].t(), dut, out_dtype=torch.float32)

    if want("bwd_three_kernel"):
        emit("bwd_three_kernel", timed(bwd_three_kernel))
    if want("bwd_du_only"):
        emit("bwd_du_only", timed(lambda i: ops.node_du_mn(hs, gs_all[i % A], lay, 0, N, dut)))
    grid, parts = ops.fused_bwd_grid(K, W)
    import os as _os
    grid_spec = _os.environ.get("BWD_GRID")
    combos = ([tuple(float(v) for v in c.split(":")) + ("all",) for c in grid_spec.split(",")]
              if grid_spec else
              [(3, 0.5, 0.9, "all"), (3, 1.0, 0.9, "all"), (4, 1.0, 0.9, "all"), (4, 2.0, 0.9, "all"),
               (3, 1.0, 0.9, "D_only"), (3, 1.0, 0.9, "D_only_nosig"), (3, 1.0, 0.9, "X_only")])
    for ring, slack, spread, mode in combos:
        ring = int(ring)
        name = f"bwd_fused_ring{ring}_slack{slack}" + ("" if spread == 0.9 else f"_spread{spread}") + \
            ("" if mode == "all" else "_" + mode)
        if not want(name):
            continue
        s = ops.fused_bwd_schedule(N, K, W, hs.device, ring=ring, slack=slack, spread=spread)
        if mode != "all":
            t = s.items.clone()
            t[t[:, 0] == (ops.IT_D if mode == "X_only" else ops.IT_X), 0] = ops.IT_NOP
            t[:, 4] = -1
            if mode.endswith("nosig"):
                t[:, 6] = -1
            s = ops.FusedSchedule(t, s.num_items, s.num_flags, s.u_rows, s.ring, s.grid)
        du = torch.zeros(s.u_rows, K * W, dtype=bf, device="cuda")
        dx_part = torch.empty(parts, N, W, device="cuda")
        dwT = torch.zeros(K * W, W, device="cuda")
        rest = torch.zeros(N, W, device="cuda")
        flags = torch.zeros(64, s.num_flags, dtype=torch.int32, device="cuda")
        st = {"n": 0}

        def run(i):
            k = st["n"] % 64
            if k == 0:
                flags.zero_()
            st["n"] += 1
            ops.node_fused_bwd(hs, gs_all[i % A], lay, xb[i % A], w3k, du, s, flags[k], dx_part, dwT, True)
            ops.sum_partials(dx_part, rest, rest)

        emit(name, timed(run), items=s.num_items)
