"""Bring-up diagnostics: run every kernel of libgkn_b200.so once against plain torch/numpy
references and print error magnitudes and timings.  Not a test (see tests/ for those): it keeps
going after a mismatch so one GPU call yields the whole picture.

    python scripts/gpu_diag.py [--skip-tc] [--big]
"""
from __future__ import annotations

import argparse
import sys
import time
import traceback
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from molecular_dynamics_neural_operator_b200 import ops  # noqa: E402
from oracle import gkn_oracle as orc  # noqa: E402

dev = torch.device("cuda:0")
FAILS = []


def report(name, ok, extra=""):
    print(f"[{'ok' if ok else 'FAIL'}] {name} {extra}", flush=True)
    if not ok:
        FAILS.append(name)


def section(fn):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception:  # noqa: BLE001
        traceback.print_exc()
        FAILS.append(fn.__name__)
        print(f"[FAIL] {fn.__name__} raised", flush=True)


def rand_graph(N, E, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.stack([torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g)])


def t_prepare():
    for N, E in [(1, 0), (5, 3), (50, 700), (504, 20000), (3000, 100000)]:
        ei = rand_graph(N, E, N + E)
        plan = ops.prepare_graph(ei.to(dev), N)
        dst = ei[1].numpy()
        order = np.argsort(dst, kind="stable")
        ok = np.array_equal(plan.perm.cpu().numpy(), order.astype(np.int32))
        ok &= np.array_equal(plan.src.cpu().numpy(), ei[0].numpy()[order])
        ok &= np.array_equal(plan.dst.cpu().numpy(), dst[order])
        cnt = np.bincount(dst, minlength=N)
        ok &= np.array_equal(plan.dst_ptr.cpu().numpy(), np.concatenate([[0], np.cumsum(cnt)]))
        ok &= np.allclose(plan.inv_cnt.cpu().numpy(), 1.0 / np.maximum(cnt, 1))
        src_sorted = ei[0].numpy()[order]
        order2 = np.argsort(src_sorted, kind="stable")
        ok &= np.array_equal(plan.src_perm.cpu().numpy(), order2.astype(np.int32))
        scnt = np.bincount(src_sorted, minlength=N)
        ok &= np.array_equal(plan.src_ptr.cpu().numpy(), np.concatenate([[0], np.cumsum(scnt)]))
        report(f"prepare N={N} E={E}", bool(ok))


def t_radius():
    rng = np.random.default_rng(0)
    for F, N, L, thr in [(1, 28, 20.0, 8.0), (3, 100, 12.0, 4.2), (2, 504, 17.1, 8.0), (1, 2000, 27.0, 4.2),
                         (4, 7, 3.0, 8.0), (1, 1, 1.0, 8.0)]:
        pos = (rng.random((F, N, 3)) * L).astype(np.float32)
        ei_ref, ea_ref = orc.radius_graph_batched(pos, thr)
        ei, ea = ops.radius_graph(torch.from_numpy(pos).to(dev), thr)
        ok = ei.shape == tuple(ei_ref.shape) and np.array_equal(ei.cpu().numpy(), ei_ref)
        ok = ok and np.array_equal(ea.cpu().numpy(), ea_ref)
        report(f"radius F={F} N={N} thr={thr}", bool(ok), f"E={ei_ref.shape[1]} got {tuple(ei.shape)}")


def t_scatter():
    for N, E, W in [(50, 700, 64), (504, 20000, 64), (300, 5000, 32), (100, 3000, 128), (40, 100, 16)]:
        ei = rand_graph(N, E, 7 * N + E)
        plan = ops.prepare_graph(ei.to(dev), N)
        g = torch.Generator().manual_seed(1)
        msg = torch.randn(E, W, generator=g)
        xb3 = torch.randn(N, W, generator=g)
        xroot = torch.randn(N, W, generator=g)
        bias = torch.randn(W, generator=g)
        order = plan.perm.cpu().long()
        src_s = ei[0][order]
        dst_s = ei[1][order]
        ref = orc.scatter_mean((msg + xb3[src_s]).double(), dst_s, N) + xroot.double() + bias.double()
        out = ops.scatter_mean(msg.to(dev), plan, xb3.to(dev), xroot.to(dev), bias.to(dev), relu=False)
        err = (out.cpu().double() - ref).abs().max().item()
        report(f"scatter_mean N={N} E={E} W={W}", err < 1e-4, f"maxerr={err:.2e}")
        out = ops.scatter_mean(msg.to(dev), plan, None, None, None, relu=True)
        ref2 = torch.relu(orc.scatter_mean(msg.double(), dst_s, N))
        err = (out.cpu().double() - ref2).abs().max().item()
        report(f"scatter_mean(relu,no extras) N={N} W={W}", err < 1e-4, f"maxerr={err:.2e}")
        # source-side sum
        dxj = torch.randn(E, W, generator=g)
        gb3 = torch.randn(N, W, generator=g)
        add = torch.randn(N, W, generator=g)
        ref3 = torch.zeros(N, W, dtype=torch.float64).index_add_(0, src_s, (dxj + gb3[dst_s]).double()) + add.double()
        out3 = ops.scatter_src_sum(dxj.to(dev), plan, gb3.to(dev), add.to(dev))
        err = (out3.cpu().double() - ref3).abs().max().item()
        report(f"scatter_src_sum N={N} E={E} W={W}", err < 1e-3, f"maxerr={err:.2e}")
        # relu grad
        gy = torch.randn(N, W, generator=g)
        y = torch.randn(N, W, generator=g)
        gg, gn = ops.relu_grad_scale(gy.to(dev), y.to(dev), plan.inv_cnt, True)
        refg = gy * (y > 0)
        ok = torch.equal(gg.cpu(), refg) and torch.allclose(gn.cpu(), refg * plan.inv_cnt.cpu()[:, None])
        report(f"relu_grad_scale N={N} W={W}", ok)
        # edge outer sum
        A = 3
        xs = torch.randn(A, N, W, generator=g)
        gs = torch.randn(A, N, W, generator=g)
        e0, e1 = E // 3, min(E, E // 3 + 257)
        S = ops.edge_outer_sum(xs.to(dev), gs.to(dev), plan, e0, e1, torch.float32)
        refS = torch.einsum("aei,aeo->eio", xs[:, src_s[e0:e1]].double(), gs[:, dst_s[e0:e1]].double()).reshape(e1 - e0, -1)
        err = (S.cpu().double() - refS).abs().max().item()
        report(f"edge_outer_sum f32 W={W}", err < 1e-4, f"maxerr={err:.2e}")
        Sb = ops.edge_outer_sum(xs.to(dev), gs.to(dev), plan, e0, e1, torch.bfloat16)
        err = ((Sb.cpu().double() - refS).abs() / (refS.abs() + 1)).max().item()
        report(f"edge_outer_sum bf16 W={W}", err < 2e-2, f"maxrel={err:.2e}")
    x = torch.randn(333, 64)
    b = torch.randn(64)
    out = ops.bias_act_cast_bf16(x.to(dev), b.to(dev), relu=True)
    ref = torch.relu(x + b).to(torch.bfloat16)
    report("bias_act_cast_bf16", torch.equal(out.cpu(), ref))


def contract_ref(h, w3, idx, node, W, kind):
    P = (h.double() @ w3.double().t()).view(-1, W, W)
    nd = node.double()[idx.long()]
    if kind == "fwd":
        return torch.einsum("ei,eio->eo", nd, P)
    return torch.einsum("eio,eo->ei", P, nd)


def t_contract(precision):
    cases = [(300, 64, 64, 40), (128 * 5, 128, 64, 77), (1000, 1024, 64, 504), (257, 256, 32, 28), (200, 128, 128, 50),
             (20000, 1024, 64, 504)]
    for E, K, W, N in cases:
        g = torch.Generator().manual_seed(E + K + W)
        h = torch.relu(torch.randn(E, K, generator=g))
        w3 = torch.randn(W * W, K, generator=g) / np.sqrt(K)
        idx = torch.randint(0, N, (E,), generator=g, dtype=torch.int32)
        node = torch.randn(N, W, generator=g)
        if precision == "bf16":
            hq, wq = h.to(torch.bfloat16), w3.to(torch.bfloat16)
            h_ref, w_ref = hq.float(), wq.float()
        else:
            hq, wq = h, w3
            h_ref, w_ref = h, w3
        for kind in ("fwd", "dx"):
            fn = ops.contract_fwd if kind == "fwd" else ops.contract_dx
            try:
                out = fn(hq.to(dev), wq.to(dev), idx.to(dev), node.to(dev), precision)
                torch.cuda.synchronize()
            except Exception as exc:  # noqa: BLE001
                report(f"contract_{kind}_{precision} E={E} K={K} W={W}", False, f"raised {exc}")
                raise
            ref = contract_ref(h_ref, w_ref, idx, node, W, kind)
            err = (out.cpu().double() - ref).abs().max().item()
            scale = ref.abs().max().item()
            report(f"contract_{kind}_{precision} E={E} K={K} W={W}", err < 2e-4 * max(scale, 1.0),
                   f"maxerr={err:.3e} scale={scale:.2e}")
            if precision == "bf16":
                ref_full = contract_ref(h, w3, idx, node, W, kind)
                rel = ((out.cpu().double() - ref_full).norm() / ref_full.norm()).item()
                print(f"      bf16 rounding vs unrounded fp64: rel-fro={rel:.3e}", flush=True)


def t_perf():
    E, K, W, N = 57000, 1024, 64, 504
    g = torch.Generator().manual_seed(0)
    h = torch.relu(torch.randn(E, K, generator=g)).to(torch.bfloat16).to(dev)
    w3 = (torch.randn(W * W, K, generator=g) / 32).to(torch.bfloat16).to(dev)
    idx = torch.randint(0, N, (E,), generator=g, dtype=torch.int32).to(dev)
    node = torch.randn(N, W, generator=g).to(dev)
    for kind, fn in (("fwd", ops.contract_fwd), ("dx", ops.contract_dx)):
        for _ in range(3):
            fn(h, w3, idx, node, "bf16")
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        iters = 10
        for _ in range(iters):
            fn(h, w3, idx, node, "bf16")
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / iters
        tf = 2.0 * E * K * W * W / (ms * 1e-3) / 1e12
        print(f"[perf] contract_{kind}_bf16 E={E} K={K} W={W}: {ms:.3f} ms  {tf:.1f} TFLOP/s", flush=True)
    hf, wf = h.float(), w3.float()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.contract_fwd(hf, wf, idx, node, "fp32")
    ev0.record()
    ops.contract_fwd(hf, wf, idx, node, "fp32")
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    print(f"[perf] contract_fwd_f32: {ms:.3f} ms  {2.0 * E * K * W * W / (ms * 1e-3) / 1e12:.1f} TFLOP/s", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-tc", action="store_true")
    ap.add_argument("--perf", action="store_true")
    args = ap.parse_args()
    print(torch.cuda.get_device_name(0), flush=True)
    t0 = time.time()
    section(t_prepare)
    section(t_radius)
    section(t_scatter)
    section(lambda: t_contract("fp32"))
    if not args.skip_tc:
        section(lambda: t_contract("bf16"))
        if args.perf:
            section(t_perf)
    print(f"done in {time.time() - t0:.1f}s; failures: {FAILS}", flush=True)
    return 1 if FAILS else 0


if __name__ == "__main__":
    sys.exit(main())
