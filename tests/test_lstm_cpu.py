"""Host side of the node encoder's LSTM path (functional._LSTMSeq) on the CPU: the two kernels of
csrc/gkn_lstm.cu are replaced by a step-by-step torch restatement of what they compute (same saved
tensors, same gradient layout: dW_ih | dW_hh | db per block), and the autograd glue around them is
compared with torch.nn.LSTM's own autograd in float64."""
import torch

from molecular_dynamics_neural_operator_b200 import functional as GF
from molecular_dynamics_neural_operator_b200 import ops
from tests.helpers import rel_err


def emu_lstm_fwd(x, w_ih, w_hh, b_ih, b_hh, save):
    T, R, D = x.shape
    h = torch.zeros(R, D)
    c = torch.zeros(R, D)
    out, gates, cells = [], [], []
    for t in range(T):
        a = x[t] @ w_ih.t() + h @ w_hh.t() + b_ih + b_hh
        i, f, g, o = a[:, :D].sigmoid(), a[:, D:2 * D].sigmoid(), a[:, 2 * D:3 * D].tanh(), a[:, 3 * D:].sigmoid()
        c = f * c + i * g
        h = o * c.tanh()
        out.append(h)
        gates.append(torch.cat([i, f, g, o], 1))
        cells.append(c)
    out = torch.stack(out)
    return (out, torch.stack(gates), torch.stack(cells)) if save else (out, None, None)


def emu_lstm_bwd(x, w_ih, w_hh, out, gates, cells, dout, need_dx):
    T, R, D = x.shape
    dh = torch.zeros(R, D)
    dc = torch.zeros(R, D)
    dw_ih, dw_hh, db = torch.zeros(4 * D, D), torch.zeros(4 * D, D), torch.zeros(4 * D)
    dx = torch.zeros(T, R, D) if need_dx else None
    for t in range(T - 1, -1, -1):
        i, f, g, o = gates[t].split(D, 1)
        tc = cells[t].tanh()
        cp = cells[t - 1] if t > 0 else torch.zeros(R, D)
        hp = out[t - 1] if t > 0 else torch.zeros(R, D)
        dht = dout[t] + dh
        dct = dc + dht * o * (1 - tc * tc)
        da = torch.cat([dct * g * i * (1 - i), dct * cp * f * (1 - f), dct * i * (1 - g * g),
                        dht * tc * o * (1 - o)], 1)
        dc = dct * f
        dw_ih += da.t() @ x[t]
        dw_hh += da.t() @ hp
        db += da.sum(0)
        dh = da @ w_hh
        if need_dx:
            dx[t] = da @ w_ih
    return dx, dw_ih, dw_hh, db


def test_lstm_autograd_glue_matches_torch_lstm(monkeypatch):
    monkeypatch.setattr(ops, "lstm_fwd", emu_lstm_fwd)
    monkeypatch.setattr(ops, "lstm_bwd", emu_lstm_bwd)
    gen = torch.Generator().manual_seed(4)
    T, R, tail = 12, 5, 3
    lstm = torch.nn.LSTM(3, 3)
    with torch.no_grad():
        for p in lstm.parameters():
            p.copy_(torch.randn(p.shape, generator=gen))
    x = torch.randn(T, R, 3, generator=gen, requires_grad=True)
    gout = torch.randn(tail, R, 3, generator=gen)
    seq = GF.lstm_sequence(lstm, x)
    (seq[-tail:] * gout).sum().backward()

    ref = torch.nn.LSTM(3, 3).double()
    ref.load_state_dict({k: v.detach().double() for k, v in lstm.state_dict().items()})
    xr = x.detach().double().requires_grad_(True)
    rseq, _ = ref(xr)
    (rseq[-tail:] * gout.double()).sum().backward()
    assert rel_err(seq, rseq) <= 1e-5
    assert rel_err(x.grad, xr.grad) <= 1e-5
    for (name, p), rp in zip(lstm.named_parameters(), ref.parameters()):
        assert rel_err(p.grad, rp.grad) <= 1e-5, name


def test_lstm_kernels_are_only_taken_for_the_reference_shape():
    x = torch.zeros(4, 2, 3)
    assert not GF.lstm_supported(torch.nn.LSTM(3, 3), x)          # CPU tensor: stock torch
    if torch.cuda.is_available():  # (not on the CPU test box)
        xc = x.cuda()
        assert GF.lstm_supported(torch.nn.LSTM(3, 3).cuda(), xc)
        assert not GF.lstm_supported(torch.nn.LSTM(3, 4).cuda(), xc)
        assert not GF.lstm_supported(torch.nn.LSTM(3, 3, num_layers=2).cuda(), xc)
