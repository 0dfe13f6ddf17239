"""Node encoder's LSTM kernels (csrc/gkn_lstm.cu, reference graph_kernel.py:278-285) against
torch.nn.LSTM evaluated on the CPU in float64.  fp32 kernels with expf / tanhf: 1e-5 relative
(max-norm) on outputs and on every gradient, the parity bar of the fp32 mode."""
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _reference(x, lstm, gout, tail):
    ref = torch.nn.LSTM(3, 3).double()
    ref.load_state_dict({k: v.detach().double().cpu() for k, v in lstm.state_dict().items()})
    xr = x.detach().double().cpu().requires_grad_(True)
    seq, _ = ref(xr)
    (seq[-tail:] * gout.double().cpu()).sum().backward()
    return seq, xr.grad, ref


@pytest.mark.parametrize("T,R,tail", [(1, 1, 1), (7, 5, 7), (160, 504, 16), (40, 300, 4)])
def test_lstm_kernels_match_torch_lstm(T, R, tail):
    from molecular_dynamics_neural_operator_b200 import functional as GF

    gen = torch.Generator().manual_seed(T * 1000 + R)
    lstm = torch.nn.LSTM(3, 3).cuda()
    with torch.no_grad():
        for p in lstm.parameters():  # beyond the default U(-1/sqrt(3), ..): saturating gates too
            p.copy_(torch.randn(p.shape, generator=gen))
    x = (torch.randn(T, R, 3, generator=gen) * 2).cuda().requires_grad_(True)
    gout = torch.randn(tail, R, 3, generator=gen).cuda()
    assert GF.lstm_supported(lstm, x)
    seq = GF.lstm_sequence(lstm, x)
    (seq[-tail:] * gout).sum().backward()
    rseq, rdx, ref = _reference(x, lstm, gout, tail)
    assert rel_err(seq, rseq) <= 1e-5
    assert rel_err(x.grad, rdx) <= 1e-5
    for (name, p), rp in zip(lstm.named_parameters(), ref.parameters()):
        assert rel_err(p.grad, rp.grad) <= 1e-5, name

    with torch.no_grad():  # inference: nothing saved, same numbers
        assert torch.equal(GF.lstm_sequence(lstm, x.detach()), seq.detach())


def test_node_encoder_is_the_same_with_and_without_the_lstm_kernels(monkeypatch):
    from molecular_dynamics_neural_operator_b200 import functional as GF
    from molecular_dynamics_neural_operator_b200.dataset import PairData
    from molecular_dynamics_neural_operator_b200.graph_kernel import KernelNN

    torch.manual_seed(3)
    B, Wn, R = 4, 10, 28
    model = KernelNN(32, 256, 1, 6, in_width=7, window_size=Wn, num_residues=R,
                     precision="fp32").cuda()
    data = PairData(x_aminoacid=torch.randint(0, 20, (B * R,)).cuda(),
                    x_position=torch.randn(B * Wn * R, 3).cuda(), edge_attr=None, edge_index=None)
    outs, grads = [], []
    for own in (True, False):
        monkeypatch.setattr(GF, "OWN_LSTM", own)
        model.zero_grad()
        h = model.node_encoder(data)
        h.square().sum().backward()
        outs.append(h.detach())
        grads.append({k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
    assert rel_err(outs[0], outs[1]) <= 1e-5
    assert set(grads[0]) == set(grads[1]) and any(k.startswith("lstm.") for k in grads[0])
    for k in grads[0]:
        assert rel_err(grads[0][k], grads[1][k]) <= 1e-4, k
