"""Device time of the node encoder's LSTM (T = window * batch steps, R rows) forward + backward:
csrc/gkn_lstm.cu against nn.LSTM (cuDNN).   python scripts/bench_lstm.py [batch] [rows]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from molecular_dynamics_neural_operator_b200 import functional as GF  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
R = int(sys.argv[2]) if len(sys.argv) > 2 else 504
T = 10 * B
lstm = torch.nn.LSTM(3, 3).cuda()
x = torch.randn(T, R, 3, device="cuda")
gout = torch.randn(B, R, 3, device="cuda")


def step(own):
    for p in lstm.parameters():
        p.grad = None
    if own:
        seq = GF.lstm_sequence(lstm, x)
    else:
        seq, _ = lstm(x)
    (seq[-B:] * gout).sum().backward()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(json.dumps({"T": T, "R": R, "own_ms": round(timed(lambda: step(True)), 4),
                  "cudnn_ms": round(timed(lambda: step(False)), 4)}))
