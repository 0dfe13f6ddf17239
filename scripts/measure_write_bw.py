"""Pure-write and pure-read HBM bandwidth (torch fill_ / sum over 2 GiB) beside the copy figure in
MEASURED_PEAKS.json: the radius-graph fill pass is a pure write stream."""
import json

import torch

x = torch.empty(1 << 29, dtype=torch.float32, device="cuda")  # 2 GiB
y = torch.empty_like(x)


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


nb = x.numel() * 4
out = {
    "write_only_fill_gbs": nb / t(lambda: x.fill_(1.0)) / 1e6,
    "memset_zero_gbs": nb / t(lambda: x.zero_()) / 1e6,
    "read_only_sum_gbs": nb / t(lambda: x.sum()) / 1e6,
    "copy_read_plus_write_gbs": 2 * nb / t(lambda: y.copy_(x)) / 1e6,
}
print(json.dumps(out))
