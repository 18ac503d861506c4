"""HBM ceilings for a write-heavy stream, measured with torch ops and CUDA events (best of 10):
copy (read+write 1:1, the MEASURED_PEAKS.json definition), fill (write only), reduce (read
only), and a 1:4 read:write broadcast copy -- the mix of the supply-chain rollout kernel
(26 MB of actions in, 118 MB of outputs out per launch).  Usage: python tools/hbm_write_peak.py"""
import json

import torch


def best(fn, n=10):
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts)


def main():
    N = 1 << 30  # bytes per buffer
    x = torch.empty(N // 4, dtype=torch.float32, device="cuda").normal_()
    y = torch.empty_like(x)
    q = torch.empty(N // 16, dtype=torch.float32, device="cuda").normal_()  # 256 MB > L2
    for _ in range(3):
        y.copy_(x); y.fill_(1.0); x.sum(); y.view(4, -1).copy_(q.view(1, -1))
    out = {
        "copy_GBps": 2 * N / best(lambda: y.copy_(x)) / 1e9,
        "fill_GBps": N / best(lambda: y.fill_(1.0)) / 1e9,
        "read_GBps": N / best(lambda: x.sum()) / 1e9,
        "mix_1r4w_GBps": (N + N // 4) / best(lambda: y.view(4, -1).copy_(q.view(1, -1))) / 1e9,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
