"""PCIe ceilings for the e2e (host-buffer) path: pinned D2H of one launch's outputs (118 MB),
pinned H2D of its actions (26 MB), alone and concurrently on two streams; best of 10, CUDA
events.  bench.py's e2e figure moves exactly these bytes per step, so
    e2e ceiling = 6 553 600 env-steps / max(t_d2h, t_h2d) when both directions overlap.
Usage: python tools/pcie_peak.py"""
import json

import torch


def main():
    d2h_bytes, h2d_bytes = 117964800, 26214400
    dev_out = torch.empty(d2h_bytes, dtype=torch.uint8, device="cuda")
    host_out = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    dev_in = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    host_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, n=10):
        best = 1e9
        for _ in range(n + 2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            fn()
            s1.synchronize(); s2.synchronize()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) * 1e-3)
        return best

    def d2h():
        with torch.cuda.stream(s1):
            host_out.copy_(dev_out, non_blocking=True)

    def h2d():
        with torch.cuda.stream(s2):
            dev_in.copy_(host_in, non_blocking=True)

    def both():
        d2h(); h2d()

    t_d, t_h, t_b = timed(d2h), timed(h2d), timed(both)
    print(json.dumps({
        "d2h_GBps": d2h_bytes / t_d / 1e9, "h2d_GBps": h2d_bytes / t_h / 1e9,
        "both_ms": t_b * 1e3, "d2h_ms": t_d * 1e3, "h2d_ms": t_h * 1e3,
        "e2e_ceiling_env_steps_per_s": 6553600 / t_b,
    }))


if __name__ == "__main__":
    main()
