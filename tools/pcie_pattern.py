#!/usr/bin/env python
"""What the link gives for the e2e pattern itself, without any kernel: 8 GiB D2H in 32 chunks of 256 MiB over 6 streams into
one pinned buffer, alone and with the 3.5 GiB of H2D chunks going the other way at the same time."""
import json
import time
import torch


def main():
    out_b, in_b, nch, nst = 8 << 30, int(3.46 * (1 << 30)), 32, 6
    d_out = torch.empty(out_b, dtype=torch.uint8, device="cuda")
    h_out = torch.empty(out_b, dtype=torch.uint8).pin_memory()
    h_in = torch.empty(in_b, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(in_b, dtype=torch.uint8, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(nst)]
    co, ci = out_b // nch, in_b // nch
    res = {}
    for name, with_h2d in (("d2h_alone", False), ("d2h_with_h2d", True), ("d2h_alone_again", False)):
        best = None
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(nch):
                with torch.cuda.stream(streams[k % nst]):
                    if with_h2d:
                        d_in[k * ci:(k + 1) * ci].copy_(h_in[k * ci:(k + 1) * ci], non_blocking=True)
                    h_out[k * co:(k + 1) * co].copy_(d_out[k * co:(k + 1) * co], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        res[name] = round(out_b / best / 1e9, 2)
    # the same bytes with ONE stream per direction: every H2D chunk on stream A, every D2H chunk on stream B, chunk k's D2H
    # waiting (event) for chunk k's H2D — what a pipeline with dedicated copy streams would issue
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    for lag in (0, 1, 2):
        best = None
        for rep in range(3):
            torch.cuda.synchronize()
            evs = [torch.cuda.Event() for _ in range(nch)]
            t0 = time.perf_counter()
            for k in range(nch + lag):
                if k < nch:
                    with torch.cuda.stream(sa):
                        d_in[k * ci:(k + 1) * ci].copy_(h_in[k * ci:(k + 1) * ci], non_blocking=True)
                        evs[k].record(sa)
                j = k - lag
                if j >= 0:
                    with torch.cuda.stream(sb):
                        sb.wait_event(evs[j])
                        h_out[j * co:(j + 1) * co].copy_(d_out[j * co:(j + 1) * co], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        res[f"one_stream_per_direction_lag{lag}"] = round(out_b / best / 1e9, 2)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
