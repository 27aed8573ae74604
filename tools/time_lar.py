#!/usr/bin/env python
"""k_lar timed alone with CUDA events on the config-5 grid (50 000 energies x 21 fields), device resident:
GB/s of algorithmic traffic (16 B in + 72 B yields + 32 B fluctuations per event) per species."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nest_b200 import capi  # noqa: E402

nl = 400 * 50000
ctx = capi.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
E = (0.1 + (1000.0 - 0.1) / 50000.0 * torch.arange(50000, dtype=torch.float64, device="cuda")).repeat(nl // 50000)
fld = torch.tensor([1.0] + [50.0 * k for k in range(1, 21)], dtype=torch.float64, device="cuda")
F = fld[torch.arange(nl // 50000, device="cuda") % fld.numel()].repeat_interleave(50000)
Y = torch.empty((9, nl), dtype=torch.float64, device="cuda")
Fl = torch.empty((4, nl), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
for sp, name in ((0, "NR"), (1, "ER"), (2, "alpha")):
    for fl, B in ((Fl.data_ptr(), 120.0), (0, 88.0)):
        for rep in range(3):
            ctx.lar_device(sp, nl, E.data_ptr(), F.data_ptr(), 1.393, Y.data_ptr(), fl, seed=1)
        ctx.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for rep in range(10):
            ctx.lar_device(sp, nl, E.data_ptr(), F.data_ptr(), 1.393, Y.data_ptr(), fl, seed=1 + rep)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print("k_lar %-5s %-12s %.3f ms per %d events = %.2f TB/s (%.2f of 6558 GB/s)" %
              (name, "yields+fluct" if fl else "yields", ms, nl, B * nl / ms / 1e9, B * nl / ms / 1e9 / 6.558))
ctx.close()
