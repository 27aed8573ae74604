#!/usr/bin/env python
"""Small driver for ncu: k_lar on the config-5 grid (50 000 energies x 21 fields), device resident."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nest_b200 import capi  # noqa: E402

species = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nl = 400 * 50000
ctx = capi.Context(0)
E = (0.1 + (1000.0 - 0.1) / 50000.0 * torch.arange(50000, dtype=torch.float64, device="cuda")).repeat(nl // 50000)
fld = torch.tensor([1.0] + [50.0 * k for k in range(1, 21)], dtype=torch.float64, device="cuda")
F = fld[torch.arange(nl // 50000, device="cuda") % fld.numel()].repeat_interleave(50000)
Y = torch.empty((9, nl), dtype=torch.float64, device="cuda")
Fl = torch.empty((4, nl), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
for rep in range(3):
    ctx.lar_device(species, nl, E.data_ptr(), F.data_ptr(), 1.393, Y.data_ptr(), Fl.data_ptr(), seed=1)
    ctx.synchronize()
print("k_lar species %d: %d events" % (species, nl))
ctx.close()
