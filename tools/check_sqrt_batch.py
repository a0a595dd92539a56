#!/usr/bin/env python
"""rolling(...).std() against sqrt(rolling(...).var()) on the device: the std kernel's batched square root (csrc/rolling.cu
sqrt_batch) must stay within one ulp of the correctly rounded root, and give what sqrt() gives for zero / negative / huge /
tiny variances.  Windows of 2 over crafted pairs make the variance take chosen values."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sdc_b200.rolling import rolling_device
rng = np.random.default_rng(0)
n = 20_000_000
for scale in (1.0, 1e-160, 1e150, 1e-3):
    x = torch.from_numpy(rng.standard_normal(n) * scale).cuda()
    for win in (2, 5, 100):
        var = rolling_device(x, "var", win, win)
        std = rolling_device(x, "std", win, win)
        want = torch.sqrt(var)
        both_nan = torch.isnan(std) & torch.isnan(want)
        ulp = (std.view(torch.int64) - want.view(torch.int64)).abs()
        ulp[both_nan] = 0
        bad = (torch.isnan(std) != torch.isnan(want)).sum().item()
        print("scale %g win %d: max ulp %d, rows off by >0 ulp %.4f %%, nan mismatches %d, zeros %d" % (
            scale, win, int(ulp.max()), 100.0 * float((ulp > 0).float().mean()), bad, int((want == 0).sum())), flush=True)
        assert bad == 0 and int(ulp.max()) <= 1
# constant runs: variance exactly 0 (and tiny negative values from cancellation -> NaN like sqrt)
x = torch.from_numpy(np.repeat(rng.standard_normal(n // 1000), 1000)).cuda()
var = rolling_device(x, "var", 100, 100)
std = rolling_device(x, "std", 100, 100)
want = torch.sqrt(var)
assert bool((torch.isnan(std) == torch.isnan(want)).all())
m = ~torch.isnan(want)
assert int((std[m].view(torch.int64) - want[m].view(torch.int64)).abs().max()) <= 1
print("constant runs: zeros", int((want == 0).sum()), "negative variances", int((var < 0).sum()), "ok")
