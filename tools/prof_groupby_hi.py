import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdc_b200.groupby import groupby_device
n = 125_000_000
K = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
g = torch.Generator(device="cuda").manual_seed(5)
vals = torch.rand(n, device="cuda", dtype=torch.float64, generator=g)
keys = torch.randint(0, K, (n,), device="cuda", generator=g)
for _ in range(2):
    groupby_device(keys, [vals], ("sum",), True, 1, K)
torch.cuda.synchronize()
