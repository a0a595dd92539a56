import numpy as np, torch, sys
sys.path.insert(0, '.')
from sdc_b200.join import join_device
r = torch.from_numpy(np.random.default_rng(2).permutation(10_000_000).astype(np.int64)).cuda() * 1000003
l = torch.from_numpy(np.random.default_rng(3).integers(0, 10_000_000, 100_000_000, dtype=np.int64)).cuda() * 1000003
for _ in range(2): o=join_device(l,r,"inner")
torch.cuda.synchronize()
