import sys, numpy as np, torch
sys.path.insert(0, '.')
from sdc_b200.rolling import rolling_device
x = torch.from_numpy(np.random.default_rng(1).random(100_000_000)).cuda()
def t(op, reps=15):
    for _ in range(3): rolling_device(x, op, 100, 100)
    torch.cuda.synchronize()
    ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); rolling_device(x, op, 100, 100); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))
for op in ("mean","sum","count","min","max","sum","mean","var","std"):
    print(op, t(op), flush=True)
# back-to-back launches without per-call sync (like bench.py steps)
for op in ("sum","mean","count","min"):
    for _ in range(3): rolling_device(x, op, 100, 100)
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): rolling_device(x, op, 100, 100)
    b.record(); torch.cuda.synchronize()
    print("b2b", op, a.elapsed_time(b)/20, flush=True)
