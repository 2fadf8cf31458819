"""Per-layer timing of the tensor-core conv kernels at the cfg2 shapes (B=96), CUDA events."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ugaitnet_b200 import ops

ctx = ops.get_ctx(0)
B = int(os.environ.get("B", "96"))
P = int(os.environ.get("P", "2"))          # planes of activations / weights
PB = int(os.environ.get("PB", "1"))        # planes of the gradient operand dz (1 = single-pass backward)
DT = torch.bfloat16 if os.environ.get("DT") == "bf16" else torch.float16
layers = [("conv1-gray", 32, 25, 60, 96, 7, True), ("conv1-of", 64, 50, 60, 96, 7, True),
          ("conv2", 96, 96, 27, 192, 5, True), ("conv3", 192, 192, 11, 512, 3, True),
          ("conv4", 512, 512, 4, 512, 2, False)]
which = sys.argv[1:] or ["fwd", "wgrad", "dgrad"]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


tot = {k: 0.0 for k in which}
for name, Cp, C, H, Co, k, pool in layers:
    Ho = H - k + 1
    Hp = Ho // 2 if pool else Ho
    x = (torch.randn(P, B, H, H, Cp, device="cuda") * 0.5).to(DT)
    w = (torch.randn(P, Co, k, k, Cp, device="cuda") * 0.05).to(DT)
    b = torch.zeros(Co, device="cuda")
    y = torch.zeros(P, B, Hp, Hp, Co, dtype=DT, device="cuda")
    idx = torch.zeros(B, Hp, Hp, Co, dtype=torch.uint8, device="cuda") if pool else None
    dz = (torch.randn(PB, B, Ho, Ho, Co, device="cuda") * 0.1).to(DT)
    dw = torch.zeros(Co, k, k, C, device="cuda")
    db = torch.zeros(Co, device="cuda")
    dx = torch.zeros(B, H, H, Cp, device="cuda")
    flops = 2.0 * B * Ho * Ho * Co * k * k * C
    mult = 3 if name.startswith("conv1") else 1        # three branches share conv2..4 shapes; conv1: 1 OF + 2 gray
    rep = {"conv1-gray": 2, "conv1-of": 1}.get(name, 3)
    line = f"{name:11s}"
    for op in which:
        if op == "fwd":
            us = timeit(lambda: ops.conv2d_fwd(ctx, x, w, b, y, idx, act=1, pool=pool))
        elif op == "wgrad":
            us = timeit(lambda: ops.conv2d_wgrad(ctx, x, dz, dw, db))
        else:
            if name.startswith("conv1"):
                line += f" | {op:5s}     n/a              "
                continue
            us = timeit(lambda: ops.conv2d_dgrad(ctx, dz, w, dx))
        tot[op] += us * rep
        line += f" | {op:5s} {us:8.1f} us {flops / us / 1e6:7.1f} TF/s"
    print(line)
ctx.check()
print("per-step totals (3 branches): " + ", ".join(f"{k} {v / 1e3:.2f} ms" for k, v in tot.items()))
