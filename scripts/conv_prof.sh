#!/bin/bash
for v in "" "UGN_NO_CONCAT=1"; do echo "== $v"; env $v UGN_CONVP_PROF=1 python - <<'PY' 2>&1 | grep -v "^$" | awk '!seen[$0]++' | head -12
import os, sys, torch
sys.path.insert(0, os.getcwd())
from ugaitnet_b200 import ops
ctx = ops.get_ctx(0)
B = 96
for name, Cp, H, Co, k in [("conv1-gray", 32, 60, 96, 7), ("conv1-of", 64, 60, 96, 7), ("conv2", 96, 27, 192, 5), ("conv3", 192, 11, 512, 3)]:
    Ho = H - k + 1; Hp = Ho // 2
    x = (torch.randn(2, B, H, H, Cp, device="cuda") * 0.5).half()
    w = (torch.randn(2, Co, k, k, Cp, device="cuda") * 0.05).half()
    b = torch.zeros(Co, device="cuda")
    y = torch.zeros(2, B, Hp, Hp, Co, dtype=torch.float16, device="cuda")
    idx = torch.zeros(B, Hp, Hp, Co, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        ops.conv2d_fwd(ctx, x, w, b, y, idx, act=1, pool=True)
    torch.cuda.synchronize()
    print(name, file=sys.stderr)
PY
done
