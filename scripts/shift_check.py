"""Experiment: does a row-shifted UMMA descriptor (start += s*rowbytes) address the rows s.. of a
128B-swizzled K-major tile, and which base_offset encoding does it need?"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import torch
    from ugaitnet_b200 import ops
    s, bo = int(sys.argv[1]), int(sys.argv[2])
    ctx = ops.get_ctx(0)
    torch.manual_seed(0)
    A = torch.randn(128, 64, device="cuda").to(torch.bfloat16)
    B = torch.randn(64, 64, device="cuda").to(torch.bfloat16)
    C = torch.zeros(128, 64, device="cuda")
    ops.gemm_bf16(ctx, A.unsqueeze(0).contiguous(), 0, B.unsqueeze(0).contiguous(), 0, C)
    ctx.check()
    ref = A.double() @ B.double().t()
    n = 128 - s
    err = float((C[:n].double() - ref[s:s + n]).abs().max())
    print(f"shift={s} base_offset_mode={bo}: max abs err on rows [0,{n}) = {err:.3e}")
else:
    for bo in (0, 1, 2):
        for s in (0, 1, 2, 3, 7, 8, 9, 60):
            env = dict(os.environ, UGN_DBG_SHIFT=str(s), UGN_DBG_BASEOFF=str(bo))
            r = subprocess.run([sys.executable, __file__, str(s), str(bo)], env=env, capture_output=True, text=True)
            print((r.stdout.strip().splitlines() or [r.stderr.strip()[-300:]])[-1])
