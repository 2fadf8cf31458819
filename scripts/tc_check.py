"""Diagnostic sweep of the tensor-core (tcgen05/TMA) kernels against fp64 references.
Prints one line per case; never asserts (so one gpurun call yields the whole picture)."""
import os, sys, time
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ugaitnet_b200 import ops

ctx = ops.get_ctx(0)
torch.manual_seed(0)
# UGN_CHECK_DT=f16 -> fp16 planes; UGN_CHECK_MIXED=1 -> backward operands dz with ONE plane against
# two-plane activations / weights (single-pass backward); UGN_CHECK_GS=<s> -> gradient scale s (dz planes
# hold s*dz, outputs must come back unscaled); UGN_CHECK_MAG=<m> -> operand magnitude (subnormal-lo test)
DT = torch.float16 if os.environ.get("UGN_CHECK_DT") == "f16" else torch.bfloat16
MIXED = bool(int(os.environ.get("UGN_CHECK_MIXED", "0")))
GS = float(os.environ.get("UGN_CHECK_GS", "1"))
MAG = float(os.environ.get("UGN_CHECK_MAG", "1"))
if GS != 1:
    ops.grad_scale_set(ctx, GS)
print(f"dtype {DT} mixed {MIXED} grad-scale {GS} magnitude {MAG}")


def planes(x, P):
    hi = x.to(DT)
    if P == 1:
        return hi.unsqueeze(0).contiguous(), hi.double()
    lo = (x - hi.float()).to(DT)
    return torch.stack([hi, lo]).contiguous(), hi.double() + lo.double()


def rel(a, b):
    return float((a.double() - b).norm() / b.norm().clamp_min(1e-30))


def gemm_case(M, N, K, a_mn, b_mn, P):
    A = torch.randn(M, K, device="cuda") * MAG
    B = torch.randn(N, K, device="cuda") * MAG
    Ap, Ae = planes(A.t().contiguous() if a_mn else A, P)
    Bp, Be = planes(B.t().contiguous() if b_mn else B, P)
    C = torch.full((M, N), float("nan"), device="cuda")
    tag = f"gemm M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} P={P}"
    try:
        ops.gemm_bf16(ctx, Ap, a_mn, Bp, b_mn, C)
        ctx.check()
        Ar = Ae.t() if a_mn else Ae
        Br = Be.t() if b_mn else Be
        ref = Ar @ Br.t()
        if P == 2:   # the kernel drops the lo*lo term
            pass
        exact = (A.double() @ B.double().t())
        print(f"{tag}: rel {rel(C, ref):.3e} vs-fp64-inputs {rel(C, exact):.3e} nan {int(torch.isnan(C).sum())}")
    except Exception as e:
        print(f"{tag}: ERROR {e}")


def conv_case(B, C, H, Co, k, pool, P, act=1):
    tag = f"conv B={B} C={C} H={H} Co={Co} k={k} pool={pool} P={P}"
    try:
        x = torch.randn(B, C, H, H, device="cuda")
        w = torch.randn(Co, C, k, k, device="cuda") * 0.1
        b = torch.randn(Co, device="cuda") * 0.1
        Cp = (C + 31) // 32 * 32
        xd = torch.zeros(P, B, H, H, Cp, dtype=DT, device="cuda")
        ops.pack_input(ctx, x, xd)
        wp = torch.zeros(P, Co, k, k, Cp, dtype=DT, device="cuda")
        ops.pack_weight(ctx, w.permute(0, 2, 3, 1).contiguous(), wp)
        xe = xd.double().sum(0)[..., :C].permute(0, 3, 1, 2)
        we = wp.double().sum(0)[..., :C].permute(0, 3, 1, 2)
        Ho = H - k + 1
        Hp = Ho // 2 if pool else Ho
        y = torch.zeros(P, B, Hp, Hp, Co, dtype=DT, device="cuda")
        idx = torch.zeros(B, Hp, Hp, Co, dtype=torch.uint8, device="cuda") if pool else None
        ops.conv2d_fwd(ctx, xd, wp, b, y, idx, act=act, alpha=0.3, pool=pool)
        ctx.check()
        xe.requires_grad_(True); we.requires_grad_(True)
        z = F.conv2d(xe, we, b.double())
        a = F.relu(z) if act == 1 else z
        ref = F.max_pool2d(a, 2) if pool else a
        got = y.double().sum(0).permute(0, 3, 1, 2)
        print(f"{tag}: fwd rel {rel(got, ref.detach()):.3e}")
        # backward pieces with a random dz (bf16 planes)
        dzf = torch.randn(B, Ho, Ho, Co, device="cuda") * (torch.rand(B, Ho, Ho, Co, device="cuda") < 0.3)
        PB = 1 if MIXED else P
        dzp, dze = planes(dzf * GS, PB)
        dze = dze / GS
        if MIXED:   # single pass on the hi planes: the reference operands are the hi planes too
            xe = xd[0].double()[..., :C].permute(0, 3, 1, 2).requires_grad_(True)
            we = wp[0].double()[..., :C].permute(0, 3, 1, 2).requires_grad_(True)
            z = F.conv2d(xe, we, b.double())
        z.backward(dze.permute(0, 3, 1, 2))
        dw = torch.zeros(Co, k, k, C, device="cuda"); db = torch.zeros(Co, device="cuda")
        ops.conv2d_wgrad(ctx, xd, dzp, dw, db)
        ctx.check()
        print(f"{tag}: wgrad rel {rel(dw.permute(0, 3, 1, 2), we.grad):.3e} db rel {rel(db, dze.sum((0, 1, 2))):.3e}")
        if Co % 64 == 0:
            dx = torch.zeros(B, H, H, Cp, device="cuda")
            ops.conv2d_dgrad(ctx, dzp, wp, dx)
            ctx.check()
            print(f"{tag}: dgrad rel {rel(dx[..., :C].permute(0, 3, 1, 2), xe.grad):.3e} pad {float(dx[..., C:].abs().max()) if Cp > C else 0.0:.1e}")
    except Exception as e:
        print(f"{tag}: ERROR {e}")


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "gemm"):
    gemm_case(128, 128, 64, 0, 0, 1)
    gemm_case(128, 128, 256, 0, 0, 1)
    gemm_case(128, 64, 64, 0, 0, 1)
    gemm_case(96, 200, 320, 0, 0, 1)
    gemm_case(128, 128, 64, 0, 1, 1)
    gemm_case(128, 128, 64, 1, 0, 1)
    gemm_case(128, 128, 128, 1, 1, 1)
    gemm_case(300, 520, 200, 1, 1, 1)
    gemm_case(128, 128, 128, 0, 0, 2)
    gemm_case(96, 4096, 4608, 0, 0, 1)
    gemm_case(96, 4608, 4096, 0, 1, 2)
    gemm_case(4096, 4608, 96, 1, 1, 2)
if which in ("all", "conv"):
    conv_case(2, 64, 12, 64, 3, False, 1)
    conv_case(2, 64, 12, 64, 3, True, 1)
    conv_case(2, 25, 20, 96, 7, True, 1)
    conv_case(3, 50, 60, 96, 7, True, 2)
    conv_case(3, 96, 27, 192, 5, True, 1)
    conv_case(5, 192, 11, 512, 3, True, 2)
    conv_case(17, 512, 4, 512, 2, False, 1)
print("done")
