"""Development aid: GaitSet step timing at the reference's shapes (25 x 60 x 60 clips, 3 modalities) with a
per-C-ABI-call CUDA-event profile.  usage: python scripts/gs_bench.py [B] [mode] [steps]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from bench import OpTimer  # noqa: E402
from ugaitnet_b200.config import MERGE_SIGNMAX, GaitSetConfig  # noqa: E402
from ugaitnet_b200.gaitset import GaitSetEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
mode = sys.argv[2] if len(sys.argv) > 2 else "f16mix"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
cfg = GaitSetConfig(in_channels=(2, 1, 1), frames=25, hw=60, nc=0, nclasses=150, merge=MERGE_SIGNMAX, wver=1.0, wid=0.1)
g = torch.Generator(device="cuda").manual_seed(1)
xs = [torch.randn(B, 25, 60, 60, c, device="cuda", generator=g) * 0.3 for c in cfg.in_channels]
fl = [torch.ones(B, 1, device="cuda") for _ in cfg.in_channels]
for i in range(B):
    if i % 4:
        fl[i % 3][i] = 0
lab = (torch.arange(B, device="cuda") // 8).int()


def run(eng, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = eng.train_step(xs, fl, lab)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


res = {"B": B, "mode": mode}
eng = GaitSetEngine(cfg, math_mode=mode, lr=1e-4, use_graph=False)
import os
if os.environ.get("GS_LITE"):          # runs under ncu: one warm-up step, one measured step, branches in sequence
    eng.multistream = False
    run(eng, 1)
    l0 = eng.ctx.launches
    torch.cuda.profiler.start()
    ms, _ = run(eng, 1)
    torch.cuda.profiler.stop()
    print(json.dumps({"lite": True, "B": B, "ms_per_step": ms, "launches_per_step": eng.ctx.launches - l0}), flush=True)
    sys.exit(0)
run(eng, 2)
eng.ctx.check()
ms, out = run(eng, steps)
res["eager_ms"] = ms
res["loss"] = [float(out["triplet"]), float(out["ce"])]
res["mem_GB"] = torch.cuda.max_memory_allocated() / 1e9
tm = OpTimer()
tm.install()
eng.multistream = False
run(eng, 2)
tm.uninstall()
agg = {}
for name, a, e0, e1 in tm.records:
    agg.setdefault(name, []).append(e0.elapsed_time(e1))
res["op_ms"] = {k: round(sum(v) / 2, 3) for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))}
res["op_calls"] = {k: len(v) // 2 for k, v in agg.items()}
half = len(tm.records) // 2
order = {"ugn_conv2d_fwd": ["a1", "a2", "b1", "b2", "a3", "a4", "b3", "b4", "a5", "a6"],
         "ugn_conv2d_wgrad": ["b4", "b3", "b2", "b1", "a6", "a5", "a4", "a3", "a2", "a1"],
         "ugn_conv2d_dgrad": ["b4", "b3", "b2", "b1", "a6", "a5", "a4", "a3", "a2"],
         "ugn_conv2d_bwd_act": ["b4", "b3", "b2", "b1", "a6", "a5", "a4", "a3", "a2", "a1"]}
for opn, names in order.items():
    calls = [e0.elapsed_time(e1) for name, a, e0, e1 in tm.records[half:] if name == opn]
    res["layers_" + opn[4:]] = {f"m{i // len(names)}_{names[i % len(names)]}": round(t, 3) for i, t in enumerate(calls)}
eng.multistream = True
import os
if os.environ.get("UGN_CONVP_PROF"):
    print(json.dumps(res), flush=True)
    sys.exit(0)
del eng
torch.cuda.empty_cache()
eng = GaitSetEngine(cfg, math_mode=mode, lr=1e-4, use_graph=True)
run(eng, 3)
ms, out = run(eng, steps)
res["graph_ms"] = ms
res["rows_per_s"] = B / ms * 1e3
# algorithmic conv FLOPs per row and modality: forward 25 frames x (a1..a6) + global branch (b1..b4)
Hs = 64
fr = lambda h, cin, co, k: 2.0 * h * h * cin * k * k * co
per_mod = lambda c: (25 * (fr(64, c, 32, 5) + fr(64, 32, 32, 3) + fr(32, 32, 64, 3) + fr(32, 64, 64, 3) + fr(16, 64, 128, 3)
                           + fr(16, 128, 128, 3)) + fr(32, 32, 64, 3) + fr(32, 64, 64, 3) + fr(16, 64, 128, 3) + fr(16, 128, 128, 3))
fwd = sum(per_mod(c) for c in cfg.in_channels)
res["fwd_gflop_per_row"] = fwd / 1e9
res["step_tflops_algorithmic"] = 3 * fwd * B / (ms * 1e-3) / 1e12
print(json.dumps(res), flush=True)
