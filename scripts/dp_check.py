"""Data-parallel check of both engines (run under torchrun, N >= 2): one SGD step (lr 1, no momentum) must move
every parameter by minus the MEAN over ranks of the local gradients, identically on every rank, eager and with the
segmented CUDA graphs, for every gradient-exchange schedule.  Prints one JSON line from rank 0."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from ugaitnet_b200.config import MERGE_SIGNMAX, GaitSetConfig  # noqa: E402
from ugaitnet_b200.gaitset import GaitSetEngine  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
pg = torch.distributed.group.WORLD
cfg = GaitSetConfig(in_channels=(2, 1, 1), frames=5, hw=28, nc=0, nclasses=20, merge=MERGE_SIGNMAX, wver=1.0, wid=0.1)
B = 8
g = torch.Generator(device="cuda").manual_seed(100 + rank)
xs = [torch.randn(B, 5, 28, 28, c, device="cuda", generator=g) * 0.3 for c in cfg.in_channels]
fl = [torch.ones(B, 1, device="cuda") for _ in cfg.in_channels]
fl[rank % 3][1] = 0
lab = (torch.arange(B, device="cuda") // 2).int()
from ugaitnet_b200.config import NetConfig  # noqa: E402
from ugaitnet_b200.net import UGaitEngine  # noqa: E402

scfg = NetConfig(in_channels=(50, 25, 25), filters_numbers=(32, 64, 64, 64), nd=128, nclasses=20, merge=MERGE_SIGNMAX,
                 wver=1.0, wid=0.1, weight_decay=0.0)
sxs = [torch.randn(B, c, 60, 60, device="cuda", generator=g) * 0.3 for c in scfg.in_channels]
res = {"world": world}
cases = [("gaitset", GaitSetEngine, cfg, xs, s) for s in ("split", "fused")] + \
        [("stacked", UGaitEngine, scfg, sxs, s) for s in ("fused", "split", "single", "bucketed")]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[4] in sys.argv[1:]]
for name, cls, c, x, sched in cases:
    for mode in ("fp32", "f16mix"):
        for graph in (False, True):
            ref = cls(c, math_mode=mode, seed=5, optimizer="sgd", lr=1.0, momentum=0.0)
            ref.loss_and_grad(x, fl, lab)
            gmean = ref.g.clone()
            torch.distributed.all_reduce(gmean)
            gmean /= world
            os.environ["UGN_DP_REDUCE"] = sched      # read at construction: "fused" allocates symmetric-memory arenas
            eng = cls(c, math_mode=mode, seed=5, optimizer="sgd", lr=1.0, momentum=0.0, process_group=pg, use_graph=graph)
            w0 = eng.w.clone()
            for _ in range(2 if graph else 1):        # graph mode: the second call is the first pure replay
                eng.w.copy_(w0)
                eng.repack_weights()
                eng.train_step(x, fl, lab)
            torch.cuda.synchronize()
            eng.sync_master_weights()
            delta = eng.w - w0
            l2 = torch.zeros_like(w0)                       # Keras L2 regulariser gradient 2*l2*w, applied in the optimiser
            for sg in eng.seg_list:
                l2[sg.off:sg.off + sg.n] = sg.l2
            want = gmean + 2.0 * l2 * w0
            err = float((delta + want).norm() / want.norm())
            wsum = eng.w.double().sum().reshape(1)
            allw = [torch.zeros_like(wsum) for _ in range(world)]
            torch.distributed.all_gather(allw, wsum)
            same = all(float(a) == float(allw[0]) for a in allw)
            res[f"{name}_{sched}_{mode}_{'graph' if graph else 'eager'}"] = {"rel_err_vs_mean_grad": err, "identical": same}
            assert err < (1e-5 if mode == "fp32" else 2e-2) and same, res
            del ref, eng
# Adam: three steps with the fused exchange (sharded optimiser state) against the NCCL all-reduce schedule
W = {}
for sched in ("split", "fused"):
    os.environ["UGN_DP_REDUCE"] = sched
    eng = UGaitEngine(scfg, math_mode="fp32", seed=5, optimizer="adam", lr=1e-3, process_group=pg, use_graph=True)
    assert eng.dp_reduce == sched
    for _ in range(3):
        out = eng.train_step(sxs, fl, lab)
    torch.cuda.synchronize()
    eng.sync_master_weights()
    W[sched] = (eng.w.clone(), float(out["reg"]))
    if sched == "fused":
        res["multicast_ptrs_in_use"] = bool(all(getattr(eng, "_mc", (0, 0))))
    del eng
d = (W["fused"][0] - W["split"][0]).norm() / (W["split"][0] - UGaitEngine(scfg, math_mode="fp32", seed=5).w).norm()
res["adam_3steps_fused_vs_split_update_rel_diff"] = float(d)
res["multicast"] = os.environ.get("UGN_DP_MULTIMEM", "1") != "0"
res["adam_reg_fused_vs_split"] = [W["fused"][1], W["split"][1]]
assert float(d) < 2e-2 and abs(W["fused"][1] - W["split"][1]) <= 1e-4 * abs(W["split"][1]), res
if rank == 0:
    print(json.dumps(res), flush=True)
torch.distributed.destroy_process_group()
