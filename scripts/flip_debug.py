"""Development aid: list the decision flips of the engine vs the fp64 oracle with their oracle windows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ugait_oracle as O
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_decisions_gpu import _cfg
from ugaitnet_b200.net import UGaitEngine
import torch.nn.functional as F

mode = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
oc = O.NetConfig(in_channels=(50, 25, 25), nd=64, nclasses=150, merge=O.MERGE_SIGNMAX, wver=1.0, wid=0.1)
xs, fl, lab = O.synth_batch(oc, base_rows=4, expand=2, seed=11)
lab = lab % 150
P = O.init_params(oc, seed=11, dtype=torch.float64)
g = torch.Generator().manual_seed(11)
for k in P:
    if k.endswith("/b"):
        P[k] = torch.randn(P[k].shape, generator=g, dtype=torch.float64) * 0.05
eng = UGaitEngine(_cfg(oc), math_mode=mode, lr=1e-4)
eng.load_params(P)
cu = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).cuda()
out = eng.loss_and_grad([cu(x) for x in xs], [cu(f) for f in fl], torch.as_tensor(lab).cuda())
B = xs[0].shape[0]
dec = eng.export_decisions(B)
p = eng._plans[(B, True)]
x64 = [torch.tensor(x, dtype=torch.float64) for x in xs]
f64 = [torch.tensor(f, dtype=torch.float64) for f in fl]
print("flags", [f.reshape(-1).tolist() for f in fl])
for m in range(3):
    bn = O.BRANCH_NAMES[m]
    h = x64[m]
    for li in range(4):
        z = F.conv2d(h, P[f"{bn}/conv{li}/w"], P[f"{bn}/conv{li}/b"])
        a = F.relu(z)
        if li < 3:
            win = O._windows2x2(a)
            idx = (win == win.max(dim=4, keepdim=True).values).to(torch.uint8).argmax(dim=4)
            sel = torch.gather(win, 4, idx.unsqueeze(4)).squeeze(4)
            eidx = dec[m][f"pool{li}"].long()
            eact = dec[m][f"act{li}"]
            ev = p.br[m].T[f"a{li + 1}"].float().sum(0).permute(0, 3, 1, 2).cpu().double()
            diff = (idx != eidx) & ((sel > 0) | eact)
            print(f"mod {m} layer {li}: pool flips {int(diff.sum())} act flips {int(((sel > 0) != eact).sum())} scale {float(sel.max()):.3f} "
                  f"value err max {float((ev - sel).abs().max()):.2e}")
            for (b, c, y, x) in diff.nonzero().tolist()[:6]:
                print(f"    b{b} c{c} y{y} x{x}: oracle win {[f'{v:.6f}' for v in win[b, c, y, x].tolist()]} oracle idx {int(idx[b,c,y,x])} "
                      f"engine idx {int(eidx[b,c,y,x])} engine val {float(ev[b,c,y,x]):.6f} flagrow {fl[m][b,0]}")
            h = sel
        else:
            eact = dec[m][f"act{li}"]
            ev = p.br[m].T[f"a{li + 1}"].float().sum(0).permute(0, 3, 1, 2).cpu().double()
            print(f"mod {m} layer {li}: act flips {int(((a > 0) != eact).sum())} value err max {float((ev - a).abs().max()):.2e}")
            h = a
