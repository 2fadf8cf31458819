import sys, os, torch, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
from oracle import ugait_oracle as O
import test_step_gpu as T
import torch.nn.functional as F
from ugaitnet_b200 import ops
name = sys.argv[1] if len(sys.argv)>1 else "real_shapes"
oc, eng, P, xs, fl, lab, masks, cmask = T.setup(name)
res, G = T.oracle_step(oc, P, xs, fl, lab, masks, cmask)
out = eng.loss_and_grad(*T.engine_inputs(xs, fl, lab, masks, cmask))
grads = eng.export_grads()
for k,g in G.items():
    ref = g - T.reg_grad(oc,k,P[k])
    d = (grads[k].double().cpu()-ref)
    print(f"{k:28s} rel {float(d.norm()/ref.norm().clamp_min(1e-30)):.2e}  |ref| {float(ref.norm()):.3e} maxabs {float(d.abs().max()):.2e}")
# standalone wgrad at gray conv0 shape
ctx = ops.get_ctx(0)
g = torch.Generator().manual_seed(0)
B,C,H,Co,k = 8,25,60,96,7
x = torch.rand(B,C,H,H,generator=g)-0.5
dz = torch.randn(B,Co,54,54,generator=g)*(torch.rand(B,Co,54,54,generator=g)<0.2)
xd = torch.zeros(B,H,H,C,device='cuda'); ops.pack_input(ctx,x.cuda(),xd)
dw = torch.zeros(Co,k,k,C,device='cuda'); db=torch.zeros(Co,device='cuda')
ops.conv2d_wgrad(ctx, xd, dz.permute(0,2,3,1).contiguous().cuda(), dw, db)
w64 = torch.zeros(Co,C,k,k,dtype=torch.float64,requires_grad=True)
y = F.conv2d(x.double(), w64); y.backward(dz.double())
d = dw.permute(0,3,1,2).double().cpu()-w64.grad
print("standalone wgrad rel", float(d.norm()/w64.grad.norm()))
