"""Where does compat model.predict spend its time?  (wall-clock sections, 96-row float64 batch)"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from ugaitnet_b200.compat import optimizers, sign_max
import ugaitnet_b200.compat.nets.mj_uwyhNets_ba as nets
nets.MATH_MODE = "f16mix"
model = nets.UWYHSemiNet3Mods.build_or_load([(50, 60, 60), (25, 60, 60), (25, 60, 60)], 4, [(7, 7), (5, 5), (3, 3), (2, 2)],
                                            [96, 192, 512, 512], 2048, 0.00005, 0.4, optimizer=optimizers.Adam(lr=1e-4),
                                            margin=0.2, nclasses=150, loss_weights=[1.0, 0.1], initnet="", fMerge=sign_max)
xs, fl, lab = bench.make_batch(1)
X = []
for x, f in zip(xs, fl):
    X += [x.astype(np.float64), f.astype(np.float64)]
for _ in range(3):
    model.predict(X)
torch.cuda.synchronize()
eng = model.engine
B = 96
hb = model._hbp[B]
def sec(name, fn, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); print(f"{name:30s} {(time.perf_counter()-t0)/n*1e3:8.3f} ms"); return r
sec("predict total", lambda: model.predict(X))
def cast():
    for m in range(3):
        src = torch.from_numpy(np.ascontiguousarray(X[2*m]))
        hb.t["x"][m].copy_(src.reshape(hb.t["x"][m].shape))
sec("cast f64->f32 pinned (torch)", cast)
tmp = [np.empty(x.shape, np.float32) for x in xs]
def cast_np():
    for m in range(3):
        np.copyto(tmp[m], X[2*m], casting="same_kind")
sec("cast numpy copyto pageable", cast_np)
print("torch threads", torch.get_num_threads())
sec("prefetch (H2D)", lambda: eng.prefetch_batch(hb, train=False))
def fwd():
    eng.prefetch_batch(hb, train=False)
    return eng.predict_prefetched("signature")
sig = sec("prefetch + forward", fwd)
sec("sig.cpu().numpy()", lambda: sig.cpu().numpy())
p = eng.plan(B, False)
sec("softmax + cpu", lambda: torch.softmax(p.logits, dim=1).cpu().numpy())
import os; print("cpus", os.cpu_count(), len(os.sched_getaffinity(0)))

class Gen:
    def __init__(self):
        self.items = [(X, [lab.astype(np.float64), np.eye(150)[lab.reshape(-1).astype(int) % 150]])] * 2
    def __len__(self): return 2
    def __getitem__(self, i): return self.items[i]
    def on_epoch_end(self): pass
model.fit(Gen(), epochs=1, steps_per_epoch=8, verbose=0)
torch.cuda.synchronize()
for i in range(6):
    t0 = time.perf_counter(); model.predict(X); print("after fit: predict call", i, f"{(time.perf_counter()-t0)*1e3:.2f} ms")
sec("cast after fit", cast)
sec("prefetch + forward after fit", fwd)
