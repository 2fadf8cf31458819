#!/usr/bin/env python
"""Benchmark of the UGaitNet hot path (BASELINE.json metric / configs[1]).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--mode f16mix|f16x3|bf16x3|bf16|fp32]

A "step" is one Keras train_function step of the 3-modality (OF+gray+depth) TUM-GAID-shaped model
(nd=2048, 150 classes, sign_max fusion, dropout 0.4, Adam) on one batch of bs=24 literal sequences
expanded x4 by the reference's missing-modality generator = 96 rows per GPU (weak scaling, batch
data-parallel; the gradient exchange is fused into the optimiser kernel over NVLink peer memory, NCCL
all-reduce as the fallback -- `config.dp_exchange` says which ran).  One JSON line is printed by rank 0:
the contract keys (value = device-resident rows/s, e2e = pinned-host inputs through the public API,
roofline = the conv-forward kernel, cpu_baseline = the oracle port on the host cores, clocks,
gpu_launches) plus two further legs: `knn` (open-world k = 3 search over a 1 M x 256 gallery, sharded over
the ranks) and, at N = 1, `gaitset` (the same step with the GaitSet branch type).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BS_LITERAL, EXPAND = 24, 4
ND, NCLASSES = 2048, 150
FLOP_FWD_ROW = {"of": 2.0214e9 + 0.0545e9, "c25": 1.3356e9 + 0.0545e9}
TRAIN_FLOP_ROW = 11.83e9                                                  # 3-mod fwd+dgrad+wgrad


def ncu_traffic(key, mode):
    """DRAM bytes per step of a kernel family from the LATEST committed ncu --set full capture
    (profiles/traffic.json, written by scripts/summarise_profiles.py from dram__bytes_read.sum + dram__bytes_write.sum
    of the launches of one step).  None when no capture of this math mode is recorded -- never a stale constant."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        ent = t.get(mode, {}).get(key)
        return None if ent is None else float(ent["bytes_per_step"])
    except Exception:
        return None


DTYPE_NAMES = {"fp32": "f32", "bf16": "bf16", "bf16x3": "bf16 (3-pass hi/lo split, fp32 accumulate)",
               "f16x3": "f16 (3-pass hi/lo split, fp32 accumulate)",
               "f16mix": "f16 (forward: 3-pass hi/lo split; backward: 1 pass, scaled fp16 gradients; fp32 accumulate)",
               "f16mix2": "f16 (forward: 2 passes, weights hi/lo split x activations hi; backward: 1 pass; fp32 accumulate)",
               "f16mix1": "f16 (forward and backward: 1 pass on 11-bit operands; fp32 accumulate)"}


# MMA passes issued per algorithmic product (forward, backward) in each math mode
MMA_PASSES = {"fp32": (1, 1), "bf16": (1, 1), "bf16x3": (3, 3), "f16x3": (3, 3), "f16mix": (3, 1), "f16mix2": (2, 1),
              "f16mix1": (1, 1)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None, "reasons": reasons}


def make_batch(seed):
    """Synthetic cfg2 batch: host numpy arrays in the reference generator's layout."""
    from oracle.ugait_oracle import NetConfig as OC, synth_batch
    oc = OC(in_channels=(50, 25, 25), nd=ND, nclasses=NCLASSES)
    xs, fl, lab = synth_batch(oc, base_rows=BS_LITERAL, expand=EXPAND, seed=seed)
    return xs, fl, lab


def engine_cfg():
    from ugaitnet_b200.config import MERGE_SIGNMAX, NetConfig
    return NetConfig(in_channels=(50, 25, 25), nd=ND, nc=0, nclasses=NCLASSES, weight_decay=5e-5,
                     merge=MERGE_SIGNMAX, margin=0.2, wver=1.0, wid=0.1, dropout=0.4)


# ------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path on the host cores.  TensorFlow is
    not installable (no network), so this is the oracle port (PyTorch-CPU fp32 restatement of the same
    graph, oneDNN, all host threads) -- the one other place bench.py may execute oracle/."""
    if rank != 0:
        return
    from oracle import ugait_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    oc = O.NetConfig(in_channels=(50, 25, 25), nd=ND, nclasses=NCLASSES, merge=O.MERGE_SIGNMAX, wver=1.0, wid=0.1)
    rows = BS_LITERAL * EXPAND      # the FULL 96-row step of the configuration (about 1.2 s per step on 16 cores)
    xs, fl, lab = O.synth_batch(oc, base_rows=rows // EXPAND, expand=EXPAND, seed=232323)
    xs = [torch.tensor(x) for x in xs]
    fl = [torch.tensor(f) for f in fl]
    lab = torch.tensor(lab)
    P = O.init_params(oc, seed=1)
    M = {k: torch.zeros_like(v) for k, v in P.items()}
    V = {k: torch.zeros_like(v) for k, v in P.items()}
    keep = 0.6
    g = torch.Generator().manual_seed(0)

    def step(t):
        masks = [(torch.rand(rows, 2 * ND, generator=g) < keep).float() / keep for _ in range(3)]
        _, G = O.loss_and_grads(xs, fl, lab, P, oc, masks, None)
        O.adam_step(P, G, M, V, t, lr=1e-4)

    for t in range(1, args.warmup + 1):
        step(t)
    t0 = time.perf_counter()
    for t in range(args.steps):
        step(args.warmup + t + 1)
    dt = (time.perf_counter() - t0) / args.steps
    val = rows / dt
    line = {"metric": "train rows/s (3-mod fwd+bwd+triplet+CE+Adam)", "value": val, "unit": "rows/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": workload_config(world),
            "cpu_baseline": {"value": val, "unit": "rows/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"all {rows} rows of one cfg2 step per CPU step (PyTorch-CPU fp32 restatement; TensorFlow is "
                                       f"absent from the image: profiles/r02a_tf_probe.log), {args.steps} steps"},
            "e2e": {"value": val, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(world):
    return {"workload": "cfg2: UGaitNet 3-modality (gray+OF+depth) TUM-GAID shape, missing-modality masking, "
                        "mergefun=sign_max, nd=2048, nclasses=150, bs=24 literal x expand 4 = 96 rows per GPU",
            "rows_per_gpu": BS_LITERAL * EXPAND, "literal_bs_per_gpu": BS_LITERAL, "parallelism": f"dp{world}",
            "l2": "working set (358 MB weights + 1.4 GB Adam state + activations) exceeds the 126 MB L2",
            "value_path": "UGaitEngine.train_step_resident: the batch sits in the engine's input block before the timed "
                          "region starts (no per-step input copy); e2e copies every step's inputs from pinned host memory"}


# ------------------------------------------------------------------------------------------
def cpu_baseline_leg():
    """Oracle port (PyTorch-CPU fp32) timed on the host cores on a bounded sample of the same step."""
    from oracle import ugait_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    oc = O.NetConfig(in_channels=(50, 25, 25), nd=ND, nclasses=NCLASSES, merge=O.MERGE_SIGNMAX, wver=1.0, wid=0.1)
    rows = BS_LITERAL * EXPAND
    xs, fl, lab = O.synth_batch(oc, base_rows=rows // EXPAND, expand=EXPAND, seed=232323)
    xs = [torch.tensor(x) for x in xs]
    fl = [torch.tensor(f) for f in fl]
    lab = torch.tensor(lab)
    P = O.init_params(oc, seed=1)
    M = {k: torch.zeros_like(v) for k, v in P.items()}
    V = {k: torch.zeros_like(v) for k, v in P.items()}
    times = []
    for t in range(1, 4):
        t0 = time.perf_counter()
        _, G = O.loss_and_grads(xs, fl, lab, P, oc)
        O.adam_step(P, G, M, V, t, lr=1e-4)
        times.append(time.perf_counter() - t0)
    dt = min(times[1:])
    return {"value": rows / dt, "unit": "rows/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"PyTorch-CPU fp32 restatement (TensorFlow absent: profiles/r02a_tf_probe.log), the full {rows}-row "
                      "cfg2 step incl. Adam, best of 2 after 1 warm-up"}


class OpTimer:
    """CUDA-event timing of every C-ABI call of a profiling pass (same stream as the launches)."""

    def __init__(self):
        self.records = []

    def install(self):
        from ugaitnet_b200 import _ffi
        self._orig = {}
        for name in _ffi.EXPORTED_SYMBOLS:
            fn = getattr(_ffi.lib, name)
            if fn.restype is not ctypes_int() or name in ("ugn_abi_version", "ugn_ctx_create", "ugn_ctx_destroy",
                                                          "ugn_ctx_check", "ugn_ctx_has_tcgen05"):
                continue
            self._orig[name] = fn
            setattr(_ffi.lib, name, self._wrap(name, fn))

    def _wrap(self, name, fn):
        def call(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            self.records.append((name, a, e0, e1))
            return rc
        return call

    def uninstall(self):
        from ugaitnet_b200 import _ffi
        for name, fn in self._orig.items():
            setattr(_ffi.lib, name, fn)


def ctypes_int():
    import ctypes
    return ctypes.c_int


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--mode", default=os.environ.get("UGN_BENCH_MODE", "f16mix"))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-knn", action="store_true")
    ap.add_argument("--no-gaitset", action="store_true", help="skip the GaitSet-branch leg (SURVEY 8f-1)")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg1 / cfg3 / cfg4 legs")
    ap.add_argument("--no-extract", action="store_true", help="skip the descriptor-extraction leg")
    ap.add_argument("--lite", action="store_true", help="timed steps only (for runs under ncu)")
    ap.add_argument("--knn-only", action="store_true", help="only the open-world k-NN leg (development aid)")
    ap.add_argument("--knn-d", type=int, default=256, help="descriptor width of the --knn-only leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
        pg = torch.distributed.group.WORLD

    if args.knn_only:
        kn = knn_leg(peaks(), rank, world, pg, D=args.knn_d)
        if rank == 0:
            print(json.dumps({"knn": kn}), flush=True)
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    from ugaitnet_b200.net import UGaitEngine
    eng = UGaitEngine(engine_cfg(), device=local, math_mode=args.mode, lr=1e-4, process_group=pg,
                      use_graph=not args.no_graph)
    xs, fl, lab = make_batch(232323 + rank)
    B = xs[0].shape[0]
    # host copies in pinned memory (e2e leg) and device-resident copies (value leg)
    hx = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in xs]
    hf = [torch.from_numpy(np.ascontiguousarray(f)).pin_memory() for f in fl]
    hl = torch.from_numpy(np.ascontiguousarray(lab.reshape(-1).astype(np.int32))).pin_memory()
    dx = [t.cuda(non_blocking=True) for t in hx]
    df = [t.cuda(non_blocking=True) for t in hf]
    dl = hl.cuda(non_blocking=True)
    h2d = sum(t.numel() * t.element_size() for t in hx + hf) + hl.numel() * 4

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t)
        return ms / steps

    # value path: the batch is resident in the engine's own input block (UGaitEngine.input_buffers) before the timed
    # region starts; a step is then exactly the hot path -- no per-step input copy of any kind
    xi, fi, li = eng.input_buffers(B)
    for m in range(3):
        xi[m].copy_(dx[m])
        fi[m].copy_(df[m].reshape(-1, 1))
    li.copy_(dl.reshape(-1).to(torch.int32))
    step_dev = lambda: eng.train_step_resident(B)
    loss_host = torch.zeros(8).pin_memory()

    def read_losses(out):
        loss_host.copy_(out["losses"], non_blocking=True)      # {triplet, count, ce, acc, reg}: ONE D2H read
        torch.cuda.current_stream().synchronize()                # the user reads the loss every step

    def step_e2e():
        read_losses(eng.train_step(hx, hf, hl))                  # 7 H2D copies of this step's inputs happen inside

    for _ in range(args.warmup):
        step_dev()
    l0 = eng.ctx.launches
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(step_dev, args.steps)
    sampler.stop_flag = True
    launches = eng.ctx.launches - l0
    if eng.use_graph:
        launches = eng.graph_launches * args.steps
    if args.lite:
        if rank == 0:
            print(json.dumps({"lite": True, "ms_per_step": ms, "rows_per_s": B * world / (ms * 1e-3),
                              "gpu_launches": int(launches)}), flush=True)
        return
    for _ in range(2):
        step_e2e()
    ms_e2e_serial = timed(step_e2e, args.steps)

    # Single-copy public API: the loader fills a pinned HostBatch (byte image of the engine's input block) in place;
    # prefetch_batch() moves it with ONE cudaMemcpyAsync on a copy stream while the previous step runs, and the step
    # reads its losses back with ONE D2H copy.  Every step still copies its own inputs from host memory and reads its
    # loss inside the timed region.
    def fill(hb, base):
        for m in range(3):
            hb.inputs[m][...] = xs[m][::EXPAND] if base else xs[m]
            hb.flags[m][...] = fl[m]
        hb.labels[...] = lab.reshape(-1).astype(np.int32)
        if base:
            hb.src_row[...] = np.repeat(np.arange(BS_LITERAL, dtype=np.int32), EXPAND)

    def pipelined(hbs):
        k = [0]

        def step():
            out = eng.train_step_prefetched()
            k[0] ^= 1
            eng.prefetch_batch(hbs[k[0]])                        # H2D of step i+1 overlaps step i
            read_losses(out)
        eng.prefetch_batch(hbs[0])
        for _ in range(3):
            step()
        return timed(step, args.steps)

    hb_full = [eng.host_batch(B), eng.host_batch(B)]
    for h in hb_full:
        fill(h, False)
    ms_e2e_full = pipelined(hb_full)
    # device-side expansion: only the 24 complete sequences + the expansion tables cross PCIe, the
    # E-fold batch with its missing-modality pattern is built inside the input pack -- same step, a quarter of the bytes
    hb_exp = [eng.host_batch(B, base_rows=BS_LITERAL), eng.host_batch(B, base_rows=BS_LITERAL)]
    for h in hb_exp:
        fill(h, True)
    ms_e2e = pipelined(hb_exp)
    h2d_exp, h2d_full = hb_exp[0].nbytes, hb_full[0].nbytes
    eng.ctx.check()

    # descriptor extraction (model_code.predict in batches of --bs 64,
    # mains/mj_testUWYHGaitNet_open_tum.py:139-148): forward to "signature" only; every rank extracts its own clips
    extract = {}
    if not args.no_extract:
        desc_host = {}
        for Bx in (64, 512):
            reps = (Bx + B - 1) // B
            ex = [torch.cat([t] * reps)[:Bx].contiguous() for t in dx]
            ef = [torch.cat([t] * reps)[:Bx].contiguous() for t in df]
            for _ in range(3):
                eng.predict(ex, ef)
            ms_x = timed(lambda: eng.predict(ex, ef), 10)
            hbx = [eng.host_batch(Bx, train=False), eng.host_batch(Bx, train=False)]
            for h in hbx:
                for m in range(3):
                    h.inputs[m][...] = ex[m].cpu().numpy()
                    h.flags[m][...] = ef[m].cpu().numpy()
            desc_host[Bx] = torch.zeros(Bx, 2048).pin_memory()
            kx = [0]

            def step_x():
                sig = eng.predict_prefetched("signature")
                kx[0] ^= 1
                eng.prefetch_batch(hbx[kx[0]], train=False)          # next batch's H2D overlaps this forward pass
                desc_host[Bx].copy_(sig, non_blocking=True)          # the descriptors are what the caller keeps
                torch.cuda.current_stream().synchronize()
            eng.prefetch_batch(hbx[0], train=False)
            for _ in range(3):
                step_x()
            ms_xe = timed(step_x, 10)
            # the same from the STORED sample integers (int16 flow / 100 * 0.1, uint8 gray & depth / 255 - 0.5,
            # data/mj_dataGeneratorMMUWYHsingle.py:313-329), decoded on the device: 0.375 of the bytes cross PCIe
            from ugaitnet_b200 import samples
            nb_f32 = int(hbx[0].nbytes)
            del hbx
            hbx = [eng.host_batch(Bx, train=False, raw=[samples.RAW_FLOW, samples.RAW_GRAY, samples.RAW_GRAY])
                   for _ in range(2)]
            for h in hbx:
                h.inputs[0][...] = np.clip(np.rint(ex[0].cpu().numpy() * 1000.0), -32768, 32767).astype(np.int16)
                for m in (1, 2):
                    h.inputs[m][...] = np.clip(np.rint((ex[m].cpu().numpy() + 0.5) * 255.0), 0, 255).astype(np.uint8)
                    h.flags[m][...] = ef[m].cpu().numpy()
                h.flags[0][...] = ef[0].cpu().numpy()
            eng.prefetch_batch(hbx[0], train=False)
            for _ in range(3):
                step_x()
            ms_xr = timed(step_x, 10)
            extract[f"B{Bx}"] = {"value": Bx * world / (ms_x * 1e-3), "unit": "rows/s", "ms_per_batch": ms_x,
                                 "model_tflops": 4.857e9 * Bx * world / (ms_x * 1e-3) / 1e12,
                                 "e2e": {"value": Bx * world / (ms_xe * 1e-3), "ms_per_batch": ms_xe,
                                         "h2d_bytes_per_batch": nb_f32, "d2h_bytes_per_batch": Bx * 2048 * 4},
                                 "e2e_raw_samples": {"value": Bx * world / (ms_xr * 1e-3), "ms_per_batch": ms_xr,
                                                     "h2d_bytes_per_batch": int(hbx[0].nbytes),
                                                     "d2h_bytes_per_batch": Bx * 2048 * 4}}
            del ex, ef, hbx
        extract["api"] = ("UGaitEngine.predict(device tensors) | host_batch(train=False[, raw=stored int16 / uint8 samples, "
                          "decoded on the device]) + prefetch_batch + predict_prefetched + D2H of the [B, 2048] "
                          "descriptors; every rank extracts its own clips")
        torch.cuda.empty_cache()
    # data-parallel sanity inside the bench itself: finite losses and bit-identical weights on every rank after the run
    assert bool(torch.isfinite(loss_host[[0, 2, 4]]).all()), f"non-finite loss on rank {rank}: {loss_host.tolist()}"
    rank_check = None
    if world > 1:
        eng.sync_master_weights()      # (owners hold the f32 masters of their arena slices; compute copies are exchanged)
        csum = torch.stack([t.float().double().sum() for t in eng.cw.values()]).sum()
        wsum = torch.stack([eng.w.double().sum(), eng.w.double().abs().sum(), csum])
        allw = [torch.zeros_like(wsum) for _ in range(world)]
        torch.distributed.all_gather(allw, wsum)
        rank_check = {"weights_identical_on_all_ranks": all(torch.equal(a, allw[0]) for a in allw),
                      "loss_finite": True}
        assert rank_check["weights_identical_on_all_ranks"], "ranks diverged"
    rows_total = B * world
    value = rows_total / (ms * 1e-3)
    e2e = rows_total / (ms_e2e * 1e-3)

    cfg_line = workload_config(world)
    if world > 1:
        cfg_line["rank_check"] = rank_check
        if eng.dp_timing_summary() is not None:
            cfg_line["dp_timing"] = eng.dp_timing_summary()
        cfg_line["dp_exchange"] = eng.dp_reduce + (" (NVSwitch multimem)" if all(getattr(eng, "_mc", (0, 0))) else "") + (
            " + deferred copy-engine weight all-gather" if getattr(eng, "_cw_mc", 0) == -1 else "")
    line = {"metric": "train rows/s (3-mod fwd+bwd+triplet+CE+Adam)", "value": value, "unit": "rows/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE_NAMES[args.mode],
            "data": "synthetic", "config": cfg_line,
            "literal_seq_per_s": value / EXPAND,
            "e2e": {"value": e2e, "unit": "rows/s", "h2d_bytes_per_step": int(h2d_exp), "d2h_bytes_per_step": 32,
                    "ms_per_step": ms_e2e, "copies_per_step": {"h2d": 1, "d2h": 1},
                    "api": "UGaitEngine.host_batch(B, base_rows) + prefetch_batch + train_step_prefetched: the 24 complete "
                           "sequences + expansion tables in ONE pinned block / ONE cudaMemcpyAsync (step i+1's copy overlaps "
                           "step i), device-side missing-modality expansion, ONE packed loss D2H",
                    "full_batch": {"value": rows_total / (ms_e2e_full * 1e-3), "ms_per_step": ms_e2e_full,
                                   "h2d_bytes_per_step": int(h2d_full),
                                   "api": "same, host-expanded 96-row batch (generator layout) in one pinned block"},
                    "serial_7_copies": {"value": rows_total / (ms_e2e_serial * 1e-3), "ms_per_step": ms_e2e_serial,
                                        "h2d_bytes_per_step": int(h2d),
                                        "api": "UGaitEngine.train_step(pinned host tensors): 7 H2D copies, not overlapped"}},
            "gpu_launches": int(launches), "gpu_launches_per_step": int(launches) // max(1, args.steps),
            "clocks": sampler.summary(),
            "model_tflops": TRAIN_FLOP_ROW * rows_total / (ms * 1e-3) / 1e12}
    if extract:
        line["extract"] = extract

    if rank == 0:
        # ---- per-op profile pass (eager, CUDA events on the launch stream) -> dominant kernel roofline
        pk = peaks()
        eager = UGaitEngine(engine_cfg(), device=local, math_mode=args.mode, lr=1e-4, use_graph=False) if eng.use_graph else eng
        if eager is not eng:
            eager.w.copy_(eng.w)
            eager.repack_weights()
        # per-op timing needs the branches in sequence: with one stream per branch the CUDA-event interval of a call
        # includes the time its kernels wait for SMs held by the other branches
        eager.multistream = False
        for _ in range(2):
            eager.train_step(dx, df, dl) if world == 1 else None
        if world == 1:
            tm = OpTimer()
            tm.install()
            nprof = 3
            for _ in range(nprof):
                eager.train_step(dx, df, dl)
            torch.cuda.synchronize()
            tm.uninstall()
            agg = {}
            for name, a, e0, e1 in tm.records:
                agg.setdefault(name, []).append(e0.elapsed_time(e1))
            per_step = {k: sum(v) / nprof for k, v in agg.items()}
            total = sum(per_step.values())
            line["op_ms_per_step"] = {k: round(v, 4) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1])}
            conv_ops = ("ugn_conv2d_fwd", "ugn_conv2d_dgrad", "ugn_conv2d_wgrad")
            conv_ms = sum(per_step.get(k, 0.0) for k in conv_ops)
            # algorithmic conv FLOPs of one step: fwd + wgrad over all layers, dgrad without conv1
            conv_fwd_row = 2.0214e9 + 2 * 1.3356e9
            conv_flops = (3 * conv_fwd_row - (1.3717e9 + 2 * 0.6858e9)) * B
            # MMA passes issued per algorithmic product (forward, backward)
            pf, pb = MMA_PASSES[args.mode]
            fwd_flops = conv_fwd_row * B
            passes = (pf * fwd_flops + pb * (conv_flops - fwd_flops)) / conv_flops
            ach = conv_flops / (conv_ms * 1e-3) / 1e12
            # dominant kernel: tc_convp_kernel = the conv FORWARD launches (ugn_conv2d_fwd, 12 per step): algorithmic
            # FLOPs of the forward convolutions / their CUDA-event time on the launch stream.  traffic = ncu
            # dram__bytes_read+write summed over the 12 launches of one step (profiles/, null until captured).
            fwd_ms = per_step.get("ugn_conv2d_fwd", 0.0)
            ach_f = fwd_flops / (fwd_ms * 1e-3) / 1e12 if fwd_ms else 0.0
            line["roofline"] = {"bound": "tensor", "kernel": "tc_convp_kernel (conv forward, 12 launches/step)",
                                "achieved": ach_f, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach_f / pk["tf_sust"],
                                "traffic": ncu_traffic("conv_fwd", args.mode),
                                "algorithmic_flop_per_step": fwd_flops, "launch_ms_per_step": fwd_ms,
                                "share_of_step": fwd_ms / total if total else None,
                                "mma_passes": pf, "issued_frac": ach_f * pf / pk["tf_sust"],
                                "note": "N=96 tiles (Cout=96) cost 73 clk per MMA vs 48 ideal (profiles/r01_umma_rate.txt); "
                                        "channel padding 25->32 / 50->64 and the 64-slot row pitch add x1.5 issued work on conv1",
                                "peak_source": pk["src"] + " bf16 sustained",
                                "all_conv": {"kernels": "conv fwd + dgrad + wgrad", "achieved": ach, "frac": ach / pk["tf_sust"],
                                             "mma_passes": passes, "issued_frac": ach * passes / pk["tf_sust"],
                                             "share_of_step": conv_ms / total if total else None}}
            line["cpu_baseline"] = cpu_baseline_leg()
        if not args.no_configs and world == 1:
            del eager
            torch.cuda.empty_cache()
            line["configs"] = config_legs(pk, args)
            line["e2e"]["compat_fit"] = compat_leg(args)
            torch.cuda.empty_cache()
            eager = None
        if not args.no_knn and world == 1:
            line["knn"] = knn_leg(pk)
            torch.cuda.empty_cache()
            line["knn_d2048"] = knn_leg(pk, D=2048)
            torch.cuda.empty_cache()
        if not args.no_gaitset and world == 1:
            del eng
            eager = None
            torch.cuda.empty_cache()
            line["gaitset"] = gaitset_leg(pk, args)
    if world > 1 and not args.no_knn:
        kn = knn_leg(peaks(), rank, world, pg)          # collective: every rank searches its gallery shard
        torch.cuda.empty_cache()
        kn2 = knn_leg(peaks(), rank, world, pg, D=2048)
        if rank == 0:
            line["knn"], line["knn_d2048"] = kn, kn2
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def compat_leg(args):
    """End to end through the DROP-IN API exactly as mains/mj_trainUWYHGaitNet_DataGen_3mods.py drives it: the model
    from UWYHSemiNet3Mods.build_or_load (reference signature), model.fit on a generator that yields the reference
    generator's batches -- float64 numpy volumes + flags, [labels, one-hot] (data/mj_dataGeneratorMMUWYHsingle.py:664-823)
    -- and model.predict on the same.  Wall-clock with the f64 -> f32 cast, the H2D copies and the per-batch loss
    read-back inside the timed region."""
    from ugaitnet_b200.compat import optimizers, sign_max
    import ugaitnet_b200.compat.nets.mj_uwyhNets_ba as nets
    nets.MATH_MODE = args.mode
    model = nets.UWYHSemiNet3Mods.build_or_load([(50, 60, 60), (25, 60, 60), (25, 60, 60)], 4,
                                                [(7, 7), (5, 5), (3, 3), (2, 2)], [96, 192, 512, 512], ND, 0.00005, 0.4,
                                                optimizer=optimizers.Adam(lr=1e-4), margin=0.2, nclasses=NCLASSES,
                                                loss_weights=[1.0, 0.1], initnet="", fMerge=sign_max)

    class Gen:
        def __init__(self):
            self.items = []
            for i in range(2):
                xs, fl, lab = make_batch(232323 + i)
                X = []
                for x, f in zip(xs, fl):
                    X += [x.astype(np.float64), f.astype(np.float64)]
                self.items.append((X, [lab.astype(np.float64), np.eye(NCLASSES)[lab.reshape(-1).astype(int) % NCLASSES]]))

        def __len__(self):
            return len(self.items)

        def __getitem__(self, i):
            return self.items[i]

        def on_epoch_end(self):
            pass

    gen = Gen()
    B = gen[0][0][0].shape[0]
    f64_bytes = sum(a.nbytes for a in gen[0][0])
    model.fit(gen, epochs=1, steps_per_epoch=4, verbose=0)           # warm-up: plans, graph capture, pinned blocks
    torch.cuda.synchronize()
    n = max(8, args.steps)
    t0 = time.perf_counter()
    hist = model.fit(gen, epochs=1, steps_per_epoch=n, verbose=0)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    X = gen[0][0]
    model.predict(X)
    torch.cuda.synchronize()
    calls = []
    for _ in range(9):
        t0 = time.perf_counter()
        sig, prob = model.predict(X)
        calls.append(time.perf_counter() - t0)
    dtp = sorted(calls)[len(calls) // 2]                             # median call (each call is synchronous)
    assert np.isfinite(hist.history["loss"][-1])
    return {"api": "compat UWYHSemiNet3Mods.build_or_load + model.fit(generator of float64 numpy batches) / model.predict",
            "math_mode": model.engine.math_mode, "rows_per_step": B,
            "fit": {"value": B / dt, "unit": "rows/s", "ms_per_step": dt * 1e3, "host_f64_bytes_per_step": int(f64_bytes),
                    "h2d_bytes_per_step": int(f64_bytes // 2), "d2h_bytes_per_step": 32,
                    "path": "f64 -> f32 cast across the host cores straight into ONE pinned input block whose prefix "
                            "pieces are copied as they are produced (the H2D of modality m under the cast of m+1), "
                            "cast + copy of batch i+1 overlapping step i, ONE packed loss read per batch"},
            "predict": {"value": B / dtp, "unit": "rows/s", "ms_per_call": dtp * 1e3,
                        "ms_calls": [round(c * 1e3, 2) for c in calls],
                        "returns": "[signature [B,2048], classprob [B,150]] as numpy"}}


def config_legs(pk, args):
    """BASELINE.json configs 1, 3 and 4 (the other training configurations; cfg2 is the headline above): device-resident
    CUDA-graph replay of the full step on one GPU, same engine and math mode.  Parity of exactly these configurations
    against the fp64 oracle: tests/test_decisions_gpu.py::test_full_size_config_against_oracle."""
    from oracle.ugait_oracle import NetConfig as OC, synth_batch
    from ugaitnet_b200.config import MERGE_SIGNMAX, NetConfig
    from ugaitnet_b200.net import UGaitEngine
    legs = {
        "cfg1": dict(desc="1-modality gray, casenet D, nclasses 150, bs 24 (no expansion)", inch=(25,), ncls=150,
                     batch=dict(base_rows=24, expand=1, kinds=("gray",)), single=True, flop_row=3.48e9),
        "cfg3": dict(desc="CASIA-B shape (gray+OF+silhouette), nclasses 74, bs 40 x expand 3 = 120 rows", inch=(50, 25, 25),
                     ncls=74, batch=dict(base_rows=40, expand=3, ids_per=10, kinds=("of", "gray", "sil")), single=False,
                     flop_row=TRAIN_FLOP_ROW),
        "cfg4": dict(desc="early-fusion BL-all (--nomissing), bs 512, every modality present", inch=(50, 25, 25), ncls=150,
                     batch=dict(base_rows=512, expand=1, ids_per=2), single=False, flop_row=TRAIN_FLOP_ROW),
    }
    out = {}
    for name, c in legs.items():
        oc = OC(in_channels=c["inch"], nd=ND, nclasses=c["ncls"], single=c["single"])
        xs, fl, lab = synth_batch(oc, seed=232323, **c["batch"])
        cfg = NetConfig(in_channels=c["inch"], nd=ND, nc=0, nclasses=c["ncls"], weight_decay=5e-5, merge=MERGE_SIGNMAX,
                        margin=0.2, wver=1.0, wid=0.1, dropout=0.4, single=c["single"])
        eng = UGaitEngine(cfg, math_mode=args.mode, lr=1e-4, use_graph=not args.no_graph)
        dx = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in xs]
        df = [torch.from_numpy(np.ascontiguousarray(f)).cuda() for f in fl]
        dl = torch.from_numpy(np.ascontiguousarray((lab % c["ncls"]).reshape(-1).astype(np.int32))).cuda()
        B = dx[0].shape[0]
        xi, fi, li = eng.input_buffers(B)                 # resident batch, as in the headline leg
        for m in range(len(dx)):
            xi[m].copy_(dx[m])
            if not c["single"]:
                fi[m].copy_(df[m].reshape(-1, 1))
        li.copy_(dl)
        for _ in range(3):
            o = eng.train_step_resident(B)
        steps = max(5, args.steps // 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            o = eng.train_step_resident(B)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        eng.ctx.check()
        assert bool(torch.isfinite(o["losses"][[0, 2, 4]]).all())
        tf = c["flop_row"] * B / (ms * 1e-3) / 1e12
        out[name] = {"workload": c["desc"], "rows_per_step": B, "value": B / (ms * 1e-3), "unit": "rows/s",
                     "ms_per_step": ms, "steps": steps, "model_tflops": tf, "frac_of_tensor_peak": tf / pk["tf_sust"],
                     "gpu_launches_per_step": int(eng.graph_launches) if eng.use_graph else None}
        del eng, dx, df, dl
        torch.cuda.empty_cache()
    return out


def gaitset_leg(pk, args):
    """GaitSet branch type (README recipes use --gaitset): the same 3-modality training
    step with UWYHSemiNet.build_gaitset_branch branches on 25 x 60 x 60 clips, 96 rows per GPU."""
    from ugaitnet_b200.config import MERGE_SIGNMAX, GaitSetConfig
    from ugaitnet_b200.gaitset import GaitSetEngine
    B, T, HW = BS_LITERAL * EXPAND, 25, 60
    cfg = GaitSetConfig(in_channels=(2, 1, 1), frames=T, hw=HW, nc=0, nclasses=NCLASSES, merge=MERGE_SIGNMAX,
                        margin=0.2, wver=1.0, wid=0.1)
    g = torch.Generator().manual_seed(232323)
    hx = [((torch.randn(B, T, HW, HW, c, generator=g) * 0.3).clamp_(-3.3, 3.3) if c == 2
           else torch.rand(B, T, HW, HW, c, generator=g) - 0.5).pin_memory() for c in cfg.in_channels]
    hf = [torch.ones(B, 1) for _ in cfg.in_channels]
    for i in range(B):                     # expansion pattern: row 4i complete, the copies lose modalities
        if i % EXPAND == 1:
            hf[i % 3][i] = 0
        elif i % EXPAND >= 2:
            keep = (i // EXPAND + i) % 3
            for m in range(3):
                hf[m][i] = 1.0 if m == keep else 0.0
    for m in range(3):
        hx[m][hf[m].reshape(-1) == 0] = 1e-9
    hf = [f.pin_memory() for f in hf]
    hl = (torch.arange(B) // (2 * EXPAND)).to(torch.int32).pin_memory()
    dx, df, dl = [t.cuda() for t in hx], [t.cuda() for t in hf], hl.cuda()
    h2d = sum(t.numel() * 4 for t in hx + hf) + hl.numel() * 4
    eng = GaitSetEngine(cfg, math_mode=args.mode, lr=1e-4, use_graph=not args.no_graph)

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    steps = max(3, args.steps // 2)
    step_dev = lambda: eng.train_step(dx, df, dl)
    loss_host = torch.zeros(2).pin_memory()

    def step_e2e():
        out = eng.train_step(hx, hf, hl)
        loss_host[0].copy_(out["triplet"], non_blocking=True)
        loss_host[1].copy_(out["ce"], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        step_dev()
    l0 = eng.ctx.launches
    ms = timed(step_dev, steps)
    launches = eng.graph_launches * steps if eng.use_graph else eng.ctx.launches - l0
    step_e2e()
    ms_e2e = timed(step_e2e, steps)
    eng.ctx.check()
    # algorithmic conv FLOPs per row: per frame a1..a6, per sequence b1..b4; training = fwd + dgrad + wgrad, no dgrad for a1
    fr = lambda h, cin, co, k: 2.0 * h * h * cin * k * k * co
    fwd_row = sum(T * (fr(64, c, 32, 5) + fr(64, 32, 32, 3) + fr(32, 32, 64, 3) + fr(32, 64, 64, 3) + fr(16, 64, 128, 3)
                       + fr(16, 128, 128, 3)) + fr(32, 32, 64, 3) + fr(32, 64, 64, 3) + fr(16, 64, 128, 3)
                  + fr(16, 128, 128, 3) for c in cfg.in_channels)
    train_row = 3 * fwd_row - sum(T * fr(64, c, 32, 5) for c in cfg.in_channels)
    out = {"workload": "UWYHSemiNet3Mods.build(gaitset=True): 3 modalities (OF 2ch, gray, depth), 25 x 60 x 60 clips, "
                       "HPP 62 parts x 256, sign_max, triplet over 62 parts + CE, Adam; 96 rows per GPU",
           "value": B / (ms * 1e-3), "unit": "rows/s", "ms_per_step": ms, "steps": steps,
           "e2e": {"value": B / (ms_e2e * 1e-3), "unit": "rows/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": 8, "api": "GaitSetEngine.train_step (pinned host inputs, serial)"},
           "gpu_launches": int(launches), "gpu_launches_per_step": int(launches) // steps,
           "model_tflops": train_row * B / (ms * 1e-3) / 1e12,
           "train_gflop_per_row": train_row / 1e9}
    # per-op pass (eager, branches in sequence) -> share of the conv kernels and their tensor roofline
    eager = GaitSetEngine(cfg, math_mode=args.mode, lr=1e-4, use_graph=False) if eng.use_graph else eng
    eager.multistream = False
    for _ in range(2):
        eager.train_step(dx, df, dl)
    tm = OpTimer()
    tm.install()
    eager.train_step(dx, df, dl)
    torch.cuda.synchronize()
    tm.uninstall()
    agg = {}
    for name, a, e0, e1 in tm.records:
        agg[name] = agg.get(name, 0.0) + e0.elapsed_time(e1)
    total = sum(agg.values())
    out["op_ms_per_step"] = {k: round(v, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])}
    conv_ms = sum(agg.get(k, 0.0) for k in ("ugn_conv2d_fwd", "ugn_conv2d_dgrad", "ugn_conv2d_wgrad"))
    tc_flops = (train_row - 2 * sum(T * fr(64, c, 32, 5) for c in cfg.in_channels)) * B     # a1 runs on the FFMA pipe
    ach = tc_flops / (conv_ms * 1e-3) / 1e12 if conv_ms else 0.0
    pf, pb = MMA_PASSES[args.mode]
    passes = (pf + 2 * pb) / 3.0
    out["roofline"] = {"bound": "tensor", "kernel": "tc_convp_kernel / tc_wgradv_kernel on the 3x3 'same' layers",
                       "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"],
                       "mma_passes": passes, "issued_frac": ach * passes / pk["tf_sust"],
                       "share_of_step": conv_ms / total if total else None, "traffic": None,
                       "note": "N = 32 / 64 channel tiles sit under the 73-clk tcgen05.mma floor and the 34 / 18 pixel "
                               "rows fill 64 / 32-slot row pitches half way: structural ceiling ~0.25 of peak issued"}
    # CPU baseline: the oracle (PyTorch-CPU fp32) on a bounded sample of the same step
    from oracle import gaitset_oracle as G
    torch.set_num_threads(os.cpu_count() or 1)
    oc = G.GaitSetConfig(in_channels=(2, 1, 1), frames=T, hw=HW, nc=0, nclasses=NCLASSES, merge=2, wver=1.0, wid=0.1)
    rows = 4
    xs, fl, lab = G.synth_batch(oc, 2, 2, seed=1)
    P = G.init_params(oc, seed=1)
    times = []
    for _ in range(2):
        t0 = time.perf_counter()
        G.loss_and_grads(xs, fl, lab, P, oc)
        times.append(time.perf_counter() - t0)
    out["cpu_baseline"] = {"value": rows / min(times), "unit": "rows/s", "cores": torch.get_num_threads(), "kind": "port",
                           "sample": f"PyTorch-CPU fp32 restatement, forward + backward of {rows} rows, best of 2"}
    return out


def knn_leg(pk, rank=0, world=1, pg=None, D=256):
    """Open-world test (BASELINE cfg5): k=3 queries/s over a synthetic 1 M x D gallery of L2-normalised
    descriptors clustered around 155 class centroids (0.1 % exact duplicate rows), row-sharded over the
    ranks; Q = 4096 (tensor-bound) and Q = 64 (gallery-streaming, HBM-bound).  D = 256: the FC1 "code" descriptor
    (casenet C); D = 2048: the `signature` descriptor of the benchmarked model (nd = 2048, no FC1)."""
    from ugaitnet_b200.knn import KNeighborsClassifier, shard_bounds
    N, k = 1_000_000, 3
    lo, hi = shard_bounds(N, rank, world)
    g = torch.Generator(device="cuda").manual_seed(5)          # same stream on every rank -> same gallery
    cent = torch.randn(155, D, device="cuda", generator=g)
    lab_all = torch.randint(0, 155, (N,), device="cuda", generator=g, dtype=torch.int32)
    G = torch.empty(hi - lo, D, device="cuda")
    step = 125_000 if D <= 256 else 25_000
    for s in range(0, N, step):                                # generate the full stream, keep this rank's rows
        blk = cent[lab_all[s:s + step].long()] + 0.35 * torch.randn(step, D, device="cuda", generator=g)
        blk = blk / blk.norm(dim=1, keepdim=True)
        a, b = max(s, lo), min(s + step, hi)
        if a < b:
            G[a - lo:b - lo] = blk[a - s:b - s]
        del blk
    dup = torch.arange(0, hi - lo - 1, 1000, device="cuda")
    G[dup + 1] = G[dup]                                        # exact duplicates (distance ties)
    lab = lab_all[lo:hi].contiguous()
    Qall = torch.randn(4096, D, device="cuda", generator=g) * 0.05
    Qall += cent[torch.randint(0, 155, (4096,), device="cuda", generator=g)] + 0.3 * torch.randn(4096, D, device="cuda", generator=g)
    Qall = Qall / Qall.norm(dim=1, keepdim=True)
    clf = KNeighborsClassifier(n_neighbors=k, process_group=pg).fit(G, lab, idx_base=lo, sharded=True)
    out = {"N": N, "D": D, "k": k, "gallery_rows_per_gpu": hi - lo, "scan": "tcgen05 fp16 hi/lo 3-pass + fused top-k"
           if clf.use_tc else "simt fp32"}

    def timed(fn, reps):
        for _ in range(3):            # first call: workspaces; second: CUDA-graph capture of the search; third: a replay
            fn()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t)
        return ms, r

    for Q in (4096, 64):
        qd = Qall[:Q].contiguous()
        ms, pred = timed(lambda: clf.predict_device(qd), 5 if D <= 256 else 3)
        flops = 2.0 * Q * N * D
        gal_bytes = (hi - lo) * clf.dp * 4 if clf.use_tc else (hi - lo) * D * 4
        out[f"Q{Q}"] = {"queries_per_s": Q / (ms * 1e-3), "ms": ms, "flagged_exact_recompute": clf.flagged_queries(),
                        "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
                        # aggregate FLOP/s over `world` GPUs against world x one GPU's peak (x3: issued MMA passes)
                        "tensor_frac_of_peak": flops / (ms * 1e-3) / 1e12 / (pk["tf_sust"] * world),
                        "tensor_issued_frac_of_peak": 3 * flops / (ms * 1e-3) / 1e12 / (pk["tf_sust"] * world),
                        "gallery_GBps_per_gpu": gal_bytes / (ms * 1e-3) / 1e9, "hbm_frac_of_peak": gal_bytes / (ms * 1e-3) / 1e9 / pk["hbm"]}
    # end to end: pinned host queries -> labels on the host
    hq = Qall.cpu().pin_memory()
    hp = torch.empty(4096, dtype=torch.int32).pin_memory()

    def e2e():
        hp.copy_(clf.predict_device(hq.cuda(non_blocking=True)), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ms, _ = timed(e2e, 3)
    out["e2e_Q4096"] = {"queries_per_s": 4096 / (ms * 1e-3), "ms": ms, "h2d_bytes": 4096 * D * 4, "d2h_bytes": 4096 * 4}
    out["queries_per_s"] = out["Q4096"]["queries_per_s"]
    if rank == 0 and world == 1 and D <= 256:
        try:   # the reference's exact call on the host cores, bounded sample of the same queries
            from sklearn.neighbors import KNeighborsClassifier as SK
            Gh, yh, Qh = G.cpu().numpy(), lab.cpu().numpy(), Qall[:128].cpu().numpy()
            sk = SK(n_neighbors=k).fit(Gh, yh)
            t0 = time.perf_counter()
            sp = sk.predict(Qh)
            dt = time.perf_counter() - t0
            pred128 = clf.predict(Qall[:128])
            out["cpu_sklearn"] = {"queries_per_s": 128 / dt, "cores": os.cpu_count(), "sample": "128 of the 4096 queries, full 1 M x 256 gallery",
                                  "labels_equal": bool((sp == pred128).all())}
        except Exception as e:  # pragma: no cover
            out["cpu_sklearn_error"] = str(e)
    return out


if __name__ == "__main__":
    main()
