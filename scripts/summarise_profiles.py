#!/usr/bin/env python
"""Condense the ncu outputs of scripts/ncu_capture.sh (gpurun_out/<tag>/) into small tracked
summaries under profiles/:
    profiles/<tag>_launches.md   per-kernel totals of one eager step (gpu__time_duration.sum list)
    profiles/<tag>_<kernel>.md   key `--set full` counters of each captured launch
usage: python scripts/summarise_profiles.py <tag>
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thr"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor (HMMA) pipe active % of elapsed"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor (HMMA) pipe active % of active"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem throughput %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
]


def read_raw(path):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    out = []
    for r in rows[hdr_i + 2:]:
        if len(r) == len(hdr):
            out.append({h: (v, u) for h, v, u in zip(hdr, r, units)})
    return out


def find(row, key):
    for h, (v, u) in row.items():
        if h == key or h.endswith("." + key) or h.endswith(key):
            return v, u
    return None, None


def launches(tag, src, dst, what="cfg2, B=96, `bench.py --no-graph --lite`"):
    rows = list(csv.reader(open(src)))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d["Metric Name"] != "gpu__time_duration.sum":
                continue
            v = float(d["Metric Value"].replace(",", ""))
            u = d["Metric Unit"]
            v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
            name = d["Kernel Name"].split("(")[0].replace("void ", "")
            agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    n = sum(len(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {tag}: launch list of ONE eager training step ({what})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are cold-cache and\n"
                "serialised, so only each kernel's SHARE of the step is comparable with the live CUDA-event numbers.\n\n")
        f.write(f"{n} launches, {tot / 1e3:.3f} ms summed\n\n| kernel | launches | sum us | max us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {len(v)} | {sum(v):.1f} | {max(v):.1f} | {sum(v) / tot:.3f} |\n")
    return agg, tot


def kernel(tag, name, src, dst):
    rows = read_raw(src)
    with open(dst, "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none` of `{name}` ({len(rows)} launches of one eager cfg2 step)\n\n")
        f.write("| # | " + " | ".join(lbl for _, lbl in KEYS) + " |\n|" + "---|" * (len(KEYS) + 1) + "\n")
        for i, r in enumerate(rows):
            cells = []
            for key, _ in KEYS:
                v, u = find(r, key)
                cells.append("-" if v is None else f"{v} {u}".strip())
            f.write(f"| {i} | " + " | ".join(cells) + " |\n")
    return rows


def to_bytes(v, u):
    v = float(str(v).replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "B": 1, "KB": 1e3, "MB": 1e6, "GB": 1e9}.get(u, 1)


def traffic(tag, mode, rows_by_kernel, dst):
    """profiles/traffic.json: DRAM bytes per step of the kernel families bench.py quotes in `roofline.traffic`
    (dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of ONE step of this capture)."""
    path = os.path.join(dst, "traffic.json")
    t = json.load(open(path)) if os.path.exists(path) else {}
    ent = t.setdefault(mode, {})

    def total(rows):
        tot = 0.0
        for r in rows:
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v, u = find(r, key)
                tot += to_bytes(v, u) if v is not None else 0.0
        return tot
    if "tc_convp_kernel" in rows_by_kernel:      # launch order of one step: 12 forward launches, then 3 input gradients
        rows = rows_by_kernel["tc_convp_kernel"]
        ent["conv_fwd"] = {"bytes_per_step": total(rows[:12]), "launches": min(12, len(rows)), "capture": tag}
    if "optim_kernel" in rows_by_kernel:
        ent["optim"] = {"bytes_per_step": total(rows_by_kernel["optim_kernel"][:1]), "launches": 1, "capture": tag}
    if "gs_convp" in rows_by_kernel:
        ent["gaitset_convp_first_modality"] = {"bytes_per_step": total(rows_by_kernel["gs_convp"]),
                                                "launches": len(rows_by_kernel["gs_convp"]), "capture": tag}
    json.dump(t, open(path, "w"), indent=1, sort_keys=True)


def main():
    tag = sys.argv[1]
    mode = sys.argv[2] if len(sys.argv) > 2 else "f16mix"
    src = os.path.join(ROOT, "gpurun_out", tag)
    dst = os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    if os.path.exists(os.path.join(src, "launches.csv")):
        launches(tag, os.path.join(src, "launches.csv"), os.path.join(dst, f"{tag}_launches.md"))
    if os.path.exists(os.path.join(src, "gs_launches.csv")):
        launches(tag, os.path.join(src, "gs_launches.csv"), os.path.join(dst, f"{tag}_gs_launches.md"),
                 "GaitSet branches, 3 modalities, 24 rows, `GS_LITE=1 scripts/gs_bench.py 24 f16mix 1`")
    rows_by_kernel = {}
    for fn in sorted(os.listdir(src)):
        if fn.startswith("prof_") and fn.endswith("_raw.csv"):
            name = fn[len("prof_"):-len("_raw.csv")]
            rows_by_kernel[name] = kernel(tag, name, os.path.join(src, fn), os.path.join(dst, f"{tag}_{name}.md"))
    traffic(tag, mode, rows_by_kernel, dst)
    for fn in ("bench_n1.json", "pytest_gpu.log"):
        p = os.path.join(src, fn)
        if os.path.exists(p):
            with open(p) as f, open(os.path.join(dst, f"{tag}_{fn}"), "w") as g:
                g.write(f.read()[-20000:])


if __name__ == "__main__":
    main()
