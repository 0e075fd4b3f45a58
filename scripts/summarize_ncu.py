"""Turns the raw ncu output of scripts/refresh_profiles.sh (gpurun_out/refresh_<tag>/) into the tracked summaries:

    profiles/<tag>_launches_bench_mt.md        every kernel of one steady-state iteration: launches, device time, share
    profiles/<tag>_conv_traffic.json           DRAM bytes / launch + tensor-pipe activity of the tcgen05 conv family
                                               (carries the source hash of the build; bench.py checks it)
    profiles/<tag>_ncu_bandwidth_kernels.md    HBM GB/s of the bandwidth-bound kernels (ROIAlign, SGD+EMA, MGD, PSM, NMS ...)
    profiles/<tag>_ncu_full_<name>.md          selected raw-page metrics of the `--set full` captures (if present)

    python scripts/summarize_ncu.py <tag>
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
SRC = os.path.join(ROOT, "gpurun_out", "refresh_" + TAG)
DST = os.path.join(ROOT, "profiles")
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6547.5


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void |at::native::|native::", "", name)
    return re.sub(r"\(.*", "", name)[:80]


def to_unit(v, unit):
    v = float(v.replace(",", ""))
    return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6, "byte": v, "Kbyte": v * 1e3, "Mbyte": v * 1e6,
            "Gbyte": v * 1e9, "%": v}.get(unit, v)


def read_launches(path):
    rows = collections.OrderedDict()
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if "Metric Name" not in r or not r.get("ID"):
            continue
        d = rows.setdefault(r["ID"], {"kernel": short(r["Kernel Name"]), "grid": r.get("Grid Size", "")})
        d[r["Metric Name"]] = to_unit(r["Metric Value"], r["Metric Unit"])
    return list(rows.values())


def main():
    src_hash = open(os.path.join(SRC, "source_hash.txt")).read().strip()
    launches = read_launches(os.path.join(SRC, "launches.csv"))
    T = "gpu__time_duration.sum"
    fam = collections.OrderedDict()
    for l in launches:
        f = fam.setdefault(l["kernel"], dict(n=0, us=0.0, rd=0.0, wr=0.0, tens=0.0))
        f["n"] += 1
        f["us"] += l.get(T, 0.0)
        f["rd"] += l.get("dram__bytes_read.sum", 0.0)
        f["wr"] += l.get("dram__bytes_write.sum", 0.0)
        f["tens"] += l.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * l.get(T, 0.0)
    total = sum(f["us"] for f in fam.values())
    ours = sum(f["us"] for k, f in fam.items() if not k.startswith(("vectorized_", "elementwise_", "index_", "mbtopk", "sbtopk",
               "radixSort", "CatArray", "at_cuda", "reduce_kernel", "unrolled_", "cub::", "Memcpy", "Memset", "flip", "bitonic",
               "vectorized_gather", "write_indices", "compute_", "DeviceScan", "DeviceRadix", "DeviceSelect", "DeviceReduce",
               "DeviceCompact", "nccl", "triu", "fill", "distribution", "indexSelect", "tensor_kernel", "gatherTopK")))
    n_launch = sum(f["n"] for f in fam.values())
    md = ["# ncu launch list — bench.py MT workload as shipped (CUDA graphs on), build %s" % src_hash, "",
          "Command: `bash scripts/refresh_profiles.sh %s ncu` (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
          "sm__pipe_tensor_cycles_active... --clock-control none -s <warm-up launches> -c <one iteration>; kernel nodes of the replayed graphs are "
          "profiled individually).  Cold-cache, serialised: compare SHARES, not absolutes." % TAG, "",
          "captured window: %d launches, %.2f ms of device time; libmmtpsm_b200 kernels: %.1f %% of it" % (n_launch, total / 1e3, 100 * ours / total), "",
          "| kernel | launches | total us | share | avg us | DRAM MB / launch | tensor pipe % (time-weighted) |", "|---|---|---|---|---|---|---|"]
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"])[:45]:
        md.append("| `%s` | %d | %.0f | %.1f%% | %.1f | %.1f | %.1f |" % (k, f["n"], f["us"], 100 * f["us"] / total, f["us"] / f["n"],
                                                                        (f["rd"] + f["wr"]) / f["n"] / 1e6, f["tens"] / max(f["us"], 1e-9)))
    open(os.path.join(DST, TAG + "_launches_bench_mt.md"), "w").write("\n".join(md) + "\n")

    conv = {k: f for k, f in fam.items() if "conv_tc_kernel" in k or "wgrad_tc_kernel" in k}
    agg = dict(n=sum(f["n"] for f in conv.values()), us=sum(f["us"] for f in conv.values()),
               bytes=sum(f["rd"] + f["wr"] for f in conv.values()), tens=sum(f["tens"] for f in conv.values()))
    out = {"source_hash": src_hash, "command": "bash scripts/refresh_profiles.sh %s ncu" % TAG,
           "note": "one steady-state MT iteration, graphs on; cold-cache serialised replay",
           "families": [{"kernel": "all tcgen05 conv launches", "launches": agg["n"], "avg_us": agg["us"] / max(agg["n"], 1),
                         "dram_bytes_per_launch": agg["bytes"] / max(agg["n"], 1),
                         "tensor_pipe_active_pct_time_weighted": agg["tens"] / max(agg["us"], 1e-9),
                         "share_of_device_time": agg["us"] / total}] +
                        [{"kernel": k, "launches": f["n"], "avg_us": f["us"] / f["n"], "dram_bytes_per_launch": (f["rd"] + f["wr"]) / f["n"],
                          "tensor_pipe_active_pct_time_weighted": f["tens"] / max(f["us"], 1e-9)} for k, f in conv.items()]}
    json.dump(out, open(os.path.join(DST, TAG + "_conv_traffic.json"), "w"), indent=1)

    bw_names = ("roi_align", "sgd_ema", "ema_kernel", "mgd_", "psm_kernel", "rank_sort", "iou_mask", "greedy_scan", "bn_relu_bwd",
                "upsample2", "maxpool", "stem_s2d", "bias_grad", "filter_dgrad", "paste_sum", "rasterize", "match_pass", "decode_clip",
                "balanced_sample", "rpn_loss", "color_views", "subsample2", "encode_kernel", "fpn_rois", "rpn_gather", "focal")
    md = ["# HBM traffic of the bandwidth / latency-bound kernels (ncu dram__bytes_read.sum + dram__bytes_write.sum, gpu__time_duration.sum), build %s" % src_hash, "",
          "Same capture as `%s_launches_bench_mt.md` (one steady-state MT iteration; cold-cache, serialised replay — DRAM GB/s here is the "
          "kernel alone on an idle GPU with cold L2, the in-step figures are in bench.py's `kernel_families`).  Peak: %.0f GB/s (MEASURED_PEAKS.json)." % (TAG, HBM), "",
          "| kernel | launches / iteration | avg us | DRAM MB / launch (read + write) | DRAM GB/s | of measured peak |", "|---|---|---|---|---|---|"]
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
        if not any(b in k for b in bw_names):
            continue
        gbs = (f["rd"] + f["wr"]) / max(f["us"], 1e-9) / 1e3
        md.append("| `%s` | %d | %.1f | %.2f | %.0f | %.2f |" % (k, f["n"], f["us"] / f["n"], (f["rd"] + f["wr"]) / f["n"] / 1e6, gbs, gbs / HBM))
    open(os.path.join(DST, TAG + "_ncu_bandwidth_kernels.md"), "w").write("\n".join(md) + "\n")

    keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct")
    for name in ("prof_conv", "prof_bw"):
        rep = os.path.join(SRC, name + ".ncu-rep")
        if not os.path.exists(rep):
            continue
        r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(r.stdout.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        cols = [i for i, h in enumerate(hdr) if h in keep or h == "Kernel Name"]
        md = ["# ncu --set full (%s), build %s" % (name, src_hash), "",
              "Raw page, selected metrics (units in the second row).  The .ncu-rep itself stays in gpurun_out/ (size).", "",
              "| " + " | ".join(hdr[i] for i in cols) + " |", "|" + "---|" * len(cols), "| " + " | ".join(units[i] for i in cols) + " |"]
        for row in rows[2:]:
            md.append("| " + " | ".join(short(row[i]) if hdr[i] == "Kernel Name" else row[i] for i in cols) + " |")
        open(os.path.join(DST, "%s_ncu_full_%s.md" % (TAG, name[5:])), "w").write("\n".join(md) + "\n")
    print("wrote summaries for", TAG, "hash", src_hash)


if __name__ == "__main__":
    main()
