"""Turns ncu outputs into the small text summaries committed under profiles/.
  python tools/ncu_summary.py launches <launch-list.csv>            -> per-kernel share of the step (gpu__time_duration.sum)
  python tools/ncu_summary.py full <report.ncu-rep>                 -> key metrics per captured launch (needs ncu on PATH)
  python tools/ncu_summary.py traffic <report.ncu-rep>              -> JSON: DRAM bytes per launch per bench stage (bench.py's roofline.traffic)"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.sum", "sm__inst_executed_pipe_tmem.sum",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
        a = agg.setdefault(row["Kernel Name"].split("(")[0], [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {sum(a[0] for a in agg.values())} launches, {tot:.1f} us total (ncu-serialised, cold caches: compare SHARES, not absolutes)")
    print(f"{'share':>7} {'us/launch':>10} {'launches':>8}  kernel")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[1] / tot * 100:6.1f}% {a[1] / a[0]:10.1f} {a[0]:8d}  {k}")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("## " + r[idx["Kernel Name"]].split("(")[0] + "   grid " + r[idx.get("launch__grid_size", 0)])
        for k in KEYS:
            if k in idx and r[idx[k]] not in ("", "n/a"):
                print(f"  {k:78s} {r[idx[k]]:>18s} {units[idx[k]]}")


# kernel-name prefix -> bench.py stage; the n-th launch of the hash forward kernel in a step decides inference (1st) vs training (2nd, when re-encoding)
STAGE_OF = [("hash_encode_forward", "encode_inference"), ("nerf_mlp_pipe_infer_kernel<1>", "mlp_inference"), ("nerf_mlp_pipe_infer_kernel<(int)1>", "mlp_inference"),
            ("nerf_mlp_pipe_train_kernel<2>", "mlp_train"), ("nerf_mlp_pipe_train_kernel<(int)2>", "mlp_train"), ("nerf_mlp_kernel<1>", "mlp_inference"),
            ("nerf_mlp_kernel<2>", "mlp_train"), ("hash_encode_backward", "encode_backward"), ("adam_ema_kernel", "optimizer"),
            ("march_words_kernel", "sampling"), ("loss_gradient_kernel", "loss")]


def traffic(path):
    import json
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    acc = collections.OrderedDict()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        for prefix, stage in STAGE_OF:
            if prefix in name.split("(")[0]:
                b = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
                a = acc.setdefault(stage, [0, 0.0, name.split("(")[0]])
                a[0] += 1; a[1] += b
                break
    out = {"source": "ncu --set full --clock-control none, " + path.split("/")[-1] + " (dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches)",
           "stages": {k: {"kernel": v[2], "launches": v[0], "dram_bytes_per_launch": v[1] / v[0]} for k, v in acc.items()}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
