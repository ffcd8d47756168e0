#!/usr/bin/env python
"""bench.py -- NeRF training throughput of the B200-native hot path (BASELINE.json: "NeRF train iters/sec").

    python bench.py [--gpus N] [--steps K] [--warmup W]                # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # the CPU restatement of the reference, host cores

Workload (BASELINE.json configs[1]): Lego-shaped synthetic scene, 100 cameras x 800x800 RGBA8, configs/nerf/base.json
(hash grid 16x2 / T=2^19, 64-wide density + rgb MLPs), target batch 2^18 compacted samples per iteration, random-init
weights, seed 1337. One "step" = one `Testbed.train(batch)` = one optimizer step incl. ray generation, marching,
inference on the uncompacted samples, compositing/loss/compaction, forward+backward, Adam/EMA, and the occupancy-grid
refresh at the reference's cadence. `value` is device-timed (CUDA events on the testbed's stream) with the dataset
resident in HBM; `e2e` goes through the public `pyngp.Testbed` calls with the dataset starting in pinned HOST memory
(its upload is inside the timed region) and the loss read back to the host every step.

Also in the line: `roofline` (dominant kernel by share of the step in the committed ncu launch list: algorithmic bytes / event time against MEASURED_PEAKS.json, DRAM traffic per launch from the
committed ncu capture profiles/r01_traffic.json, and the same for every stage under `per_stage`), `cpu_baseline` (the oracle's whole-iteration restatement on the
host cores, bounded to --cpu-seconds), `render` (classic Testbed.render Msamples/s and, single GPU only, the Blender path request_nerf_render_sync from a snapshot),
`clocks` (nvidia-smi sampled during the timed region), `gpu_launches`.

Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "blender-ngp_b200"))

METRIC = "nerf_train_iters_per_sec"
UNIT = "it/s (batch 2^18 compacted samples per iteration)"
BATCH = 1 << 18
N_IMAGES, RES = 100, 800

# Algorithmic bytes / flops per unit (restated in DESIGN.md s4)
HASH_FWD_BYTES = 12 + 16 * 8 * 4 + 16 * 2 * 2      # 588 B/sample: position + 128 gathers x 4 B + 32 fp16 features out
HASH_BWD_BYTES = 12 + 64 + 16 * 8 * 4               # 588 B/sample: position + dL/dy + 128 scatter-adds of one fp16 pair each -- the REFERENCE's traffic
HASH_BWD_BYTES_AS_EXECUTED = 12 + 64 + 16 * 8 * 2 * 4   # 1100 B/sample: this build accumulates fp32 pairs (8 B per scatter-add); reported next to the figure above, never instead of it
MLP_FWD_FLOPS = 20480                               # padded widths as executed
MLP_TRAIN_FLOPS = 61440
STAGE_ALGO = {  # stage -> (bound, per-unit quantity, unit of `achieved`)
    "encode_inference": ("hbm", HASH_FWD_BYTES, "GB/s"),
    "encode_train": ("hbm", HASH_FWD_BYTES, "GB/s"),
    "encode_backward": ("hbm", HASH_BWD_BYTES, "GB/s"),
    "mlp_inference": ("tensor", MLP_FWD_FLOPS, "TFLOP/s"),
    "mlp_train": ("tensor", MLP_TRAIN_FLOPS, "TFLOP/s"),
}
# kernel-name fragment -> stage (to place the top line of the committed ncu launch list, profiles/r02_launches_summary.txt)
KERNEL_STAGE = [("hash_encode_forward", "encode_inference"), ("hash_encode_backward", "encode_backward"), ("adam_ema", "optimizer"),
                ("nerf_mlp_pipe_train", "mlp_train"), ("reduce_partials", "mlp_train"), ("nerf_mlp_pipe_infer", "mlp_inference"),
                ("march_words", "sampling"), ("training_samples", "sampling"), ("training_rays", "sampling"),
                ("loss_", "loss"), ("rollover", "loss")]
# Stages whose algorithmic bytes depend on more than one unit count; evaluated in stage_bytes() below:
#   sampling (K1)   28 B per marched sample written + 36 B per ray
#   loss (K6 + K7)  72 B per marched sample (8 B network output + 28 B coordinate in, 28 B coordinate + 8 B gradient out) + 56 B per ray
#   optimizer (K15) 10 B per parameter (gradient memset 2 + Adam's gradient read 2 + EMA 6) + 34 B per parameter whose gradient is non-zero;
#                   the touched count is MEASURED in the run (per-parameter step counters before / after one step), not fitted
OPT_FIXED_BYTES, OPT_TOUCHED_BYTES = 10, 34
L2_PEAK_BYTES_PER_CLK = 6300.0  # B300_MICROARCH.md "LTS throughput cap ~6300 B/cyc full-chip" (measured on B300; same L2 design): x the SM clock of the run


def stage_bytes(name, rays, samples, batch, n_params, touched):
    if name == "sampling":
        return samples * 28 + rays * 36
    if name == "loss":
        return samples * 72 + rays * 56
    if name == "optimizer":
        return n_params * OPT_FIXED_BYTES + touched * OPT_TOUCHED_BYTES
    return None


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--preroll", type=int, default=512, help="untimed training steps before warm-up, so the timed steps run in the steady regime "
                   "(occupancy grid has culled empty space, refresh every 16 steps; the first 256 steps refresh 2M cells every step)")
    p.add_argument("--batch", type=int, default=BATCH)
    p.add_argument("--n-images", type=int, default=N_IMAGES)
    p.add_argument("--res", type=int, default=RES)
    p.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline leg")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="N > 1: weak = every rank trains --batch samples per step (global batch N x B, "
                   "the default and what the driver's scaling run measures); strong = the global batch stays --batch, every rank trains --batch / N")
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), tensor_burst=float(d["bf16_tflops"]), tensor_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, src="fallback")


class ClockSampler:
    """SM clock and throttle reasons polled through NVML every ~1 ms from a thread while the timed region runs (the timed call releases the GIL).
    nvidia-smi's own polling (50 ms at best) is too coarse for a timed region of tens of milliseconds."""

    def __init__(self, index):
        import threading
        self.index, self.samples, self.reasons, self.stop_flag, self.thread = index, [], set(), threading.Event(), None
        self.max_mhz, self.power = None, []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
                if len(self.samples) % 16 == 0:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
            except Exception:
                break
            time.sleep(0.001)

    def start(self):
        if self.nv is None:
            return
        import threading
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=self.max_mhz, reasons=[])
        if self.thread is None:
            return out
        self.stop_flag.set()
        self.thread.join(timeout=2)
        if self.samples:
            out.update(sm_mhz=float(np.median(self.samples)), sm_min_mhz=float(min(self.samples)), samples=len(self.samples), reasons=sorted(self.reasons),
                       power_w_max=float(max(self.power)) if self.power else None, how="NVML polled every ~1 ms during the timed region")
        return out


def make_scene(n_images, res, device):
    import synthetic
    return synthetic.make_lego_scene(n_images, res, device=device, seed=0, as_numpy=False)


# ---------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's whole-iteration restatement (oracle/ngp_trainer.cpp) on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_baseline(images_np, xforms, fx, fy, seconds, steps_cap=10 ** 9, warmup=1, batch_cpu=1 << 14):
    """Times the CPU restatement of Testbed::train at a reduced batch in the steady-state regime (step counter 257: occupancy
    refresh every 16 steps; occupancy grid set from the scene geometry), and scales samples/s to the metric's batch."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as orc
    from conftest import scene_occupancy_bitfield
    cores = orc.set_num_threads(os.cpu_count() or 1)  # all host threads, whatever OMP_NUM_THREADS says (torchrun exports 1)
    imgs = orc.make_images(images_np, xforms, fx, fy)
    grid, _ = scene_occupancy_bitfield(orc)
    t = orc.Trainer(imgs, aabb_scale=1, seed=1337)
    t.set_state(257, 0, grid)
    for _ in range(warmup + 2):  # lets rays_per_batch settle (it adapts in two steps)
        t.train(batch_cpu)
    n, t0 = 0, time.perf_counter()
    per_step = []
    while n < steps_cap:
        s0 = time.perf_counter()
        t.train(batch_cpu)
        per_step.append(time.perf_counter() - s0)
        n += 1
        if time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    it_s_cpu_batch = n / dt
    return dict(value=it_s_cpu_batch * batch_cpu / BATCH, unit=UNIT, cores=cores, kind="port",
                sample=f"{n} training iterations at batch 2^{int(np.log2(batch_cpu))} ({it_s_cpu_batch:.2f} it/s), OpenMP over {cores} host threads, "
                       f"steady-state regime (step counter 257, occupancy grid from scene geometry); value = it/s x 2^{int(np.log2(batch_cpu))}/2^18",
                ms_per_step_cpu_batch=1e3 * float(np.mean(per_step))), per_step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        import torch
        dev = "cuda" if torch.cuda.is_available() else "cpu"
    except Exception:
        dev = "cpu"
    n_images = args.n_images if dev == "cuda" else min(args.n_images, 8)  # scene synthesis on the CPU costs ~4 s per 800x800 image
    scene = make_scene(n_images, args.res, dev)
    images_np = scene["images"].cpu().numpy()
    steps = max(1, args.steps)
    # each step is a bounded sample of the workload: one iteration at batch 2^14 (1/16 of the metric's batch)
    cb, per_step = cpu_baseline(images_np, scene["xforms"], scene["fx"], scene["fy"], seconds=1e9, steps_cap=steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": cb["ms_per_step_cpu_batch"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16 storage / fp32 accumulate",
        "data": "synthetic",
        "config": {"workload": f"NeRF Lego-shaped synthetic scene ({n_images} cams {args.res}x{args.res}), configs/nerf/base.json, CPU restatement of the reference "
                               "(the reference has no CPU path of its own, SURVEY.md s8d); bounded sample per step: one iteration at batch 2^14",
                   "batch": 1 << 14, "metric_batch": BATCH},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import pyngp
    import synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = max(1, args.steps), max(3, args.warmup)
    global_batch = args.batch * world if args.scaling == "weak" else args.batch
    if args.scaling == "strong" and world > 1:
        args.batch = max(128, args.batch // world // 128 * 128)  # this rank's share of the fixed global batch
    scene = make_scene(args.n_images, args.res, f"cuda:{local_rank}")
    images_dev = scene["images"]
    host_images = torch.empty(images_dev.shape, dtype=torch.uint8, pin_memory=True)
    host_images.copy_(images_dev)
    del images_dev
    torch.cuda.synchronize()
    images_np = host_images.numpy()
    dataset_bytes = int(images_np.nbytes)

    def new_testbed():
        tb = pyngp.Testbed(device=local_rank)
        if world > 1:
            tb.init_data_parallel(rank, world)
        return tb

    # ---- device-timed arm: dataset resident in HBM ----
    tb = new_testbed()
    tb.load_training_images(list(images_np), scene["xforms"], scene["fx"], scene["fy"])
    tb.train_n(args.preroll, args.batch)
    tb.train_n(W, args.batch)
    stream = torch.cuda.ExternalStream(tb.stream, device=torch.device("cuda", local_rank))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = tb.stats()["gpu_launches"]
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    cuda_profiler = os.environ.get("NGPB_BENCH_CUDA_PROFILER") == "1"  # ncu --profile-from-start off: capture the timed region only
    if cuda_profiler:
        torch.cuda.profiler.start()
    ev0.record(stream)
    tb.train_n(K, args.batch)
    ev1.record(stream)
    barrier()
    if cuda_profiler:
        torch.cuda.profiler.stop()
    clk = clocks.stop()
    ms = ev0.elapsed_time(ev1)
    st = tb.stats()
    launches = st["gpu_launches"] - launches0
    loss = tb.loss
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / K
    # whole-job aggregate: every rank processes `batch` compacted samples per step on its own ray shard
    value = world * (args.batch / BATCH) * 1e3 / ms_per_step

    # ---- per-stage device times for the roofline (separate pass with event brackets around every stage) ----
    tb._set("overlap_sampling", 0.0)  # stage times are taken with the stages serialised on one stream (no co-running kernel)
    tb.profile_stages(True)
    tb.stage_times(reset=True)
    n_prof = min(K, 64)
    tb.train_n(n_prof, args.batch)
    stages = tb.stage_times(reset=True)
    tb.profile_stages(False)
    pk = peaks()
    # parameters whose gradient was non-zero in one step, measured: Adam advances a per-parameter step counter only for those (adam.h:76-79,:104)
    n_params = tb.n_params
    touched = None
    if world == 1:
        ps0 = np.empty(n_params, np.uint32); ps1 = np.empty(n_params, np.uint32)
        pyngp.check(pyngp.lib().ngpb_testbed_get_optimizer_state(tb._h, None, None, ps0.ctypes.data_as(C.c_void_p)))
        tb.train(args.batch)
        pyngp.check(pyngp.lib().ngpb_testbed_get_optimizer_state(tb._h, None, None, ps1.ctypes.data_as(C.c_void_p)))
        touched = int((ps0 != ps1).sum())
    stage_report = {}
    total_stage_ms = sum(v[0] for v in stages.values()) or 1.0
    rays_per_call = stages["sampling"][2] / max(stages["sampling"][1], 1)
    samples_per_call = stages["encode_inference"][2] / max(stages["encode_inference"][1], 1)
    sm_hz = (clk.get("sm_mhz") or pk.get("sm_max_mhz", 1965.0)) * 1e6
    for name, (sms, calls, units) in stages.items():
        if calls == 0:
            continue
        rep = dict(ms_per_call=sms / calls, calls_per_step=calls / n_prof, share=sms / total_stage_ms, units_per_call=units / calls)
        sec = sms * 1e-3 / calls
        if name in STAGE_ALGO and sms > 0:
            bound, per_unit, unit = STAGE_ALGO[name]
            achieved = (units / calls) * per_unit / sec / (1e9 if bound == "hbm" else 1e12)
            peak = pk["hbm"] if bound == "hbm" else pk["tensor_sustained"]
            rep.update(bound=bound, achieved=achieved, peak=peak, unit=unit, frac=achieved / peak, algorithmic_per_unit=per_unit)
            if name == "encode_backward":  # the fp32 scatter-adds this build executes, next to (never instead of) the reference's fp16 traffic model
                rep.update(achieved_as_executed=(units / calls) * HASH_BWD_BYTES_AS_EXECUTED / sec / 1e9, frac_as_executed=(units / calls) * HASH_BWD_BYTES_AS_EXECUTED / sec / 1e9 / peak)
            if name.startswith("encode"):  # the 24 MB table is L2-resident: the same bytes against the L2's measured ceiling (HBM is the wrong roof for this kernel)
                l2_peak = L2_PEAK_BYTES_PER_CLK * sm_hz / 1e9
                rep.update(l2=dict(achieved=achieved, peak=l2_peak, unit="GB/s", frac=achieved / l2_peak, peak_source="6300 B/clk (B300_MICROARCH.md LTS cap) x SM clock of this run"))
        else:
            nbytes = stage_bytes(name, rays_per_call, samples_per_call, args.batch, n_params, touched if touched is not None else 0)
            if nbytes is not None and sms > 0 and not (name == "optimizer" and touched is None):
                achieved = nbytes / sec / 1e9
                rep.update(bound="hbm", achieved=achieved, peak=pk["hbm"], unit="GB/s", frac=achieved / pk["hbm"], algorithmic_bytes_per_launch=nbytes)
                if name == "optimizer":
                    rep.update(touched_params=touched, touched_fraction=touched / n_params, model="10 B x n_params + 34 B x touched (SURVEY.md s8d); touched measured from the per-parameter step counters")
        stage_report[name] = rep
    # Dominant KERNEL of the step: the first line of the committed ncu launch list of this same command (profiles/r02_launches_summary.txt, shares of the
    # serialised kernel time), mapped to the stage it belongs to -- a stage such as `sampling` is five kernels, none of which is the step's largest. The
    # stage's live event time in THIS run then gives `achieved`; every stage, `sampling` included, keeps its own entry in per_stage. Without the file: the
    # single-kernel stage with the largest share of this run.
    dom, dom_kernel, dom_share_ncu = None, None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_launches_summary.txt")) as f:
            for line in f:
                parts = line.split()
                if len(parts) >= 4 and parts[0].endswith("%"):
                    kname = " ".join(parts[3:]).replace("void ", "")
                    stage = next((s_ for pat, s_ in KERNEL_STAGE if pat in kname), None)
                    if stage in stage_report and "frac" in stage_report[stage]:
                        dom, dom_kernel, dom_share_ncu = stage, kname, float(parts[0].rstrip("%")) / 100.0
                    break
    except (OSError, ValueError):
        pass
    if dom is None:
        single = [n for n in stage_report if "frac" in stage_report[n] and n in ("encode_inference", "encode_backward", "optimizer", "mlp_inference")]
        dom = max(single or [n for n in stage_report if "frac" in stage_report[n]], key=lambda n: stage_report[n]["share"])
    d = stage_report[dom]
    # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the stages' kernels from the committed `ncu --set full` capture of this same
    # command (profiles/r02_traffic.json, written by tools/ncu_summary.py): a profiler number, read from the committed file because nothing timed may run under ncu
    traffic, traffic_src = None, None
    for tf in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", tf)) as f:
                tj = json.load(f)
            for n2 in stage_report:
                if n2 in tj.get("stages", {}):
                    stage_report[n2].setdefault("traffic", float(tj["stages"][n2]["dram_bytes_per_launch"]))
            if dom in tj.get("stages", {}) and traffic is None:
                traffic, traffic_src = float(tj["stages"][dom]["dram_bytes_per_launch"]), f"profiles/{tf}: " + str(tj.get("source"))
        except (OSError, ValueError, KeyError):
            pass
    roofline = dict(kernel=dom if dom_kernel is None else f"{dom_kernel} (stage {dom})", kernel_share_ncu=dom_share_ncu,
                    largest_stage=max((n for n in stage_report if "frac" in stage_report[n]), key=lambda n: stage_report[n]["share"]), bound=d["bound"], achieved=d["achieved"], peak=d["peak"], unit=d["unit"], frac=d["frac"], traffic=traffic, traffic_source=traffic_src,
                    algorithmic_bytes_per_launch=d.get("algorithmic_bytes_per_launch", (d["units_per_call"] * STAGE_ALGO[dom][1]) if dom in STAGE_ALGO and d["bound"] == "hbm" else None),
                    peak_source=f"MEASURED_PEAKS.json ({pk['src']}; {'hbm_gbs' if d['bound'] == 'hbm' else 'bf16_tflops_sustained'})",
                    share_of_step=d["share"], per_stage=stage_report)
    if "l2" in d:
        roofline["l2"] = d["l2"]

    # ---- render leg (second half of BASELINE.json's metric: "render Msamples/sec"): one 800x800 frame of the trained model, device-timed ----
    render = None
    try:
        import math
        cam = synthetic.nerf_matrix_to_ngp(synthetic.hemisphere_cameras(7, seed=3)[2])
        tb.camera_matrix = cam
        tb.fov_axis = 0
        tb.fov = math.degrees(synthetic.CAMERA_ANGLE_X)
        tb.render(args.res, args.res, 1, True)  # warm-up (workspace allocation)
        ms_r, ns_r = [], []
        for _ in range(5):
            tb.render(args.res, args.res, 1, True)
            ms_r.append(tb.last_render_ms); ns_r.append(tb.last_render_samples)
        render = dict(metric="nerf_render_msamples_per_sec", value=float(np.median(ns_r)) / (float(np.median(ms_r)) * 1e-3) / 1e6, unit="Msamples/s",
                      ms_per_frame=float(np.median(ms_r)), samples_per_frame=int(np.median(ns_r)), resolution=[args.res, args.res], spp=1,
                      note="network-evaluated samples of live rays / device time of Testbed.render (CUDA events around the whole call, incl. the frame's device-to-host copy)")
    except Exception as e:  # the render leg never invalidates the training number
        render = dict(error=str(e))
    # Blender path (request_nerf_render_sync, K18): the same frame through a snapshot -> RenderRequest with one NeRF, host wall clock around the call
    # (the call is synchronous and ends with the frame in host memory; the snapshot load is cached after the warm-up frame)
    if world == 1 and isinstance(render, dict) and "error" not in render:  # (single process only: a snapshot of a data-parallel run is a collective)
        try:
            snap_path = os.path.join(tempfile.gettempdir(), f"ngpb_bench_{os.getpid()}.msgpack")
            tb.save_snapshot(snap_path)
            res2 = (args.res, args.res)
            out_p = pyngp.RenderOutputProperties(res2, pyngp.DownsampleInfo.MakeFromMip(res2, 0), 1, pyngp.ColorSpace.SRGB, pyngp.TonemapCurve.Identity, 0.0, [0, 0, 0, 0], False)
            cam_p = pyngp.RenderCameraProperties(cam, pyngp.CameraModel.Perspective, scene["fx"] / RES * args.res, 0.0, 0.0, 1.0, None, None)
            box = pyngp.BoundingBox([0, 0, 0], [1, 1, 1])
            rq = pyngp.RenderRequest(out_p, cam_p, pyngp.RenderModifiers([]), [pyngp.NerfDescriptor(snap_path, box, np.eye(4), pyngp.RenderModifiers([]), 1.0)], box)
            tb.request_nerf_render_sync(rq)
            ms_b, ns_b = [], []
            for _ in range(5):
                tb0 = time.perf_counter()
                tb.request_nerf_render_sync(rq)
                ms_b.append(1e3 * (time.perf_counter() - tb0)); ns_b.append(tb.last_render_samples)
            render["blender"] = dict(metric="blender_render_msamples_per_sec", value=float(np.median(ns_b)) / (float(np.median(ms_b)) * 1e-3) / 1e6, unit="Msamples/s",
                                     ms_per_frame=float(np.median(ms_b)), samples_per_frame=int(np.median(ns_b)), waves_launches=tb.last_render_launches,
                                     note="Testbed.request_nerf_render_sync, one NeRF from a snapshot, host wall clock around the synchronous call")
            os.remove(snap_path)
        except Exception as e:
            render["blender"] = dict(error=str(e))

    # ---- PSNR leg (third part of BASELINE.json's metric: "PSNR vs ref"): the reference's own evaluation procedure (scripts/run.py:216-303) on held-out views of
    # the model as trained so far: black background, snap_to_pixel_centers, 8 spp, render_min_transmittance 1e-4, linear render, PSNR of the sRGB-clipped images ----
    psnr = None
    try:
        import math
        def to_srgb(x):
            return np.where(x < 0.0031308, 12.92 * x, 1.055 * np.maximum(x, 1e-8) ** (1 / 2.4) - 0.055)
        def to_linear(x):
            return np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)
        steps_trained = tb.training_step
        tb.background_color = [0.0, 0.0, 0.0, 1.0]
        tb.snap_to_pixel_centers = True
        tb.nerf.render_min_transmittance = 1e-4
        tb.fov_axis = 0
        tb.fov = math.degrees(synthetic.CAMERA_ANGLE_X)
        fx_eval = 0.5 * args.res / math.tan(0.5 * synthetic.CAMERA_ANGLE_X)
        views, vals, mses = synthetic.hemisphere_cameras(4, seed=11), [], []
        for c in views:
            ngp_cam = synthetic.nerf_matrix_to_ngp(c)
            gt8 = synthetic.render_image(ngp_cam, args.res, fx_eval, fx_eval, synthetic.lego_boxes(), device=f"cuda:{local_rank}").cpu().numpy().astype(np.float32) / 255.0
            ref_lin = to_linear(gt8[..., :3]) * gt8[..., 3:4]  # scripts/common.py read_image: sRGB -> linear, premultiplied (black background)
            tb.camera_matrix = ngp_cam
            img = tb.render(args.res, args.res, 8, True)
            A, R = np.clip(to_srgb(img[..., :3]), 0.0, 1.0), np.clip(to_srgb(ref_lin), 0.0, 1.0)
            mse = float(np.mean((A - R) ** 2))
            mses.append(mse); vals.append(10.0 * math.log10(1.0 / max(mse, 1e-12)))
        psnr = dict(metric="psnr_db_held_out_views", value=float(np.mean(vals)), min=float(min(vals)), max=float(max(vals)), psnr_of_mean_mse=10.0 * math.log10(1.0 / max(float(np.mean(mses)), 1e-12)),
                    views=len(vals), resolution=[args.res, args.res], spp=8, steps_trained=steps_trained,
                    procedure="scripts/run.py:216-303 (black background, snap_to_pixel_centers, render_min_transmittance 1e-4, PSNR of sRGB-clipped frames) against the analytic ground truth of the synthetic scene")
        tb.snap_to_pixel_centers = False
        tb.nerf.render_min_transmittance = 0.01
    except Exception as e:  # the PSNR leg never invalidates the training number
        psnr = dict(error=str(e))
    # the unmodified reference on the same GPU model, same scene and config (oracle/gen_golden_full.py `big`, committed numbers: a builder-run context figure, not a bench arm)
    reference_on_b200 = None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_reference_full.json")) as f:
            rj = json.load(f)
        reference_on_b200 = dict(it_per_s=rj["reference"]["it_per_s"], ms_per_step=rj["reference"]["ms_per_step"], after_steps=rj["steps"],
                                 psnr_db_held_out_views=rj.get("mean_psnr_reference_vs_gt"), psnr_this_repo_same_run=rj.get("mean_psnr_ours_trained_vs_gt"),
                                 psnr_this_repo_vs_reference_render_of_the_same_snapshot=rj.get("mean_psnr_ours_vs_reference_same_weights"),
                                 l1_this_repo_vs_reference_render_of_the_same_snapshot=rj.get("mean_l1_ours_vs_reference_same_weights"),
                                 source="profiles/r02_reference_full.json: ngp::Testbed compiled headless from /root/reference for sm_100 (oracle/Makefile.full), run on a B200 by oracle/gen_golden_full.py")
    except (OSError, ValueError, KeyError):
        pass

    # ---- e2e arm: public pyngp surface, dataset starts in pinned host memory, loss read back every step ----
    # (same Testbed object: reloading a same-sized dataset re-uploads it and re-initialises the model without new allocations)
    tb._set("overlap_sampling", 1.0)
    barrier()
    t0 = time.perf_counter()
    tb.load_training_images(list(images_np), scene["xforms"], scene["fx"], scene["fy"])  # H2D of the whole dataset, inside the timed region
    t_load = time.perf_counter() - t0
    tb.train_n(args.preroll, args.batch)  # untimed pre-roll to the same regime as above
    for _ in range(W):
        tb.train(args.batch); _ = tb.loss
    s0 = tb.stats()
    barrier()
    t1 = time.perf_counter()
    e2e_losses = []
    for _ in range(K):
        tb.train(args.batch)      # one optimizer step through the public call; syncs and reads the counters (+ loss every 16th step) back
        e2e_losses.append(tb.loss)  # host-side read of the step's result
    torch.cuda.synchronize()
    t_steps = time.perf_counter() - t1
    s1 = tb.stats()
    e2e_s = t_load + t_steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = dict(value=world * (args.batch / BATCH) * K / e2e_s, unit=UNIT,
               h2d_bytes_per_step=(dataset_bytes + 0.0) / K, d2h_bytes_per_step=(s1["d2h_bytes"] - s0["d2h_bytes"]) / K,
               dataset_upload_ms=1e3 * t_load, ms_per_step_host=1e3 * t_steps / K, mean_loss=float(np.mean(e2e_losses)) if e2e_losses else None)
    del tb

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_baseline(images_np, scene["xforms"], scene["fx"], scene["fy"], seconds=args.cpu_seconds)
        cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "fp16 storage / fp32 accumulate (tcgen05 kind::f16)",
            "data": "synthetic",
            "config": {"workload": f"NeRF Lego-shaped synthetic scene ({args.n_images} cams {args.res}x{args.res} RGBA8), configs/nerf/base.json, "
                                   f"batch 2^{int(np.log2(args.batch))} compacted samples/iteration/GPU, seed 1337",
                       "batch": args.batch, "global_batch": global_batch, "optimizer_steps_per_sec": 1e3 / ms_per_step, "rays_per_batch": st["rays_per_batch"], "samples_before_compaction": st["measured_batch_size_before_compaction"],
                       "samples_per_sec": value * BATCH, "pre_trained_steps": args.preroll + W, "final_loss": loss,
                       "parallelism": "single GPU" if world == 1 else f"dp{world}: ray-sharded replicas; gradients reduce-scattered (bf16), Adam on 1/{world} of the parameters per rank, "
                                                                                f"fp16 weights all-gathered ({'peer-memory kernels over NVLink' if os.environ.get('NGPB_DP_EXCHANGE') == '1' else 'NCCL'})",
                       "l2": "no flush: the per-iteration working set (256 MB images + 293 MB parameter/optimizer state + ~150 MB sample buffers) exceeds the 126 MB L2",
                       "schedule": "every stage of the reference's iteration runs every step (K1, inference on all marched samples, K6, forward+backward on the compacted batch, "
                                   "Adam/EMA, occupancy refresh at its cadence); the training pass reads the hash-grid features the inference pass computed for the same samples "
                                   "with the same weights instead of re-encoding them (bit-identical, reuse_encoding=0 restores the second encode)"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clk, "roofline": roofline, "render": render, "psnr": psnr,
        }
        if reference_on_b200 is not None:
            line["reference_on_b200"] = reference_on_b200
        if cb is not None:
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
