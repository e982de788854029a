#!/usr/bin/env python
"""bench.py -- tracker.update() frames/sec at 256 dets/frame (BASELINE.json metric), B200 arm and CPU reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config.workload): BASELINE.json configs[1] -- BoT-SORT + OSNet_x0_25 ReID inside update(), one
1280x720 stream of 256 detections per frame per GPU, reference bench generator (benchmark_fps.py:60-94),
YAML-default parameters with CMC off, float32 kernels.  N GPUs = N independent streams, one process per GPU
(weak scaling, no collective on the frame path; NCCL only for the barrier and the max-over-ranks gather).

One "step" = one frame through the whole hot path (crop staging, ReID CNN, appearance cost, Kalman
predict/update, three assignment rounds, lifecycle, output rows).
  value : frames/s with frames and detections resident in HBM (ring of distinct frames larger than L2),
          timed with CUDA events on the engine's stream, max over ranks.
  e2e   : same metric through the public API `MultiStreamTracker.update(dets, imgs)` with HOST numpy buffers:
          every step copies the frame + detections host->device and reads the result rows back.
  roofline : the dominant kernel class by device time, from a CUDA-event profiling pass inside this script.
  cpu_baseline : the oracle port of the reference path (numpy/scipy/torch-CPU restatement pinned to the reference
          by tests/golden) on this box's host cores, on a bounded sample of the same workload.
`--impl reference` runs that CPU arm alone and prints the same line shape.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

LOCAL_RANK = int(os.environ.get("LOCAL_RANK", "0"))
RANK = int(os.environ.get("RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
if WORLD > 1 and "BENCH_KEEP_VISIBLE" not in os.environ:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    ids = vis.split(",") if vis else [str(i) for i in range(64)]
    os.environ["CUDA_VISIBLE_DEVICES"] = ids[LOCAL_RANK % len(ids)]  # one GPU per process, device 0 inside it

import numpy as np  # noqa: E402

N_DETS = 256
IMG_HW = (720, 1280)
RING = 64  # distinct frames in the input ring: 64 x 2.76 MB = 177 MB > 126 MB of L2
BOTSORT = dict(
    track_high_thresh=0.6296854875023994, track_low_thresh=0.1014392537025336,
    new_track_thresh=0.6246494191492591, track_buffer=40, match_thresh=0.7722224024589055,
    proximity_thresh=0.6084297894561342, appearance_thresh=0.6188818853936099,
    unconfirmed_emb_scale=2.5445206391993294, second_match_thresh=0.28795081514328974,
    unconfirmed_match_thresh=0.41148010638233784, removed_stracks_buffer=329, fuse_first_associate=True,
    frame_rate=30, with_reid=True)
METRIC = "tracker.update() frames/sec at 256 dets/frame"
WORKLOAD = "BoT-SORT + OSNet_x0_25 ReID in update(), 1 stream x 256 dets/frame per GPU, 1280x720, CMC off"
CLASSES = ["crop_resize_norm", "stem_conv7x7", "maxpool", "pointwise_gemm", "lightconv", "gates", "avgpool", "head",
           "association"]


# ------------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------------
def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=float(d["hbm_gbs"]), tensor_tflops=float(d["bf16_tflops_sustained"]), source="measured")
    return dict(hbm_gbs=6650.0, tensor_tflops=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self):
        self.rows = []
        self.proc = None

    def start(self):
        gpu = os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0]
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", gpu], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def osnet_class_costs(ch=(16, 64, 96, 128), feat=512):
    """Algorithmic MACs and compulsory float32 bytes PER CROP for each kernel class of csrc/reid_model.cu."""
    mac = {c: 0.0 for c in CLASSES}
    byt = {c: 0.0 for c in CLASSES}
    mac["stem_conv7x7"] = 128 * 64 * 147 * ch[0]
    byt["stem_conv7x7"] = 4 * (256 * 128 * 3 + 128 * 64 * ch[0])
    byt["crop_resize_norm"] = 70 * 140 * 3 + 4 * 256 * 128 * 3
    byt["maxpool"] = 4 * (128 * 64 * ch[0] + 64 * 32 * ch[0])
    hw = 64 * 32
    for s in range(3):
        for j in range(2):
            cin = ch[s] if j == 0 else ch[s + 1]
            cout = ch[s + 1]
            mid = cout // 4
            mac["pointwise_gemm"] += hw * cin * mid
            byt["pointwise_gemm"] += 4 * hw * (cin + mid)
            mac["lightconv"] += 10 * hw * (mid * mid + 9 * mid)
            byt["lightconv"] += 10 * 4 * hw * 2 * mid
            k = mid + (cin if cin != cout else 0)
            mac["pointwise_gemm"] += hw * k * cout + 4 * hw * mid
            byt["pointwise_gemm"] += 4 * hw * (4 * mid + cin + cout)
            byt["gates"] += 4 * 4 * mid * 2
        if s < 2:
            c = ch[s + 1]
            mac["pointwise_gemm"] += hw * c * c
            byt["pointwise_gemm"] += 4 * hw * 2 * c
            byt["avgpool"] += 4 * hw * c * 1.25
            hw //= 4
    mac["pointwise_gemm"] += hw * ch[3] * ch[3]
    byt["pointwise_gemm"] += 4 * hw * 2 * ch[3]
    mac["head"] = ch[3] * feat
    byt["head"] = 4 * (hw * ch[3] + feat)
    return mac, byt


def make_inputs(stream_index: int, frames: int):
    from boxmot_b200.synthetic import bench_stream

    _, dets = bench_stream(N_DETS, frames, hw=IMG_HW, stream=stream_index)
    rng = np.random.default_rng(9000 + stream_index)
    imgs = rng.integers(0, 255, size=(RING, IMG_HW[0], IMG_HW[1], 3), dtype=np.uint8)
    return imgs, dets


def make_blob(tmpdir: Path) -> Path:
    from boxmot_b200.synthetic import make_osnet_state
    from boxmot_b200.weights import export_blob

    return export_blob(make_osnet_state("osnet_x0_25", seed=0), tmpdir / "osnet_x0_25_synthetic.b200reid")


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_arm(sample_frames: int, warm: int, budget_s: float = 25.0, stream_index: int = 0, threads: int = 0):
    """Oracle port of the reference path on the host cores, bounded by wall-clock: the host of a GPU box can be
    anything from 8 fast cores to a heavily shared 128-thread part, so the sample is 'as many frames as fit in
    `budget_s` seconds' (at least one), after a thread-count probe that is itself time-bounded."""
    import torch

    from boxmot_b200.synthetic import make_osnet_state
    from oracle import reid as orid
    from oracle.trackers import BotSortOracle

    sd = make_osnet_state("osnet_x0_25", seed=0)
    imgs, dets = make_inputs(stream_index, warm + sample_frames + 1)
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    probe_x = orid.get_crops(dets[0][:16, :4], imgs[0])
    best = (min(avail, 8), 1e30)
    t_probe = time.perf_counter()
    # `threads` > 0: one of several concurrent stream workers with a fixed share of the cores (no probe)
    for th in ([threads] if threads > 0 else sorted({min(avail, 8), min(avail, 16), min(avail, 32), avail})):
        torch.set_num_threads(th)
        t0 = time.perf_counter()
        orid.osnet_forward(sd, probe_x)
        dt = time.perf_counter() - t0
        if dt < best[1]:
            best = (th, dt)
        if time.perf_counter() - t_probe > 8.0:
            break
    cores = best[0]
    torch.set_num_threads(cores)
    trk = BotSortOracle(reid_model=orid.OracleReID(sd), **BOTSORT)
    t_start = time.perf_counter()
    done_warm = 0
    for f in range(warm):
        trk.update(dets[f], imgs[f % RING])
        done_warm += 1
        if time.perf_counter() - t_start > budget_s / 2:
            break
    t0 = time.perf_counter()
    n = 0
    for f in range(done_warm, done_warm + sample_frames):
        trk.update(dets[f], imgs[f % RING])
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": "port", "frames": n, "warm": done_warm,
            "sample": f"{n} frames of the same 256-det stream after {done_warm} warm-up frame(s) (time-bounded to "
                      f"~{budget_s:.0f} s), oracle port (numpy/scipy/lapjv-C + torch-CPU OSNet fp32, {cores} of {avail} "
                      f"usable threads, fastest of a bounded thread-count probe)",
            "ms_per_frame": 1e3 * dt / n}


def _cpu_stream_worker(job):
    steps, stream_index, threads = job
    return cpu_arm(steps, 1, budget_s=40.0, stream_index=stream_index, threads=threads)


def run_reference(args):
    """The reference path on the host cores for the SAME workload as the B200 arm at this --gpus: one 256-detection
    stream per GPU, i.e. N independent streams.  N == 1: one tracker with the fastest thread count of a bounded probe.
    N > 1: N concurrent tracker processes (the reference's own replay parallelism is one process per sequence,
    engine/eval/replay.py:27-115), each with an equal share of the usable cores; value = sum of the streams' rates."""
    if RANK != 0:
        return
    steps = max(2, min(args.steps, 12))
    n_streams = max(1, int(args.gpus))
    if n_streams == 1:
        base = cpu_arm(steps, 1, budget_s=40.0)
        note = "single stream on the host cores; steps bounded to keep the run short"
    else:
        import multiprocessing as mp

        avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        threads = max(1, avail // n_streams)
        with mp.get_context("spawn").Pool(n_streams) as pool:
            parts = pool.map(_cpu_stream_worker, [(steps, i, threads) for i in range(n_streams)])
        value = sum(p["value"] for p in parts)
        frames = min(p["frames"] for p in parts)
        base = {"value": value, "unit": "frames/s", "cores": threads * n_streams, "kind": "port", "frames": frames,
                "warm": min(p["warm"] for p in parts), "ms_per_frame": 1e3 / value,
                "sample": f"{n_streams} concurrent stream processes x {threads} threads, {[p['frames'] for p in parts]} frames "
                          f"each after a warm-up frame (time-bounded to ~40 s), oracle port (numpy/scipy/lapjv-C + "
                          f"torch-CPU OSNet fp32); value = sum of the per-stream rates"}
        note = f"{n_streams} independent streams on the host cores (one process per stream), like one stream per GPU"
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "frames/s", "n_gpus": args.gpus,
            "steps": base["frames"], "warmup": base["warm"], "ms_per_step": base["ms_per_frame"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": note},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch

    import boxmot_b200 as bb
    from boxmot_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: boxmot_b200 has no CPU fallback")
    torch.cuda.set_device(0)
    dist = None
    if WORLD > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda:0"))
    lib = _lib.require_device()
    K, Wm = args.steps, max(3, args.warmup)
    tmp = Path(tempfile.mkdtemp(prefix="b200bench_"))
    blob = make_blob(tmp)
    S = 1
    imgs, dets = make_inputs(RANK, Wm + K + 8)
    n_first = float(np.mean([(d[:, 4].astype(np.float64) > BOTSORT["track_high_thresh"]).sum() for d in dets]))

    def new_tracker():
        return bb.MultiStreamTracker("botsort", n_streams=S, cap_tracks=1024, cap_dets=N_DETS, feat_dim=512,
                                     reid_blob=str(blob), **BOTSORT)

    # ---------------- value: inputs resident in HBM ----------------
    trk = new_tracker()
    d_imgs = torch.from_numpy(imgs).cuda()                          # [RING][H][W][3] u8
    # frame f uses dets[f] (the stream is a sequence); frames cycle through the ring of distinct images
    d_dets = torch.from_numpy(np.stack([np.pad(d, ((0, N_DETS - len(d)), (0, 0))) for d in dets])[:, None].astype(np.float32)).cuda()
    rows = (ctypes.c_int * S)(N_DETS)
    torch.cuda.synchronize()
    H, Wd = IMG_HW
    img_bytes = H * Wd * 3

    def dev_step(f, sync=0):
        ok = lib.boxmot_b200_tracker_update_device(trk.handle, d_dets[f].data_ptr(), rows, None,
                                                   d_imgs[f % RING].data_ptr(), H, Wd, sync)
        if not ok:
            raise RuntimeError(_lib.last_error(lib))

    for f in range(Wm):
        dev_step(f, 1)
    launches_per_step = trk.last_launches()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    clocks = ClockSampler()
    clocks.start()
    lib.boxmot_b200_tracker_mark(trk.handle, 0)
    for f in range(Wm, Wm + K):
        dev_step(f, 0)
    lib.boxmot_b200_tracker_mark(trk.handle, 1)
    ms = ctypes.c_double(0)
    if not lib.boxmot_b200_tracker_elapsed_ms(trk.handle, ctypes.byref(ms)):
        raise RuntimeError(_lib.last_error(lib))
    torch.cuda.synchronize()
    out_rows = (ctypes.c_int * S)()
    if not lib.boxmot_b200_tracker_fetch(trk.handle, None, None, out_rows):   # surfaces device-side errors
        raise RuntimeError(_lib.last_error(lib))
    value_ms = ms.value

    # ---------------- roofline: profiling pass (events around every launch) ----------------
    phase = (ctypes.c_longlong * 16)()
    lib.boxmot_b200_tracker_phase_clocks(trk.handle, 0, phase, 1)
    lib.boxmot_b200_tracker_profile(trk.handle, 1)
    P = 16
    for f in range(Wm + K - P, Wm + K):
        dev_step(f % (Wm + K), 1)
    cls_ms = (ctypes.c_double * 9)()
    cls_n = (ctypes.c_int * 9)()
    lib.boxmot_b200_tracker_profile_read(trk.handle, cls_ms, cls_n)
    lib.boxmot_b200_tracker_profile(trk.handle, 0)
    lib.boxmot_b200_tracker_phase_clocks(trk.handle, 0, phase, 1)
    phase_names = ["split+predict", "cost1", "assign1", "update1", "round2", "round3", "births+lists", "dups+output"]
    assoc_phases = {n: phase[i] / P for i, n in enumerate(phase_names)}
    clock_info = clocks.stop()
    prof = {CLASSES[i]: {"ms_per_step": cls_ms[i] / P, "launches_per_step": cls_n[i] / P} for i in range(9)}
    trk.close()

    # ---------------- e2e: public API, host buffers ----------------
    # the frame ring lives in page-locked host memory (as a capture pipeline would hand frames over): update()
    # copies each frame host->device inside the timed region, straight from that buffer
    imgs_pinned = torch.from_numpy(imgs).pin_memory()
    imgs = imgs_pinned.numpy()
    trk = new_tracker()
    for f in range(Wm):
        trk.update([dets[f]], [imgs[f % RING]])
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    last = None
    for f in range(Wm, Wm + K):
        last = trk.update([dets[f]], [imgs[f % RING]])
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    n_out = int(len(last[0]))
    trk.close()

    # ---------------- reduce over ranks ----------------
    from boxmot_b200 import sharding

    value_ms, e2e_ms = sharding.reduce_max([value_ms, e2e_ms], dist, device="cuda")  # slowest rank defines the job
    frames_per_rank = sharding.gather_counts(S * K, dist, device="cuda")
    if RANK != 0:
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    mac, byt = osnet_class_costs()
    dom = max((c for c in CLASSES if c != "association"), key=lambda c: prof[c]["ms_per_step"])
    crops = n_first
    dom_bytes = byt[dom] * crops
    dom_s = prof[dom]["ms_per_step"] * 1e-3
    reid_ms = sum(prof[c]["ms_per_step"] for c in CLASSES if c != "association")
    fps = sum(frames_per_rank) / (value_ms * 1e-3)
    e2e_fps = sum(frames_per_rank) / (e2e_ms * 1e-3)
    total_flop = 2 * sum(mac.values()) * crops
    traffic = None   # DRAM bytes of the dominant class per step, from the committed ncu capture (same shapes)
    try:
        tj = json.loads((ROOT / "profiles" / "traffic.json").read_text())
        if dom in tj and tj[dom].get("crops_per_step") == crops:
            traffic = tj[dom]["dram_bytes_per_step"]
    except Exception:
        traffic = None
    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": WORLD, "steps": K, "warmup": Wm,
        "ms_per_step": value_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "streams_per_gpu": S, "dets_per_frame": N_DETS,
                   "first_round_crops_per_frame": crops, "reid": "osnet_x0_25 random-init (seed 0), fp32 kernels",
                   "l2": f"input ring of {RING} distinct frames ({RING * img_bytes / 1e6:.0f} MB) and a per-chunk "
                         f"activation workspace larger than L2; no explicit flush", "parallelism": f"streams x{WORLD}",
                   "value_path": "update_device, device-resident inputs, no per-frame sync: ReID of frame f+1 overlaps "
                                 "the association of frame f on two CUDA streams",
                   "e2e_path": "MultiStreamTracker.update per frame: frame H2D from page-locked host memory + dets H2D, "
                               "full sync, rows D2H"},
        "e2e": {"value": e2e_fps, "unit": "frames/s", "ms_per_step": e2e_ms / K,
                "h2d_bytes_per_step": S * (img_bytes + N_DETS * 6 * 4 + 4),
                "d2h_bytes_per_step": S * (N_DETS * 8 * 4 + 16 * 4), "rows_last_frame": n_out},
        "gpu_launches": launches_per_step * K,
        "launches_per_step": launches_per_step,
        "clocks": clock_info,
        "roofline": {"kernel": dom, "bound": "hbm", "achieved": dom_bytes / dom_s / 1e9, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": dom_bytes / dom_s / 1e9 / peaks["hbm_gbs"], "traffic": traffic,
                     "peak_source": peaks["source"], "algorithmic_bytes_per_step": dom_bytes,
                     "ms_per_step": prof[dom]["ms_per_step"], "share_of_reid": prof[dom]["ms_per_step"] / max(reid_ms, 1e-9)},
        "reid_conv_roofline": {"achieved_tflops": total_flop / (reid_ms * 1e-3) / 1e12, "peak_tflops": peaks["tensor_tflops"],
                               "frac_of_tensor_peak": total_flop / (reid_ms * 1e-3) / 1e12 / peaks["tensor_tflops"],
                               "flop_per_crop": 2 * sum(mac.values()), "reid_ms_per_step": reid_ms},
        "kernel_classes": prof,
        "association_phase_sm_clocks_per_step": assoc_phases,
    }
    if WORLD == 1 and not args.skip_cpu:
        line["cpu_baseline"] = cpu_arm(args.cpu_frames, 1)
        line["speedup_e2e_vs_cpu"] = e2e_fps / line["cpu_baseline"]["value"]
    print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=8)
    ap.add_argument("--skip-cpu", action="store_true", help="kernel A/B experiments only: omit the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
