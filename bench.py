#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json on synthetic yuv420p.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1] — 1080p30 H.264 encode of synthetic
yuv420p, GOP=60, CAVLC, I+P frames, deblocking on, constant QP.  A "step" is one pass of the
hot path (K1 colour/pad -> K2 motion search -> K3 transform/quant/recon -> K4 deblock -> K5
CAVLC + NAL pack) over one batch of `--gops` closed GOPs per GPU.
  value : encoded frames/s, whole job, raw frames already resident in HBM, CUDA-event timed
  e2e   : the same through the public Session API with pinned HOST buffers: H2D of the frames and
          D2H of the bitstream inside the timed region
  roofline     : the dominant kernel of the step, algorithmic bytes / CUDA-event time
  cpu_baseline : the CPU oracle (a port; the reference's libx264 is not in the image) on a bounded
                 sample of the same workload, timed on this box's host cores
Multi-GPU: GOPs are independent -> each rank encodes its own GOPs, no data-path collective,
weak scaling (per-GPU batch fixed).  Time = max over ranks.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # see video_codec_pipeline_b200/csrc/encoder.cu

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from video_codec_pipeline_b200 import synth  # noqa: E402

W, H, FPS, GOP = 1920, 1080, 30, 60
QP_I, QP_P = 25, 27
SEED = 1080
ENTROPY = 0
SLICES = 1
T8X8 = 0
CODEC = 0          # 0 H.264, 1 HEVC (--codec hevc: BASELINE.json configs[3], one GPU's GOP shard)
HEVC_SAO = 0       # --hevc-sao: sample adaptive offset (bit-exact; its first kernel is not yet tuned)
METRIC = "1080p H.264 encode fps (GOP=60, CAVLC, I+P)"
WORKLOAD = "configs[1]: 1080p30 yuv420p, GOP=60, CAVLC, I+P, deblock, CQP 25/27"


def select_workload(name: str, entropy: int, slices: int = -1, codec: str = "h264"):
    """Default = BASELINE.json configs[1].  `4k` = the single-GPU shard of configs[2]: 4K60, High
    profile (CABAC + 8x8 transform) unless --entropy 0 asks for the CAVLC/Baseline variant."""
    global W, H, FPS, SEED, ENTROPY, METRIC, WORKLOAD, SLICES, T8X8, CODEC
    if codec == "hevc":
        # configs[3]: the h265-* presets' path.  Stream structure of csrc/k6_hevc.cu: Main profile, 16x16 coding
        # units, 8x8 transforms, half-sample motion, CABAC, in-loop deblocking (+ SAO with --hevc-sao; DESIGN.md 1, "HEVC")
        CODEC = 1
        if name == "4k":
            W, H, FPS, SEED = 3840, 2160, 60, 2160
        ENTROPY, T8X8 = 1, 0
        mbh = (H + 15) // 16
        SLICES = slices if slices >= 0 else max(1, mbh // 17)
        METRIC = "%s HEVC encode fps (GOP=60, Main profile, I+P)" % ("4K" if name == "4k" else "1080p")
        WORKLOAD = "%s: %dx%d@%d yuv420p, HEVC Main, GOP=60, CABAC, %d slice%s, I+P, half-sample motion, deblock%s, CQP %d/%d" % (
            "configs[3] (one GPU's GOP shard)" if name == "4k" else "configs[3] at 1080p", W, H, FPS, SLICES, "" if SLICES == 1 else "s",
            ", SAO" if HEVC_SAO else "", QP_I, QP_P)
        return
    if name == "4k":
        W, H, FPS, SEED = 3840, 2160, 60, 2160
        ENTROPY = 1 if entropy < 0 else entropy
        T8X8 = 1 if ENTROPY else 0
    else:
        ENTROPY = 0 if entropy < 0 else entropy
    coder = "CABAC" if ENTROPY else "CAVLC"
    mbh = (H + 15) // 16
    # slices: the encoder's own choice unless given (vcp_algo.h vcp_auto_slices: CAVLC 1; CABAC one per ~17 rows)
    SLICES = slices if slices >= 0 else (max(1, mbh // 17) if ENTROPY else 1)
    METRIC = "%s H.264 encode fps (GOP=60, %s%s, I+P)" % ("4K" if name == "4k" else "1080p", "High profile, " if T8X8 else "", coder)
    WORKLOAD = "%s: %dx%d@%d yuv420p, GOP=60, %s, %d slice%s, I+P, deblock, CQP %d/%d" % (
        ("configs[2] (one GPU's GOP shard, %s profile)" % ("High" if T8X8 else "Baseline")) if name == "4k" else "configs[1]", W, H, FPS, coder,
        SLICES, "" if SLICES == 1 else "s", QP_I, QP_P)


def make_workload(gops: int) -> np.ndarray:
    """`gops` closed GOPs of S1080-style content.  Two distinct GOPs are synthesised (numpy is
    slow at 1080p) and cycled; every GOP is encoded independently so repetition does not make
    the work any cheaper.  Inputs (>= 1.4 GB at 8 GOPs) are far larger than the 126 MB L2."""
    a = synth.make_clip(W, H, GOP, seed=SEED, start=0)
    b = synth.make_clip(W, H, GOP, seed=SEED + 1, start=GOP) if gops > 1 else a
    return np.concatenate([a if (g % 2 == 0) else b for g in range(gops)], axis=0)


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def algorithmic_bytes_per_frame(kernel: str) -> float:
    """compulsory HBM bytes per frame for the stage a kernel belongs to."""
    mbw, mbh = (W + 15) // 16, (H + 15) // 16
    P = 256 * mbw * mbh      # coded luma samples
    nmb = mbw * mbh
    return {
        "csc": 3.0 * W * H,                    # K1: read 1.5WH + write 1.5WH
        "me_prepass": 2.0 * P + 8 * nmb,       # K2: cur + ref luma, vector out
        "me_refine": 2.0 * P + 8 * nmb,
        "p_recon": 7.5 * P,                    # K3: cur 1.5P + ref 1.5P + recon 1.5P + levels 3P
        "i_recon": 7.5 * P,
        "deblock": 3.0 * P,                    # K4
        "pad": 0.0, "mbinfo": 0.0, "rc": 0.0,
        "cavlc_count": 3.0 * P, "cavlc_scan": 0.0, "cavlc_write_pack": 3.0 * P,   # K5: levels 3P
        "hpel": 4.0 * P,                       # K2c: read the reconstruction, write its three half-sample planes
        "cabac_bins": 3.0 * P, "cabac_code": 0.0,   # K5 (CABAC): levels 3P; the coder itself reads 2 B per bin
    }[kernel]


def cpu_port_fps(frames: np.ndarray, threads: int, frames_per_gop: int):
    """Oracle (CPU port) on a bounded sample: `threads` GOP-prefixes in parallel (ctypes drops the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle
    fb = frames.shape[1]
    ngop_avail = frames.shape[0] // GOP
    jobs = []
    for k in range(threads):
        g = k % max(1, ngop_avail)
        jobs.append(frames[g * GOP: g * GOP + frames_per_gop])

    def one(fr):
        p = pyoracle.make_params(W, H, fps=FPS, gop=GOP, qp_i=QP_I, qp_p=QP_P, entropy=ENTROPY, slices=SLICES, transform8x8=T8X8, codec=CODEC,
                                 hevc_subpel=1 if CODEC else 0, hevc_sao=HEVC_SAO if CODEC else 0)
        if CODEC:
            return len(pyoracle.encode_hevc(p, fr)["stream"])
        return len(pyoracle.encode(p, fr, want_recon=False)["stream"])

    pyoracle.lib()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(one, jobs))
    dt = time.perf_counter() - t0
    return threads * frames_per_gop / dt, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path is libx264 inside a system
    ffmpeg (cmd/consumer.go:382); neither exists in this image and the reference has no codec
    source to compile, so this arm times the oracle port on all host cores."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 32))
    fpg = 30                               # IDR + 29 P per GOP prefix, per thread and step
    base = make_workload(2)
    times, nframes = [], threads * fpg
    for i in range(args.warmup + args.steps):
        fps, dt = cpu_port_fps(base, threads, fpg)
        if i >= args.warmup:
            times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    value = nframes / (ms / 1000.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": nframes, "seed": SEED},
        "cpu_baseline": {"value": round(value, 3), "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d threads x first %d frames of a GOP (IDR+P), oracle/%s; libx264/libx265/ffmpeg absent from image" % (threads, fpg, "hevc_oracle.inc.c" if CODEC else "h264_oracle.c")},
        "e2e": {"value": round(value, 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gops", type=int, default=32, help="closed GOPs per GPU per step (weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="1080p", choices=["1080p", "4k"], help="default: BASELINE.json configs[1]")
    ap.add_argument("--codec", default="h264", choices=["h264", "hevc"], help="hevc: BASELINE.json configs[3] (use with --workload 4k)")
    ap.add_argument("--hevc-sao", action="store_true", help="HEVC: sample adaptive offset on (default off: its first kernel is slow)")
    ap.add_argument("--entropy", type=int, default=-1, help="0 CAVLC, 1 CABAC (default: what the workload names)")
    ap.add_argument("--slices", type=int, default=-1, help="slices per picture (default: the encoder's choice)")
    ap.add_argument("--e2e-threads", type=int, default=2, help="host threads (sessions) of the end-to-end pipeline, like consumer -j")
    ap.add_argument("--deblock-idc", type=int, default=0, help="experiments only: 1 switches the in-loop filter off")
    args = ap.parse_args()
    global HEVC_SAO
    HEVC_SAO = 1 if args.hevc_sao else 0
    select_workload(args.workload, args.entropy, args.slices, args.codec)
    if args.workload == "4k" and args.gops == 32:
        args.gops = 16                     # 960 frames of 4K = 12 GB of raw input per GPU

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    from video_codec_pipeline_b200 import api
    if not torch.cuda.is_available() or api.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the encoder has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    frames = make_workload(args.gops)
    n = frames.shape[0]
    fb = frames.shape[1]
    p = api.default_params(W, H, fps=FPS, gop=GOP, qp_i=QP_I, qp_p=QP_P, slices=SLICES, deblock_idc=args.deblock_idc,
                           first_gop=rank * args.gops, entropy=ENTROPY, transform8x8=T8X8, codec=CODEC, hevc_subpel=1 if CODEC else 0, hevc_sao=HEVC_SAO if CODEC else 0)
    host = torch.from_numpy(frames).pin_memory()
    dev = host.to("cuda", non_blocking=False)
    out_host = torch.empty(n * fb // 2 + (1 << 20), dtype=torch.uint8).pin_memory()
    out_np = out_host.numpy()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with api.Session(p, n, device=local_rank) as s:
        # ---- device-resident: K1..K5, CUDA events on the session's launching stream ----
        for _ in range(args.warmup):
            s.upload_device(dev.data_ptr(), n)
            s.encode()
        sampler = ClockSampler(local_rank)
        l0 = s.launch_count()
        barrier()
        sampler.start()
        dev_ms = 0.0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            dev_ms += s.upload_device(dev.data_ptr(), n)
            dev_ms += s.encode()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1000.0
        sampler.stop_flag.set()
        launches = s.launch_count() - l0
        res = s.download(out=out_np)
        stream_bytes = int(res["stream"].size)
        # per-kernel CUDA-event times: same step again, right after the timed region, with every
        # launch on ONE stream and an event pair around it (the timed steps overlap GOP groups on
        # several streams, which would smear per-kernel durations)
        s.profile(True)
        prof_steps = 2
        for _ in range(prof_steps):
            s.upload_device(dev.data_ptr(), n)
            s.encode()
        stats = s.kernel_stats()
        s.profile(False)

    # ---- end to end through the public API: pinned host frames -> bitstream in host memory ----
    # Two sessions driven by two host threads, the way a consumer with `-j 2` (cmd/consumer.go:123)
    # drives the executor: the H2D copy of one batch overlaps the kernels of the other, and inside a
    # batch the upload is streamed (each GOP group's chain starts when its frames have landed).  Every
    # batch still pays its own H2D of all frames and D2H of the whole bitstream inside the timed region.
    import threading as _th
    nthreads = max(1, args.e2e_threads)
    sessions = [api.Session(p, n, device=local_rank) for _ in range(nthreads)]
    outs = [torch.empty(n * fb // 2 + (1 << 20), dtype=torch.uint8).pin_memory().numpy() for _ in range(nthreads)]
    errors = []
    phase = [[0.0, 0.0, 0.0] for _ in range(nthreads)]   # host wall time in upload / encode / download

    def e2e_worker(i, count):
        try:
            for _ in range(count):
                t_a = time.perf_counter()
                sessions[i].upload(host.data_ptr(), n, wait=False)   # streamed: GOP groups start as their frames land
                t_b = time.perf_counter()
                sessions[i].encode()
                t_c = time.perf_counter()
                sessions[i].download(out=outs[i])
                t_d = time.perf_counter()
                phase[i][0] += t_b - t_a; phase[i][1] += t_c - t_b; phase[i][2] += t_d - t_c
        except Exception as ex:  # noqa: BLE001
            errors.append(ex)

    def e2e_round(total_steps):
        per = [total_steps // nthreads + (1 if i < total_steps % nthreads else 0) for i in range(nthreads)]
        ths = [_th.Thread(target=e2e_worker, args=(i, per[i])) for i in range(nthreads)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    e2e_round(2)                       # warm-up
    phase = [[0.0, 0.0, 0.0] for _ in range(nthreads)]
    barrier()
    t0 = time.perf_counter()
    e2e_round(args.steps)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1000.0
    for ss in sessions:
        ss.close()
    if errors:
        raise errors[0]

    times = torch.tensor([dev_ms, wall_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, e2e_ms = [float(x) for x in times.cpu()]
    total_frames = n * world * args.steps
    value = total_frames / (dev_ms / 1000.0)
    e2e_value = total_frames / (e2e_ms / 1000.0)

    if rank == 0:
        # roofline of the dominant kernel: largest share of CUDA-event time among the kernels whose work is
        # counted in bytes per frame.  (The CABAC arithmetic coder is a latency chain of 2 B per bin that runs
        # beside the step on side streams; it is listed in kernels_ms_per_step_single_stream, not ranked here.)
        ranked = [k for k in stats if stats[k]["launches"] and algorithmic_bytes_per_frame(k) > 0]
        top = max(ranked or stats, key=lambda k: stats[k]["ms"])
        st = stats[top]
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, which = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
        else:
            peak, which = 6650.0, "fallback"
        frames_per_launch = n * prof_steps / max(1, st["launches"])
        bytes_per_launch = algorithmic_bytes_per_frame(top) * frames_per_launch
        avg_ms = st["ms"] / max(1, st["launches"])
        achieved = bytes_per_launch / (avg_ms / 1000.0) / 1e9 if avg_ms > 0 else 0.0
        tot_ms = sum(v["ms"] for v in stats.values())
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):
            per_frame = json.load(open(tpath)).get(("hevc_" + top) if CODEC and top in ("p_recon", "i_recon", "cabac_bins") else top)
            if per_frame:
                traffic = int(per_frame * frames_per_launch)   # ncu --set full capture, scaled to this launch size
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dev_ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD + ", --verify-able Annex-B",
                       "frames_per_step_per_gpu": n, "gops_per_gpu": args.gops, "seed": SEED,
                       "l2": "inputs (%.1f GB/GPU) larger than L2" % (n * fb / 1e9),
                       "realtime_x": round(value / FPS, 1), "wall_ms_per_step": round(wall_ms / args.steps, 3),
                       "bitstream_bytes_per_step": stream_bytes},
            "e2e": {"value": round(e2e_value, 2), "unit": "frames/s", "h2d_bytes_per_step": n * fb,
                    "d2h_bytes_per_step": stream_bytes, "threads": nthreads,
                    "host_ms_per_step_upload_encode_download": [round(1000.0 * sum(ph[k] for ph in phase) / max(1, args.steps), 1) for k in range(3)]},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": top, "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 5), "traffic": traffic, "peak_source": which,
                         "share_of_step": round(st["ms"] / tot_ms, 4) if tot_ms else None,
                         "avg_launch_ms": round(avg_ms, 4), "launches": st["launches"]},
            "kernels_ms_per_step_single_stream": {k: round(v["ms"] / prof_steps, 3) for k, v in stats.items() if v["launches"]},
        }
        if not args.no_cpu_baseline and world == 1:
            cores = max(1, min(os.cpu_count() or 1, 32))
            fps, dt = cpu_port_fps(frames, cores, GOP)
            line["cpu_baseline"] = {"value": round(fps, 3), "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": "%d threads x one whole GOP (IDR+59P) of the same clip each, %.1f s; oracle/%s (libx264/libx265/ffmpeg absent from image)" % (cores, dt, "hevc_oracle.inc.c" if CODEC else "h264_oracle.c")}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
