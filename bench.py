#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json on synthetic yuv420p.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Headline workload (config.workload): BASELINE.json configs[1] — 1080p30 H.264 encode of synthetic
yuv420p, GOP=60, CAVLC, I+P frames, deblocking on, constant QP.  A "step" is one pass of the
hot path (K1 colour/pad -> K2 motion search -> K3 transform/quant/recon -> K4 deblock -> K5
CAVLC + NAL pack) over one batch of `--gops` closed GOPs per GPU.
  value : encoded frames/s, whole job, raw frames already resident in HBM, CUDA-event timed
  e2e   : the same through the public Session API with pinned HOST buffers: H2D of the frames and
          D2H of the bitstream inside the timed region
  verified     : the bitstream of the LAST TIMED step, one GOP per GOP group: byte-identical to the CPU
                 oracle's stream for that GOP, and the FFmpeg decoder's output of it equals the oracle's
                 reconstruction (outside the timed region)
  roofline     : the dominant kernel of the step, algorithmic bytes / CUDA-event time
  cpu_baseline : the CPU oracle (a port; the reference's libx264 is not in the image) on a bounded
                 sample of the same workload, timed on this box's host cores
  h2d          : what the box delivers when ONLY the e2e leg's copies run (same pinned buffer, same
                 bytes, two streams): the ceiling the e2e number can be read against
  e2e_transcode: the reference-facing call itself, vcpenc_transcode(path, path, argv), y4m on /dev/shm
                 -> mp4 on /dev/shm, two worker threads (consumer -j 2)
  extra        : the same measurements (3 short steps) for what the built-in presets actually run
                 (h264-cpu as parsed: High profile, CABAC), a hard-content clip, the 4K High shard of
                 configs[2] and the 4K HEVC shard of configs[3]
Multi-GPU: GOPs are independent -> each rank encodes its own GOPs, no data-path collective,
weak scaling (per-GPU batch fixed).  Time = max over ranks.  N>1 also proves the host concatenation:
sha256 of rank 0's + rank 1's streams concatenated == sha256 of one unsharded encode of the same GOPs.
"""
from __future__ import annotations

import argparse
import hashlib
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

GOP = 60
QP_I, QP_P = 25, 27
H264_CPU_PRESET = "-c:v libx264 -preset medium -crf 23 -c:a aac -b:a 128k -movflags +faststart"   # internal/config/config.go:49
H265_CPU_PRESET = "-c:v libx265 -preset medium -crf 28 -c:a aac -b:a 128k -movflags +faststart"   # internal/config/config.go:50


class Workload:
    """One encoder configuration + clip.  `entropy`/`slices`/... follow vcpenc_params."""

    def __init__(self, name, w=1920, h=1080, fps=30, entropy=0, slices=-1, t8x8=0, codec=0, content="std", gops=32,
                 seed=1080, qp_i=QP_I, qp_p=QP_P, hevc_sao=0, hevc_subpel=2, label=None):
        self.name, self.w, self.h, self.fps = name, w, h, fps
        self.codec, self.content, self.gops, self.seed = codec, content, gops, seed
        self.qp_i, self.qp_p, self.hevc_sao = qp_i, qp_p, hevc_sao
        self.hevc_subpel = hevc_subpel if codec else 0
        if codec:
            entropy, t8x8 = 1, 0
        self.entropy, self.t8x8 = entropy, t8x8
        mbh = (h + 15) // 16
        # slices: the encoder's own choice unless given (vcp_algo.h vcp_auto_slices: CAVLC 1; CABAC one per ~17 rows)
        self.slices = slices if slices >= 0 else (max(1, mbh // 17) if entropy else 1)
        res = "4K" if w >= 3840 else "%dp" % h
        if codec:
            self.metric = "%s HEVC encode fps (GOP=60, Main profile, I+P)" % res
            tools = "HEVC Main, GOP=60, CABAC, %d slice%s, I+P, %s-sample motion, deblock%s" % (
                self.slices, "" if self.slices == 1 else "s", ("full", "half", "quarter")[self.hevc_subpel], ", SAO" if hevc_sao else "")
        else:
            coder = "CABAC" if entropy else "CAVLC"
            self.metric = "%s H.264 encode fps (GOP=60, %s%s, I+P)" % (res, "High profile, " if t8x8 else "", coder)
            tools = "GOP=60, %s%s, %d slice%s, I+P, deblock" % ("High profile, " if t8x8 else "", coder, self.slices,
                                                               "" if self.slices == 1 else "s")
        self.workload = "%s: %dx%d@%d yuv420p%s, %s, CQP %d/%d" % (
            label or name, w, h, fps, " (hard content: fractional pan + noise)" if content == "hard" else "", tools, qp_i, qp_p)

    def params(self, api, first_gop=0, **kw):
        return api.default_params(self.w, self.h, fps=self.fps, gop=GOP, qp_i=self.qp_i, qp_p=self.qp_p, slices=self.slices,
                                  first_gop=first_gop, entropy=self.entropy, transform8x8=self.t8x8, codec=self.codec,
                                  hevc_subpel=self.hevc_subpel, hevc_sao=self.hevc_sao if self.codec else 0, **kw)

    def oracle_params(self, pyoracle, first_gop=0):
        return pyoracle.make_params(self.w, self.h, fps=self.fps, gop=GOP, qp_i=self.qp_i, qp_p=self.qp_p, entropy=self.entropy,
                                    slices=self.slices, transform8x8=self.t8x8, codec=self.codec, first_gop=first_gop,
                                    hevc_subpel=self.hevc_subpel, hevc_sao=self.hevc_sao if self.codec else 0)

    def make_frames(self, gops=None):
        """`gops` closed GOPs.  Two distinct GOPs are synthesised (numpy is slow at these sizes) and cycled;
        every GOP is encoded independently, so repetition does not make the work any cheaper.  Inputs
        (>= 1.4 GB at 8 GOPs of 1080p) are far larger than the 126 MB L2."""
        gops = gops or self.gops
        gen = synth.make_hard_clip if self.content == "hard" else synth.make_clip
        a = gen(self.w, self.h, GOP, seed=self.seed, start=0)
        if gops == 1:
            return a
        if self.w >= 3840:
            # 4K: the second GOP is the first one upside down (one synthesis pass of 60 4K frames costs ~17 s)
            cw, ch = self.w // 2, self.h // 2
            ysz, csz = self.w * self.h, cw * ch
            b = np.empty_like(a)
            b[:, :ysz] = a[:, :ysz].reshape(GOP, self.h, self.w)[:, ::-1, :].reshape(GOP, ysz)
            b[:, ysz:ysz + csz] = a[:, ysz:ysz + csz].reshape(GOP, ch, cw)[:, ::-1, :].reshape(GOP, csz)
            b[:, ysz + csz:] = a[:, ysz + csz:].reshape(GOP, ch, cw)[:, ::-1, :].reshape(GOP, csz)
        else:
            b = gen(self.w, self.h, GOP, seed=self.seed + 1, start=GOP)
        return np.concatenate([a if (g % 2 == 0) else b for g in range(gops)], axis=0)


def headline_workload(args):
    if args.codec == "hevc":
        if args.workload == "4k":
            return Workload("hevc4k", 3840, 2160, 60, codec=1, gops=args.gops or 16, seed=2160, slices=args.slices, hevc_sao=args.hevc_sao,
                            hevc_subpel=args.hevc_subpel, label="configs[3] (one GPU's GOP shard)")
        return Workload("hevc1080", codec=1, gops=args.gops or 32, slices=args.slices, hevc_sao=args.hevc_sao, hevc_subpel=args.hevc_subpel, label="configs[3] at 1080p")
    if args.workload == "4k":
        ent = 1 if args.entropy < 0 else args.entropy
        return Workload("4k", 3840, 2160, 60, entropy=ent, t8x8=1 if ent else 0, gops=args.gops or 16, seed=2160, slices=args.slices,
                        label="configs[2] (one GPU's GOP shard, %s profile)" % ("High" if ent else "Baseline"))
    ent = 0 if args.entropy < 0 else args.entropy
    return Workload("1080p", entropy=ent, t8x8=args.t8x8, gops=args.gops or 32, slices=args.slices, content=args.content, label="configs[1]")


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (rank 0 only)."""

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


def algorithmic_bytes_per_frame(kernel: str, wl: Workload) -> float:
    """compulsory HBM bytes per frame for the stage a kernel belongs to."""
    mbw, mbh = (wl.w + 15) // 16, (wl.h + 15) // 16
    P = 256 * mbw * mbh      # coded luma samples
    nmb = mbw * mbh
    return {
        "csc": 3.0 * wl.w * wl.h,              # K1: read 1.5WH + write 1.5WH
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


def prepass_absdiffs_per_frame(wl: Workload) -> float:
    """Pixel-absdiffs of the motion-search pre-pass per P picture: 625 half-res candidates x 64 px + 25
    full-res candidates x 256 px per macroblock (csrc/k2_me.cu)."""
    return ((wl.w + 15) // 16) * ((wl.h + 15) // 16) * (625 * 64 + 25 * 256)


def cpu_port_fps(wl: Workload, frames: np.ndarray, threads: int, frames_per_gop: int):
    """Oracle (CPU port) on a bounded sample: `threads` GOP-prefixes in parallel (ctypes drops the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle
    ngop_avail = frames.shape[0] // GOP
    jobs = []
    for k in range(threads):
        g = k % max(1, ngop_avail)
        jobs.append(frames[g * GOP: g * GOP + frames_per_gop])

    def one(fr):
        p = wl.oracle_params(pyoracle)
        if wl.codec:
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
    wl = headline_workload(args)
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 32))
    fpg = 30                               # IDR + 29 P per GOP prefix, per thread and step
    base = wl.make_frames(2)
    times, nframes = [], threads * fpg
    for i in range(args.warmup + args.steps):
        fps, dt = cpu_port_fps(wl, base, threads, fpg)
        if i >= args.warmup:
            times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    value = nframes / (ms / 1000.0)
    line = {
        "impl": "reference", "metric": wl.metric, "value": round(value, 3), "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": wl.workload, "frames_per_step": nframes, "seed": wl.seed},
        "cpu_baseline": {"value": round(value, 3), "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d threads x first %d frames of a GOP (IDR+P), oracle/%s; libx264/libx265/ffmpeg absent from image" % (threads, fpg, "hevc_oracle.inc.c" if wl.codec else "h264_oracle.c")},
        "e2e": {"value": round(value, 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------
# verification of a timed bitstream (outside every timed region)
# --------------------------------------------------------------------------------------------------------
def split_gop_streams(stream: np.ndarray, info, gop_ids):
    """Annex-B bytes of the given GOPs out of a session download (info = (offset, size, idr, qp) per frame)."""
    out = {}
    for g in gop_ids:
        a = info[g * GOP][0]
        last = info[min(len(info), (g + 1) * GOP) - 1]
        out[g] = stream[a: last[0] + last[1]].tobytes()
    return out


def verify_timed(wl: Workload, frames: np.ndarray, stream: np.ndarray, info, gop_ids, first_gop, use_oracle=True):
    """For each GOP in gop_ids of the timed output: FFmpeg decodes it (frame count must match) and, with
    use_oracle, the bytes equal the CPU oracle's stream for the same input and the decoder's pictures equal
    the oracle's reconstruction.  Without the oracle (4K: a GOP costs the scalar port ~1 min) the decoded
    pictures must be a plausible encode of the source (luma PSNR > 30 dB).  Returns a small report."""
    from concurrent.futures import ThreadPoolExecutor
    from video_codec_pipeline_b200 import arbiter
    rep = {"gops_checked": [int(g) for g in gop_ids], "decoder": None, "oracle_identical": None, "decoder_equals_recon": None}
    parts = split_gop_streams(stream, info, gop_ids)
    fb = frames.shape[1]

    def check(g):
        r = {}
        src = frames[g * GOP:(g + 1) * GOP]
        data = parts[g]
        if arbiter.available():
            dec = arbiter.decode_annexb_hevc(data) if wl.codec else arbiter.decode_annexb(data)
            r["decoded"] = len(dec) == src.shape[0]
            flat = [np.concatenate([pl.ravel() for pl in d]) for d in dec]
        else:
            r["decoded"] = None
            flat = None
        if use_oracle:
            from oracle import pyoracle
            p = wl.oracle_params(pyoracle, first_gop=first_gop + g)
            ref = pyoracle.encode_hevc(p, src) if wl.codec else pyoracle.encode(p, src, want_recon=True)
            r["identical"] = ref["stream"] == data
            if flat is not None and r["decoded"]:
                r["recon"] = all(np.array_equal(flat[i], ref["recon"][i]) for i in range(src.shape[0]))
        elif flat is not None and r["decoded"]:
            ysz = wl.w * wl.h
            mse = np.mean([(flat[i][:ysz].astype(np.float32) - src[i][:ysz].astype(np.float32)) ** 2 for i in (0, len(flat) // 2, len(flat) - 1)])
            r["psnr_y"] = float(10 * np.log10(255.0 ** 2 / max(mse, 1e-9)))
        return r

    with ThreadPoolExecutor(max(1, len(gop_ids))) as ex:
        res = list(ex.map(check, gop_ids))
    assert fb == frames.shape[1]
    if res and res[0].get("decoded") is not None:
        rep["decoder"] = all(r["decoded"] for r in res)
    if use_oracle:
        rep["oracle_identical"] = all(r["identical"] for r in res)
        if all("recon" in r for r in res):
            rep["decoder_equals_recon"] = all(r["recon"] for r in res)
        rep["ok"] = bool(rep["oracle_identical"] and rep["decoder"] is not False and rep["decoder_equals_recon"] is not False)
    else:
        ps = [r.get("psnr_y") for r in res if "psnr_y" in r]
        rep["psnr_y_min"] = round(min(ps), 2) if ps else None
        rep["ok"] = bool(rep["decoder"] and ps and min(ps) > 30.0)
    return rep


# --------------------------------------------------------------------------------------------------------
# measurement of one workload on this rank
# --------------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, torch, api, dist, rank, world, local_rank):
        self.torch, self.api, self.dist, self.rank, self.world, self.local_rank = torch, api, dist, rank, world, local_rank

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]


def measure(cx: Ctx, wl: Workload, frames: np.ndarray, steps: int, warmup: int, e2e_threads: int, profile=True, sampler=None,
            deblock_idc=0, e2e_steps=None):
    """Device-resident value, e2e through the Session API, per-kernel profile; returns a dict (times already
    max-reduced over ranks) plus the stream + frame index of the LAST TIMED device-resident step."""
    torch, api = cx.torch, cx.api
    n, fb = frames.shape
    p = wl.params(api, first_gop=cx.rank * wl.gops, deblock_idc=deblock_idc)
    host = torch.from_numpy(frames).pin_memory()
    dev = host.to("cuda", non_blocking=False)
    out_cap = n * fb // 10 + (1 << 20)      # bitstream buffer: the hard clip at CRF 23 needs 1.9 % of the raw size; page-locked, so not larger than needed
    out_np = torch.empty(out_cap, dtype=torch.uint8).pin_memory().numpy()
    res = {}
    with api.Session(p, n, device=cx.local_rank) as s:
        # ---- device-resident: K1..K5, CUDA events on the session's launching stream ----
        for _ in range(warmup):
            s.upload_device(dev.data_ptr(), n)
            s.encode()
        l0 = s.launch_count()
        cx.barrier()
        if sampler is not None:
            sampler.start()
        dev_ms = 0.0
        t0 = time.perf_counter()
        for _ in range(steps):
            dev_ms += s.upload_device(dev.data_ptr(), n)
            dev_ms += s.encode()
        cx.barrier()
        wall_ms = (time.perf_counter() - t0) * 1000.0
        if sampler is not None:
            sampler.stop_flag.set()
        res["launches"] = s.launch_count() - l0
        dl = s.download(out=out_np)                 # the output of the last TIMED step
        res["stream"] = dl["stream"].copy()
        res["info"] = dl["info"]
        stats = None
        if profile:
            # per-kernel CUDA-event times: same step again, right after the timed region, with every
            # launch on ONE stream and an event pair around it (the timed steps overlap GOP groups on
            # several streams, which would smear per-kernel durations)
            s.profile(True)
            res["prof_steps"] = 2
            for _ in range(res["prof_steps"]):
                s.upload_device(dev.data_ptr(), n)
                s.encode()
            stats = s.kernel_stats()
            s.profile(False)
        res["stats"] = stats
    del dev
    torch.cuda.empty_cache()

    # ---- h2d-only: the copies of the e2e leg and nothing else (same pinned buffer, two streams) ----
    dst = torch.empty(n * fb, dtype=torch.uint8, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(2)]
    half = (n // 2) * fb
    flat_host = host.view(-1)

    def h2d_once():
        with torch.cuda.stream(streams[0]):
            dst[:half].copy_(flat_host[:half], non_blocking=True)
        with torch.cuda.stream(streams[1]):
            dst[half:].copy_(flat_host[half:], non_blocking=True)
    h2d_once()
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        h2d_once()
    cx.barrier()
    h2d_ms = (time.perf_counter() - t0) * 1000.0 / 3
    del dst
    torch.cuda.empty_cache()

    # ---- end to end through the public API: pinned host frames -> bitstream in host memory ----
    # `e2e_threads` sessions driven by as many host threads, the way a consumer with `-j 2`
    # (cmd/consumer.go:123) drives the executor: the H2D copy of one batch overlaps the kernels of the
    # other, and inside a batch the upload is streamed (each GOP group's chain starts when its frames have
    # landed).  Every batch still pays its own H2D of all frames and D2H of the whole bitstream inside the
    # timed region.
    nthreads = max(1, e2e_threads)
    e2e_steps = e2e_steps or steps
    sessions = [api.Session(p, n, device=cx.local_rank) for _ in range(nthreads)]
    outs = [out_np] + [torch.empty(out_cap, dtype=torch.uint8).pin_memory().numpy() for _ in range(nthreads - 1)]
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
        ths = [threading.Thread(target=e2e_worker, args=(i, per[i])) for i in range(nthreads)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    e2e_round(2)                       # warm-up
    for ph in phase:
        ph[:] = [0.0, 0.0, 0.0]
    cx.barrier()
    t0 = time.perf_counter()
    e2e_round(e2e_steps)
    cx.barrier()
    e2e_ms = (time.perf_counter() - t0) * 1000.0
    for ss in sessions:
        ss.close()
    if errors:
        raise errors[0]
    del host
    dev_ms, wall_ms, e2e_ms, h2d_ms = cx.max_over_ranks([dev_ms, wall_ms, e2e_ms, h2d_ms])
    res.update(n=n, fb=fb, dev_ms=dev_ms, wall_ms=wall_ms, e2e_ms=e2e_ms, h2d_ms=h2d_ms, steps=steps, e2e_steps=e2e_steps,
               value=n * cx.world * steps / (dev_ms / 1000.0), e2e_value=n * cx.world * e2e_steps / (e2e_ms / 1000.0),
               h2d_fps=n * cx.world / (h2d_ms / 1000.0), h2d_gbs_per_gpu=n * fb / (h2d_ms / 1000.0) / 1e9,
               phase_ms=[round(1000.0 * sum(ph[k] for ph in phase) / max(1, e2e_steps), 1) for k in range(3)], nthreads=nthreads)
    return res


def roofline_of(wl: Workload, res, peak, which):
    """Dominant kernel of the single-stream profile: largest share of CUDA-event time among the kernels whose
    work SURVEY 8d counts in bytes per frame.  The CABAC arithmetic coder is a 2 B/bin dependency chain: it is
    ranked separately (bins/s), see `cabac_coder` in the JSON line."""
    stats = res["stats"]
    ranked = [k for k in stats if stats[k]["launches"] and algorithmic_bytes_per_frame(k, wl) > 0]
    top = max(ranked or stats, key=lambda k: stats[k]["ms"])
    st = stats[top]
    frames_per_launch = res["n"] * res["prof_steps"] / max(1, st["launches"])
    bytes_per_launch = algorithmic_bytes_per_frame(top, wl) * frames_per_launch
    avg_ms = st["ms"] / max(1, st["launches"])
    achieved = bytes_per_launch / (avg_ms / 1000.0) / 1e9 if avg_ms > 0 else 0.0
    tot_ms = sum(v["ms"] for v in stats.values())
    traffic = None
    for tname in ("r02_traffic.json", "r01_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            per_frame = json.load(open(tpath)).get(("hevc_" + top) if wl.codec and top in ("p_recon", "i_recon", "cabac_bins") else top)
            if per_frame:
                traffic = int(per_frame * frames_per_launch * (wl.w * wl.h) / (1920 * 1080))   # ncu --set full capture at 1080p, scaled to this launch
                break
    return {"bound": "hbm", "kernel": top, "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 5), "traffic": traffic, "peak_source": which,
            "share_of_step": round(st["ms"] / tot_ms, 4) if tot_ms else None,
            "avg_launch_ms": round(avg_ms, 4), "launches": st["launches"]}


def write_y4m(path, wl: Workload, frames: np.ndarray):
    with open(path, "wb") as f:
        f.write(b"YUV4MPEG2 W%d H%d F%d:1 Ip A1:1 C420\n" % (wl.w, wl.h, wl.fps))
        for i in range(frames.shape[0]):
            f.write(b"FRAME\n")
            f.write(memoryview(frames[i]))


def e2e_transcode(cx: Ctx, wl: Workload, frames2: np.ndarray, tokens: str, clips):
    """The reference-facing call: vcpenc_transcode(input, output, argv) from `-j 2` worker threads, y4m in and
    mp4 out on /dev/shm (falls back to the temp dir).  clips = [(label, gops)]; one warm-up task per thread."""
    import tempfile
    api = cx.api
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    out = {"tokens": tokens, "dir": base, "threads": 2}
    for label, gops in clips:
        n = gops * GOP
        need = n * frames2.shape[1] + (256 << 20)
        st = os.statvfs(base)
        if st.f_bavail * st.f_frsize < need:
            out[label] = {"skipped": "not enough space on %s (%.1f GB needed)" % (base, need / 1e9)}
            continue
        src = os.path.join(base, "vcp_bench_%d_%s.y4m" % (os.getpid(), label))
        clip = np.concatenate([frames2[(g % 2) * GOP:(g % 2 + 1) * GOP] for g in range(gops)], axis=0)
        write_y4m(src, wl, clip)
        del clip
        dsts = [os.path.join(base, "vcp_bench_%d_%s_%d.mp4" % (os.getpid(), label, i)) for i in range(2)]
        errs = []
        # what the file system delivers: the same bytes pread() into page-locked memory by the same number of threads
        # (2 tasks x the library's read workers), nothing else -- the ceiling of any file-to-file transcode on this box
        read_fps = None
        try:
            import concurrent.futures as cf
            fsz = os.path.getsize(src)
            nthr = 2 * max(1, min(8, (os.cpu_count() or 2) // 2))
            pin = cx.torch.empty(min(fsz, 1 << 30), dtype=cx.torch.uint8).pin_memory().numpy()
            piece = 64 << 20
            jobs = [(o, min(piece, fsz - o)) for o in range(0, fsz, piece)]
            fd = os.open(src, os.O_RDONLY)

            def rd(job):
                o, ln = job
                mv = memoryview(pin)[(o % (pin.size - piece + 1)) if pin.size > piece else 0:][:ln]
                got = 0
                while got < ln:
                    k = os.preadv(fd, [mv[got:]], o + got)
                    if k <= 0:
                        break
                    got += k
                return got
            t0 = time.perf_counter()
            with cf.ThreadPoolExecutor(nthr) as ex:
                tot = sum(ex.map(rd, jobs + jobs))                 # two tasks' worth
            dt = time.perf_counter() - t0
            os.close(fd)
            read_fps = round(2 * n / dt, 1) if tot >= 2 * fsz - 16 else None
            del pin
        except Exception:  # noqa: BLE001
            read_fps = None

        def task(i, reps):
            try:
                api.set_thread_device(cx.local_rank)
                for _ in range(reps):
                    api.transcode(src, dsts[i], tokens)
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)

        def round_(reps):
            ths = [threading.Thread(target=task, args=(i, reps)) for i in range(2)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        try:
            round_(1)                                  # warm-up: sessions + pinned buffers are created and cached per thread
            reps = 3 if gops <= 8 else 1
            t0 = time.perf_counter()
            round_(reps)
            dt = time.perf_counter() - t0
            if errs:
                raise errs[0]
            ok = all(os.path.getsize(d) > 0 for d in dsts)
            try:
                api.verify(dsts[0])
            except Exception:  # noqa: BLE001
                ok = False
            out[label] = {"frames_per_task": n, "tasks": 2 * reps, "fps": round(2 * reps * n / dt, 1), "seconds": round(dt, 3),
                          "mp4_bytes": os.path.getsize(dsts[0]), "verify": ok, "file_read_ceiling_fps": read_fps,
                          "of_read_ceiling": round(2 * reps * n / dt / read_fps, 3) if read_fps else None}
            if gops <= 8 and ok:
                # what the producer really forwards (cmd/producer.go:485-488): a CONTAINER file.  The mp4 just written goes
                # back in: demux + CPU decode (libavcodec, frame threads) in the reader thread, encode on the GPU.
                src2 = os.path.join(base, "vcp_bench_%d_%s_in.mp4" % (os.getpid(), label))
                os.replace(dsts[0], src2)
                try:
                    def task2(i, reps2):
                        try:
                            api.set_thread_device(cx.local_rank)
                            for _ in range(reps2):
                                api.transcode(src2, dsts[i], tokens)
                        except Exception as ex:  # noqa: BLE001
                            errs.append(ex)
                    for r2 in (1, 2):                       # warm-up round, timed round
                        t0 = time.perf_counter()
                        ths = [threading.Thread(target=task2, args=(i, r2)) for i in range(2)]
                        for t in ths:
                            t.start()
                        for t in ths:
                            t.join()
                        dt2 = time.perf_counter() - t0
                    if errs:
                        raise errs[0]
                    api.verify(dsts[0])
                    out[label + "_mp4_in"] = {"input": "the H.264 mp4 of the leg above (%d bytes)" % os.path.getsize(src2), "frames_per_task": n, "tasks": 4,
                                              "fps": round(4 * n / dt2, 1), "seconds": round(dt2, 3), "verify": True,
                                              "note": "bound by the CPU decode of the input (libavcodec h264, frame threads), not by the encoder"}
                except Exception as ex:  # noqa: BLE001
                    out[label + "_mp4_in"] = {"error": str(ex)[:200]}
                finally:
                    if os.path.exists(src2):
                        os.remove(src2)
        except Exception as ex:  # noqa: BLE001
            out[label] = {"error": str(ex)[:200]}
        finally:
            for f in [src] + dsts:
                if os.path.exists(f):
                    os.remove(f)
    return out


def concat_parity(cx: Ctx, wl: Workload, frames: np.ndarray, my_stream: np.ndarray):
    """N>1: the path shards by closed GOPs and concatenates on the host.  Every rank hashes the stream it timed;
    rank 0 gathers ranks 0 and 1, concatenates in rank order and compares with ONE unsharded encode of the same
    2 x gops GOPs (outside the timed region).  Also checks that all ranks produced a stream of the same hash as
    the rank with the same idr_pic_id parity (identical synthetic GOPs, first_gop = rank * gops)."""
    dist, torch, api = cx.dist, cx.torch, cx.api
    my_hash = hashlib.sha256(my_stream.tobytes()).hexdigest()
    hashes = [None] * cx.world
    dist.all_gather_object(hashes, my_hash)
    parts = [None] * cx.world
    dist.all_gather_object(parts, my_stream.tobytes() if cx.rank < 2 else b"")
    rep = None
    if cx.rank == 0:
        n = frames.shape[0]
        both = np.concatenate([frames, frames], axis=0)
        p = wl.params(api, first_gop=0)
        with api.Session(p, 2 * n, device=cx.local_rank) as s:
            s.upload(both)
            s.encode()
            whole = s.download()["stream"].tobytes()
        cat = parts[0] + parts[1]
        rep = {"ranks_concatenated": [0, 1], "sha256_concat": hashlib.sha256(cat).hexdigest()[:16],
               "sha256_unsharded": hashlib.sha256(whole).hexdigest()[:16], "equal": cat == whole,
               "rank_hashes_consistent": all(hh == hashes[r % 2] for r, hh in enumerate(hashes)) if (wl.gops % 2) else len(set(hashes)) == 1}
    dist.barrier()
    return rep


def host_topology(local_rank):
    """Where this rank's GPU and CPUs sit: the e2e path is a host -> device feed, NUMA placement matters."""
    info = {"cpus_allowed": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None}
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=5).stdout.strip().lower()
        if bus.startswith("0000"):
            bus = bus[4:]
        node = open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip()
        info["gpu_numa_node"] = int(node)
        info["gpu_local_cpulist"] = open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read().strip()
    except Exception:  # noqa: BLE001
        pass
    try:
        info["numa_nodes_online"] = open("/sys/devices/system/node/online").read().strip()
        for ln in open("/proc/self/status"):
            if ln.startswith("Mems_allowed_list") or ln.startswith("Cpus_allowed_list"):
                info[ln.split(":")[0].lower()] = ln.split(":")[1].strip()
    except Exception:  # noqa: BLE001
        pass
    return info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gops", type=int, default=0, help="closed GOPs per GPU per step (weak scaling; default 32, 4K 16)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra workloads (presets as parsed, hard content, 4K, HEVC)")
    ap.add_argument("--no-verify", action="store_true", help="skip the decoder + oracle check of the timed bitstream")
    ap.add_argument("--no-transcode", action="store_true", help="skip the vcpenc_transcode (plugin call) leg")
    ap.add_argument("--workload", default="1080p", choices=["1080p", "4k"], help="default: BASELINE.json configs[1]")
    ap.add_argument("--codec", default="h264", choices=["h264", "hevc"], help="hevc: BASELINE.json configs[3] (use with --workload 4k)")
    ap.add_argument("--content", default="std", choices=["std", "hard"], help="hard: fractional full-frame pan + noise")
    ap.add_argument("--hevc-subpel", type=int, default=2, help="HEVC motion precision: 0 full, 1 half, 2 quarter samples (what the h265 presets parse to)")
    ap.add_argument("--hevc-sao", action="store_true", help="HEVC: sample adaptive offset on (default off: its first kernel is slow)")
    ap.add_argument("--entropy", type=int, default=-1, help="0 CAVLC, 1 CABAC (default: what the workload names)")
    ap.add_argument("--t8x8", type=int, default=0, help="1: High profile 8x8 transform (1080p workload)")
    ap.add_argument("--slices", type=int, default=-1, help="slices per picture (default: the encoder's choice)")
    ap.add_argument("--e2e-threads", type=int, default=2, help="host threads (sessions) of the end-to-end pipeline, like consumer -j")
    ap.add_argument("--deblock-idc", type=int, default=0, help="experiments only: 1 switches the in-loop filter off")
    args = ap.parse_args()
    args.hevc_sao = 1 if args.hevc_sao else 0

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
    cx = Ctx(torch, api, dist, rank, world, local_rank)

    wl = headline_workload(args)
    frames = wl.make_frames()
    sampler = ClockSampler(local_rank) if rank == 0 else None     # one sampler per job, not per rank
    res = measure(cx, wl, frames, args.steps, args.warmup, args.e2e_threads, profile=True, sampler=sampler, deblock_idc=args.deblock_idc)
    n, fb = res["n"], res["fb"]

    # the timed output, checked outside the timed region: one GOP per GOP group (both synthetic GOPs occur)
    verified = None
    if not args.no_verify and rank == 0:
        ng = min(4, wl.gops)
        ids = sorted(set(min(wl.gops - 1, (wl.gops * k) // ng + (k % 2)) for k in range(ng)))
        verified = verify_timed(wl, frames, res["stream"], res["info"], ids, first_gop=rank * wl.gops, use_oracle=wl.w < 3840)
    parity = concat_parity(cx, wl, frames, res["stream"]) if world > 1 else None

    line = None
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, which = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
        else:
            peak, which = 6650.0, "fallback"
        stats = res["stats"]
        line = {
            "metric": wl.metric, "value": round(res["value"], 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(res["dev_ms"] / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": wl.workload,
                       "frames_per_step_per_gpu": n, "gops_per_gpu": wl.gops, "seed": wl.seed,
                       "l2": "inputs (%.1f GB/GPU) larger than L2" % (n * fb / 1e9),
                       "realtime_x": round(res["value"] / wl.fps, 1), "wall_ms_per_step": round(res["wall_ms"] / args.steps, 3),
                       "bitstream_bytes_per_step": int(res["stream"].size)},
            "e2e": {"value": round(res["e2e_value"], 2), "unit": "frames/s", "h2d_bytes_per_step": n * fb,
                    "d2h_bytes_per_step": int(res["stream"].size), "threads": res["nthreads"],
                    "host_ms_per_step_upload_encode_download": res["phase_ms"],
                    "of_h2d_ceiling": round(res["e2e_value"] / res["h2d_fps"], 3)},
            "h2d": {"gb_per_s_per_gpu": round(res["h2d_gbs_per_gpu"], 2), "fps_ceiling": round(res["h2d_fps"], 1),
                    "what": "the e2e leg's host->device copies alone (same pinned buffer, 2 streams, all ranks at once, max over ranks)",
                    "topology": host_topology(local_rank)},
            "verified": verified["ok"] if verified else None,
            "verify": verified,
            "gpu_launches": int(res["launches"]),
            "clocks": sampler.summary(),
            "roofline": roofline_of(wl, res, peak, which),
            "kernels_ms_per_step_single_stream": {k: round(v["ms"] / res["prof_steps"], 3) for k, v in stats.items() if v["launches"]},
        }
        # integer-pipe roofline of the motion-search pre-pass (north-star: "integer-pipe utilisation for motion search")
        if stats["me_prepass"]["launches"]:
            pp_s = stats["me_prepass"]["ms"] / res["prof_steps"] / 1000.0
            pframes = n - wl.gops
            ad = prepass_absdiffs_per_frame(wl) * pframes / pp_s
            line["me_prepass_int_pipe"] = {"pixel_absdiffs_per_s": float("%.4g" % ad), "peak": 7.37e13, "frac": round(ad / 7.37e13, 4),
                                           "peak_source": "tools/sad_peak.cu, measured round 1 (63.4 VABSDIFF4/clk/SM at 1965 MHz)"}
        if stats["cabac_code"]["launches"]:
            line["cabac_coder"] = {"ms_per_step_single_stream": round(stats["cabac_code"]["ms"] / res["prof_steps"], 3)}
        if parity is not None:
            line["concat_parity"] = parity
    stream_keep = None
    del res
    torch.cuda.empty_cache()

    # ---- the plugin call itself ----
    if not args.no_transcode and world == 1 and wl.name == "1080p" and rank == 0:
        tokens = "-c:v libx264 -profile:v %s -coder %d -g 60 -qp %d -slices %d" % ("high" if wl.t8x8 else ("main" if wl.entropy else "baseline"),
                                                                                    wl.entropy, QP_P, wl.slices)
        try:
            line["e2e_transcode"] = e2e_transcode(cx, wl, frames[:2 * GOP], tokens, [("S1080", 5), ("S1080L", wl.gops)])
        except Exception as ex:  # noqa: BLE001
            line["e2e_transcode"] = {"error": str(ex)[:300]}

    if not args.no_cpu_baseline and world == 1 and rank == 0:
        cores = max(1, min(os.cpu_count() or 1, 32))
        fps, dt = cpu_port_fps(wl, frames, cores, GOP)
        line["cpu_baseline"] = {"value": round(fps, 3), "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "%d threads x one whole GOP (IDR+59P) of the same clip each, %.1f s; oracle/%s (libx264/libx265/ffmpeg absent from image)" % (cores, dt, "hevc_oracle.inc.c" if wl.codec else "h264_oracle.c")}
    elif rank == 0:
        line["cpu_baseline"] = None

    # ---- extra: what the presets actually run, hard content, 4K High, 4K HEVC (3 short steps each) ----
    if not args.no_extra and wl.name == "1080p" and args.content == "std":
        extra = {}
        pc = api.parse_args(H264_CPU_PRESET.split())            # h264-cpu as vcpenc_parse_args sees it
        ph = api.parse_args(H265_CPU_PRESET.split())
        todo = [
            ("h264-cpu preset as parsed (1080p)", Workload("h264cpu", entropy=pc.entropy, t8x8=pc.transform8x8, qp_i=pc.qp_i, qp_p=pc.qp_p,
                                                            label="h264-cpu (config.go:49) as parsed"), frames),
            ("hard content 1080p CAVLC", Workload("hard", content="hard", label="configs[1] settings on hard content"), None),
            ("hard content 1080p h264-cpu preset", Workload("hardcpu", content="hard", entropy=pc.entropy, t8x8=pc.transform8x8, qp_i=pc.qp_i, qp_p=pc.qp_p,
                                                             label="h264-cpu as parsed on hard content"), "prev"),
            ("configs[2] 4K60 High CABAC shard", Workload("4k", 3840, 2160, 60, entropy=1, t8x8=1, gops=16, seed=2160,
                                                          label="configs[2] (one GPU's GOP shard, High profile)"), None),
            ("configs[3] 4K60 HEVC shard (h265-cpu as parsed)", Workload("hevc4k", 3840, 2160, 60, codec=1, gops=16, seed=2160, qp_i=ph.qp_i, qp_p=ph.qp_p, hevc_subpel=ph.hevc_subpel,
                                                                         label="configs[3] (one GPU's GOP shard), h265-cpu (config.go:50) as parsed"), "prev"),
        ]
        prev = None
        for key, w2, fr in todo:
            try:
                f2 = frames if fr is frames else (prev if isinstance(fr, str) and prev is not None else w2.make_frames())
                prev = f2
                r2 = measure(cx, w2, f2, 3, 2, args.e2e_threads, profile=(w2.name in ("h264cpu", "hard", "hardcpu")), e2e_steps=2)
                if rank == 0:
                    ent = {"workload": w2.workload, "metric": w2.metric, "value": round(r2["value"], 1), "e2e": round(r2["e2e_value"], 1),
                           "unit": "frames/s", "ms_per_step": round(r2["dev_ms"] / 3, 2), "frames_per_step_per_gpu": r2["n"], "steps": 3,
                           "realtime_x": round(r2["value"] / w2.fps, 1), "bitstream_bytes_per_step": int(r2["stream"].size),
                           "h2d_fps_ceiling": round(r2["h2d_fps"], 1)}
                    if r2["stats"]:
                        ent["kernels_ms_per_step_single_stream"] = {k: round(v["ms"] / r2["prof_steps"], 2) for k, v in r2["stats"].items() if v["launches"]}
                    if not args.no_verify:
                        ids = [0, min(w2.gops - 1, w2.gops // 2 + 1)]
                        v2 = verify_timed(w2, f2, r2["stream"], r2["info"], ids, first_gop=0, use_oracle=w2.w < 3840)
                        ent["verified"] = v2["ok"]
                        ent["verify"] = v2
                    if w2.content == "hard" and r2["stats"]:
                        # how much of the refine kernel's work the content lets the early-out skip: time per step relative to the standard clip
                        ent["me_refine_ms_vs_std"] = [ent["kernels_ms_per_step_single_stream"].get("me_refine"),
                                                      line["kernels_ms_per_step_single_stream"].get("me_refine")]
                    extra[key] = ent
                del r2
                torch.cuda.empty_cache()
            except Exception as ex:  # noqa: BLE001
                if rank == 0:
                    extra[key] = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}
        if rank == 0:
            line["extra"] = extra

    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    _ = stream_keep


if __name__ == "__main__":
    main()
