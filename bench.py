#!/usr/bin/env python
"""bench.py — framed→ADΔER transcode throughput (Mpixels/s) on B200, beside the CPU path.

Headline workload (BASELINE.json configs[4], the largest single-GPU configuration and the one north_star's scaling
target is quoted on): 7680x4320 gray 8-bit, 2000 frames, static scene with rare one-frame changes (p = 2/256 per pixel
per frame, SURVEY.md §8(d) cfg 5), ref_time 256, delta_t_max 2^20 (long-integration mode), crf 3, FramePerfect /
Collapse / AbsoluteT, chunk_rows 1, fresh pixel state at the start of every step.  One step = the whole 2000-frame
sequence through the hot path.  STRONG scaling: at N GPUs the one 8K frame is split into N row bands (sharding.band_of,
SURVEY.md §8(e)); rank g permanently owns band g's pixel state; rank-order concatenation of the bands' event streams is
the reference's raster order.

  value     device-resident: frames already in HBM, events left in HBM (adder_b200_video_integrate_frames_device:
            one persistent kernel launch per run of 250 frames), CUDA-event timed on the launching stream, max over ranks.
  e2e       the same step through adder_b200_video_integrate_frames_host: pinned HOST frames in, all events out to
            pinned HOST memory, copies inside the timed region.
  gather    (N > 1) the same step with the whole frame's events delivered in raster order into ONE consumer rank's
            HBM after every frame (the exchange step a single downstream encoder needs, video.rs:736-740).
  roofline  algorithmic bytes of the integrate kernel (counted exactly by the instrumented twin of the kernel in an
            untimed pass) / CUDA-event time of the timed region; traffic = DRAM bytes measured in this
            run by an `ncu --metrics dram__bytes_*` side pass over a short launch of the same workload.
  workloads (N = 1) the other BASELINE configs on the same kernel: cfg 1, cfg 2, cfg 3 for every c_thresh 0..10, cfg 4.
  cpu_baseline  the C oracle (a port: the Rust reference cannot be built here) on the host cores, bounded sample.

`--impl reference` times that CPU port alone, all host cores, on the same config.
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KIND_GRADIENT, KIND_NOISE, KIND_JITTER, KIND_STATIC = 0, 1, 2, 3
SEED = 0xADDE5
METRIC = "Mpixels/s framed->ADDER transcode (bit-exact events)"

# the headline: BASELINE configs[4]
W, H, C, NF = 7680, 4320, 1, 2000
REF, DTM, CRF = 256, 1 << 20, 3
WORKLOAD = ("7680x4320 gray 8-bit synthetic static scene + rare one-frame changes (p=2/256), 2000 frames, crf 3 (c 2..7, velocity 7), "
            "ref 256, dtm 2^20, FramePerfect/Collapse/AbsoluteT; one frame strong-scaled over N row bands")
BATCH = 250        # frames per integrate launch (events of one batch stay in HBM until the next batch overwrites them)
EV_PER_PX = 0.25   # event records per pixel per frame the device buffer holds (overflow is an error, not a silent drop)


def env_int(k, d):
    return int(os.environ.get(k, d))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.p = None
        self.t0 = self.t1 = None

    def mark_begin(self):
        """the timed region starts now (the sampler itself is started earlier: nvidia-smi needs up to a second to come up)"""
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        import datetime

        rows = []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[9], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), f[5:9]))
            except ValueError:
                continue
        window = "timed region"
        inside = [r for r in rows if self.t0 is not None and self.t1 is not None and self.t0 - 0.02 <= r[0] <= self.t1 + 0.02]
        if not inside:  # a timed region shorter than the sampling period: the warm-up steps just before it ran the same work
            inside, window = rows, "warm-up + timed region"
        sm, mx, reasons = [r[1] for r in inside], [r[2] for r in inside], set()
        for r in inside:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm_sorted = sorted(sm)  # under load = the upper half of the samples (the sampler also sees the idle edges)
        load = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm), "window": window}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---- measured-in-run DRAM traffic: an ncu side pass over a short launch of the same workload ------------------------

def ncu_dram_bytes_per_px_frame(extra_args, frames, px_per_frame, timeout_s=240):
    """Runs tools/profile_run.py under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` (one pass, no replay) and
    returns (dram bytes per px-frame of the last integrate launch, counted algorithmic bytes per px-frame of that same
    launch) or None when ncu is not usable here."""
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--csv",
           "-k", "regex:integrate_frame", sys.executable, os.path.join(ROOT, "tools", "profile_run.py"), "--batch", "--count",
           "--reps", "1", "--frames", str(frames)] + [str(x) for x in extra_args]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, cwd=ROOT)
    except Exception:
        return None
    if r.returncode != 0:
        return None
    rows = [row for row in csv.reader(io.StringIO(r.stdout)) if len(row) > 5]
    hdr = next((row for row in rows if "Metric Name" in row), None)
    if hdr is None:
        return None
    i_id, i_name, i_unit, i_val = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    per_launch = {}
    for row in rows:
        if row is hdr or not row[i_name].startswith("dram__bytes"):
            continue
        try:
            per_launch.setdefault(int(row[i_id]), 0.0)
            per_launch[int(row[i_id])] += float(row[i_val].replace(",", "")) * scale.get(row[i_unit], 1.0)
        except ValueError:
            continue
    if not per_launch:
        return None
    dram = per_launch[max(per_launch)]  # the last launch: the timed rep (the first is the counting twin)
    counted = None
    for line in r.stdout.splitlines():
        if line.startswith("counted:"):
            counted = float(line.split()[1])
    return dram / (px_per_frame * frames), counted


# ---- one device-resident workload -----------------------------------------------------------------------------------

class DeviceWorkload:
    """A plane (or this rank's band of it), its frames resident in HBM, and the step that runs them through the hot path."""

    def __init__(self, A, S, *, w, h, c, kind, frames, ref, dtm, crf=None, manual=None, multi=None, rank=0, world=1, device=0,
                 batch=BATCH, ev_per_px=2.0, offsets=True):
        self.A = A
        self.w, self.c, self.kind, self.nf, self.ref = w, c, kind, frames, ref
        self.crf, self.manual = crf, manual
        self.row0, self.rows = S.band_of(h, 1, rank, world)
        self.v = A.Video(w, self.rows, c, A.MODE_FRAME_PERFECT, device=device)
        v = self.v
        assert v.time_parameters(ref * 30, ref, dtm, None)
        if multi is not None:
            v.write_out(None, multi)
        v.set_row_offset(self.row0)
        self.P = w * self.rows * c
        self.batch = min(batch, frames)
        self.ev_stride = int(self.P * ev_per_px)
        self.d_frames = v.device_alloc(self.P * frames)
        v.synth_frames(self.d_frames, 0, frames, kind, SEED)  # a band gets its rows of the undivided frame
        self.d_events = v.device_alloc(self.ev_stride * 12 * self.batch)
        self.d_off = v.device_alloc((v.n_chunks + 1) * 4 * self.batch) if offsets else None
        self.quality()
        v.sync()

    def quality(self):
        if self.manual is not None:
            c = self.manual
            self.v.update_quality_manual(c, c, 30, 1, 0.0)  # SURVEY.md §8(d) cfg 3: quality_manual(c, c, 30, 1, 0.0)
        elif self.crf is not None:
            self.v.update_crf(self.crf)

    def step(self):
        v = self.v
        v.reset_state()
        self.quality()
        for f0 in range(0, self.nf, self.batch):
            n = min(self.batch, self.nf - f0)
            v.integrate_frames_device(self.d_frames.ptr + f0 * self.P, self.P, n, float(self.ref), self.d_events.ptr, self.ev_stride,
                                      self.d_off.ptr if self.d_off else None)

    def count(self):
        """One untimed step with the counting twin of the kernel: exact algorithmic bytes of the workload."""
        v = self.v
        v.set_counting(True)
        self.step()
        v.sync()
        cnt = v.read_counters()
        v.set_counting(False)
        pxf = self.P * self.nf
        alg = (1 + 8 + 8) * pxf + 16 * (cnt["node_loads"] + cnt["node_stores"]) + cnt["display_writes"] + 12 * cnt["events"]
        return cnt, alg

    def timed(self, steps, warmup, barrier=lambda: None, sampler=None):
        v = self.v
        for _ in range(warmup):
            self.step()
        v.sync()
        barrier()
        l0 = v.launch_count
        if sampler:
            sampler.mark_begin()
        v.timer_start()
        for _ in range(steps):
            self.step()
        ms = v.timer_stop()
        if sampler:
            sampler.mark_end()
        v.sync()
        barrier()
        return ms, v.launch_count - l0

    def free(self):
        for b in (self.d_frames, self.d_events, self.d_off):
            if b is not None:
                b.free()
        self.d_frames = self.d_events = self.d_off = None
        self.v.close()


def survey_formula(cnt, pxf):
    """SURVEY.md §8(d)'s planning formula on the same run: 1 + 2*(12 + 16*L) + 1[display] + 12*E."""
    L_mean = (cnt["live_nodes_in"] + cnt["live_nodes_out"]) / (2.0 * pxf)
    E_mean = cnt["events"] / pxf
    return 1 + 2 * (12 + 16 * L_mean) + 1 + 12 * E_mean, L_mean, E_mean


def e2e_cfg2_legs(A, args):
    """BASELINE configs[1] end to end through the host-buffer forms of the ABI (pinned host frames in, results out to pinned
    host memory, copies inside the timed region): 12-byte records, the raw .adder body, the compact form, and the compact
    form expanded back to records on the host's threads (VERDICT r1 task 5)."""
    w, h, c, nf, ref = 1920, 1080, 3, (min(300, args.frames) if args.frames else 300), 255
    P = w * h * c
    v = A.Video(w, h, c)
    assert v.time_parameters(ref * 30, ref, 7650, None)
    d = v.device_alloc(P * nf)
    v.synth_frames(d, 0, nf, KIND_NOISE, SEED)
    hf = np.asarray(A.pinned_empty((nf, h, w, c), np.uint8))
    d.to_host_into(hf, P * nf)
    d.free()
    sub = 20
    cap = int(P * sub * 1.2) + 4096
    out = {}
    n_threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for form in ("records", "raw", "compact", "compact_expanded"):
        if form == "records":
            buf = np.asarray(A.pinned_empty((cap,), A.EVENT_DTYPE))
        else:
            buf = np.asarray(A.pinned_empty((cap * 11,), np.uint8))
        rec = np.asarray(A.pinned_empty((cap,), A.EVENT_DTYPE)) if form == "compact_expanded" else None

        def step():
            v.reset_state()
            v.update_crf(3)
            n_ev, n_bytes = 0, 0
            for f0 in range(0, nf, sub):
                fr = hf[f0:f0 + sub]
                if form == "records":
                    ev, fc, cc = v.integrate_frames_host(fr, float(ref), buf)
                    n_bytes += ev.nbytes
                elif form == "raw":
                    body, fc, cc = v.integrate_frames_host_raw(fr, float(ref), buf)
                    n_bytes += len(body)
                else:
                    body, fc, cc = v.integrate_frames_host_compact(fr, float(ref), buf)
                    n_bytes += len(body)
                    if rec is not None:
                        pos, at = 0, 0
                        for k in range(len(fr)):
                            nb = A.compact_frame_bytes(P, int(fc[k]))
                            A.expand_compact(w, h, c, 0, body[pos:pos + nb], int(fc[k]), rec[at:], n_threads)
                            pos += nb
                            at += int(fc[k])
                n_ev += int(fc.sum())
            return n_ev, n_bytes

        step()
        t0 = time.perf_counter()
        n_ev, n_bytes = step()
        dt = time.perf_counter() - t0
        out[form] = {"value": P * nf / dt / 1e6, "unit": "Mpx/s", "d2h_bytes_per_step": n_bytes + (h + 1) * 4 * nf, "h2d_bytes_per_step": P * nf,
                     "events_per_step": n_ev}
        del buf, rec
    out["compact_expanded"]["host_threads"] = n_threads
    out["api"] = ("adder_b200_video_integrate_frames_host / _host_raw / _host_compact in runs of 20 frames, pinned host buffers; "
                  "compact_expanded adds adder_b200_expand_compact (12-byte records rebuilt on the host's threads) inside the timed region")
    v.close()
    return out


def run_side_workloads(A, S, args, peak):
    """The other BASELINE configs on one GPU (VERDICT r1 task 2): value, counted bytes, roofline fraction each."""
    out = []
    specs = [("cfg1: 640x480 gray gradient, 30 frames, API defaults (c 10), dtm = ref 255, one frame per launch",
              dict(w=640, h=480, c=1, kind=KIND_GRADIENT, frames=30, ref=255, dtm=255, batch=1, ev_per_px=4.0), 20, None),
             ("cfg2: 1920x1080 RGB noise, 300 frames, crf 3, ref 255, dtm 7650",
              dict(w=1920, h=1080, c=3, kind=KIND_NOISE, frames=300, ref=255, dtm=7650, crf=3, batch=300, ev_per_px=2.0), 3,
              ["--w", 1920, "--h", 1080, "--c", 3, "--kind", 1, "--crf", 3, "--cap", 2])]
    for c in range(11):
        specs.append((f"cfg3: 3840x2160 gray jitter, 1000 frames, quality_manual(c={c}), ref 255, dtm 7650",
                      dict(w=3840, h=2160, c=1, kind=KIND_JITTER, frames=1000, ref=255, dtm=7650, manual=c, batch=250, ev_per_px=2.0), 2,
                      ["--w", 3840, "--h", 2160, "--c", 1, "--kind", 2, "--manual", c, "--cap", 2] if c in (0, 5, 10) or args.traffic == "all" else None))
    specs.append(("cfg4: 3840x2160 RGB noise, 1000 frames, crf 3, ref 255, dtm 7650 (the whole plane on one GPU)",
                  dict(w=3840, h=2160, c=3, kind=KIND_NOISE, frames=1000, ref=255, dtm=7650, crf=3, batch=100, ev_per_px=1.25), 2,
                  ["--w", 3840, "--h", 2160, "--c", 3, "--kind", 1, "--crf", 3, "--cap", 1.25] if args.traffic == "all" else None))
    jitter_frames = None
    for name, kw, steps, ncu_args in specs:
        nf = kw["frames"] if not args.frames else min(kw["frames"], args.frames)
        kw = dict(kw, frames=nf)
        traffic = None
        if ncu_args is not None and args.traffic != "off":
            traffic = ncu_dram_bytes_per_px_frame(ncu_args, 16, kw["w"] * kw["h"] * kw["c"])
        wl = DeviceWorkload(A, S, **kw)
        cnt, alg = wl.count()
        ms, launches = wl.timed(steps, 1)
        pxf = wl.P * nf
        b_survey, L_mean, E_mean = survey_formula(cnt, pxf)
        achieved = alg * steps / (ms * 1e-3) / 1e9
        e = {"workload": name, "value": pxf * steps / (ms * 1e-3) / 1e6, "unit": "Mpx/s", "us_per_frame": ms * 1e3 / (steps * nf),
             "frames": nf, "steps": steps, "launches": launches, "bytes_per_px_frame": alg / pxf, "achieved_gbs": achieved,
             "frac": achieved / peak, "events_per_px_frame": E_mean, "live_nodes_mean": L_mean,
             "survey_formula_bytes_per_px_frame": b_survey, "survey_formula_frac": b_survey * pxf * steps / (ms * 1e-3) / 1e9 / peak}
        if traffic is not None:
            e["dram_bytes_per_px_frame"] = traffic[0]
            e["dram_over_counted"] = traffic[0] / traffic[1] if traffic[1] else None
            e["traffic_source"] = "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum side pass in this run, a 16-frame launch of the same plane from a fresh state"
        out.append(e)
        wl.free()
        del wl
        if name.startswith("cfg2"):
            e["e2e"] = e2e_cfg2_legs(A, args)
    return out


def oracle_video(O, w, h, c, ref, dtm, crf):
    ov = O.Video(w, h, c, O.MODE_FRAME_PERFECT)
    assert ov.time_parameters(ref * 30, ref, dtm, None)
    ov.update_crf(crf)
    return ov


def cpu_baseline_run(O, n_threads, budget_s, max_frames, min_frames=2):
    """Times the oracle port on `n_threads` host threads over the first frames of the headline workload (whole 8K plane,
    fresh state) until about budget_s seconds are spent.  Returns (Mpx/s, frames used, seconds)."""
    ov = oracle_video(O, W, H, C, REF, DTM, CRF)
    t_total, n = 0.0, 0
    for f in range(max_frames):
        fr = O.synth_frame(KIND_STATIC, SEED, f, W, H, C)
        t0 = time.perf_counter()
        ov.integrate_matrix_count_only(fr, float(REF), n_threads)
        t_total += time.perf_counter() - t0
        n += 1
        if t_total >= budget_s and n >= min_frames:
            break
    return W * H * C * n / t_total / 1e6, n, t_total


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port, all host cores) on the headline config."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    from oracle import oracle_py as O

    n_threads = O.host_threads()  # not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its workers
    sample_frames = 2
    ov = oracle_video(O, W, H, C, REF, DTM, CRF)
    P = W * H * C
    f = 0
    for _ in range(args.warmup):
        for _ in range(sample_frames):
            ov.integrate_matrix_count_only(O.synth_frame(KIND_STATIC, SEED, f, W, H, C), float(REF), n_threads)
            f += 1
    dt = 0.0
    for _ in range(args.steps):
        pre = [O.synth_frame(KIND_STATIC, SEED, f + k, W, H, C) for k in range(sample_frames)]  # frames resident before the clock starts
        t0 = time.perf_counter()
        for fr in pre:
            ov.integrate_matrix_count_only(fr, float(REF), n_threads)
        dt += time.perf_counter() - t0
        f += sample_frames
    value = P * sample_frames * args.steps / dt / 1e6
    sample = (f"{sample_frames} consecutive whole 8K frames of the workload per step (state carried across steps: frames "
              f"{args.warmup * sample_frames}..{f - 1}), frames resident in host memory")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpx/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mpx/s", "cores": n_threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "C restatement of the reference's Rust path (oracle/, OpenMP over the reference's row chunks); the Rust crate cannot be built in this image",
    }
    print(json.dumps(line))
    return 0


def bind_to_gpu_numa_node(local):
    """Run this rank on the CPUs next to its GPU (NVML's ideal affinity), so that its page-locked buffers are allocated
    on that NUMA node and the e2e legs' copies do not cross sockets.  Returns the number of CPUs bound, 0 if unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return 0


def gather_leg(A, dist, wl, rank, world, total_chunks, out_stride, gb, steps, barrier):
    """The step of `wl` with the exchange on: after every integrate launch of gb frames each band pushes its records into
    rank 0's whole-frame ring (adder_b200_comm_push_frames); rank 0 waits for the frames and releases the slots.
    Two batches in flight (two sets of band buffers), host wall clock with all streams drained, max over ranks."""
    import torch

    v, P = wl.v, wl.P
    d_ev2 = [v.device_alloc(wl.ev_stride * 12 * gb) for _ in range(2)]
    d_off2 = [v.device_alloc((v.n_chunks + 1) * 4 * gb) for _ in range(2)]
    cons = A.Exchange.consumer(v, world, total_chunks, 2 * gb, out_stride) if rank == 0 else None
    blob = [cons.export() if rank == 0 else None]
    dist.broadcast_object_list(blob, src=0)
    prod = cons.attach(v) if rank == 0 else A.Exchange.open(v, blob[0])
    seq = [0]

    def step():
        v.reset_state()
        wl.quality()
        for k, f0 in enumerate(range(0, wl.nf, gb)):
            n = min(gb, wl.nf - f0)
            b = k & 1
            v.integrate_frames_device(wl.d_frames.ptr + f0 * P, P, n, float(wl.ref), d_ev2[b].ptr, wl.ev_stride, d_off2[b].ptr)
            prod.push_frames(rank, wl.row0, d_ev2[b].ptr, wl.ev_stride, d_off2[b].ptr, n, seq[0])
            if cons is not None:  # the consumer: wait for the whole frames, "consume" them, give the slots back
                cons.wait_frames(seq[0], n)
                cons.release_frames(seq[0] + n)
            seq[0] += n

    def drain():
        v.sync()
        prod.sync()
        if cons is not None:
            cons.sync()

    step()
    drain()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    drain()
    secs = time.perf_counter() - t0
    barrier()
    t = torch.tensor([secs], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    for b in d_ev2 + d_off2:
        b.free()
    prod.close()
    if cons is not None:
        cons.close()
    return {"seconds": t[0].item(), "steps": steps, "frames_per_push": gb}


def run_ours(args):
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    nf = args.frames or NF
    n_cpus = bind_to_gpu_numa_node(local) if world > 1 and not args.no_numa else 0
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import adder_codec_rs_b200 as A
    from adder_codec_rs_b200 import sharding as S

    if A.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU port)")
    peak, peak_src = measured_peak_gbs()

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- side passes first (they need the GPU's memory for themselves): measured DRAM traffic of the headline kernel
    traffic = None
    workloads = None
    if world == 1 and rank == 0:
        if args.traffic != "off":
            # aged stacks: the launch measured is frames 592..607 of the sequence (the first frames have two-node stacks)
            traffic = ncu_dram_bytes_per_px_frame(["--w", W, "--h", H, "--c", C, "--kind", KIND_STATIC, "--crf", CRF, "--ref", REF, "--dtm", DTM,
                                                   "--cap", EV_PER_PX, "--warm-frames", 592], 16, W * H * C)
        if not args.no_workloads:
            workloads = run_side_workloads(A, S, args, peak)

    # ---- the headline: this rank's band of the 8K frame, 2000 frames resident in HBM ------------------------------
    wl = DeviceWorkload(A, S, w=W, h=H, c=C, kind=KIND_STATIC, frames=nf, ref=REF, dtm=DTM, crf=CRF, rank=rank, world=world,
                        device=local, batch=BATCH, ev_per_px=EV_PER_PX)
    v, P = wl.v, wl.P
    n_chunks = v.n_chunks
    sampler = ClockSampler(local)
    sampler.start()  # well before the timed region: with eight GPUs on the box nvidia-smi needs a second to come up
    cnt, alg_bytes_step = wl.count()  # untimed: exact algorithmic bytes of this band's step (also the first warm-up)
    ev0 = v.events_emitted()
    ms, launches = wl.timed(args.steps, max(args.warmup, 3), barrier, sampler)
    clocks = sampler.stop()
    events_per_step = (v.events_emitted() - ev0) // (args.steps + max(args.warmup, 3))
    assert events_per_step == cnt["events"], (events_per_step, cnt)

    # ---- e2e: host buffers through the C ABI -------------------------------------------------------------------
    # The step's frames come from pinned host memory in runs of `sub` frames; the pinned run is refilled from the
    # frames in HBM between calls, outside the timed calls (2000 8K frames would be 66 GB of page-locked memory).
    sub = min(100, nf)
    host_frames = np.asarray(A.pinned_empty((sub, wl.rows, W, C), np.uint8))
    wl.d_events.free()
    wl.d_events = None
    max_sub_events = int(events_per_step / nf * sub * 3) + (1 << 20)
    host_events = np.asarray(A.pinned_empty((max_sub_events,), A.EVENT_DTYPE))

    def step_host():
        v.reset_state()
        wl.quality()
        total, secs = 0, 0.0
        for f0 in range(0, nf, sub):
            n = min(sub, nf - f0)
            wl.d_frames.to_host_into(host_frames, n * P, offset=f0 * P)  # refill (untimed)
            t0 = time.perf_counter()
            ev, fc, cc = v.integrate_frames_host(host_frames[:n], float(REF), host_events)
            secs += time.perf_counter() - t0
            total += len(ev)
        return total, secs

    e2e_steps = max(1, min(args.steps, 2))
    step_host()
    barrier()
    tot, e2e_s = 0, 0.0
    for _ in range(e2e_steps):
        t, s = step_host()
        tot += t
        e2e_s += s
    barrier()
    assert tot == events_per_step * e2e_steps, (tot, events_per_step)

    # ---- gather legs (N > 1): the whole frame's events, in order, into rank 0's HBM after every batch of frames ---------
    gather = gather4 = None
    if dist is not None and not args.no_gather:
        host_events = None
        gather = gather_leg(A, dist, wl, rank, world, H, int(W * H * C * EV_PER_PX), min(50, nf), max(1, min(args.steps, 3)), barrier)
        # BASELINE configs[3] — the one that names the event all-gather: 3840x2160 RGB noise, row bands, dense event stream
        nf4 = min(200, nf)
        wl4 = DeviceWorkload(A, S, w=3840, h=2160, c=3, kind=KIND_NOISE, frames=nf4, ref=255, dtm=7650, crf=3, rank=rank, world=world,
                             device=local, batch=min(100, nf4), ev_per_px=1.25)
        ms4, _ = wl4.timed(2, 1, barrier)
        wl4.d_events.free()
        wl4.d_events = None
        g4 = gather_leg(A, dist, wl4, rank, world, 2160, int(3840 * 2160 * 3 * 1.25), min(20, nf4), 2, barrier)
        import torch

        t4 = torch.tensor([ms4], dtype=torch.float64, device="cuda")
        dist.all_reduce(t4, op=dist.ReduceOp.MAX)
        px4 = 3840 * 2160 * 3 * nf4
        gather4 = {"workload": f"cfg4: 3840x2160 RGB noise, {nf4} frames, crf 3, {world} row bands (dense event stream: ~0.93 events per px-frame)",
                   "value_gather_off": px4 * 2 / (t4[0].item() * 1e-3) / 1e6, "value_gather_on": px4 * g4["steps"] / g4["seconds"] / 1e6, "unit": "Mpx/s",
                   "frames_per_push": g4["frames_per_push"],
                   "limiter": "rank 0's NVLink ingress: ~280 MB of records per frame land on one GPU (~900 GB/s per direction), against "
                              f"{490 / world:.0f} us per frame for the band kernels"}
        wl4.free()

    # ---- reduce over ranks -----------------------------------------------------------------------------------
    if dist is not None:
        import torch

        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = t[0].item(), t[1].item()
        s = torch.tensor([float(alg_bytes_step), float(events_per_step), float(launches), float(P)], dtype=torch.float64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        alg_bytes_all, events_all, launches, P_all = s[0].item(), s[1].item(), int(s[2].item()), int(s[3].item())
    else:
        alg_bytes_all, events_all, P_all = float(alg_bytes_step), float(events_per_step), P
    assert P_all == W * H * C

    if rank == 0:
        px_step = W * H * C * nf  # the whole frame, whatever N
        value = px_step * args.steps / (ms * 1e-3) / 1e6
        e2e_value = px_step * e2e_steps / e2e_s / 1e6
        # the integrate kernel of THIS rank (its band): achieved bandwidth per GPU
        achieved = alg_bytes_step * args.steps / (ms * 1e-3) / 1e9
        pxf = P * nf
        b_survey, L_mean, E_mean = survey_formula(cnt, pxf)
        n_launch = (nf + wl.batch - 1) // wl.batch
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "kernel": "integrate_frame_kernel<8,false,true,false,true,true> (offset-form node stacks)" if wl.v.state_form == 1 else "integrate_frame_kernel<8,false,true,false,true>",
                "state_form": "offset" if wl.v.state_form == 1 else "eager",
                "frames_per_launch": wl.batch, "launches_per_step": n_launch,
                "algorithmic_bytes_per_launch": alg_bytes_step / n_launch, "algorithmic_bytes_per_px_frame": alg_bytes_step / pxf,
                "node_loads_per_px_frame": cnt["node_loads"] / pxf, "node_stores_per_px_frame": cnt["node_stores"] / pxf,
                "survey_formula": {"bytes_per_px_frame": b_survey, "L_mean_live_nodes": L_mean, "E_events_per_px_frame": E_mean,
                                   "frac": b_survey * pxf * args.steps / (ms * 1e-3) / 1e9 / peak},
                "note": "per GPU (rank 0's band); time = CUDA events around the whole timed region on the launching stream "
                        "(the integrate launches + 2 reset kernels per step); bytes counted exactly by the instrumented twin of the kernel"}
        if traffic is not None:
            roof["traffic"] = traffic[0] * W * H * C * wl.batch  # per launch of `frames_per_launch` whole-plane frames
            roof["traffic_bytes_per_px_frame"] = traffic[0]
            roof["traffic_over_algorithmic"] = traffic[0] / traffic[1] if traffic[1] else None
            roof["traffic_source"] = ("measured in this run: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum side pass over a 16-frame "
                                      "launch of the same plane with aged stacks (frames 592..607), scaled to frames_per_launch; the ratio is "
                                      "against the bytes counted for that same launch")
        line = {
            "metric": METRIC, "value": value, "unit": "Mpx/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "plane": f"{W}x{H}x{C}", "plane_per_gpu": f"{W}x{wl.rows}x{C}", "frames_per_step": nf,
                       "sharding": (f"{world} row bands of one frame, no data-path collective in `value`; each rank bound to the {n_cpus} CPUs next to its GPU"
                                    if n_cpus else f"{world} row bands of one frame, no data-path collective in `value`") if world > 1 else "single GPU",
                       "l2": "per-frame working set (headers 265 MB + live node levels + frame) exceeds the 126 MB L2; no flush needed",
                       "events_per_step": events_all, "events_per_px_frame": events_all / px_step},
            "e2e": {"value": e2e_value, "unit": "Mpx/s", "h2d_bytes_per_step": px_step, "d2h_bytes_per_step": int(events_all * 12 + (H + world) * 4 * nf),
                    "steps": e2e_steps,
                    "api": f"adder_b200_video_integrate_frames_host in runs of {sub} frames: pinned host frames in, all events out to pinned host memory; "
                           "timed = the calls (the pinned run is refilled from HBM between calls, untimed)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
        }
        if gather is not None:
            line["gather"] = {"on": True, "value": px_step * gather["steps"] / gather["seconds"] / 1e6, "unit": "Mpx/s", "steps": gather["steps"],
                              "consumer_rank": 0, "event_bytes_per_step": int(events_all * 12), "frames_per_push": gather["frames_per_push"],
                              "api": "adder_b200_comm_push_frames after every integrate launch: each band stores its compacted records at its offset in rank 0's "
                                     "whole-frame ring over NVLink peer memory (offset = inter-GPU look-back over the lower bands' totals); rank 0 waits and releases; "
                                     "host wall clock, streams drained, max over ranks",
                              "limiter": "rank 0's NVLink ingress (~900 GB/s) for dense event streams; this workload's stream is sparse, so the leg runs at the kernels' pace"}
        if gather4 is not None:
            line["gather_cfg4"] = gather4
        if workloads is not None:
            line["workloads"] = workloads
        # ---- CPU baseline on this box's host cores (bounded sample) -----------------------------------------
        if world == 1 and not args.no_cpu:
            from oracle import oracle_py as O

            nt = O.host_threads()
            mt, n_mt, s_mt = cpu_baseline_run(O, nt, args.cpu_seconds, 64)
            st, n_st, s_st = cpu_baseline_run(O, 1, min(args.cpu_seconds, 6.0), 4, min_frames=1)
            line["cpu_baseline"] = {"value": mt, "unit": "Mpx/s", "cores": nt, "kind": "port",
                                    "sample": f"first {n_mt} whole 8K frames of the same workload from a fresh state ({s_mt:.1f} s), frames resident in host memory",
                                    "single_thread": {"value": st, "frames": n_st, "seconds": s_st}}
        print(json.dumps(line))
        if workloads is not None:  # the table, readable in the tail of the log
            sys.stderr.write("workload table (1 GPU, device-resident):\n")
            for e in workloads:
                sys.stderr.write(f"  {e['value'] / 1e3:7.2f} Gpx/s  {e['us_per_frame']:8.1f} us/frame  {e['bytes_per_px_frame']:6.1f} B/px-frame  "
                                 f"frac {e['frac']:.3f}  dram/counted {e.get('dram_over_counted') or float('nan'):.2f}  {e['workload']}\n")
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=0, help="frames per step instead of the config's 2000 (debugging: not a bench number)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-workloads", action="store_true", help="skip the table of the other BASELINE configs (N = 1)")
    ap.add_argument("--traffic", default="some", choices=["off", "some", "all"], help="ncu side passes for measured DRAM traffic: headline + cfg 2 + cfg 3 c in {0,5,10} (some), every workload (all)")
    ap.add_argument("--no-gather", action="store_true", help="skip the gather leg (N > 1)")
    ap.add_argument("--no-numa", action="store_true", help="do not bind each rank to the CPUs next to its GPU (N > 1)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    world = env_int("WORLD_SIZE", 1)
    if args.impl == "ours" and args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
