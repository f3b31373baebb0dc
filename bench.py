#!/usr/bin/env python
"""bench.py — framed→ADΔER transcode throughput (Mpixels/s) on B200, beside the CPU path.

Workload (BASELINE.json configs[1]): 1920x1080 RGB 8-bit synthetic uniform noise, 300 frames,
crf 3 (c_thresh baseline 2 / max 7 / velocity 7), ref_time 255, delta_t_max 7650, FramePerfect,
PixelMultiMode::Collapse, TimeMode::AbsoluteT, chunk_rows 1, fresh pixel state at the start of every
step.  One step = the whole 300-frame sequence through the hot path (adder_b200_video_integrate_frames_device:
one persistent kernel launch for the 300 frames, the tail of each frame overlapping the head of the next).
At N GPUs the frame is N row bands of 1080 rows (weak scaling: each rank permanently owns one band's
state); there is no data-path collective — rank-order concatenation of the bands'
event streams is the reference's raster order.

  value     device-resident: frames already in HBM, events left in HBM, CUDA-event timed.
  e2e       the same step through adder_b200_video_integrate_frames_host: pinned HOST frames in,
            all events out to pinned HOST memory, copies inside the timed region.
  roofline  algorithmic bytes of the integrate kernel (counted exactly by the instrumented twin of
            the kernel in an untimed pass) / CUDA-event time of the timed region.
  cpu_baseline  the C oracle (a port: the Rust reference cannot be built here) on the host cores,
            on a bounded sample of the same workload.

`--impl reference` times that CPU port alone, all host threads, on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H, C, NF = 1920, 1080, 3, 300
REF, DTM, CRF = 255, 7650, 3
KIND_NOISE, SEED = 1, 0xADDE5
METRIC = "Mpixels/s framed->ADDER transcode (bit-exact events)"
WORKLOAD = "1920x1080 RGB 8-bit synthetic noise, 300 frames, crf 3 (c 2..7, velocity 7), ref 255, dtm 7650, FramePerfect/Collapse/AbsoluteT"


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
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
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
        # under load = the upper half of the samples (the sampler also sees the idle edges)
        sm_sorted = sorted(sm)
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


def ncu_traffic():
    """dram bytes per launch of the integrate kernel from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def make_oracle_video(n_rows=H):
    from oracle import oracle_py as O

    ov = O.Video(W, n_rows, C, O.MODE_FRAME_PERFECT)
    assert ov.time_parameters(REF * 30, REF, DTM, None)
    ov.update_crf(CRF)
    return ov, O


def cpu_baseline_run(n_threads, budget_s, frames_fn, max_frames=NF):
    """Times the oracle port on `n_threads` host threads over the first frames of the workload until
    about budget_s seconds are spent.  Returns (Mpx/s, frames used, seconds)."""
    ov, O = make_oracle_video()
    P = W * H * C
    t_total, n = 0.0, 0
    f = 0
    while f < max_frames:
        fr = frames_fn(f)
        t0 = time.perf_counter()
        ov.integrate_matrix_count_only(fr, float(REF), n_threads)
        t_total += time.perf_counter() - t0
        n += 1
        f += 1
        if t_total >= budget_s and n >= 4:
            break
    return P * n / t_total / 1e6, n, t_total


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port, all host threads) on the same config."""
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return 0
    from tests import synth
    from oracle import oracle_py as O

    n_threads = O.max_threads()
    sample_frames = 8
    ov, _ = make_oracle_video()
    P = W * H * C
    cache = {}

    def frame(f):
        f = f % NF
        if f not in cache:
            cache[f] = synth.frame(KIND_NOISE, SEED, f, W, H, C)
        return cache[f]

    f = 0
    for _ in range(args.warmup):
        for _ in range(sample_frames):
            fr = frame(f)
            ov.integrate_matrix_count_only(fr, float(REF), n_threads)
            f += 1
    pre = [frame(f + k) for k in range(args.steps * sample_frames)]
    t0 = time.perf_counter()
    for fr in pre:
        ov.integrate_matrix_count_only(fr, float(REF), n_threads)
    dt = time.perf_counter() - t0
    value = P * len(pre) / dt / 1e6
    sample = (f"{sample_frames} consecutive frames of the workload per step (state carried across steps, frames "
              f"{args.warmup * sample_frames}..{args.warmup * sample_frames + len(pre) - 1}), frames resident in host memory")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpx/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
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


def run_ours(args):
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    n_cpus = bind_to_gpu_numa_node(local) if world > 1 and not args.no_numa else 0
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import adder_codec_rs_b200 as A

    if A.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU port)")
    P = W * H * C
    v = A.Video(W, H, C, A.MODE_FRAME_PERFECT, device=local)
    assert v.time_parameters(REF * 30, REF, DTM, None)
    v.update_crf(CRF)
    v.set_row_offset(rank * H)  # this rank's band of the N*1080-row frame
    n_chunks = v.n_chunks

    # ---- inputs resident in HBM: this band's 300 frames --------------------------------------
    d_frames = v.device_alloc(P * NF)
    v.synth_frames(d_frames, 0, NF, KIND_NOISE, SEED + rank)
    ev_stride = P * 2  # records per frame; overflow would be reported by sync()
    d_events = v.device_alloc(ev_stride * 12 * NF)
    d_off = v.device_alloc((n_chunks + 1) * 4 * NF)

    def step_device():
        v.reset_state()
        v.update_crf(CRF)
        v.integrate_frames_device(d_frames.ptr, P, NF, float(REF), d_events.ptr, ev_stride, d_off.ptr)

    def barrier():
        if dist is not None:
            dist.barrier()

    sampler = ClockSampler(local)
    sampler.start()  # well before the timed region: with eight GPUs on the box nvidia-smi needs a second to come up
    # untimed counting pass: exact algorithmic bytes of this workload (also the first warm-up)
    v.set_counting(True)
    step_device()
    v.sync()
    cnt = v.read_counters()
    v.set_counting(False)
    for _ in range(max(args.warmup, 3)):
        step_device()
    v.sync()

    barrier()
    ev0, l0 = v.events_emitted(), v.launch_count
    sampler.mark_begin()
    v.timer_start()
    for _ in range(args.steps):
        step_device()
    ms = v.timer_stop()
    sampler.mark_end()
    clocks = sampler.stop()
    v.sync()
    barrier()
    launches = v.launch_count - l0
    events_per_step = (v.events_emitted() - ev0) // args.steps
    assert events_per_step == cnt["events"], (events_per_step, cnt)

    # algorithmic bytes of one step of the integrate kernel:
    #   per px-frame: 1 sample + 8 header read + 8 header write; per node load/store 16; per display write 1; per event 12
    alg_bytes_step = (1 + 8 + 8) * P * NF + 16 * (cnt["node_loads"] + cnt["node_stores"]) + cnt["display_writes"] + 12 * cnt["events"]

    # ---- e2e: host buffers through the C ABI --------------------------------------------------
    sub = 20
    host_frames = A.pinned_empty((NF, H, W, C), np.uint8)
    hf = np.asarray(host_frames)
    hf.reshape(-1)[:] = d_frames.to_host()
    d_events.free()
    d_off.free()
    max_sub_events = int(events_per_step / NF * sub * 1.25) + 4096
    host_events = np.asarray(A.pinned_empty((max_sub_events,), A.EVENT_DTYPE))

    def step_host():
        v.reset_state()
        v.update_crf(CRF)
        total = 0
        for f0 in range(0, NF, sub):
            ev, fc, cc = v.integrate_frames_host(hf[f0:f0 + sub], float(REF), host_events)
            total += len(ev)
        return total

    e2e_steps = max(1, min(args.steps, 3))
    step_host()
    barrier()
    t0 = time.perf_counter()
    tot = 0
    for _ in range(e2e_steps):
        tot += step_host()
    e2e_s = time.perf_counter() - t0
    barrier()
    assert tot == events_per_step * e2e_steps

    # the same step delivering the raw .adder body (11-byte wire records for RGB) instead of 12-byte records:
    # what Framed::consume + Encoder<RawOutput>::ingest_event produce.  One timed step.
    esize = v.raw_event_size
    del host_events
    host_bytes = np.asarray(A.pinned_empty((max_sub_events * esize,), np.uint8))

    def step_host_raw():
        v.reset_state()
        v.update_crf(CRF)
        total = 0
        for f0 in range(0, NF, sub):
            body, fc, cc = v.integrate_frames_host_raw(hf[f0:f0 + sub], float(REF), host_bytes)
            total += len(body)
        return total

    step_host_raw()
    barrier()
    t0 = time.perf_counter()
    raw_bytes = step_host_raw()
    raw_s = time.perf_counter() - t0
    barrier()
    assert raw_bytes == events_per_step * esize

    # ---- reduce over ranks -------------------------------------------------------------------
    if dist is not None:
        import torch

        t = torch.tensor([ms, e2e_s, raw_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, raw_s = t[0].item(), t[1].item(), t[2].item()
        s = torch.tensor([float(alg_bytes_step), float(events_per_step), float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        alg_bytes_all, events_all, launches = s[0].item(), s[1].item(), int(s[2].item())
    else:
        alg_bytes_all, events_all = float(alg_bytes_step), float(events_per_step)

    if rank == 0:
        px_step = P * NF * world
        value = px_step * args.steps / (ms * 1e-3) / 1e6
        e2e_value = px_step * e2e_steps / e2e_s / 1e6
        peak, peak_src = measured_peak_gbs()
        # per-GPU achieved bandwidth of the integrate kernel
        achieved = alg_bytes_step * args.steps / (ms * 1e-3) / 1e9
        traffic = ncu_traffic()
        # SURVEY.md §8(d)'s planning formula on the same run, for comparison: 1 + 2*(12 + 16*L) + 1[display] + 12*E with L the
        # mean live nodes per px at frame entry and exit, E the events per px-frame (both counted above).  It charges a
        # 12-byte header and every live node; `achieved` uses the smaller, exactly counted figure.
        L_mean = (cnt["live_nodes_in"] + cnt["live_nodes_out"]) / (2.0 * P * NF)
        E_mean = cnt["events"] / (P * NF)
        b_survey = 1 + 2 * (12 + 16 * L_mean) + 1 + 12 * E_mean
        survey = {"bytes_per_px_frame": b_survey, "L_mean_live_nodes": L_mean, "E_events_per_px_frame": E_mean,
                  "achieved": b_survey * P * NF * args.steps / (ms * 1e-3) / 1e9,
                  "frac": b_survey * P * NF * args.steps / (ms * 1e-3) / 1e9 / peak}
        line = {
            "metric": METRIC, "value": value, "unit": "Mpx/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "plane_per_gpu": f"{W}x{H}x{C}", "frames_per_step": NF,
                       "sharding": (f"row bands, no collective; each rank bound to the {n_cpus} CPUs next to its GPU" if n_cpus else "row bands, no collective") if world > 1 else "single GPU",
                       "l2": "per-frame working set (state 350 MB + frame + events) exceeds the 126 MB L2; no flush needed",
                       "events_per_step": events_all, "events_per_px_frame": events_all / px_step},
            "e2e": {"value": e2e_value, "unit": "Mpx/s", "h2d_bytes_per_step": P * NF * world, "d2h_bytes_per_step": int(events_all * 12 + (n_chunks + 1) * 4 * NF * world),
                    "steps": e2e_steps, "api": "adder_b200_video_integrate_frames_host, pinned host frames in, all events out to pinned host memory"},
            "e2e_raw": {"value": px_step / raw_s / 1e6, "unit": "Mpx/s", "h2d_bytes_per_step": P * NF * world, "d2h_bytes_per_step": int((raw_bytes + (n_chunks + 1) * 4 * NF) * world),
                        "steps": 1, "api": "adder_b200_video_integrate_frames_host_raw: the raw .adder stream body (wire records serialised on the device) to pinned host memory"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ((traffic or {}).get("dram_bytes_per_frame") or 0) * NF or None, "peak_source": peak_src,
                         "kernel": "integrate_frame_kernel<8,false,true>", "frames_per_launch": NF, "algorithmic_bytes_per_launch": alg_bytes_step,
                         "algorithmic_bytes_per_px_frame": alg_bytes_step / (P * NF),
                         "node_loads_per_px_frame": cnt["node_loads"] / (P * NF), "node_stores_per_px_frame": cnt["node_stores"] / (P * NF),
                         "survey_formula": survey,
                         "note": "one integrate launch spans the step's 300 frames; time = CUDA events around the whole timed region on the launching stream (that launch + 2 reset kernels per step); traffic = ncu dram bytes per frame of a 16-frame launch x 300"},
        }
        # ---- CPU baseline on this box's host cores (bounded sample) -------------------------
        if world == 1 and not args.no_cpu:
            from oracle import oracle_py as O

            nt = O.max_threads()
            mt, n_mt, s_mt = cpu_baseline_run(nt, args.cpu_seconds, lambda f: hf[f])
            st, n_st, s_st = cpu_baseline_run(1, min(args.cpu_seconds, 6.0), lambda f: hf[f], max_frames=8)
            line["cpu_baseline"] = {"value": mt, "unit": "Mpx/s", "cores": nt, "kind": "port",
                                    "sample": f"first {n_mt} frames of the same workload from a fresh state ({s_mt:.1f} s), frames resident in host memory",
                                    "single_thread": {"value": st, "frames": n_st, "seconds": s_st}}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
