"""Row-band sharding on the GPU: bands concatenated in rank order must be the whole-frame oracle
stream.  The one-device test runs anywhere; the NCCL test needs two GPUs (gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest

import adder_codec_rs_b200 as A
from adder_codec_rs_b200 import sharding as S
from oracle import oracle_py as O
from tests import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", ["cfg2_rgb_noise_crf3", "ragged_37x13x3_chunk4", "jitter_dtm4_normal"])
def test_bands_on_one_device_concatenate_to_the_whole_frame(name, world):
    case = cases.CASES_BY_NAME[name]
    n_dev = A.device_count()
    bands = []
    for r in range(world):
        try:
            bv = S.BandedVideo(case.w, case.h, case.c, r, world, device=r % n_dev, chunk_rows=case.chunk_rows)
        except ValueError:
            continue  # fewer chunks than ranks
        cases.configure(bv, case)
        bands.append(bv)
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(ov, case)
    frames = case.frames()
    for f in range(case.n_frames):
        parts = [bv.integrate_matrix(frames[f], case.time) for bv in bands]
        eo, co = ov.integrate_matrix(frames[f], case.time)
        assert np.concatenate([p[0] for p in parts]).tobytes() == eo.tobytes(), f"frame {f}"
        assert np.array_equal(np.concatenate([p[1] for p in parts]), co), f"frame {f}"
    disp = np.concatenate([bv.running_intensities() for bv in bands], axis=0)
    assert np.array_equal(disp, ov.running_intensities())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        case = cases.Case("mid_noise", 320, 96, 3, 1, 10, crf=3)
        bv = S.BandedVideo(case.w, case.h, case.c, rank, world, device=rank)
        cases.configure(bv, case)
        P = case.w * bv.rows * case.c
        d_frames = bv.device_alloc(P)
        stride = P * 3
        d_events = bv.device_alloc(stride * 12)
        d_off = bv.device_alloc((bv.n_chunks + 1) * 4)
        ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT) if rank == 0 else None
        if ov is not None:
            cases.configure(ov, case)
        frames = case.frames()
        for f in range(case.n_frames):
            d_frames.from_host(np.ascontiguousarray(bv.band(frames[f])))
            bv.integrate_frames_device(d_frames.ptr, P, 1, case.time, d_events.ptr, stride, d_off.ptr)
            bv.sync()
            off = S.device_bytes_as_tensor(d_off.ptr, (bv.n_chunks + 1) * 4, rank).view(torch.int32).to(torch.int64)
            n = int(off[-1])
            ev = S.device_bytes_as_tensor(d_events.ptr, n * 12, rank)  # the band's records, still in HBM
            g_ev, g_cc = S.gather_events(ev, off[1:] - off[:-1], dst=0)  # NCCL over NVLink
            if rank == 0:
                eo, co = ov.integrate_matrix(frames[f], case.time)
                assert S.events_from_bytes(g_ev).tobytes() == eo.tobytes(), f"frame {f}"
                assert np.array_equal(g_cc.cpu().numpy(), co.astype(np.int64))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_nccl_gather_of_device_resident_events(tmp_path):
    if A.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_nccl_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(2))
