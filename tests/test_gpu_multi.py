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


# ---- the exchange step inside the library (adder_b200_comm_*): bands push into the consumer's ring over peer memory ----

def _exchange_case(name, world, n_dev):
    case = cases.CASES_BY_NAME[name]
    bands = []
    for r in range(world):
        bv = S.BandedVideo(case.w, case.h, case.c, r, world, device=r % n_dev, chunk_rows=case.chunk_rows)
        cases.configure(bv, case)
        bands.append(bv)
    ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT)
    cases.configure(ov, case)
    return case, bands, ov


@pytest.mark.parametrize("name,world,batch", [("cfg2_rgb_noise_crf3", 3, 1), ("cfg3_jitter_c10", 2, 7), ("ragged_37x13x3_chunk4", 2, 3),
                                               ("jitter_dtm4_normal", 4, 5)])
def test_exchange_in_one_process_delivers_the_whole_frame_in_order(name, world, batch):
    """All bands and the consumer in this process (bands spread over the GPUs present; peer access or one device).
    Frames go through integrate_frames_device in batches, every band pushes, the consumer waits, reads each frame from
    its ring — more frames than ring slots, so slots are released and reused — and finds the oracle's whole-frame stream."""
    n_dev = A.device_count()
    case, bands, ov = _exchange_case(name, world, n_dev)
    total_chunks = ov.n_chunks
    slots = batch + 2
    cons = A.Exchange.consumer(bands[0].video, world, total_chunks, slots, case.w * case.h * case.c * 4)
    prods = [cons.attach(b.video) for b in bands]
    frames = case.frames()
    bufs = []
    for b in bands:
        P = case.w * b.rows * case.c
        bufs.append((P, b.device_alloc(P * batch), b.device_alloc(P * 4 * 12 * batch), b.device_alloc((b.n_chunks + 1) * 4 * batch)))
    seq = 0
    for f0 in range(0, case.n_frames, batch):
        n = min(batch, case.n_frames - f0)
        for r in range(world):  # inside one process the bands are queued in rank order: a push waits only for pushes queued before it,
            # so streams that share a hardware queue (CUDA_DEVICE_MAX_CONNECTIONS) cannot block each other
            b, (P, d_fr, d_ev, d_off) = bands[r], bufs[r]
            d_fr.from_host(np.ascontiguousarray(frames[f0:f0 + n, b.row0:b.row0 + b.rows]))
            b.integrate_frames_device(d_fr.ptr, P, n, case.time, d_ev.ptr, P * 4, d_off.ptr)
            prods[r].push_frames(r, b.row0 // case.chunk_rows, d_ev.ptr, P * 4, d_off.ptr, n, seq)
        cons.wait_frames(seq, n)
        cons.sync()
        for p in prods:
            p.sync()
        for k in range(n):
            eo, co = ov.integrate_matrix(frames[f0 + k], case.time)
            ev, off = cons.read_frame(seq + k, total_chunks)
            assert int(off[-1]) == len(eo), f"frame {f0 + k}"
            assert np.array_equal(np.diff(off), co), f"frame {f0 + k}: chunk lengths"
            assert ev.tobytes() == eo.tobytes(), f"frame {f0 + k}: stream"
        seq += n
        cons.release_frames(seq)
    for b in bands:
        b.sync()


def _ipc_worker(rank, world, port, out_dir):
    """One process per GPU: the consumer's ring is opened through its exported blob (CUDA IPC), the blob travels over
    torch.distributed (gloo is enough: 256 bytes)."""
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        case = cases.Case("mid_noise", 320, 96, 3, 1, 24, crf=3)
        bv = S.BandedVideo(case.w, case.h, case.c, rank, world, device=rank)
        cases.configure(bv, case)
        P = case.w * bv.rows * case.c
        batch, slots = 4, 6
        cons = A.Exchange.consumer(bv.video, world, case.h, slots, case.w * case.h * case.c * 3) if rank == 0 else None
        blob = [cons.export() if rank == 0 else None]
        dist.broadcast_object_list(blob, src=0)
        prod = cons.attach(bv.video) if rank == 0 else A.Exchange.open(bv.video, blob[0])
        d_frames = bv.device_alloc(P * batch)
        d_events = bv.device_alloc(P * 3 * 12 * batch)
        d_off = bv.device_alloc((bv.n_chunks + 1) * 4 * batch)
        ov = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT) if rank == 0 else None
        if ov is not None:
            cases.configure(ov, case)
        frames = case.frames()
        for f0 in range(0, case.n_frames, batch):
            d_frames.from_host(np.ascontiguousarray(frames[f0:f0 + batch, bv.row0:bv.row0 + bv.rows]))
            bv.integrate_frames_device(d_frames.ptr, P, batch, case.time, d_events.ptr, P * 3, d_off.ptr)
            prod.push_frames(rank, bv.row0, d_events.ptr, P * 3, d_off.ptr, batch, f0)
            prod.sync()  # the band's buffers are reused by the next batch
            if rank == 0:
                cons.wait_frames(f0, batch)
                cons.sync()
                for k in range(batch):
                    eo, co = ov.integrate_matrix(frames[f0 + k], case.time)
                    ev, off = cons.read_frame(f0 + k, case.h)
                    assert ev.tobytes() == eo.tobytes(), f"frame {f0 + k}"
                    assert np.array_equal(np.diff(off), co)
                cons.release_frames(f0 + batch)
        bv.sync()
        dist.barrier()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_exchange_between_processes_over_nvlink(tmp_path, world):
    if A.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_ipc_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
