"""Row-band sharding host logic on CPU: band arithmetic, and the world_size-2 gloo run of the event
gather (each rank transcodes its band with the ORACLE — the checker standing in for a GPU — and rank 0
must end up with exactly the whole-frame oracle stream and chunk lengths)."""
import os
import socket

import numpy as np
import pytest

from adder_codec_rs_b200 import sharding as S
from tests import cases, synth


def test_band_of_covers_the_plane_in_whole_chunks():
    for h, cr, world in [(1080, 1, 8), (1080, 1, 7), (2160, 4, 8), (13, 4, 2), (13, 4, 3), (5, 64, 2), (4320, 1, 8), (7, 1, 8)]:
        rows = [S.band_of(h, cr, r, world) for r in range(world)]
        assert rows[0][0] == 0
        for (r0, n), (r1, _) in zip(rows, rows[1:]):
            assert r0 + n == r1 or (n == 0 and r0 == r1) or r1 == h
        assert sum(n for _, n in rows) == h
        for r0, n in rows:
            assert r0 % cr == 0 or n == 0  # bands start on a chunk boundary
        chunks = [(n + cr - 1) // cr for _, n in rows]
        assert max(chunks) - min(chunks) <= 1  # as even as whole chunks allow
    with pytest.raises(ValueError):
        S.band_of(10, 1, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case_name, dst, out_dir):
    import torch
    import torch.distributed as dist

    from oracle import oracle_py as O

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        case = cases.CASES_BY_NAME[case_name]
        row0, rows = S.band_of(case.h, case.chunk_rows, rank, world)
        ov = O.Video(case.w, rows, case.c, O.MODE_FRAME_PERFECT)
        cases.configure(ov, case)
        full = O.Video(case.w, case.h, case.c, O.MODE_FRAME_PERFECT) if rank == 0 or dst is None else None
        if full is not None:
            cases.configure(full, case)
        frames = case.frames()
        for f in range(case.n_frames):
            ev, cc = ov.integrate_matrix(np.ascontiguousarray(frames[f, row0:row0 + rows]), case.time)
            ev = ev.copy()
            ev["y"] += row0  # what adder_b200_video_set_row_offset does on the device
            t_ev = torch.from_numpy(ev.view(np.uint8).copy())
            t_cc = torch.from_numpy(cc.astype(np.int64))
            g_ev, g_cc = S.gather_events(t_ev, t_cc, dst=dst)
            if full is not None:
                want_ev, want_cc = full.integrate_matrix(frames[f], case.time)
                assert g_ev is not None
                assert S.events_from_bytes(g_ev).tobytes() == want_ev.tobytes(), f"rank {rank} frame {f}: gathered stream differs"
                assert np.array_equal(g_cc.numpy(), want_cc.astype(np.int64)), f"rank {rank} frame {f}: chunk lengths differ"
            else:
                assert g_ev is None and g_cc is None
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case_name,dst", [("cfg2_rgb_noise_crf3", 0), ("ragged_37x13x3_chunk4", 0), ("cfg5_static_normal", None)])
def test_two_rank_gloo_gather_equals_whole_frame_oracle(case_name, dst, tmp_path):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case_name, dst, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def _ragged_worker(rank, world, port, dst, out_dir):
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        sizes = [0, 7, 3]  # records per rank: one band with nothing to give
        chunks = [2, 1, 3]
        for rnd in range(3):
            mine = torch.full((sizes[rank] * 12,), 10 * rnd + rank, dtype=torch.uint8)
            cc = torch.arange(chunks[rank], dtype=torch.int64) + 100 * rank + rnd
            ev, gc = S.gather_events(mine, cc, dst=dst)
            if dst is None or rank == dst:
                want = torch.cat([torch.full((sizes[g] * 12,), 10 * rnd + g, dtype=torch.uint8) for g in range(world)])
                want_c = torch.cat([torch.arange(chunks[g], dtype=torch.int64) + 100 * g + rnd for g in range(world)])
                assert torch.equal(ev, want) and torch.equal(gc, want_c)
            else:
                assert ev is None and gc is None
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dst", [1, None])
def test_three_rank_gather_with_an_empty_band_and_a_consumer_that_is_not_rank_0(dst, tmp_path):
    import torch.multiprocessing as mp

    world = 3
    mp.spawn(_ragged_worker, args=(world, _free_port(), dst, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
