"""Row-band sharding of the framed→ADΔER path over the GPUs of one box.

Pixels are independent in the reference's hot loop (video.rs:697-731), so a frame splits into bands of
whole chunks (`chunk_rows` rows each, video.rs:677-692); rank g permanently owns band g's pixel state
and nothing but frames in and events out ever moves.  Every band emits global `y` coordinates
(adder_b200_video_set_row_offset), so concatenating the bands' streams — and their per-chunk lengths —
in rank order IS the reference's `Vec<Vec<Event>>` for the whole frame.

There is no collective on the data path.  `gather_events` is the one exchange step a single downstream
consumer needs when it wants the whole frame's events in order on one rank (the reference feeds its
serial encoder that way, video.rs:736-740): an all-gather of the G counts, then a gather of the
compacted records only (padded to the largest band, trimmed on arrival).  It runs on whatever
`torch.distributed` backend the process group has: NCCL over NVLink for device tensors, gloo for the
CPU tests.  torch is plumbing here (process group + collectives); the kernels do not use it.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from . import binding as B


def band_of(height: int, chunk_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """(first_row, n_rows) of `rank`'s band: whole chunks, as even as possible, earlier ranks take the
    remainder.  A rank can end up with no rows when there are fewer chunks than ranks."""
    if not (0 <= rank < world) or height <= 0 or chunk_rows <= 0:
        raise ValueError("bad band request")
    n_chunks = (height + chunk_rows - 1) // chunk_rows
    base, extra = divmod(n_chunks, world)
    c0 = rank * base + min(rank, extra)
    nc = base + (1 if rank < extra else 0)
    row0 = min(c0 * chunk_rows, height)
    row1 = min((c0 + nc) * chunk_rows, height)
    return row0, row1 - row0


class BandedVideo:
    """This rank's band of a `width` x `height` x `channels` plane: a `Video` of the band's rows whose
    events carry frame coordinates.  Same setter names as `Video` (forwarded)."""

    def __init__(self, width: int, height: int, channels: int, rank: int, world: int, device: Optional[int] = None,
                 chunk_rows: int = 1, pixel_tree_mode: int = B.MODE_FRAME_PERFECT, max_depth: int = 0):
        self.rank, self.world = rank, world
        self.full_height = height
        self.row0, self.rows = band_of(height, chunk_rows, rank, world)
        if self.rows == 0:
            raise ValueError(f"rank {rank} of {world} gets no rows of a {height}-row plane with chunk_rows {chunk_rows}")
        self.video = B.Video(width, self.rows, channels, pixel_tree_mode, rank if device is None else device, max_depth)
        if chunk_rows != 1:
            self.video.chunk_rows(chunk_rows)
        self.video.set_row_offset(self.row0)

    def __getattr__(self, name):  # builder / setters / getters of Video
        return getattr(self.video, name)

    def band(self, frame: np.ndarray) -> np.ndarray:
        """This rank's rows of a whole (H, W, C) frame."""
        assert frame.shape[0] == self.full_height
        return frame[self.row0:self.row0 + self.rows]

    def integrate_matrix(self, frame: np.ndarray, time_spanned: float):
        """Whole frame in, this band's (events, chunk_counts) out."""
        return self.video.integrate_matrix(np.ascontiguousarray(self.band(frame)), time_spanned)


class _CudaBytes:
    """A device allocation of the C ABI seen through __cuda_array_interface__ (zero-copy into torch)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def device_bytes_as_tensor(ptr: int, nbytes: int, device: int):
    import torch

    return torch.as_tensor(_CudaBytes(ptr, max(nbytes, 1)), device=torch.device("cuda", device))[:nbytes]


def gather_events(events, chunk_counts, group=None, dst: Optional[int] = 0):
    """The exchange step: every rank passes its band's records (`events`: a uint8 torch tensor of n*12
    bytes, on the GPU for NCCL or on the CPU for gloo) and its per-chunk lengths (int64 tensor on the
    same device).  Returns (events, chunk_counts) of the whole frame in raster order on rank `dst`
    (on every rank when dst is None) and (None, None) elsewhere."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = events.device
    assert events.dtype == torch.uint8 and events.numel() % 12 == 0
    meta = torch.tensor([events.numel(), chunk_counts.numel()], dtype=torch.int64, device=dev)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)  # G counts: 16 bytes per rank
    sizes = [int(m[0]) for m in metas]
    n_chunks = [int(m[1]) for m in metas]
    if dst is None:  # every rank wants the frame: padded all-gather
        pad_e, pad_c = max(max(sizes), 1), max(max(n_chunks), 1)
        send_e = torch.zeros(pad_e, dtype=torch.uint8, device=dev)
        send_e[:events.numel()] = events
        send_c = torch.zeros(pad_c, dtype=torch.int64, device=dev)
        send_c[:chunk_counts.numel()] = chunk_counts
        recv_e = [torch.empty_like(send_e) for _ in range(world)]
        recv_c = [torch.empty_like(send_c) for _ in range(world)]
        dist.all_gather(recv_e, send_e, group=group)
        dist.all_gather(recv_c, send_c, group=group)
        ev = torch.cat([recv_e[g][:sizes[g]] for g in range(world)])  # rank order == raster order
        cc = torch.cat([recv_c[g][:n_chunks[g]] for g in range(world)])
        return ev, cc
    # one consumer: every band goes straight from where the kernel left it into its place in the consumer's frame
    # buffer (exact sizes, no padding, no staging copy): rank order == raster order
    if rank != dst:
        ops = []
        if sizes[rank]:
            ops.append(dist.P2POp(dist.isend, events, dst, group=group))
        if n_chunks[rank]:
            ops.append(dist.P2POp(dist.isend, chunk_counts.contiguous(), dst, group=group))
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        return None, None
    ev = torch.empty(sum(sizes), dtype=torch.uint8, device=dev)
    cc = torch.empty(sum(n_chunks), dtype=torch.int64, device=dev)
    e0 = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    c0 = np.concatenate([[0], np.cumsum(n_chunks)]).astype(np.int64)
    ops = []
    for g in range(world):
        if g == dst:
            continue
        if sizes[g]:
            ops.append(dist.P2POp(dist.irecv, ev[e0[g]:e0[g + 1]], g, group=group))
        if n_chunks[g]:
            ops.append(dist.P2POp(dist.irecv, cc[c0[g]:c0[g + 1]], g, group=group))
    works = dist.batch_isend_irecv(ops) if ops else []
    ev[e0[dst]:e0[dst + 1]] = events  # this rank's own band, while the others arrive
    cc[c0[dst]:c0[dst + 1]] = chunk_counts
    for w in works:
        w.wait()
    return ev, cc


def events_from_bytes(t) -> np.ndarray:
    """uint8 torch tensor (n*12 bytes) -> numpy structured array of adder_event_t."""
    return t.cpu().numpy().view(B.EVENT_DTYPE)
