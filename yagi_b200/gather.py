"""Optional gather of per-channel outputs across time-sharded ranks.

NOT on the hot path: the channelizer itself needs no collective.  After a time-sharded analysis every
rank holds `y_local[frames_r][M]` for its own contiguous frame range; a consumer that wants whole
channel time series on every rank (or on one rank) gathers them with one NCCL collective over
NVLink/NVSwitch (gloo on CPU for tests).  Frame ranges may differ by a couple of frames between ranks
(shards start on even frames), so the tensors are padded to the longest shard for the collective.
"""
from __future__ import annotations


def all_gather_frames(y_local, n_frames_per_rank, M: int, group=None):
    """All-gather frame-major outputs.  `y_local`: complex64 tensor of n_frames_per_rank[rank] * M samples.
    Returns a tensor [sum(n_frames_per_rank), M] (frame-major, global frame order) on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(n_frames_per_rank) == world
    assert y_local.numel() == n_frames_per_rank[rank] * M
    longest = max(n_frames_per_rank)
    send = torch.view_as_real(y_local.reshape(-1, M))
    if n_frames_per_rank[rank] < longest:
        pad = torch.zeros(longest - n_frames_per_rank[rank], M, 2, dtype=send.dtype, device=send.device)
        send = torch.cat([send, pad], dim=0)
    recv = torch.empty(world * longest, M, 2, dtype=send.dtype, device=send.device)      # concatenated along dim 0
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    recv = recv.view(world, longest, M, 2)
    parts = [torch.view_as_complex(recv[r, : n_frames_per_rank[r]].contiguous()) for r in range(world)]
    return torch.cat(parts, dim=0)


def channel_major(y_frames, out=None):
    """[frames][M] -> [M][frames]: one contiguous time series per channel (tiled transpose kernel of
    libyagi_b200.so, asynchronous on torch's current stream; CPU tensors -- the gloo tests -- are transposed by torch)."""
    import ctypes as C

    import torch

    from . import _buffers as B
    from . import _lib
    K, M = y_frames.shape
    if not y_frames.is_cuda:
        return y_frames.transpose(0, 1).contiguous()
    if y_frames.dtype != torch.complex64 or not y_frames.is_contiguous():
        raise B.ValueError_("y_frames must be a contiguous complex64 [frames][M] tensor")
    if out is None:
        out = torch.empty(M, K, dtype=torch.complex64, device=y_frames.device)
    elif out.shape != (M, K) or out.dtype != torch.complex64 or not out.is_contiguous() or out.device != y_frames.device:
        raise B.ValueError_("out must be a contiguous complex64 [M][frames] tensor on the same device")
    with torch.cuda.device(y_frames.device):
        _lib.check(_lib.lib().yg_channel_major_dev(C.c_void_p(y_frames.data_ptr()), K, M, C.c_void_p(out.data_ptr()),
                                                   B.cur_stream(y_frames)))
    return out
