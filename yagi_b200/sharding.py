"""Time-block and stream sharding of the channelizer path across GPUs.

The path needs no reduction and no exchange: a shard is made independent by replicating
filter history (the reference's own precedent: src/filter/resampler/rresamp.rs:198-237,
a second object primed with history).  Pure host logic, no device code.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class TimeShard:
    rank: int
    frame_begin: int      # first output frame of this shard (global, even)
    frame_end: int        # one past the last output frame
    sample_begin: int     # first NEW input sample (global index)
    sample_end: int
    halo_begin: int       # first history sample needed (may be negative: zeros before reset)
    halo_len: int         # = (4m-1) * M/2

    @property
    def n_frames(self) -> int:
        return self.frame_end - self.frame_begin

    @property
    def n_samples(self) -> int:
        return self.sample_end - self.sample_begin


def firpfbch2_time_shards(n_frames: int, M: int, m: int, world_size: int) -> list[TimeShard]:
    """Split `n_frames` analysis frames into `world_size` contiguous shards starting on even frames.

    Frame k needs input samples t_k - 2Mm + 1 .. t_k with t_k = (k+1) M/2 - 1, so a shard that
    starts at (even) frame k0 needs (4m-1) M/2 samples of history before sample k0 M/2.
    """
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    if M < 2 or M % 2 or m < 1:
        raise ValueError("invalid channelizer geometry")
    M2 = M // 2
    halo = (4 * m - 1) * M2
    pairs = n_frames // 2
    odd = n_frames % 2
    base, rem = divmod(pairs, world_size)
    shards = []
    k = 0
    for r in range(world_size):
        nf = 2 * (base + (1 if r < rem else 0))
        if r == world_size - 1:
            nf += odd
        shards.append(TimeShard(r, k, k + nf, k * M2, (k + nf) * M2, k * M2 - halo, halo))
        k += nf
    assert k == n_frames
    return shards


def stream_shards(n_streams: int, world_size: int) -> list[range]:
    """Contiguous split of independent streams (config #5); no halo."""
    base, rem = divmod(n_streams, world_size)
    out, s = [], 0
    for r in range(world_size):
        n = base + (1 if r < rem else 0)
        out.append(range(s, s + n))
        s += n
    return out
