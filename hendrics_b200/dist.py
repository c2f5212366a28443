"""Multi-GPU sharding of the trial grid.

One process per GPU (``torch.distributed``; NCCL over NVLink on B200, ``gloo`` in the CPU tests).
The event list is replicated on every rank; rank r searches a contiguous slice of the frequency
axis (exact fold) or of the centre-frequency list (``--fast``); the only data-path collective is ONE
all-gather of the fp64 statistic slices.  There is nothing to fuse it with: it moves a few MB
after seconds of compute.
"""

from __future__ import annotations

import numpy as np

__all__ = ["rank", "world_size", "shard_bounds", "all_gather_concat", "sharded_search", "sharded_rows"]


def _dist():
    import torch.distributed as dist

    return dist


def is_distributed():
    try:
        d = _dist()
        return d.is_available() and d.is_initialized() and d.get_world_size() > 1
    except Exception:
        return False


def rank():
    return _dist().get_rank() if is_distributed() else 0


def world_size():
    return _dist().get_world_size() if is_distributed() else 1


def shard_bounds(n, world, r):
    """Contiguous, nearly equal split of ``range(n)``: the first ``n % world`` ranks get one more."""
    base, extra = divmod(int(n), int(world))
    lo = r * base + min(r, extra)
    return lo, lo + base + (1 if r < extra else 0)


def all_gather_concat(local, counts, axis=-1):
    """All-gather ragged fp64 slices (``counts[r]`` entries along ``axis`` on rank r) and
    concatenate them in rank order.  Uses the backend's device (cuda for NCCL, cpu for gloo)."""
    import torch

    d = _dist()
    local = np.moveaxis(np.asarray(local, dtype=np.float64), axis, -1)
    lead = local.shape[:-1]
    cmax = int(max(counts))
    use_cuda = d.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    buf = torch.zeros(lead + (cmax,), dtype=torch.float64, device=dev)
    if local.shape[-1] > 0:
        buf[..., : local.shape[-1]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    parts = [torch.empty_like(buf) for _ in range(d.get_world_size())]
    d.all_gather(parts, buf)
    out = np.concatenate([p[..., : int(c)].cpu().numpy() for p, c in zip(parts, counts)], axis=-1)
    return np.moveaxis(out, -1, axis)


def sharded_search(func, times, trial_freqs, fdots=0, **kwargs):
    """``func(times, trial_freqs, fdots=..., **kwargs)`` with the frequency axis sharded over ranks.

    ``func`` follows the z_n_search / epoch_folding_search convention: returns ``(fgrid, stats)`` for a
    single fdot, else ``(fgrid, fdgrid, stats)`` with stats of shape ``(n_fdot, n_freq)``."""
    trial_freqs = np.atleast_1d(np.asarray(trial_freqs, dtype=np.float64))
    ws = world_size()
    if ws == 1 or trial_freqs.size < ws:
        return func(times, trial_freqs, fdots=fdots, **kwargs)
    r = rank()
    bounds = [shard_bounds(trial_freqs.size, ws, i) for i in range(ws)]
    lo, hi = bounds[r]
    res = func(times, trial_freqs[lo:hi], fdots=fdots, **kwargs)
    stats = all_gather_concat(res[-1], [b[1] - b[0] for b in bounds], axis=-1)
    fd = np.atleast_1d(np.asarray(fdots, dtype=np.float64))
    if len(res) == 2:
        return trial_freqs.copy(), stats
    fgrid, fdgrid = np.meshgrid(trial_freqs, fd)
    return fgrid, fdgrid, stats


def sharded_rows(local_fn, rows):
    """Apply ``local_fn(rows[lo:hi]) -> array (hi-lo, ...)`` on this rank's contiguous block of
    ``rows`` and all-gather the blocks along axis 0 (the ``--fast`` centre list)."""
    rows = np.asarray(rows)
    ws = world_size()
    if ws == 1 or rows.shape[0] < ws:
        return np.asarray(local_fn(rows))
    bounds = [shard_bounds(rows.shape[0], ws, i) for i in range(ws)]
    lo, hi = bounds[rank()]
    local = np.asarray(local_fn(rows[lo:hi]))
    return all_gather_concat(local, [b[1] - b[0] for b in bounds], axis=0)
