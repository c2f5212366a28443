"""Multi-GPU sharding of the trial grid.

One process per GPU (``torch.distributed``; NCCL over NVLink on B200, ``gloo`` in the CPU tests).
The event list is replicated on every rank; rank r searches a contiguous slice of the frequency
axis (exact fold) or of the centre-frequency list (``--fast``); the only data-path collective is ONE
all-gather of the fp64 statistic slices.  There is nothing to fuse it with: it moves a few MB
after seconds of compute.
"""

from __future__ import annotations

import numpy as np

__all__ = ["rank", "world_size", "shard_bounds", "all_gather_concat", "sharded_search", "sharded_rows",
           "upload_replicated"]

# Host arrays at least this large are uploaded in shards (see upload_replicated); smaller ones are not worth a collective
REPLICATED_UPLOAD_MIN_BYTES = 32 << 20


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


def _to_host(x):
    """Device tensor -> numpy through ONE copy into page-locked memory (torch's caching host allocator; the array
    keeps the block alive)."""
    import torch

    if x.device.type != "cuda":
        return x.contiguous().numpy()
    host = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    host.copy_(x)
    return host.numpy()


def all_gather_concat(local, counts, axis=-1):
    """All-gather ragged fp64 slices (``counts[r]`` entries along ``axis`` on rank r) and concatenate them in rank
    order.  ``local`` is a numpy array or a torch tensor already on the backend's device (cuda for NCCL, cpu for
    gloo).  The statistics never visit the host on the way: one ``all_gather_into_tensor`` into a contiguous
    (world, ..., max count) buffer, the rank-order concatenation on the device, one device-to-host copy."""
    import torch

    d = _dist()
    ws = d.get_world_size()
    use_cuda = d.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    if not isinstance(local, torch.Tensor):
        local = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    local = local.to(device=dev, dtype=torch.float64).movedim(axis, -1)
    lead = tuple(local.shape[:-1])
    cmax = int(max(counts))
    if local.shape[-1] == cmax:
        buf = local.contiguous()
    else:
        buf = torch.zeros(lead + (cmax,), dtype=torch.float64, device=dev)
        buf[..., : local.shape[-1]] = local
    parts = torch.empty((ws,) + lead + (cmax,), dtype=torch.float64, device=dev)
    d.all_gather_into_tensor(parts.view(-1), buf.view(-1))  # flat views: accepted by NCCL and gloo alike
    if all(int(c) == cmax for c in counts) and not lead:
        out = parts.reshape(-1)
    else:
        out = torch.cat([parts[r][..., : int(c)] for r, c in enumerate(counts)], dim=-1)
    return np.moveaxis(_to_host(out), -1, axis)


def upload_replicated(a, device=None, min_bytes=None):
    """Host fp64 array that is IDENTICAL on every rank (the replicated event list) -> device tensor on every rank.

    One process: a plain H2D copy.  Several NCCL ranks and a large array: every rank copies only ITS 1/N of the array
    over its own PCIe link and one NVLink all-gather assembles the whole list on every GPU -- 8 x 0.1 GB of H2D in
    parallel plus ~2 ms of NVSwitch traffic instead of 8 concurrent 0.8 GB copies out of the same host memory (measured
    on C4 / C5 at 8 GPUs: the replicated copies took 30-50 ms per step).  The ranks first agree that they hold the
    same array (size, ends, a strided checksum): a mismatch raises on every rank instead of silently mixing inputs.
    Collective: every rank must call it."""
    import torch

    from .search import _device

    dev = _device(device)
    a = np.ascontiguousarray(a, dtype=np.float64)
    limit = REPLICATED_UPLOAD_MIN_BYTES if min_bytes is None else min_bytes
    if not is_distributed() or _dist().get_backend() != "nccl" or a.nbytes < limit:
        return torch.from_numpy(a).to(dev)
    d = _dist()
    ws, r = d.get_world_size(), d.get_rank()
    probe = a[:: max(1, a.size // 4096)]
    sig = torch.tensor(np.nan_to_num([float(a.size), float(a[0]), float(a[-1]), float(probe.sum())]), dtype=torch.float64,
                       device=dev)
    both = torch.cat([sig, -sig])
    d.all_reduce(both, op=d.ReduceOp.MAX)  # max(x) and -min(x) in one collective
    if not torch.equal(both[:4], -both[4:]):
        raise ValueError("upload_replicated: the ranks hold different arrays (size / ends / checksum differ)")
    per = (a.size + ws - 1) // ws
    per += per & 1  # even: every shard starts 16-byte aligned
    lo, hi = min(a.size, r * per), min(a.size, (r + 1) * per)
    mine = torch.zeros(per, dtype=torch.float64, device=dev)
    if hi > lo:
        mine[: hi - lo].copy_(torch.from_numpy(a[lo:hi]), non_blocking=False)
    full = torch.empty(per * ws, dtype=torch.float64, device=dev)
    d.all_gather_into_tensor(full, mine)
    return full[: a.size]


def _raise_together(err):
    """One rank failing (e.g. HxRangeError for its slice) must not leave the others blocked in the all-gather:
    agree on an error flag first and raise on every rank."""
    import torch

    d = _dist()
    use_cuda = d.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    flag = torch.tensor([1 if err is not None else 0], dtype=torch.int32, device=dev)
    d.all_reduce(flag, op=d.ReduceOp.MAX)
    if err is not None:
        raise err
    if int(flag.item()):
        raise RuntimeError("sharded search: another rank failed on its slice of the trial grid")


def _call_local(fn):
    try:
        res, err = fn(), None
    except Exception as e:  # noqa: BLE001 - re-raised by _raise_together on every rank
        res, err = None, e
    _raise_together(err)
    return res


def sharded_search(func, times, trial_freqs, fdots=0, **kwargs):
    """``func(times, trial_freqs, fdots=..., **kwargs)`` with the frequency axis sharded over ranks.

    ``func`` follows the z_n_search / epoch_folding_search convention: returns ``(fgrid, stats)`` for a
    single fdot, else ``(fgrid, fdgrid, stats)`` with stats of shape ``(n_fdot, n_freq)``.  This package's own
    searches (marked ``_hx_device_stats``) hand their slice over as a device tensor, so the statistics go
    kernel -> NCCL -> one D2H copy."""
    trial_freqs = np.atleast_1d(np.asarray(trial_freqs, dtype=np.float64))
    ws = world_size()
    if ws == 1 or trial_freqs.size < ws:
        return func(times, trial_freqs, fdots=fdots, **kwargs)
    r = rank()
    bounds = [shard_bounds(trial_freqs.size, ws, i) for i in range(ws)]
    lo, hi = bounds[r]
    if getattr(func, "_hx_device_stats", False) and _dist().get_backend() == "nccl":
        kwargs = dict(kwargs, _device_stats=True)
        from .search import _is_slow_branch

        t_np = np.asanyarray(times)
        if (t_np.dtype == np.float64 and t_np.ndim == 1 and t_np.nbytes >= REPLICATED_UPLOAD_MIN_BYTES
                and not _is_slow_branch(kwargs.get("expocorr", False), kwargs.get("weights", 1))):
            # the event list is the same on every rank: upload 1/N each, all-gather over NVLink
            times = _call_local(lambda: upload_replicated(t_np))
    res = _call_local(lambda: func(times, trial_freqs[lo:hi], fdots=fdots, **kwargs))
    fd = np.atleast_1d(np.asarray(fdots, dtype=np.float64))
    local = res[-1]
    if len(res) == 2 and fd.size == 1:
        local = local.reshape(1, -1)
    stats = all_gather_concat(local, [b[1] - b[0] for b in bounds], axis=-1)
    if len(res) == 2:
        return trial_freqs.copy(), stats.reshape(-1)
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
    local = _call_local(lambda: local_fn(rows[lo:hi]))
    return all_gather_concat(local, [b[1] - b[0] for b in bounds], axis=0)
