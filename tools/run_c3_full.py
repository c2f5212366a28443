"""The WHOLE config 3 (BASELINE.json configs[2]) through the public API: Z^2_4, nbin = 128, 1e7 events x 1e6 trial
frequencies x 64 fdots = 6.4e14 event*trials, `hendrics_b200.efsearch.folding_search` with the frequency axis sharded
over the ranks of this torchrun job (the call shape of hendrics/efsearch.py:1850-1863).  Rank 0 prints the wall time,
the top-10 candidates and a 64-trial spot check against the oracle (CPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_c3_full.py
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

from hendrics_b200 import _cabi, z_n_search
from hendrics_b200.base import find_peaks_in_image
from hendrics_b200.efsearch import folding_search
from hendrics_b200.fake import simulate_events
from hendrics_b200.io import EventList

n_f = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
T, step, fdstep, n_fd = 1e5, 5e-6, 5e-11, 64
times = simulate_events(10_000_000, T, seed=3) + 8e7
ev = EventList(time=times, gti=np.array([[8e7, 8e7 + T]]))
fmin = 7.4 if n_f == 1_000_000 else 9.90099 - step * (n_f // 2)
fmax, fdmin = fmin + step * (n_f - 1), -1.6e-9
fdmax = fdmin + fdstep * (n_fd - 1)

# warm-up on a tiny grid: CUDA context, NCCL communicator, pinned staging buffers
folding_search(ev, fmin, fmin + 63 * step, step=step, func=z_n_search, fdotmin=fdmin, fdotmax=fdmax, fdotstep=fdstep,
               nharm=4, nbin=128)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
fg, fdg, stats, step_out, fdstep_out, length = folding_search(
    ev, fmin, fmax, step=step, func=z_n_search, fdotmin=fdmin, fdotmax=fdmax, fdotstep=fdstep, nharm=4, nbin=128)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t1 = time.perf_counter()
kernel_s = _cabi.fold_last_kernel()[0] / 1e3
if rank == 0:
    n_used = int(np.count_nonzero((times - 8e7) < (times[-1] - 8e7)))  # events at the last time stamp are dropped (A2)
    units = float(n_used) * stats.size
    peaks = find_peaks_in_image(stats, n=10)
    t2 = time.perf_counter()
    top = [dict(f=float(fg[tuple(p)]), fdot=float(fdg[tuple(p)]), stat=float(stats[tuple(p)]),
                idx=[int(p[0]), int(p[1])]) for p in peaks]
    # spot check: 64 trials spread over the plane (+ the best one) against the oracle's C port
    from oracle import c_oracle as C
    rng = np.random.default_rng(0)
    ii = np.append(rng.integers(0, stats.shape[0], 63), peaks[0][0]); jj = np.append(rng.integers(0, stats.shape[1], 63), peaks[0][1])
    ts = (times - 8e7).astype(np.float64); ts = (ts - ts[0]); ts = ts[ts < ts[-1]]
    worst = 0.0
    for i, j in zip(ii, jj):
        ref = C.fold_search(ts, np.array([fg[i, j]]), np.array([fdg[i, j]]), 128, 4)[0, 0]
        worst = max(worst, abs(stats[i, j] / ref - 1))
    rec = dict(what="full C3 through efsearch.folding_search", n_gpus=world, n_events=int(times.size), grid=list(stats.shape),
               event_trials=units, wall_s=t1 - t0, value_event_trials_per_s=units / (t1 - t0),
               last_rank0_kernel_s=kernel_s, topk_s=t2 - t1, top10=top, spot_check_trials=int(ii.size),
               spot_check_max_rel_diff_vs_oracle=worst, stats_finite=bool(np.isfinite(stats).all()),
               median_stat=float(np.median(stats)))
    print(json.dumps(rec))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rec, open(f"gpurun_out/c3_full_{world}gpu.json", "w"), indent=1)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
