#!/usr/bin/env python
"""Benchmark of the HENDRICS periodicity-search hot path on B200 ("Measurement").

    python bench.py --gpus N --steps K --warmup W [--workload c3|c1|c2|c4|c5|zn|w1|w2|ffa] [--scaling strong|weak]
                    [--impl reference]

One "step" = one pass of the search over the trial grid on a synthetic HENfake-style event list.
Default workload = the configuration BASELINE.json quotes its metric and target on (configs[2], "C3"):
Z^2_4, nbin = 128, 1e7 events x 64 fdots x a 16384-frequency slab of the 1e6-frequency grid (one GPU's
full share would take 26 s per step).  Default scaling is STRONG: the slab is FIXED and its frequency axis
is sharded over the N ranks (events replicated, one all-gather of the statistic slices per step), so
value(N) / value(1) is the fixed-problem speed-up the north star asks for.  `--scaling weak` gives every
GPU 2048 frequencies instead; `--workload c2` is BASELINE.json configs[1] (epoch folding, nbin = 16).

Prints ONE JSON line (rank 0).  `value` = event*trials/s with inputs resident in HBM (CUDA events);
`e2e` = the same metric through the public API `hendrics_b200.epoch_folding_search` / `z_n_search`
from host arrays (H2D + kernels + D2H inside the timed region); `roofline` = K1 fold kernel against
the FP64-pipe issue rate measured in the same process; `cpu_baseline` = the oracle port on this
box's host cores.  `--impl reference` times only that CPU port.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Z_n^2 / epoch-folding search throughput (event*trials/s)"
UNIT = "event*trials/s"

# fixed frequency-grid width of the strong-scaling mode (split over the ranks); weak mode: n_freq per GPU below
STRONG_N_FREQ = {"c1": 10_000, "c2": 1_000_000, "c3": 16_384, "c5": 100_000}

# name -> (description, n_events, T, n_freq per GPU, f_start, f_step, n_fdot, fdot0, fdot_step, nbin, nharm, seed)
WORKLOADS = {
    "c1": ("C1: HENzsearch Z^2_2, 1e5 events x 1e4 trial frequencies, nbin=128, no fdot",
           100_000, 1e5, 10_000, 9.876, 5e-6, 1, 0.0, 0.0, 128, 2, 1),
    "c2": ("C2: HENefsearch epoch folding nbin=16, 1e6 events x 1e6 trial frequencies",
           1_000_000, 1e5, 1_000_000, 7.4, 5e-6, 1, 0.0, 0.0, 16, 0, 2),
    "c3": ("C3 slab: Z^2_4 nbin=128, 1e7 events x 64 fdots x a slab of the 1e6-frequency grid "
           "(full config: all 1e6 frequencies sharded over 8 GPUs)",
           10_000_000, 1e5, 2048, 7.4, 5e-6, 64, -1.6e-9, 5e-11, 128, 4, 3),
    "c5": ("C5: Fermi-like long baseline, Z^2_2 nbin=128, 1e8 events over 1e8 s x 1e5 trial frequencies at step 1/(2T) "
           "(the whole config; weak mode: 12500 frequencies per GPU)",
           100_000_000, 1e8, 12_500, 9.9009, 5e-9, 1, 0.0, 0.0, 128, 2, 5),
}


# C4: HENzsearch --fast (quasi-fast folding): K2 streaming 2-D histogram + K3 shift-and-sum.
# (description, n_events, T, centres per GPU, f0, f1, nbin, nharm)
C4 = ("C4: HENzsearch --fast -f 9.85 -F 9.95 -N 2 -n 128 (binned quasi-fast folding + fdot shift search), 1e8 events, "
      "nprof=2048, 16x16 (L,Q) shifts, all 1252 centre frequencies of the band", 100_000_000, 1e5, 1252, 9.85, 9.95, 128, 2)


def algorithmic_ops_per_evtr(n_fd_nonzero):
    """6 FP64 ops per event*trial with fdot, 4 without (each op counts 1, no FMA)."""
    return 6.0 if n_fd_nonzero else 4.0


def default_scaling(wl):
    return "strong" if wl in ("c3", "c5") else "weak"


def make_inputs(wl, rank, world, scaling="weak"):
    """Returns (desc, times, all_freqs, (lo, hi), fdots, nbin, nharm): the whole job's frequency grid and this
    rank's contiguous slice of it.  weak: n_f frequencies per GPU; strong: STRONG_N_FREQ[wl] in total."""
    from hendrics_b200.dist import shard_bounds
    from hendrics_b200.fake import simulate_events

    desc, n_ev, T, n_f, f0, fstep, n_fd, fd0, fdstep, nbin, nharm, seed = WORKLOADS[wl]
    times = simulate_events(n_ev, T, seed=seed)  # identical on every rank (replicated event list)
    n_total = STRONG_N_FREQ[wl] if scaling == "strong" else n_f * world
    all_freqs = f0 + fstep * np.arange(n_total, dtype=np.float64)
    fdots = fd0 + fdstep * np.arange(n_fd, dtype=np.float64) if n_fd > 1 else np.zeros(1)
    return desc, times, all_freqs, shard_bounds(n_total, world, rank), fdots, nbin, nharm


def fold_config(desc, nbin, nharm, n_events, n_freq_total, n_fdot, n_gpus, scaling):
    """The `config` object of the JSON line -- built by BOTH arms from the same arguments, so the two lines
    describe the same job key for key."""
    return {"workload": desc, "nbin": int(nbin), "nharm": int(nharm), "n_events": int(n_events),
            "n_freq_total": int(n_freq_total), "n_fdot": int(n_fdot), "scaling": scaling,
            "parallelism": f"frequency axis of the trial grid sharded x{n_gpus}, events replicated, "
                           "one all-gather of the statistic slices per step",
            "l2": "flushed between timed iterations (256 MiB write)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]),
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows), "reasons": reasons}


def cpu_port_throughput(times, freqs, fdots, nbin, nharm, target_s=12.0):
    """The oracle's C port (oracle/hx_oracle.c, all host threads) on a bounded sample of the same
    workload: all events x the first S trial frequencies.  Returns (event*trials/s, cores, sample)."""
    from oracle import c_oracle as C

    cores = C.max_threads()
    n_ev = times.size
    ts = (times - times[0]).astype(np.float64)
    ts = ts[ts < ts[-1]]
    s = max(cores * 4, 8)
    t0 = time.perf_counter()
    C.fold_search(ts, freqs[:s], fdots[:1], nbin, nharm)
    dt = max(time.perf_counter() - t0, 1e-4)
    rate = ts.size * s / dt
    n_tr = int(min(freqs.size * fdots.size, max(s, target_s * rate / ts.size)))
    n_fd = min(fdots.size, max(1, n_tr // freqs.size)) if n_tr > freqs.size else 1
    n_f = min(freqs.size, max(1, n_tr // n_fd))
    t0 = time.perf_counter()
    C.fold_search(ts, freqs[:n_f], fdots[:n_fd], nbin, nharm)
    dt = time.perf_counter() - t0
    sample = (f"oracle C port (pthreads, {cores} threads), {ts.size} events x {n_f} freqs x {n_fd} fdots "
              f"of the workload, {dt:.1f} s")
    return ts.size * n_f * n_fd / dt, cores, sample


def run_reference(args, wl):
    """`--impl reference`: the reference's CPU algorithm (restated oracle; stingray itself is not
    installable here) on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if wl == "c4":
        return run_reference_c4(args)
    scaling = args.scaling or default_scaling(wl)
    desc, times, freqs, _, fdots, nbin, nharm = make_inputs(wl, 0, max(args.gpus, 1), scaling)
    from oracle import c_oracle as C

    cores = C.max_threads()
    ts = (times - times[0]).astype(np.float64)
    ts = ts[ts < ts[-1]]
    # calibrate a ~2 s step
    s = max(cores * 4, 8)
    t0 = time.perf_counter()
    C.fold_search(ts, freqs[:s], fdots[:1], nbin, nharm)
    rate = ts.size * s / max(time.perf_counter() - t0, 1e-4)
    n_fd = fdots.size if fdots.size > 1 else 1
    n_f = int(min(freqs.size, max(cores, 2.0 * rate / ts.size / n_fd)))
    for _ in range(args.warmup):
        C.fold_search(ts, freqs[:n_f], fdots[:n_fd], nbin, nharm)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        C.fold_search(ts, freqs[:n_f], fdots[:n_fd], nbin, nharm)
    dt = time.perf_counter() - t0
    value = ts.size * n_f * n_fd * args.steps / dt
    sample = f"per step: {ts.size} events x {n_f} freqs x {n_fd} fdots (bounded sample of the workload)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": fold_config(desc, nbin, nharm, ts.size, freqs.size, fdots.size, max(args.gpus, 1), scaling),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def c4_inputs(rank, world):
    from hendrics_b200.efsearch import _shift_grids, qffa_centres
    from hendrics_b200.fake import simulate_events

    desc, n_ev, T, ncent, f0, f1, nbin, n = C4
    times = simulate_events(n_ev, T, seed=4, amp=0.0 if n_ev > 20_000_000 else 0.25)  # thinning 1e8 events is slow; uniform
    tc = times - (times.max() + times.min()) / 2
    # weak scaling: rank r searches the band shifted by r * (f1 - f0)
    cents, _ = qffa_centres(f0 + rank * (f1 - f0), f1 + rank * (f1 - f0), T, 2)
    cents = np.asarray(cents[:ncent])
    L, Q = _shift_grids(nbin, 2, 4 * n, True)
    return desc, tc, cents, L, Q, nbin, 16 * nbin, n


def c4_cpu_step(tc, cents, L, Q, nbin, nprof, n):
    """The reference's in-tree --fast step, restated in C: phases + serial 2-D histogram (as the
    reference: one thread) + shift-and-sum on all threads."""
    from oracle import c_oracle as C

    for (mf, mfd, mfdd) in cents:
        prof = C.qffa_profiles(tc, mf, mfd, mfdd, nbin, nprof)
        C.fast_step(prof, L, Q, n)


def run_reference_c4(args):
    from oracle import c_oracle as C

    desc, tc, cents, L, Q, nbin, nprof, n = c4_inputs(0, 1)
    sample_ev = min(tc.size, 10_000_000)  # bounded sample: the first 1e7 events, one centre per step
    ts = np.ascontiguousarray(tc[:sample_ev])
    for _ in range(args.warmup):
        c4_cpu_step(ts, cents[:1], L, Q, nbin, nprof, n)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c4_cpu_step(ts, cents[:1], L, Q, nbin, nprof, n)
    dt = time.perf_counter() - t0
    value = sample_ev * args.steps / dt
    sample = f"per step: {sample_ev} events x 1 centre frequency (hist2d serial as in the reference, shift-and-sum on {C.max_threads()} threads)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc + " per GPU; one trial = one centre-frequency step", "nbin": nbin, "nharm": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": C.max_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_c4(args):
    """--fast path: value/e2e in event*centre-steps/s, roofline of K2 against measured HBM bandwidth."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from hendrics_b200 import _cabi
    from hendrics_b200 import dist as hdist
    from hendrics_b200.efsearch import _qffa_device_call

    desc, tc, cents, L, Q, nbin, nprof, n = c4_inputs(rank, world)
    dev = torch.device("cuda", local_rank)
    td = torch.from_numpy(tc).to(dev)
    units_local = float(tc.size) * len(cents)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gathered = [torch.empty((len(cents), L.size), dtype=torch.float64, device=dev) for _ in range(world)]

    def step_resident():
        stats = _qffa_device_call(td, tc[0], tc[-1], cents, nbin, nprof, L, Q, n)
        if world > 1:
            dist.all_gather(gathered, torch.from_numpy(stats.reshape(len(cents), -1)).to(dev))
        return stats

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _cabi.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms, t_wall0 = [], time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        step_resident()
        ev[k][1].record()
        kern_ms.append(_cabi.qffa_last_kernel())
    barrier()
    t_wall1 = time.perf_counter()
    launches = _cabi.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    value = units_local * world * args.steps / total_s
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    # e2e: host array in (H2D of the 0.8 GB event list every step), host stats out
    pin = torch.from_numpy(tc).pin_memory()
    pin_np = pin.numpy()
    from hendrics_b200.fastfold import _qffa_host_call

    def step_e2e():
        # what search_with_qffa does.  N ranks: each uploads 1/N of the (replicated) list + one NVLink all-gather, then
        # the device entry; one GPU: the host entry, which streams the list up while K2 folds the chunks already there
        if world > 1:
            d = hdist.upload_replicated(pin_np)
            return _qffa_device_call(d, tc[0], tc[-1], cents, nbin, nprof, L, Q, n)
        return _qffa_host_call(pin_np, cents, nbin, nprof, L, Q, n)
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = units_local * world * args.steps / float(e2e_s.item())

    if rank == 0:
        k_ms = float(np.mean([m for m, _ in kern_ms]))
        n_launch = int(kern_ms[-1][1])
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_gbs = json.load(open(peaks_file)).get("hbm_gbs") if os.path.exists(peaks_file) else 6650.0
        alg_bytes = 8.0 * units_local  # 8 B per event*centre-step
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "qffa_c4_ncu_summary.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        mix_ceiling = _cabi.microbench(5, reps=3, device=local_rank)  # DMUL + DADD.RM + LOP3 + ATOMS units/s
        roofline = {
            "bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
            "traffic": traffic, "kernel": "hx::qffa_lane_kernel", "kernel_ms": k_ms / n_launch,
            "mix_ceiling_units_per_s": mix_ceiling,
            "frac_of_mix_ceiling": units_local / (k_ms * 1e-3) / mix_ceiling,
            "kernel_launches_per_step": n_launch, "kernel_share_of_step": k_ms * args.steps / sum(step_ms),
            "algorithmic_bytes_per_event_centre_step": 8,
            "streamed_gbs": 8.0 * tc.size * n_launch / (k_ms * 1e-3) / 1e9,
            "note": "hundreds of centre frequencies (lane = centre) share each pass over the events, so DRAM traffic "
                    "is a small fraction of the algorithmic 8 B per event*centre-step and frac exceeds 1: fused this way "
                    "the path is no longer HBM-bound but bound by the shared-memory atomic pipe, see frac_of_mix_ceiling "
                    "(fraction of the DMUL+DADD+LOP3+ATOMS issue ceiling measured in this process); streamed_gbs is "
                    "the bandwidth actually drawn",
            "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if os.path.exists(peaks_file) else "fallback",
        }
        cpu = None
        if not args.no_cpu_baseline:
            sample_ev = 10_000_000
            ts = np.ascontiguousarray(tc[:sample_ev])
            t0 = time.perf_counter()
            c4_cpu_step(ts, cents[:2], L, Q, nbin, nprof, n)
            dt = time.perf_counter() - t0
            from oracle import c_oracle as C
            cpu = {"value": sample_ev * 2 / dt, "unit": UNIT, "cores": C.max_threads(), "kind": "port",
                   "sample": f"{sample_ev} events x 2 centre steps, {dt:.1f} s; hist2d serial (as the reference), "
                             f"shift-and-sum on {C.max_threads()} threads"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc + " per GPU; one trial = one centre-frequency step (each yields 256 (f, fdot) grid points)",
                       "nbin": nbin, "nharm": n, "n_events": int(tc.size), "n_centres_per_gpu": len(cents),
                       "parallelism": f"centre list sharded x{world}, events replicated, one all-gather of stats",
                       "l2": "inputs (0.8 GB) exceed L2; L2 also flushed between timed iterations"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(tc.nbytes), "d2h_bytes_per_step": int(len(cents) * L.size * 8)},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --ffa (widening row: hendrics/ffa.py, efsearch.py:1092-1138): the whole `search_with_ffa` of a C4-like event list.
# (description, n_events, T, f0, f1, nbin)
FFA = ("FFA: HENzsearch --ffa -f 9.89 -F 9.95 -n 128 on 1e8 events over 1e5 s: 1.27e8-bin light curve, 2^20 tables x "
       "128..129 bins, 20 folding steps per trial bin period", 100_000_000, 1e5, 9.89, 9.95, 128)
FFA_UNIT = "bin*steps/s"
FFA_METRIC = "fast-folding search throughput (light-curve bins x folding steps per second)"


def ffa_inputs():
    from hendrics_b200.fake import simulate_events

    desc, n_ev, T, f0, f1, nbin = FFA
    ts = simulate_events(n_ev, T, seed=6)
    return desc, ts, f0, f1, nbin


def ffa_units(n_bins, nbin, f0, f1):
    """table elements x folding steps summed over the integer bin periods of the search (ffa.py:366-395, no rebinning
    for f1/f0 < 2)"""
    p0, p1 = 1 / f1, 1 / f0
    dt = p0 / nbin
    tot = 0.0
    for bp in range(int(np.floor(p0 / dt)), int(np.ceil(p1 / dt)) + 1):
        ntab = int(2 ** np.ceil(np.log2(n_bins // bp + 1)))
        tot += float(ntab) * bp * np.log2(ntab)
    return tot


def run_ffa_reference(args):
    """CPU arm: the oracle's restatement of search_with_ffa (numba, as the reference) on a bounded sample."""
    from oracle import hx_oracle as O

    desc, n_ev, T, f0, f1, nbin = FFA
    n_s, T_s = 1_000_000, 1e3  # same rate, 1 % of the exposure: 2^13 tables, 13 steps
    from hendrics_b200.fake import simulate_events
    ts = simulate_events(n_s, T_s, seed=6)
    O.search_with_ffa(ts[:20000], f0, f1, nbin=nbin)  # numba compilation
    times = []
    for _ in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.search_with_ffa(ts, f0, f1, nbin=nbin)
        times.append(time.perf_counter() - t0)
    dt = float(np.mean(times[args.warmup:]))
    n_bins = int(np.rint((ts[-1] - ts[0]) / (1 / f1 / nbin)))
    v = ffa_units(n_bins, nbin, f0, f1) / dt
    line = {"impl": "reference", "metric": FFA_METRIC, "value": v, "unit": FFA_UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "config": {"workload": desc, "nbin": nbin},
            "cpu_baseline": {"value": v, "unit": FFA_UNIT, "cores": 1, "kind": "port",
                             "sample": f"search_with_ffa on {n_s} events over {T_s:g} s (serial numba loops, like the reference)"},
            "e2e": {"value": v, "unit": FFA_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ffa(args):
    """--ffa: value = the FFAs of the search on a device-resident light curve; e2e = search_with_ffa from host events."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    from hendrics_b200 import _cabi
    from hendrics_b200.efsearch import search_with_ffa
    from hendrics_b200.ffa import ffa_search

    desc, ts, f0, f1, nbin = ffa_inputs()
    dev = torch.device("cuda", local_rank)
    p0, p1 = 1 / f1, 1 / f0
    dt = p0 / nbin
    length = float(ts[-1] - ts[0])
    bins = int(np.rint(length / dt))
    x = torch.from_numpy(ts - ts[0]).to(dev)
    counts = torch.empty(bins, dtype=torch.int32, device=dev)
    _cabi.check(_cabi.lib().hx_hist1d_dev(x.data_ptr(), x.numel(), 0.0, 0.0, length, bins, counts.data_ptr(), 0))
    torch.cuda.synchronize()
    units = ffa_units(bins, nbin, f0, f1)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    kern = []
    def step_resident():
        # the loop of ffa_search, keeping the device time of every hx_ffa_dev call
        ms = 0.0
        from hendrics_b200.ffa import _ffa_device
        for bp in range(int(np.floor(p0 / dt)), int(np.ceil(p1 / dt)) + 1):
            _ffa_device(counts, bp, reuse_host_buffer=True)  # as ffa_search calls it
            k = _cabi.ffa_last_kernel()
            ms += k[0]
        kern.append((ms, k[1], k[2], k[3]))

    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()
    kern.clear()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _cabi.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        step_resident()
        ev[k][1].record()
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    launches = _cabi.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_s = sum(step_ms) * 1e-3
    value = units * world * args.steps / total_s
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    pin = torch.from_numpy(ts).pin_memory().numpy()
    search_with_ffa(pin, f0, f1, nbin=nbin)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fr, st, _, _ = search_with_ffa(pin, f0, f1, nbin=nbin)
    torch.cuda.synchronize()
    e2e_value = units * world * args.steps / (time.perf_counter() - t0)

    if rank == 0:
        k_ms = float(np.mean([m for m, *_ in kern]))
        passes, steps, kk = kern[-1][1:]
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_gbs = json.load(open(peaks_file)).get("hbm_gbs") if os.path.exists(peaks_file) else 6650.0
        achieved = 8.0 * units / (k_ms * 1e-3) / 1e9  # one read + one write of every table element per step
        roofline = {
            "bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs, "traffic": None,
            "kernel": "hx::ffa_pass_kernel", "kernel_ms": k_ms, "kernel_share_of_step": k_ms * args.steps / sum(step_ms),
            "algorithmic_bytes_per_bin_step": 8, "passes_over_hbm_per_ffa": int(passes), "steps_per_ffa": int(steps),
            "steps_fused_per_pass": int(kk), "streamed_gbs": achieved * passes / steps,
            "note": "k folding steps are fused per pass in shared memory, so HBM sees 8 B per bin and PASS, not per step: frac "
                    "counts the step-by-step algorithm's bytes and exceeds the traffic actually drawn (streamed_gbs); the "
                    "fused kernel is bound by the shared-memory pipe / instruction issue of the in-SM fold (DESIGN.md K5)",
            "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if os.path.exists(peaks_file) else "fallback",
        }
        line = {
            "metric": FFA_METRIC, "value": value, "unit": FFA_UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": desc, "nbin": nbin, "n_events": int(ts.size), "light_curve_bins": bins,
                       "parallelism": "replicas only (a search has 2-3 bin periods; nothing to shard)",
                       "l2": "light curve (0.5 GB) exceeds L2; L2 also flushed between timed iterations"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": FFA_UNIT, "h2d_bytes_per_step": int(ts.nbytes), "d2h_bytes_per_step": int(st.nbytes)},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": None,
        }
        if not args.no_cpu_baseline:
            from oracle import hx_oracle as O
            from hendrics_b200.fake import simulate_events
            tsamp = simulate_events(1_000_000, 1e3, seed=6)
            O.search_with_ffa(tsamp[:20000], f0, f1, nbin=nbin)
            t0 = time.perf_counter()
            O.search_with_ffa(tsamp, f0, f1, nbin=nbin)
            dtc = time.perf_counter() - t0
            nb = int(np.rint((tsamp[-1] - tsamp[0]) / dt))
            line["cpu_baseline"] = {"value": ffa_units(nb, nbin, f0, f1) / dtc, "unit": FFA_UNIT, "cores": 1, "kind": "port",
                                    "sample": f"search_with_ffa on 1e6 events over 1e3 s (same rate, 1 % of the exposure), {dtc:.1f} s; "
                                              "serial numba loops like the reference"}
        print(json.dumps(line), flush=True)


# Weighted / exposure-corrected fold (K1w): light-curve input (weights = counts per bin) and `--expocorr`.
# name -> (description, n_bins_or_events, T, n_freq per GPU, f_start, f_step, nbin, nharm, weights, n_gti, seed)
WEIGHTED = {
    "w1": ("W1: HENzsearch from a light curve (weights = counts per bin), Z^2_2 nbin=64, 2e6 time bins x 1e5 trial frequencies",
           2_000_000, 1e5, 100_000, 9.4, 5e-6, 64, 2, "counts", 0, 11),
    "w2": ("W2: HENefsearch --expocorr with fractional weights, nbin=32, 1e6 events x 1e5 trial frequencies, 50 GTIs",
           1_000_000, 1e5, 100_000, 9.4, 5e-6, 32, 0, "float", 50, 12),
}


def weighted_inputs(wl, rank, world):
    from hendrics_b200.fake import simulate_events

    desc, n_ev, T, n_f, f0, fstep, nbin, nharm, wkind, n_gti, seed = WEIGHTED[wl]
    rng = np.random.default_rng(seed)
    if wkind == "counts":  # a light curve: photons of a pulsed source binned at dt = T / n_ev
        ph = simulate_events(6 * n_ev, T, seed=seed)
        edges = np.linspace(0.0, T, n_ev + 1)
        weights = np.histogram(ph, bins=edges)[0].astype(np.float64)
        times = (edges[:-1] + edges[1:]) / 2
    else:
        times = simulate_events(n_ev, T, seed=seed)
        weights = rng.uniform(0.5, 1.5, n_ev)
    gti = None
    if n_gti:
        e = np.linspace(0.0, T, n_gti + 1)
        gti = np.stack([e[:-1], e[1:] - 0.37 * (e[1] - e[0])], axis=1)
        keep = np.zeros(times.size, bool)
        for g0, g1 in gti:
            keep |= (times >= g0) & (times < g1)
        times, weights = times[keep], weights[keep]
    freqs = f0 + fstep * (rank * n_f + np.arange(n_f, dtype=np.float64))
    return desc, times, weights, gti, freqs, nbin, nharm


def weighted_cpu_rate(times, weights, gti, freqs, nbin, nharm, target_s=10.0):
    """The oracle's restatement of upstream's fold_events branch (numpy, 1 thread -- upstream's trial loop is
    serial Python) on the first S trial frequencies."""
    from oracle import hx_oracle as O

    kw = dict(nbin=nbin, weights=weights, expocorr=gti is not None, gti=gti)
    fn = (lambda f: O.epoch_folding_search(times, f, **kw)) if nharm == 0 else (lambda f: O.z_n_search(times, f, nharm=nharm, **kw))
    t0 = time.perf_counter()
    fn(freqs[:2])
    per = (time.perf_counter() - t0) / 2
    s = int(max(2, min(freqs.size, target_s / per)))
    t0 = time.perf_counter()
    fn(freqs[:s])
    dt = time.perf_counter() - t0
    return (times.size - 1) * s / dt, 1, f"oracle numpy restatement of fold_events (np.histogram per trial, 1 thread), {times.size - 1} events x {s} freqs, {dt:.1f} s"


def run_weighted_reference(args):
    wl = args.workload
    desc, times, weights, gti, freqs, nbin, nharm = weighted_inputs(wl, 0, 1)
    if int(os.environ.get("RANK", "0")) != 0:
        return
    vals = []
    for k in range(args.warmup + args.steps):
        v, cores, sample = weighted_cpu_rate(times, weights, gti, freqs, nbin, nharm, target_s=8.0)
        if k >= args.warmup:
            vals.append(v)
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc + " per GPU", "nbin": nbin, "nharm": nharm, "n_events": int(times.size)},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_weighted(args):
    """K1w: value/e2e in (weighted) event*trials/s; roofline of wfold_kernel against the FP64 pipe like K1."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from hendrics_b200 import _cabi, epoch_folding_search, z_n_search
    from hendrics_b200.search import wfold_stats_device

    wl = args.workload
    desc, times, weights, gti, freqs, nbin, nharm = weighted_inputs(wl, rank, world)
    dev = torch.device("cuda", local_rank)
    # event selection of the public API (A1/A2) done once for the resident leg
    sel = (times - times[0]) < (times[-1] - times[0])
    ts_host, w_host = (times - times[0])[sel], weights[sel]
    ts_d, w_d = torch.from_numpy(ts_host).to(dev), torch.from_numpy(w_host).to(dev)
    f_d = torch.from_numpy(freqs).to(dev)
    fd_d = torch.zeros(1, dtype=torch.float64, device=dev)
    kw = dict(gti=gti, start_time=ts_host[0], stop_time=ts_host[-1] + abs(ts_host[1] - ts_host[0]))
    units_local = float(ts_host.size) * freqs.size
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gathered = [torch.empty((1, freqs.size), dtype=torch.float64, device=dev) for _ in range(world)]

    def step_resident():
        stats = wfold_stats_device(ts_d, w_d, 1.0, f_d, fd_d, nbin, nharm, **kw)
        if world > 1:
            dist.all_gather(gathered, stats)
        return stats

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _cabi.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms, t_wall0 = [], time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        step_resident()
        ev[k][1].record()
        kern_ms.append(_cabi.wfold_last_kernel()[0])
    barrier()
    t_wall1 = time.perf_counter()
    launches = _cabi.launch_count() - launches0
    cell = _cabi.wfold_last_kernel()[1]
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    value = units_local * world * args.steps / total_s
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    pin_t, pin_w, pin_f = (torch.from_numpy(x).pin_memory() for x in (times, weights, freqs))
    t_np, w_np, f_np = pin_t.numpy(), pin_w.numpy(), pin_f.numpy()
    api_kw = dict(nbin=nbin, weights=w_np, expocorr=gti is not None, gti=gti)

    def step_e2e():
        res = epoch_folding_search(t_np, f_np, **api_kw) if nharm == 0 else z_n_search(t_np, f_np, nharm=nharm, **api_kw)
        stats = res[-1]
        if world > 1:
            from hendrics_b200 import dist as hdist
            stats = hdist.all_gather_concat(stats, [freqs.size] * world, axis=-1)
        return stats

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = units_local * world * args.steps / float(e2e_s.item())

    if rank == 0:
        k_ms = float(np.mean(kern_ms))
        peak_ops = _cabi.microbench(1, reps=5, device=local_rank)
        mix_ceiling = _cabi.microbench(5, reps=3, device=local_rank)
        achieved = 4.0 * units_local / (k_ms * 1e-3)
        traffic = None
        prof = os.path.join(ROOT, "profiles", f"wfold_{wl}_ncu_summary.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {
            "bound": "fp64_pipe", "achieved": achieved * 1e-12, "peak": peak_ops * 1e-12, "unit": "TFLOP/s",
            "frac": achieved / peak_ops, "traffic": traffic, "kernel": "hx::wfold_kernel",
            "cells": {0: "fp64 (LDS.64 + DADD + STS.64 on private columns)", 1: "uint32 (red.shared.add.u32, integer weights)",
                      2: "uint32 (unit increments)"}[cell],
            "kernel_ms": k_ms, "kernel_share_of_step": k_ms * args.steps / sum(step_ms),
            "algorithmic_fp64_ops_per_event_trial": 4.0,
            "mix_ceiling_units_per_s": mix_ceiling, "frac_of_mix_ceiling": units_local / (k_ms * 1e-3) / mix_ceiling,
            "peak_source": "measured in this process: hx_microbench kind 1 (DMUL+DADD issue rate, non-FMA op/s)",
        }
        cpu = None
        if not args.no_cpu_baseline:
            v, cores, sample = weighted_cpu_rate(times, weights, gti, freqs, nbin, nharm)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc + " per GPU", "nbin": nbin, "nharm": nharm, "n_events": int(ts_host.size),
                       "n_freq_per_gpu": int(freqs.size), "n_gti": 0 if gti is None else int(len(gti)),
                       "parallelism": f"trial-grid sharding x{world}, events replicated, one all-gather of stats",
                       "l2": "flushed between timed iterations (256 MiB write)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(times.nbytes + weights.nbytes + freqs.nbytes),
                    "d2h_bytes_per_step": int(freqs.size * 8)},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# K4: unbinned Z^2_n (the north star's own kernel sketch: one sin/cos per event*trial + angle-addition recurrence)
# ---------------------------------------------------------------------------------------------
# (description, n_events, T, n_freq per GPU, f0, f_step, n_fdot, fdot0, fdot_step, nharm, seed)
ZN = {
    "zn": ("ZN: unbinned Z^2_2 (z_n_fast over an (f, fdot) grid), 1e6 events x 8192 frequencies x 8 fdots per GPU",
           1_000_000, 1e5, 8192, 9.88, 5e-6, 8, -2e-10, 5e-11, 2, 7),
    "zn4": ("ZN4: unbinned Z^2_4, 1e6 events x 8192 frequencies x 8 fdots per GPU",
            1_000_000, 1e5, 8192, 9.88, 5e-6, 8, -2e-10, 5e-11, 4, 7),
    "zn1": ("ZN1: unbinned Z^2_2 on the C1 shape, 1e5 events x 1e4 frequencies, no fdot",
            100_000, 1e5, 10_000, 9.876, 5e-6, 1, 0.0, 0.0, 2, 1),
}


def zn_inputs(wl, rank, world):
    from hendrics_b200.fake import simulate_events

    desc, n_ev, T, n_f, f0, fstep, n_fd, fd0, fdstep, nharm, seed = ZN[wl]
    times = simulate_events(n_ev, T, seed=seed)
    all_freqs = f0 + fstep * np.arange(n_f * world, dtype=np.float64)
    fdots = fd0 + fdstep * np.arange(n_fd, dtype=np.float64) if n_fd > 1 else np.zeros(1)
    return desc, times, all_freqs, (rank * n_f, (rank + 1) * n_f), fdots, nharm


def zn_config(desc, nharm, n_events, n_freq_total, n_fdot, n_gpus):
    return {"workload": desc, "nharm": int(nharm), "n_events": int(n_events), "n_freq_total": int(n_freq_total),
            "n_fdot": int(n_fdot), "scaling": "weak",
            "parallelism": f"frequency axis of the trial grid sharded x{n_gpus}, events replicated, "
                           "one all-gather of the statistic slices per step",
            "l2": "flushed between timed iterations (256 MiB write)"}


def zn_cpu_rate(times, freqs, fdots, nharm, target_s=8.0):
    """The oracle's unbinned Z^2_n (numpy, one thread: the reference's z_n_fast is a serial numba function) on the
    first S trials of the workload."""
    from oracle import hx_oracle as O

    def one(f, fd):
        ph = times * f + 0.5 * times * times * fd
        ph -= np.floor(ph)
        return O.z_n_events(ph, nharm)

    t0 = time.perf_counter()
    one(freqs[0], fdots[0])
    per = max(time.perf_counter() - t0, 1e-5)
    n = int(max(2, min(freqs.size, target_s / per)))
    t0 = time.perf_counter()
    for j in range(n):
        one(freqs[j], fdots[0])
    dt = time.perf_counter() - t0
    return times.size * n / dt, 1, f"oracle z_n_events (numpy, 1 thread), {times.size} events x {n} trials, {dt:.1f} s"


def run_zn_reference(args, wl):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    desc, times, freqs, _, fdots, nharm = zn_inputs(wl, 0, max(args.gpus, 1))
    rate, cores, sample = zn_cpu_rate(times, freqs, fdots, nharm, target_s=3.0 * max(args.steps, 1))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": zn_config(desc, nharm, times.size, freqs.size, fdots.size, max(args.gpus, 1)),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_zn(args, wl):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from hendrics_b200 import _cabi, z_n_search_unbinned

    desc, times, all_freqs, (lo, hi), fdots, nharm = zn_inputs(wl, rank, world)
    freqs = np.ascontiguousarray(all_freqs[lo:hi])
    dev = torch.device("cuda", local_rank)
    ts_d = torch.from_numpy(times).to(dev)
    f_d = torch.from_numpy(freqs).to(dev)
    fd_d = torch.from_numpy(fdots).to(dev)
    stats_d = torch.empty((fdots.size, freqs.size), dtype=torch.float64, device=dev)
    gathered = torch.empty((world, fdots.size, freqs.size), dtype=torch.float64, device=dev)
    units_local = float(times.size) * freqs.size * fdots.size
    units_total = units_local * world
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    lib = _cabi.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        _cabi.check(lib.hx_zn_unbinned_dev(ts_d.data_ptr(), ts_d.numel(), f_d.data_ptr(), f_d.numel(), fd_d.data_ptr(),
                                           fd_d.numel(), nharm, stats_d.data_ptr(), torch.cuda.current_stream().cuda_stream))
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), stats_d.view(-1))

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _cabi.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms, ops = [], 0
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        step_resident()
        ev[k][1].record()
        ms, ops = _cabi.zn_last_kernel()
        kern_ms.append(ms)
    barrier()
    t_wall1 = time.perf_counter()
    launches = _cabi.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    value = units_total * args.steps / total_s
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    kern_avg_ms = float(np.mean(kern_ms))

    # e2e: the public API with host arrays (H2D of events + grid, kernels, D2H of the statistics), this rank's slice
    pin_t = torch.from_numpy(times).pin_memory().numpy()
    for _ in range(2):
        z_n_search_unbinned(pin_t, freqs, nharm=nharm, fdots=fdots)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = z_n_search_unbinned(pin_t, freqs, nharm=nharm, fdots=fdots)[-1]
    barrier()
    assert out.size == freqs.size * fdots.size
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = units_total * args.steps / float(e2e_s.item())

    if rank == 0:
        peak_dfma = _cabi.microbench(0, reps=5, device=local_rank)  # FP64 instruction issue rate (DFMA), inst/s
        exec_rate = ops * units_local / (kern_avg_ms * 1e-3)
        ncu = ncu_record(wl)
        roofline = {
            "bound": "fp64_pipe", "achieved": exec_rate * 1e-12, "peak": peak_dfma * 1e-12, "unit": "TFLOP/s",
            "frac": exec_rate / peak_dfma,
            "note": "achieved = FP64 INSTRUCTIONS issued per second (an FMA counts 1, like a DMUL or a DADD: the pipe "
                    "issues one per lane per slot whatever the opcode); peak = measured DFMA issue rate",
            "executed_fp64_ops_per_event_trial": ops,
            "ncu": ncu, "traffic": (ncu or {}).get("dram_bytes_per_launch"), "traffic_source": (ncu or {}).get("source"),
            "kernel": "hx::zn_unbinned_kernel", "kernel_ms": kern_avg_ms,
            "kernel_share_of_step": kern_avg_ms * args.steps / sum(step_ms),
            "peak_source": "measured in this process: hx_microbench kind 0 (DFMA issue rate); MEASURED_PEAKS.json has no "
                           "FP64 entry",
        }
        cpu = None
        if not args.no_cpu_baseline:
            v, cores, sample = zn_cpu_rate(times, all_freqs, fdots, nharm)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": zn_config(desc, nharm, times.size, all_freqs.size, fdots.size, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(times.nbytes + freqs.nbytes + fdots.nbytes),
                    "d2h_bytes_per_step": int(freqs.size * fdots.size * 8),
                    "api": "hendrics_b200.z_n_search_unbinned(times, freqs, nharm, fdots) per rank on its slice"},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# orbital-demodulation fold (folding_orbital_search's T0 loop, one parameter row per launch)
# ---------------------------------------------------------------------------------------------
ORB = ("ORB: folding_orbital_search, one (freq, fdot, X, Porb) row: 1e6 events x 4736 trial epochs T0 per GPU, Z^2_2 nbin=64",
       1_000_000, 1e5, 4736, 9.9, 1e-10, 3.0, 86400.0, 64, 2)
ORB_UNIT = "event*trials/s"


def orb_inputs(rank):
    from hendrics_b200.fake import simulate_events

    desc, n_ev, T, n_t0, freq, fdot, X, Porb, nbin, nharm = ORB
    times = simulate_events(n_ev, T, seed=6)
    T0s = np.random.default_rng(100 + rank).uniform(0, Porb, n_t0)  # every rank scans its own epochs (weak scaling)
    return desc, times, T0s, freq, fdot, X, Porb, nbin, nharm


def orb_config(desc, n_events, n_t0, nbin, nharm, n_gpus):
    return {"workload": desc, "nbin": int(nbin), "nharm": int(nharm), "n_events": int(n_events), "n_t0_per_gpu": int(n_t0),
            "scaling": "weak", "parallelism": f"trial epochs sharded x{n_gpus}, events replicated, no collective",
            "l2": "flushed between timed iterations (256 MiB write)"}


def run_orb_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import hx_oracle as O

    desc, times, T0s, freq, fdot, X, Porb, nbin, nharm = orb_inputs(0)
    n = max(2, min(T0s.size, 8 * max(args.steps, 1)))
    t0 = time.perf_counter()
    O.orbital_stats(times, freq, fdot, X, Porb, T0s[:n], fun=O.z_n_search, nbin=nbin, nharm=nharm)
    dt = time.perf_counter() - t0
    rate = times.size * n / dt
    sample = f"oracle orbital_stats (numpy, 1 thread: the reference loops over T0 in Python), {times.size} events x {n} T0, {dt:.1f} s"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": ORB_UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": orb_config(desc, times.size, T0s.size, nbin, nharm, max(args.gpus, 1)),
            "cpu_baseline": {"value": rate, "unit": ORB_UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": ORB_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_orb(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from hendrics_b200 import _cabi
    from hendrics_b200.search import orbital_fold_stats

    desc, times, T0s, freq, fdot, X, Porb, nbin, nharm = orb_inputs(rank)
    dev = torch.device("cuda", local_rank)
    td = torch.from_numpy(times).to(dev)
    t0d = torch.from_numpy(T0s).to(dev)
    stats = torch.empty(T0s.size, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    lib = _cabi.lib()
    units_local = float(times.size) * T0s.size

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        _cabi.check(lib.hx_orbital_search_dev(td.data_ptr(), td.numel(), freq, fdot, X, Porb, t0d.data_ptr(), T0s.size, nbin,
                                              nharm, stats.data_ptr(), None, torch.cuda.current_stream().cuda_stream))

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _cabi.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        step_resident()
        ev[k][1].record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = _cabi.launch_count() - launches0
    total_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    value = units_local * world * args.steps / total_s
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    pin = torch.from_numpy(times).pin_memory().numpy()
    for _ in range(2):
        orbital_fold_stats(pin, freq, fdot, X, Porb, T0s, nbin=nbin, nharm=nharm)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = orbital_fold_stats(pin, freq, fdot, X, Porb, T0s, nbin=nbin, nharm=nharm)
    barrier()
    assert out.size == T0s.size
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = units_local * world * args.steps / float(e2e_s.item())
    if rank == 0:
        # the kernel is a long chain of dependent instructions per unit (two sin(), two IEEE divisions, the fold): its
        # ceiling is the issue rate.  Instructions per unit come from the committed ncu capture (smsp__inst_executed.sum
        # over the units of that launch), labelled with its file; the issue rate is 4 warp instructions per clk and SM.
        n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
        inst_per_unit, src = None, None
        f = os.path.join(ROOT, "profiles", "r02_ncu_orbital_top.json")
        if os.path.exists(f):
            try:
                d = json.load(open(f))
                inst_per_unit = float(str(d["smsp__inst_executed.sum"]).split()[0]) * 32.0 / (1_000_000 * 4736)
                src = "profiles/r02_ncu_orbital_top.json (ncu --set full on tools/quick_orbital.py, same shape)"
            except Exception:
                inst_per_unit = None
        clk = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
        peak_inst = n_sm * 4 * 32 * clk  # thread-instructions per second
        achieved = (value / world) * inst_per_unit if inst_per_unit else None
        roofline = {"bound": "issue_slots", "achieved": achieved and achieved * 1e-12, "peak": peak_inst * 1e-12,
                    "unit": "T thread-instructions/s", "frac": achieved and achieved / peak_inst,
                    "instructions_per_event_trial": inst_per_unit, "instructions_source": src, "traffic": None,
                    "kernel": "hx::orbital_fold_kernel", "kernel_ms": total_s / args.steps * 1e3}
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import hx_oracle as O

            t1 = time.perf_counter()
            O.orbital_stats(times, freq, fdot, X, Porb, T0s[:16], fun=O.z_n_search, nbin=nbin, nharm=nharm)
            dtc = time.perf_counter() - t1
            cpu = {"value": times.size * 16 / dtc, "unit": ORB_UNIT, "cores": 1, "kind": "port",
                   "sample": f"oracle orbital_stats (numpy, 1 thread), {times.size} events x 16 T0, {dtc:.1f} s"}
        line = {"metric": METRIC, "value": value, "unit": ORB_UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": orb_config(desc, times.size, T0s.size, nbin, nharm, world),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": ORB_UNIT, "h2d_bytes_per_step": int(times.nbytes + T0s.nbytes),
                        "d2h_bytes_per_step": int(T0s.size * 8), "api": "hendrics_b200.search.orbital_fold_stats per rank"},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ncu_record(wl):
    """Counters of the committed `ncu --set full` capture of this workload's bench command (profiles/), labelled with
    their source file: DRAM bytes per launch (roofline.traffic) and the pipe utilisations.  They are NOT measured
    in this run (a number taken under a profiler is never a bench value, and ncu cannot run inside the bench)."""
    for rnd in ("r02", "r01"):
        f = os.path.join(ROOT, "profiles", f"{rnd}_ncu_bench_{wl}_top.json")
        if not os.path.exists(f):
            continue
        try:
            d = json.load(open(f))
        except Exception:
            continue

        def num(key):
            v = str(d.get(key, "")).split()
            if not v:
                return None
            try:
                x = float(v[0])
            except ValueError:
                return None
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(v[1], 1.0) if len(v) > 1 else 1.0
            return x * mult

        rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
        return {"source": f"ncu --set full, profiles/{os.path.basename(f)} (same bench command, earlier run)",
                "kernel": d.get("Kernel Name"),
                "dram_bytes_per_launch": (rd + wr) if rd is not None and wr is not None else None,
                "fp64_pipe_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                "smem_pipe_pct": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active")}
    return None


def run_fold(args, wl):
    """Exact-fold workloads (K1): c1, c2, c3 (default), c5."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from hendrics_b200 import _cabi, epoch_folding_search, z_n_search
    from hendrics_b200 import dist as hdist
    from hendrics_b200.search import fold_stats_device

    scaling = args.scaling or default_scaling(wl)
    desc, times, all_freqs, (lo, hi), fdots, nbin, nharm = make_inputs(wl, rank, world, scaling)
    freqs = np.ascontiguousarray(all_freqs[lo:hi])
    counts = [b - a for a, b in (hdist.shard_bounds(all_freqs.size, world, r) for r in range(world))]
    has_fd = bool(np.any(fdots != 0))
    dev = torch.device("cuda", local_rank)
    # event selection of the public API (A1/A2), done once: resident inputs for the `value` leg
    ts_host = (times - times[0]).astype(np.float64)
    ts_host = ts_host[ts_host < ts_host[-1]]
    ts_d = torch.from_numpy(ts_host).to(dev)
    f_d = torch.from_numpy(freqs).to(dev)
    fd_d = torch.from_numpy(fdots).to(dev)
    units_local = float(ts_host.size) * freqs.size * fdots.size
    units_total = float(ts_host.size) * all_freqs.size * fdots.size
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cmax = max(counts)
    gathered = torch.empty((world, fdots.size, cmax), dtype=torch.float64, device=dev)

    def step_resident():
        stats = fold_stats_device(ts_d, f_d, fd_d, nbin, nharm)
        if world > 1:  # the one data-path collective: statistic slices (padded to the widest slice)
            if stats.shape[1] != cmax:
                stats = torch.nn.functional.pad(stats, (0, cmax - stats.shape[1]))
            dist.all_gather_into_tensor(gathered.view(-1), stats.contiguous().view(-1))
        return stats

    # ---- value: inputs resident in HBM -----------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _cabi.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms = []
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()  # evict the events from L2 between timed iterations (outside the timed interval)
        ev[k][0].record()
        step_resident()
        ev[k][1].record()
        kern_ms.append(_cabi.fold_last_kernel()[0])
    barrier()
    t_wall1 = time.perf_counter()
    launches = _cabi.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    value = units_total * args.steps / total_s
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    kms, grid, block, smem = _cabi.fold_last_kernel()
    kern_avg_ms = float(np.mean(kern_ms))

    # ---- e2e: the public API, host arrays in, host arrays out (every rank ends up with the full grid) -------
    # z_n_search / epoch_folding_search on one GPU; the same functions through dist.sharded_search -- what
    # hendrics_b200.efsearch.folding_search calls -- on N: H2D of the events + trial grid, kernels, NCCL all-gather of
    # the device-resident statistic slices, one D2H copy.
    pin_t = torch.from_numpy(times).pin_memory()
    pin_f = torch.from_numpy(all_freqs).pin_memory()
    t_np, f_np = pin_t.numpy(), pin_f.numpy()
    func = epoch_folding_search if nharm == 0 else z_n_search
    kw = dict(nbin=nbin) if nharm == 0 else dict(nbin=nbin, nharm=nharm)

    def step_e2e():
        return hdist.sharded_search(func, t_np, f_np, fdots=fdots, **kw)[-1]

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tq = time.perf_counter()
        stats_host = step_e2e()
        if os.environ.get("HX_BENCH_DEBUG"):
            print(f"[debug] rank {rank} e2e step wall {time.perf_counter() - tq:.4f} s, kernel {_cabi.fold_last_kernel()}", file=sys.stderr)
    barrier()
    assert stats_host.size == all_freqs.size * fdots.size
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = units_total * args.steps / float(e2e_s.item())
    h2d = times.nbytes + freqs.nbytes + fdots.nbytes  # per rank
    d2h = all_freqs.size * fdots.size * 8              # per rank: the whole statistic grid

    if rank == 0:
        # ---- roofline of the dominant kernel (K1 fold_kernel) ----------------------------------
        peak_ops = _cabi.microbench(1, reps=5, device=local_rank)  # DMUL+DADD issue rate, inst/s
        # attainable rate of the inner loop's instruction mix (no event loads, full occupancy), units/s
        mix_ceiling = _cabi.microbench(4 if has_fd else 5, reps=3, device=local_rank)
        alg_per_unit = algorithmic_ops_per_evtr(has_fd)
        alg_rate = alg_per_unit * units_local / (kern_avg_ms * 1e-3)
        shape = _cabi.fold_last_kernel_shape()  # instantiation of the timed launches
        exec_rate = shape["fp64_per_unit"] * units_local / (kern_avg_ms * 1e-3)
        # shared-memory data pipe: 1 wavefront/clk/SM; K1 needs one per increment (ATOMS) plus the broadcast event loads
        n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
        smem_peak_wf = n_sm * ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
        smem_wf = shape["smem_wavefronts_per_unit"] * units_local / 32.0 / (kern_avg_ms * 1e-3)
        ncu = ncu_record(wl)
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_gbs = json.load(open(peaks_file)).get("hbm_gbs") if os.path.exists(peaks_file) else 6650.0
        roofline = {
            # what the hardware did: FP64 instructions issued by the inner loop / measured FP64 issue rate (<= 1)
            "bound": "fp64_pipe", "achieved": exec_rate * 1e-12, "peak": peak_ops * 1e-12, "unit": "TFLOP/s",
            "frac": exec_rate / peak_ops,
            "executed_fp64_ops_per_event_trial": shape["fp64_per_unit"],
            "binding_resource": "shared-memory data pipe: one ATOMS per event*trial (26 of 32 lanes/clk/SM sustained) plus the "
                                "broadcast event loads; see smem_pipe_frac_model and ncu.smem_pipe_pct -- the FP64 pipe is "
                                "NOT the limiter of this kernel, frac says how busy it is",
            "smem_pipe_frac_model": smem_wf / smem_peak_wf,
            "smem_wavefronts_per_event_trial_model": shape["smem_wavefronts_per_unit"],
            "ncu": ncu, "traffic": (ncu or {}).get("dram_bytes_per_launch"),
            "traffic_source": (ncu or {}).get("source"),
            # SURVEY 8d's algorithmic count (6 FP64 ops per event*trial with fdot, 4 without); the kernel issues fewer
            # instructions than that (pre-scaled constants, magic-add floor, shared fdot term), hence > executed
            "algorithmic": {"fp64_ops_per_event_trial": alg_per_unit, "achieved": alg_rate * 1e-12,
                            "frac": alg_rate / peak_ops},
            "kernel": "hx::fold_kernel", "kernel_ms": kern_avg_ms, "kernel_share_of_step": kern_avg_ms * args.steps / sum(step_ms),
            "kernel_variant": {k: shape[k] for k in ("R", "mode", "rowfd", "m2t")},
            "mix_ceiling_units_per_s": mix_ceiling,
            "frac_of_mix_ceiling": units_local / (kern_avg_ms * 1e-3) / mix_ceiling,
            "peak_source": "measured in this process: hx_microbench kind 1 (DMUL+DADD issue rate, non-FMA op/s); "
                           "MEASURED_PEAKS.json has no FP64 entry",
            "hbm_peak_gbs": hbm_gbs, "hbm_peak_source": "measured" if os.path.exists(peaks_file) else "fallback",
            "launch": {"grid": grid, "block": block, "dyn_smem_bytes": smem},
        }
        cpu = None
        if not args.no_cpu_baseline:
            v, cores, sample = cpu_port_throughput(times, all_freqs, fdots, nbin, nharm)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": fold_config(desc, nbin, nharm, ts_host.size, all_freqs.size, fdots.size, world, scaling),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "hendrics_b200.dist.sharded_search(z_n_search | epoch_folding_search, ...) (= the call inside "
                           "efsearch.folding_search); bytes are per rank"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS) + ["c4", "ffa", "orb"] + sorted(WEIGHTED) + sorted(ZN))
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"],
                    help="exact-fold workloads: strong = fixed trial grid split over the ranks (default for c3), "
                         "weak = fixed grid per GPU (default otherwise)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    wl = args.workload
    if wl == "ffa":
        return run_ffa_reference(args) if args.impl == "reference" else run_ffa(args)
    if wl in WEIGHTED:
        return run_weighted_reference(args) if args.impl == "reference" else run_weighted(args)
    if wl in ZN:
        return run_zn_reference(args, wl) if args.impl == "reference" else run_zn(args, wl)
    if wl == "orb":
        return run_orb_reference(args) if args.impl == "reference" else run_orb(args)
    if args.impl == "reference":
        return run_reference(args, wl)
    if wl == "c4":
        return run_c4(args)
    return run_fold(args, wl)


if __name__ == "__main__":
    main()
