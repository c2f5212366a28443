"""Under torchrun (NCCL): sharded exact-fold and --fast searches equal the single-GPU result."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist

lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from hendrics_b200 import dist as hd, z_n_search, epoch_folding_search
from hendrics_b200.efsearch import search_with_qffa, folding_search
from hendrics_b200.fake import simulate_events
from hendrics_b200.io import EventList

ts = simulate_events(50000, 1000.0, seed=3) + 8e7
ev = EventList(time=ts, gti=np.array([[8e7, 8e7 + 1000.0]]))
full = folding_search(ev, 9.85, 9.95, func=z_n_search, oversample=2, nbin=64, nharm=2, fdotmin=-1e-6, fdotmax=1e-6)
assert hd.world_size() == int(os.environ["WORLD_SIZE"]) > 1
# single-GPU reference on every rank (bypasses the sharding)
t = (ts - 8e7).astype(np.float64)
f = np.arange(9.85, 9.95 + 1e-8 * full[3], full[3])
fd = np.arange(-1e-6, 1e-6 + 1e-2 * full[4], full[4])
ref = z_n_search(t, f, nharm=2, nbin=64, fdots=fd)
assert np.array_equal(full[0], ref[0]) and np.array_equal(full[1], ref[1])
assert np.array_equal(full[2], ref[2]), np.abs(full[2] - ref[2]).max()
# weighted (light-curve) and exposure-corrected branches shard the same way
from hendrics_b200.io import Lightcurve
edges = np.arange(8e7, 8e7 + 1000.0 + 1e-3, 5e-3)
lc = Lightcurve(time=(edges[:-1] + edges[1:]) / 2, counts=np.histogram(ts, bins=edges)[0], gti=ev.gti, dt=5e-3)
wfull = folding_search(lc, 9.85, 9.95, func=epoch_folding_search, oversample=2, nbin=32)
wref = epoch_folding_search((lc.time - 8e7).astype(np.float64), wfull[0], nbin=32, weights=lc.counts)
assert np.array_equal(wfull[1], wref[1]), np.abs(wfull[1] - wref[1]).max()
xfull = folding_search(ev, 9.85, 9.95, func=z_n_search, oversample=2, nbin=64, nharm=2, expocorr=True)
xref = z_n_search(t, xfull[0], nharm=2, nbin=64, expocorr=True, gti=(ev.gti - 8e7).astype(np.float64))
assert np.array_equal(xfull[1], xref[1])
# the replicated event list goes up in shards + one NVLink all-gather: same bits on every rank, ragged sizes included
for n in (1, 7, 1000, 50001, 1 << 20):
    a = np.random.default_rng(n).uniform(0, 1e5, n)
    got = hd.upload_replicated(a, min_bytes=0)
    assert got.shape == (n,) and np.array_equal(got.cpu().numpy(), a), n
bad = np.arange(1000.0) + (1.0 if dist.get_rank() == 1 else 0.0)
try:
    hd.upload_replicated(bad, min_bytes=0)
    raise SystemExit("upload_replicated accepted different arrays")
except ValueError:
    pass
hd.REPLICATED_UPLOAD_MIN_BYTES = 0  # force the sharded upload inside the searches below
again = folding_search(ev, 9.85, 9.95, func=z_n_search, oversample=2, nbin=64, nharm=2, fdotmin=-1e-6, fdotmax=1e-6)
assert np.array_equal(again[2], full[2])
np.random.seed(0)
q = search_with_qffa(ts, 9.85, 9.95, nbin=16, n=2, silent=True)
np.random.seed(0)
hd.REPLICATED_UPLOAD_MIN_BYTES = 32 << 20
q2 = search_with_qffa(ts, 9.85, 9.95, nbin=16, n=2, silent=True)
assert np.array_equal(q[2], q2[2])
os.environ["HX_NO_SHARD"] = "1"
print(f"rank {dist.get_rank()}: sharded search ok, stats {full[2].shape}, qffa {q[2].shape}, best f {full[0].flat[np.argmax(full[2])]:.4f}", flush=True)
dist.barrier()
dist.destroy_process_group()
