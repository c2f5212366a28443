"""Wall time of the whole HENzsearch / HENefsearch CLI mirror on a realistic observation (1e6 events, 1e5 s)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hendrics_b200.efsearch import main_zsearch, main_efsearch
from hendrics_b200.fake import simulate_events
from hendrics_b200.io import EventList, save_events

d = tempfile.mkdtemp()
os.chdir(d)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
ts = simulate_events(n, 1e5, seed=1) + 8e7
save_events(EventList(time=ts, gti=np.array([[8e7, 8e7 + 1e5]]), mjdref=56000), "obs_ev.p")
for name, fn, argv in (
    ("HENzsearch (exact, + dynamical pass)", main_zsearch, ["obs_ev.p", "-f", "9.85", "-F", "9.95", "-N", "2"]),
    ("HENzsearch again (warm)", main_zsearch, ["obs_ev.p", "-f", "9.85", "-F", "9.95", "-N", "2"]),
    ("HENefsearch --fdotmin/max", main_efsearch, ["obs_ev.p", "-f", "9.85", "-F", "9.95", "-n", "64", "--fdotmin", " -1e-9", "--fdotmax", "1e-9"]),
    ("HENzsearch --fast", main_zsearch, ["obs_ev.p", "-f", "9.85", "-F", "9.95", "-N", "2", "--fast"]),
    ("HENzsearch --expocorr", main_zsearch, ["obs_ev.p", "-f", "9.85", "-F", "9.95", "-N", "2", "--expocorr"]),
):
    t0 = time.perf_counter()
    out = fn(argv)
    print(f"{name:42s} {time.perf_counter() - t0:8.3f} s  -> {os.path.basename(out[0]) if out else out}", flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
main_zsearch(["obs_ev.p", "-f", "9.85", "-F", "9.95", "-N", "2"])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
