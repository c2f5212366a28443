"""Turn the ncu outputs under gpurun_out/ into the committed summaries under profiles/."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)

def launch_summary(w, tag):
    rows = list(csv.reader(open(f'gpurun_out/launches_bench_{w}.csv')))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', '')) * {'us': 1e-3, 'ns': 1e-6, 's': 1e3, 'ms': 1}.get(r[ui], 1)
        a = agg.setdefault(r[ki], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(f'profiles/{tag}_launches_bench_{w}_summary.txt', 'w') as f:
        f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none -c 600  python bench.py --workload {w} --steps 1 --warmup 3 --no-cpu-baseline\n")
        f.write("(per-launch times are cold-cache/serialised under ncu: compare shares, not absolutes)\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
            f.write(f"{t:12.3f} ms {100*t/tot:6.2f}%  x{n:3d}  {k[:110]}\n")
        f.write(f"total {tot:.3f} ms over {len(data)} launches\n")
    top = sorted(agg.items(), key=lambda kv: -kv[1][1])[:3]
    print(w, [(k[:44], n, round(t, 2), round(100 * t / tot, 1)) for k, (n, t) in top])

def num(s):
    p = s.split(); v = float(p[0]); u = p[1] if len(p) > 1 else ''
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(u, 1)

tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
which = sys.argv[2].split(',') if len(sys.argv) > 2 else ['c2', 'c3', 'c4']
for w, name in [x for x in (('c2', 'fold_c2'), ('c3', 'fold_c3'), ('c4', 'qffa_c4'), ('w1', 'wfold_w1'), ('w2', 'wfold_w2')) if x[0] in which]:
    launch_summary(w, tag)
    subprocess.run([sys.executable, 'tools/ncu_summary.py', f'gpurun_out/bench_{w}_top.ncu-rep', f'profiles/{tag}_ncu_bench_{w}_top.json'], stdout=subprocess.DEVNULL)
    d = json.load(open(f'profiles/{tag}_ncu_bench_{w}_top.json'))
    out = {"dram_bytes_per_launch": num(d['dram__bytes_read.sum']) + num(d['dram__bytes_write.sum']),
           "source": f"profiles/{tag}_ncu_bench_{w}_top.json (ncu --set full on `bench.py --workload {w}`)",
           "kernel": d.get("Kernel Name"), "gpu_time": d.get("gpu__time_duration.sum"),
           "fp64_pipe_active_pct": d.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
           "smem_wavefront_pct": d.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
           "issue_active_pct": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "dram_pct": d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
           "stall_per_issue": d.get("stall_per_issue")}
    json.dump(out, open(f'profiles/{name}_ncu_summary.json', 'w'), indent=1)
    print(w, {k: out[k] for k in ('kernel', 'gpu_time', 'dram_bytes_per_launch', 'fp64_pipe_active_pct', 'smem_wavefront_pct', 'issue_active_pct', 'dram_pct')})
