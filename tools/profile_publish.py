"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/ (launch-list shares, ncu text summary,
narrow_traffic.json read by bench.py)."""
import collections, csv, io, json, re, subprocess, sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launch_csv, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    name = re.sub(r'<.*', '', re.sub(r'\(.*', '', r[ki])); v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e6 if u == 'ns' else v / 1e3 if u == 'us' else v * 1e3 if u == 's' else v
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
out = ["# ncu launch list summary (%s): python bench.py --steps 2 --warmup 3 --no-cpu (1M floes, 1 GPU); per-launch times are cold-cache and serialised: compare SHARES" % tag,
       "%-45s %6s %12s %8s %12s" % ("kernel", "n", "total ms", "share", "ms/launch")]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append('%-45s %6d %12.3f %8.4f %12.4f' % (k[:45], v[0], v[1], v[1] / tot, v[1] / v[0]))
open(os.path.join(R, 'profiles', '%s_launches_1M_summary.txt' % tag), 'w').write('\n'.join(out) + '\n')
import shutil; shutil.copy(launch_csv, os.path.join(R, 'profiles', '%s_launches_1M.csv' % tag))
print('\n'.join(out[:8]))
subprocess.run([sys.executable, os.path.join(R, 'tools', 'ncu_summary.py'), rep, '40'], stdout=open(os.path.join(R, 'profiles', '%s_top_kernels_1M.txt' % tag), 'w'))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw))); Hh, U = r[0], r[1]
V = next((row for row in r[2:] if "narrow_convex_kernel" in row[Hh.index("Kernel Name")]), r[2])      # the class C launch of a multi-kernel capture
g = lambda n: (float(V[Hh.index(n)].replace(',', '')), U[Hh.index(n)])
mul = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'Tbyte': 1e12}
rd, ru = g('dram__bytes_read.sum'); wr, wu = g('dram__bytes_write.sum'); t, tu = g('gpu__time_duration.sum')
tms = t if tu == 'ms' else t / 1e3 if tu == 'us' else t * 1e3 if tu == 's' else t / 1e6
sys.path.insert(0, R)
from subzero_b200.build import kernel_stamp
d = {"kernel": "narrow_convex_kernel<PairS>", "kernel_stamp": kernel_stamp(), "workload": "bench.py default (1M floes, 1 GPU)", "dram_bytes_per_launch": rd * mul[ru] + wr * mul[wu], "dram_read": rd * mul[ru],
     "dram_write": wr * mul[wu], "gpu_time_ms_under_ncu": tms, "dram_gb_per_s": (rd * mul[ru] + wr * mul[wu]) / tms / 1e6, "source": "profiles/%s_top_kernels_1M.txt (ncu --set full, one launch per kernel)" % tag}
json.dump(d, open(os.path.join(R, 'profiles', 'narrow_traffic.json'), 'w'), indent=1); print(json.dumps(d))
