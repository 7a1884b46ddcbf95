"""Prints the kernels of the last full train step from an ncu --csv launch list (gpu__time_duration.sum)."""
import csv, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines); hdr = next(r); idx = {h: i for i, h in enumerate(hdr)}
data = list(r)
# one bce_kernel per train step: the window between the last two holds exactly one step's kernels (backward of step k, forward of k+1)
ups = [i for i, d in enumerate(data) if 'bce_kernel' in d[idx['Kernel Name']]]
a, b = ups[-2], ups[-1]
tot = 0
agg = {}
for d in data[a:b]:
    n = d[idx['Kernel Name']]
    n = re.sub(r'regat::<unnamed>::', '', n); n = re.sub(r'\(.*', '', n); n = n.replace('void ', '')
    t = int(d[idx['Metric Value']]); tot += t
    if len(sys.argv) < 3:
        print(f"{d[idx['ID']]:>4} {n[:60]:60s} {t/1000:8.1f} {d[idx['Grid Size']]:16s} {d[idx['Block Size']]} s{d[idx['Stream']]}")
    k = re.sub(r'<.*', '', n)
    agg[k] = agg.get(k, [0, 0]); agg[k][0] += t; agg[k][1] += 1
print('total us', tot / 1000, 'launches', b - a)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:32s} {v[0]/1000:8.1f} us  x{v[1]}  {100*v[0]/tot:5.1f}%")
