#!/bin/bash
# Turn the files gpurun brought back (gpurun_out/) into the committed summaries under profiles/.
set -e
cd "$(dirname "$0")/.."
export R=${ROUND:-r02}
for k in integrate_final integrate_paged render render_long; do
  [ -f gpurun_out/prof_$k.ncu-rep ] && python scripts/summarize_ncu.py gpurun_out/prof_$k.ncu-rep profiles/${R}_${k}_ncu.txt > /dev/null
done
python - <<'PY'
import collections, csv, json, os
R = os.environ['R']
rows = [r for r in csv.reader(l for l in open('gpurun_out/launches_bench.csv') if not l.startswith('=='))]
h = rows[0]; ik, ig, iv = h.index('Kernel Name'), h.index('Grid Size'), h.index('Metric Value')
with open(f'profiles/{R}_launches_bench.csv', 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --strong-res 0\n')
    f.write('# (cold-cache, serialised: compare shares, not absolutes).  columns: kernel, grid, duration_ns\n')
    tot = collections.Counter()
    for r in rows[1:]:
        f.write(f'{r[ik][:100]}, {r[ig]}, {r[iv]}\n')
        tot[r[ik].split('(')[0].strip()] += int(r[iv])
s = sum(tot.values())
with open(f'profiles/{R}_launch_shares.txt', 'w') as f:
    f.write('# share of device time per kernel over the whole `bench.py --steps 2 --warmup 1` run under ncu\n')
    for k, v in tot.most_common(12):
        f.write(f'{100 * v / s:6.2f}%  {v / 1e6:10.3f} ms  {k}\n')
t = [r for r in csv.reader(l for l in open('gpurun_out/traffic_paged.csv') if not l.startswith('=='))]
h = t[0]; im, iv, iid = h.index('Metric Name'), h.index('Metric Value'), h.index('ID')
first = {r[im]: int(r[iv]) for r in t[1:] if r[iid] == '0'}
old = json.load(open('profiles/r01_traffic.json'))   # key names only
key = [k for k in old if 'MODE_PAGED' in k][0]
old[key].update(dram_bytes_read=first['dram__bytes_read.sum'], dram_bytes_write=first['dram__bytes_write.sum'],
                gpu_time_ns_under_ncu=first['gpu__time_duration.sum'])
for line in open(f'profiles/{R}_integrate_final_ncu.txt'):
    if line.startswith('dram__bytes_read.sum'):
        rd = float(line.split()[1]) * (1e6 if 'Mbyte' in line else 1e9)
    if line.startswith('dram__bytes_write.sum'):
        wr = float(line.split()[1]) * (1e6 if 'Mbyte' in line else 1e9)
k2 = [k for k in old if 'MODE_FINAL' in k][0]
old[k2].update(dram_bytes_read=int(rd), dram_bytes_write=int(wr), source=f'profiles/{R}_integrate_final_ncu.txt')
json.dump(old, open(f'profiles/{R}_traffic.json', 'w'), indent=1)
for f in (f'bench_{R}_n1.json', f'bench_{R}_ref.json', 'parity_report.txt'):
    if os.path.exists('gpurun_out/' + f):
        open('profiles/' + (R + '_' + f if not f.startswith('bench') else f), 'w').write(open('gpurun_out/' + f).read())
PY
echo refreshed
