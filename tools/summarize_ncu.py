"""Turn the ncu outputs a gpurun call brought back (gpurun_out/) into the small text summaries kept under profiles/.
   python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/launches_r1.md
   python tools/summarize_ncu.py full gpurun_out/gemm3x.ncu-rep profiles/gemm3x_r1.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum']


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = [(x['Kernel Name'], float(x['Metric Value'].replace(',', '')), x.get('Metric Unit', 'ns')) for x in csv.DictReader(lines)]
    idx = [i for i, x in enumerate(rows) if 'pool_kernel' in x[0]]
    if len(idx) >= 3:
        s, e, n = idx[-3], idx[-1], 2
    else:
        s, e, n = 0, len(rows), 1
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, val, u in rows[s:e]:
        short = re.sub(r'\(.*', '', name).replace('void ', '').replace('nar::', '')[:80]
        agg[short][0] += 1
        agg[short][1] += val
    tot = sum(v[1] for v in agg.values())
    with open(dst, 'w') as f:
        f.write('# kernel launch list of `python bench.py --steps 3 --warmup 3` under\n')
        f.write('# `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES)\n')
        f.write('# last %d complete steps; total %.1f us per step\n\n' % (n, tot / n / 1000))
        f.write('| kernel | launches/step | us/step | share |\n|---|---:|---:|---:|\n')
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %.1f | %.1f | %.1f %% |\n' % (k, v[0] / n, v[1] / n / 1000, 100 * v[1] / tot))
    print('wrote', dst)


UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def full(src, dst, traffic_key=None):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic = None
    with open(dst, 'w') as f:
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            f.write('kernel: %s\n' % d.get('Kernel Name', '?'))
            tot = 0.0
            for h, u, v in zip(hdr, units, vals):
                if h in KEEP:
                    f.write('  %-86s %-14s %s\n' % (h, u, v))
                if h in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                    tot += float(v.replace(',', '')) * UNIT.get(u, 1.0)
            f.write('\n')
            traffic = tot
    print('wrote', dst)
    if traffic_key:
        import json
        import os
        p = os.path.join(os.path.dirname(os.path.abspath(dst)), 'ncu_traffic.json')
        d = json.load(open(p)) if os.path.exists(p) else {}
        d[traffic_key] = traffic
        json.dump(d, open(p, 'w'), indent=1, sort_keys=True)
        print('traffic[%s] = %.0f bytes -> %s' % (traffic_key, traffic, p))


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](*sys.argv[2:])
