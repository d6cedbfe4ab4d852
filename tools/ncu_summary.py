"""Turn ncu output into the markdown summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_T96.csv            > profiles/<round>_launches.md
    python tools/ncu_summary.py full gpurun_out/full_rec.ncu-rep [more.ncu-rep] > profiles/<round>_ncu_full.md

`launches` aggregates the gpu__time_duration.sum launch list per kernel (count, total, share).  `full` reads the
raw page of an `ncu --set full` report (needs ncu on PATH) and prints the metrics B200_PROFILING.md names."""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r'nabu::<unnamed>::|void |unnamed>::', '', name)
    name = re.sub(r'\(.*', '', name)
    return name[:70]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        try:
            ns = float(r[14])
        except ValueError:
            continue
        k = short(r[4])
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + ns)
    tot = sum(t for _, t in agg.values())
    print('| kernel | launches | total us | share |')
    print('|---|---|---|---|')
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| %s | %d | %.1f | %.3f |' % (k, c, t / 1e3, t / tot))
    print('\n%d launches, %.1f us in total (cold-cache, serialised: compare shares, not absolutes).' % (len(rows), tot / 1e3))


METRICS = [
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %'),
    ('sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active', 'tensor pipe inst %'),
    ('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor pipe active %'),
    ('sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active', 'TMEM pipe %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy %'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('launch__cluster_dim_x', 'cluster'),
    ('smsp__inst_executed.sum', 'warp insts'),
]


def full(paths):
    print('| kernel | ' + ' | '.join(n for _, n in METRICS) + ' |')
    print('|---|' + '---|' * len(METRICS))
    for path in paths:
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        kn = hdr.index('Kernel Name')
        for r in rows[2:]:
            cells = []
            for key, _ in METRICS:
                idx = [i for i, h in enumerate(hdr) if h == key or h.endswith('.' + key)]
                if not idx:
                    cells.append('-')
                    continue
                v = r[idx[0]]
                try:
                    v = '%.4g' % float(v)
                except ValueError:
                    pass
                cells.append('%s %s' % (v, units[idx[0]]) if units[idx[0]] not in ('', '%') else v)
            print('| %s | ' % short(r[kn]) + ' | '.join(cells) + ' |')


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
