"""profiles/r2_traffic.json from an ncu launch list with DRAM byte counters (tools/gpu_ncu_traffic.sh):
per recurrence kernel of the default bench command, dram__bytes_read.sum + dram__bytes_write.sum per launch.

    python tools/ncu_traffic.py gpurun_out/traffic.csv dblstm_ctc > profiles/r2_traffic.json
"""
import csv
import json
import re
import sys


def label(name):
    if 'fwd_cluster_tc' in name:
        return 'blstm_rec_fwd_cluster_tc'
    m = re.search(r'bwd_chain_kernel<\D*\d+,\s*\D*\d+,\s*\D*(\d+)>', name)
    if m:
        return 'blstm_rec_bwd_chain%s' % m.group(1)
    if 'bwd_cluster8' in name:
        return 'blstm_rec_bwd_cluster8'
    if 'dec_attn_step_kernel' in name:
        return 'dec_attn_step'
    return None


def main(path, workload):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    cols = {n: i for i, n in enumerate(rows[hdr])}
    per = {}
    for r in rows[hdr + 1:]:
        if len(r) <= cols['Metric Value']:
            continue
        lab = label(r[cols['Kernel Name']])
        if lab is None:
            continue
        key = (lab, r[cols['ID']])
        d = per.setdefault(key, {})
        val = float(r[cols['Metric Value']].replace(',', ''))
        unit = r[cols['Metric Unit']]
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3,
                 'msecond': 1e6, 'second': 1e9}.get(unit, 1)
        d[r[cols['Metric Name']]] = val * scale
    out = {}
    for (lab, _), d in per.items():
        e = out.setdefault(lab, {'launches': 0, 'read': 0.0, 'write': 0.0, 'ns': 0.0})
        e['launches'] += 1
        e['read'] += d.get('dram__bytes_read.sum', 0.0)
        e['write'] += d.get('dram__bytes_write.sum', 0.0)
        e['ns'] += d.get('gpu__time_duration.sum', 0.0)
    table = {}
    for lab, e in out.items():
        n = e['launches']
        table[lab] = {'dram_bytes_per_launch': (e['read'] + e['write']) / n, 'dram_read_per_launch': e['read'] / n,
                      'dram_write_per_launch': e['write'] / n, 'launches_captured': n, 'ncu_ms_per_launch': e['ns'] / n / 1e6,
                      'source': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on '
                                '`python bench.py --steps 1 --warmup 3` (the shipped configuration, full T), tools/gpu_ncu_traffic.sh'}
    print(json.dumps({workload: table}, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else 'dblstm_ctc')
