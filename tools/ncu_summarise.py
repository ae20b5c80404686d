"""Turns ncu CSV output into the tables committed under profiles/.
  python tools/ncu_summarise.py launches <launch list csv (--metrics gpu__time_duration.sum)> <out prefix> "<title>"
  python tools/ncu_summarise.py full <--page raw --csv of a --set full capture> <out prefix> "<title>"
"""
import collections
import csv
import json
import re
import sys


def read(path):
    rows = [r for r in csv.reader(open(path, errors='replace')) if r and not r[0].startswith('==')]
    hdr = rows[0]
    units = rows[1] if rows[1] and not rows[1][0].isdigit() else None
    body = rows[2:] if units else rows[1:]
    return hdr, units, body


def short(name):
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\(.*$', '', name)
    return name.replace('mnn::', '', 1) if name.startswith('mnn::tc') else name


def launches(path, out, title):
    hdr, units, body = read(path)
    ik = hdr.index('Kernel Name')
    if 'Metric Value' in hdr:                  # long format: one row per (launch, metric)
        iv, iu = hdr.index('Metric Value'), hdr.index('Metric Unit')
        per = [(short(r[ik]), float(r[iv].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[r[iu]]) for r in body]
    else:
        iv = hdr.index('gpu__time_duration.sum')
        scale = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'nsecond': 1e-6, 'msecond': 1.0, 'ms': 1.0, 'second': 1e3}[units[iv]]
        per = [(short(r[ik]), float(r[iv].replace(',', '')) * scale) for r in body]
    agg = collections.OrderedDict()
    for k, ms in per:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    with open(out + '_summary.csv', 'w') as f:
        f.write(f'# {title}\n# source: ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES\n')
        f.write('kernel,launches,total_ms,share_pct\n')
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{k[:70]}",{n},{ms:.3f},{100 * ms / tot:.2f}\n')
        f.write(f'"TOTAL",{sum(a[0] for a in agg.values())},{tot:.3f},100.0\n')


FULL = [('time_ms', 'gpu__time_duration.sum', 1e-6), ('dram_read_GB', 'dram__bytes_read.sum', None),
        ('dram_write_GB', 'dram__bytes_write.sum', None), ('dram_pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 1),
        ('tensor_pipe_pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 1),
        ('xu_pipe_pct', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 1),
        ('fma_pipe_pct', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 1),
        ('issue_active_pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 1),
        ('sm_throughput_pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 1),
        ('occupancy_pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 1),
        ('regs', 'launch__registers_per_thread', 1), ('warp_insts', 'smsp__inst_executed.sum', 1)]
BYTES = {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0, 'Tbyte': 1e3}
TIME = {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3, 'second': 1e3}


def full(path, out, title):
    hdr, units, body = read(path)
    ik, ig, ib = hdr.index('Kernel Name'), hdr.index('Grid Size'), hdr.index('Block Size')
    cols = []
    for name, metric, _ in FULL:
        cols.append(hdr.index(metric) if metric in hdr else None)
    traffic = collections.OrderedDict()
    with open(out + '.csv', 'w') as f:
        f.write(f'# {title}\n# ncu --set full --clock-control none; one row per launch\n')
        f.write('kernel,grid,block,' + ','.join(n for n, _, _ in FULL) + '\n')
        for r in body:
            vals = []
            for (name, metric, _), c in zip(FULL, cols):
                if c is None or r[c] in ('', 'n/a'):
                    vals.append('')
                    continue
                v = float(r[c].replace(',', ''))
                u = units[c] if units else ''
                if name == 'time_ms':
                    v *= TIME.get(u, 1e-6)
                elif name.startswith('dram_') and name.endswith('GB'):
                    v *= BYTES.get(u, 1e-9)
                vals.append(f'{v:.6g}')
            k = short(r[ik])
            f.write(f'"{k[:70]}","{r[ig]}","{r[ib]}",' + ','.join(vals) + '\n')
            if vals[0] and vals[1] and vals[2]:
                t = traffic.setdefault(k.split('<')[0], dict(launches=0, dram_bytes=0.0, ms=0.0))
                t['launches'] += 1
                t['dram_bytes'] += (float(vals[1]) + float(vals[2])) * 1e9
                t['ms'] += float(vals[0])
    detail = {k: dict(launches=t['launches'], dram_bytes_per_launch=t['dram_bytes'] / t['launches'],
                      dram_bytes_per_step=t['dram_bytes'], ms_per_step=t['ms']) for k, t in traffic.items()}
    gem = [t for k, t in traffic.items() if 'gemm_tc' in k]
    js = {'source': f'{out}.csv ({title}; dram__bytes_read.sum + dram__bytes_write.sum)',
          'gemm_tc': (sum(t['dram_bytes'] for t in gem) / max(sum(t['launches'] for t in gem), 1)) if gem else None,
          'detail': detail}
    json.dump(js, open(out + '_traffic.json', 'w'), indent=1)


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else '')
