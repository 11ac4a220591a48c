"""Summarise an `ncu --page raw --csv` export (tools/round2_profile.sh): per launch the device time,
DRAM bytes and achieved GB/s, DRAM-throughput %, tensor-pipe and FMA-pipe activity; optionally writes
the DRAM bytes per launch of the classifier tap GEMM as profiles/rNN_tap_gemm_traffic.json (the
`roofline.traffic` figure of bench.py).

    python tools/summarize_ncu_raw.py profiles/r02_ncu_gemms_raw.csv [--traffic profiles/r02_tap_gemm_traffic.json]
"""
import csv
import json
import re
import sys


def num(s):
    try:
        return float(s.replace(',', ''))
    except ValueError:
        return float('nan')


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Gbyte': 1e9}
    print('# %s' % path)
    print('%-46s %9s %10s %10s %8s %7s %7s %7s %5s' % ('kernel', 'us', 'DRAM MB', 'GB/s', 'dram%', 'tensor%', 'fma%', 'sm%', 'regs'))
    tap = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r'\(.*$', '', r[col['Kernel Name']]).replace('void ', '').replace('dmc::', '')
        us = num(r[col['gpu__time_duration.sum']]) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(
            units[col['gpu__time_duration.sum']], 1.0)
        rd = num(r[col['dram__bytes_read.sum']]) * scale.get(units[col['dram__bytes_read.sum']], 1.0)
        wr = num(r[col['dram__bytes_write.sum']]) * scale.get(units[col['dram__bytes_write.sum']], 1.0)
        print('%-46s %9.1f %10.1f %10.0f %8.1f %7.1f %7.1f %7.1f %5.0f' % (
            name[:46], us, (rd + wr) / 1e6, (rd + wr) / us / 1e3,
            num(r[col['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']]),
            num(r[col['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']]),
            num(r[col['sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']]),
            num(r[col['sm__throughput.avg.pct_of_peak_sustained_elapsed']]),
            num(r[col['launch__registers_per_thread']])))
        if name.startswith('tap_gemm_ws_kernel'):
            tap.append({'kernel': name, 'us': us, 'dram_bytes': rd + wr})
    if '--traffic' in sys.argv and tap:
        out = sys.argv[sys.argv.index('--traffic') + 1]
        with open(out, 'w') as f:
            json.dump({'source': path, 'launches': tap,
                       'dram_bytes_per_launch_avg': sum(t['dram_bytes'] for t in tap) / len(tap),
                       'note': 'dram__bytes_read.sum + dram__bytes_write.sum per launch of tap_gemm_ws_kernel, '
                               'ncu --set full, config 2 at B=64 (first launches of the classifier forward)'}, f, indent=1)


if __name__ == '__main__':
    main()
