"""Summarise an ncu launch list (--csv --metrics gpu__time_duration.sum[,dram__bytes_*]):
per kernel name: launches, total time, share, DRAM bytes and achieved DRAM GB/s.

    python tools/summarise_launches.py gpurun_out/launches.csv [peak_gbs]
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6544.7
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ik, im, iu, iv = hdr.index('Kernel Name'), hdr.index('Metric Name'), \
        hdr.index('Metric Unit'), hdr.index('Metric Value')
    iid = hdr.index('ID')
    per = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r'\(.*', '', r[ik]).replace('void ', '').replace('h3d::', '')
        d = per.setdefault((r[iid], name), {})
        val = float(r[iv].replace(',', ''))
        unit = r[iu]
        scale = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'msecond': 1.0, 'ms': 1.0,
                 'nsecond': 1e-6, 'second': 1e3, 'byte': 1.0, 'Kbyte': 1e3,
                 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1.0)
        d[r[im]] = val * scale
    agg = collections.OrderedDict()
    for (_, name), d in per.items():
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get('gpu__time_duration.sum', 0.0)
        a[2] += d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
    tot = sum(a[1] for a in agg.values())
    print('total %.2f ms over %d launches' % (tot, sum(a[0] for a in agg.values())))
    print('%-44s %6s %10s %6s %10s %9s %6s' % ('kernel', 'n', 'ms', 'share', 'DRAM MB', 'GB/s', '%peak'))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = a[2] / (a[1] * 1e-3) / 1e9 if a[1] else 0
        print('%-44s %6d %10.3f %5.1f%% %10.1f %9.0f %5.1f%%' % (
            name[:44], a[0], a[1], 100 * a[1] / tot, a[2] / 1e6, gbs, 100 * gbs / peak))


if __name__ == '__main__':
    main()
