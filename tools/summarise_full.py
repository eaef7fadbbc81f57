"""Summarise an ``ncu --set full`` report (one launch per kernel of interest):
time, DRAM traffic, occupancy, issue rate, active lanes, FP64-pipe utilisation
and the EXECUTED FP64 operations (dfma / dmul / dadd thread instructions) as a
rate against the measured FP64 peak.

    ncu -i rep.ncu-rep --page raw --csv > raw.csv
    python tools/summarise_full.py raw.csv [fp64_peak_tflops] [json_out]
"""
import csv
import json
import re
import sys


def val(row, h, name, default=0.0):
    if name not in h:
        return default
    v = row[h.index(name)].replace(',', '')
    try:
        return float(v)
    except ValueError:
        return default


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h, units = rows[0], rows[1]
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else 34.2
    out = {}
    for r in rows[2:]:
        name = re.sub(r'\(.*', '', r[h.index('Kernel Name')]).replace('void ', '') \
            .replace('h3d::', '')
        unit = units[h.index('gpu__time_duration.sum')]
        ms = val(r, h, 'gpu__time_duration.sum') * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0,
                                                    'nsecond': 1e-6, 'usecond': 1e-3,
                                                    'msecond': 1.0, 'second': 1e3}.get(unit, 1e-6)
        def scaled(metric):
            u = units[h.index(metric)] if metric in h else ''
            return val(r, h, metric) * {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9,
                                        'Tbyte': 1e12, 'byte': 1.0}.get(u, 1.0)
        dram = scaled('dram__bytes_read.sum') + scaled('dram__bytes_write.sum')
        # thread instructions per elapsed cycle (summed over the sub-partitions)
        # x elapsed cycles
        cyc = val(r, h, 'smsp__cycles_elapsed.avg') or val(r, h, 'sm__cycles_elapsed.avg')
        dfma = val(r, h, 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed') * cyc
        dmul = val(r, h, 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed') * cyc
        dadd = val(r, h, 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed') * cyc
        flop = 2 * dfma + dmul + dadd
        stalls = sorted(((val(r, h, c), c.split('issue_stalled_')[1].split('_per_')[0])
                         for c in h if c.startswith('smsp__average_warps_issue_stalled_')
                         and c.endswith('_per_issue_active.ratio')), reverse=True)[:5]
        rec = dict(
            grid=r[h.index('Grid Size')], block=r[h.index('Block Size')],
            regs=int(val(r, h, 'launch__registers_per_thread')),
            ms=round(ms, 4), dram_gb=round(dram / 1e9, 4),
            dram_gbs=round(dram / (ms * 1e-3) / 1e9, 1) if ms else 0,
            warps_active_pct=round(val(r, h, 'sm__warps_active.avg.pct_of_peak_sustained_active'), 1),
            issue_active_pct=round(val(r, h, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'), 1),
            lanes_per_inst=round(val(r, h, 'smsp__thread_inst_executed_per_inst_executed.ratio'), 2),
            fp64_pipe_active_pct=round(val(r, h, 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'), 1),
            fp64_thread_inst=dict(dfma=int(dfma), dmul=int(dmul), dadd=int(dadd)),
            executed_fp64_tflops=round(flop / (ms * 1e-3) / 1e12, 2) if ms else 0,
            executed_fp64_frac_of_peak=round(flop / (ms * 1e-3) / 1e12 / peak, 3) if ms else 0,
            stalls_per_issue={k: round(v, 2) for v, k in stalls})
        out[name] = rec
        print('--- %s  (grid %s x %s, %d regs)' % (name, rec['grid'], rec['block'], rec['regs']))
        print('  %.4f ms; DRAM %.3f GB (%.0f GB/s); warps active %.1f %%; issue active %.1f %%; '
              'lanes/inst %.2f' % (rec['ms'], rec['dram_gb'], rec['dram_gbs'],
                                   rec['warps_active_pct'], rec['issue_active_pct'],
                                   rec['lanes_per_inst']))
        print('  FP64 pipe active %.1f %%; executed FP64: dfma %.3g dmul %.3g dadd %.3g thread-inst '
              '-> %.2f TFLOP/s = %.1f %% of the measured %.1f TFLOP/s'
              % (rec['fp64_pipe_active_pct'], dfma, dmul, dadd, rec['executed_fp64_tflops'],
                 100 * rec['executed_fp64_frac_of_peak'], peak))
        print('  stalls per issue: ' + ', '.join('%s %.2f' % (k, v) for k, v in rec['stalls_per_issue'].items()))
    if len(sys.argv) > 3:
        json.dump(out, open(sys.argv[3], 'w'), indent=1)


if __name__ == '__main__':
    main()
