"""Builds profiles/<tag>_hw.json, the hardware-measured numbers bench.py quotes
next to its model figures, from the two ncu passes of tools/run_prof_r02.sh:

    python tools/make_profile_json.py <launches.csv> <full_raw.csv> <out.json> [fp64_peak]

  launches.csv   ncu --metrics gpu__time_duration.sum,dram__bytes_* of ONE timed
                 step (every launch)
  full_raw.csv   ncu -i <rep> --page raw --csv of the --set full capture (first
                 launch of the kernels of interest)
"""
import collections
import csv
import json
import re
import subprocess
import sys


def launch_totals(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ik, im, iu, iv, iid = [h.index(c) for c in ('Kernel Name', 'Metric Name', 'Metric Unit',
                                                'Metric Value', 'ID')]
    per = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r'\(.*', '', r[ik]).replace('void ', '').replace('h3d::', '')
        name = re.sub(r'<.*', '', name)
        scale = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'msecond': 1.0, 'ms': 1.0, 'nsecond': 1e-6,
                 'second': 1e3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(r[iu], 1.0)
        per.setdefault((r[iid], name), {})[r[im]] = float(r[iv].replace(',', '')) * scale
    agg = collections.OrderedDict()
    for (_, name), d in per.items():
        a = agg.setdefault(name, dict(launches=0, ms=0.0, dram_bytes=0.0, nonempty=0))
        a['launches'] += 1
        a['ms'] += d.get('gpu__time_duration.sum', 0.0)
        b = d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
        a['dram_bytes'] += b
        a['nonempty'] += d.get('gpu__time_duration.sum', 0.0) > 0.02       # > 20 us
    return agg


def main():
    launches, full, out = sys.argv[1:4]
    peak = sys.argv[4] if len(sys.argv) > 4 else '34.2'
    agg = launch_totals(launches)
    tmp = out + '.kernels.tmp'
    subprocess.run([sys.executable, __file__.replace('make_profile_json.py', 'summarise_full.py'),
                    full, peak, tmp], check=True, stdout=subprocess.DEVNULL)
    kernels = {re.sub(r'<.*', '', k): v for k, v in json.load(open(tmp)).items()}
    subprocess.run(['rm', '-f', tmp])
    bh = [k for k in agg if k.startswith(('sort_', 'bh_'))]
    prep = [k for k in agg if k.startswith(('union_', 'bias_filter', 'ratio_scatter', 'median_select',
                                            'sf_table', 'scale_filter', 'mask_', 'group_'))]
    doc = dict(
        source=dict(launch_list=launches, full_capture=full,
                    note='ncu times are cold-cache and serialised: shares, not absolutes'),
        step_total_ms=round(sum(a['ms'] for a in agg.values()), 3),
        per_kernel_step={k: dict(launches=a['launches'], ms=round(a['ms'], 3),
                                 dram_bytes=int(a['dram_bytes']))
                         for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['ms'])[:24]},
        equalize_kernel=dict(
            dram_bytes_per_step=int(agg['equalize_kernel']['dram_bytes']),
            launches_with_work=int(agg['equalize_kernel']['nonempty']),
            first_launch=kernels.get('equalize_kernel')),
        nll_kernel=dict(first_launch=kernels.get('nll_kernel')),
        lrt_fused_kernel=dict(first_launch=kernels.get('lrt_fused_kernel')),
        bh=dict(dram_bytes_per_step=int(sum(agg[k]['dram_bytes'] for k in bh)),
                ms=round(sum(agg[k]['ms'] for k in bh), 3), kernels=bh,
                sort_pass_first_launch=kernels.get('sort_pass_kernel')),
        prepare_data=dict(dram_bytes_per_step=int(sum(agg[k]['dram_bytes'] for k in prep)),
                          ms_sum_of_kernels=round(sum(agg[k]['ms'] for k in prep), 3)),
        pool=dict(dram_bytes_per_step=int(sum(agg[k]['dram_bytes'] for k in agg
                                              if k.startswith('pool_'))),
                  pool_pull_first_launch=kernels.get('pool_pull_kernel')))
    json.dump(doc, open(out, 'w'), indent=1)
    print(json.dumps({k: doc[k] for k in ('step_total_ms',)}))


if __name__ == '__main__':
    main()
