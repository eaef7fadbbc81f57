"""Wall-clock of the drop-in class on files (file patterns in -> .npy on disk),
per step: SURVEY.md section 8(d) "end-to-end" clock for BASELINE configs[0]
(chr18 + chr19, mouse-sized, 10 kb, 2-vs-2).

    python tools/time_class.py [workdir] [--chroms=chr18,chr19 | --chroms=all] [--reuse]

``--chroms=all``: BASELINE configs[1], the whole mouse genome (20 chromosomes,
48.4 M union pixels, ~2.3 GB of .npz in, ~8.2 GB of .npy out): only the
one-call ``run_to_qvalues`` is timed (twice: cold, then with a warm page cache).
"""
import json
import os
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    from hic3defdr_b200 import HiC3DeFDR
    from hic3defdr_b200.synth import MM10_10KB, write_dataset
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    root = args[0] if args else tempfile.mkdtemp(prefix='h3d_class_')
    chroms = ['chr18', 'chr19']
    genome = False
    for a in sys.argv[1:]:
        if a.startswith('--chroms'):
            chroms = a.split('=', 1)[1].split(',')
            if chroms == ['all']:
                chroms, genome = list(MM10_10KB), True
    t0 = time.perf_counter()
    kw = write_dataset(os.path.join(root, 'in'),
                       {c: MM10_10KB[c] for c in chroms}, n_reps=4,
                       dist_max=200, config=1, n_jobs=os.cpu_count() or 1,
                       generate='--reuse' not in sys.argv or
                       not os.path.isdir(os.path.join(root, 'in')))
    t_gen = time.perf_counter() - t0
    out = {}
    for trial in range(0 if genome else 2):   # second pass: page cache and CUDA context warm
        outdir = os.path.join(root, 'out%d' % trial)
        h = HiC3DeFDR(outdir=outdir, dist_thresh_max=200, **kw)
        times = {}
        t_all = time.perf_counter()
        for name, fn in (('prepare_data', h.prepare_data),
                         ('estimate_disp', h.estimate_disp),
                         ('lrt', h.lrt), ('bh', h.bh)):
            t = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            times[name] = round(time.perf_counter() - t, 3)
        times['total'] = round(time.perf_counter() - t_all, 3)
        n_px = sum(len(h.load_data('row', c)) for c in chroms)
        nbytes = sum(os.path.getsize(os.path.join(outdir, f))
                     for f in os.listdir(outdir))
        times['prepare_host'] = {k: round(v, 3) for k, v in h.timings.items()
                                 if k.startswith('prepare/')}
        out['trial%d' % trial] = dict(times, union_pixels=n_px,
                                      output_bytes=nbytes,
                                      pixels_per_s=round(n_px / times['total']))
    # the whole pipeline in one call: writers of a step overlap the next step
    import gc
    import shutil
    threads = [-1]
    for a in sys.argv[1:]:
        if a.startswith('--threads='):     # I/O thread settings to compare
            threads = [int(v) for v in a.split('=', 1)[1].split(',')]
    for n_threads, trial in [(n, t) for n in threads for t in range(2)]:
        outdir = os.path.join(root, 'out_all')
        shutil.rmtree(outdir, ignore_errors=True)
        h = HiC3DeFDR(outdir=outdir, dist_thresh_max=200, **kw)
        # host wall time of the four steps inside the one call (no
        # synchronisation added: the writers of a step overlap the next one,
        # "flush" is the wait for the last files at the end)
        host = {}

        def timed(name, fn):
            def run(*a, **k):
                t0 = time.perf_counter()
                try:
                    return fn(*a, **k)
                finally:
                    host[name] = round(host.get(name, 0.0) +
                                       time.perf_counter() - t0, 3)
            return run
        for name in ('estimate_disp', 'lrt', 'bh', '_flush_writes'):
            setattr(h, name, timed(name, getattr(h, name)))
        prepare = h.prepare_data
        h.prepare_data = lambda *a, **k: (
            timed('prepare_data', prepare)(*a, **k) if k.get('chrom') is None
            else prepare(*a, **k))
        # the wrappers above make the previous instance a reference cycle:
        # collect it (device cache, writer threads, pinned buffers) now, not
        # somewhere inside the timed call
        gc.collect()
        t = time.perf_counter()
        h.run_to_qvalues(n_threads=n_threads)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        n_px = sum(len(h.load_data('row', c)) for c in chroms)
        nbytes = sum(os.path.getsize(os.path.join(outdir, f))
                     for f in os.listdir(outdir))
        in_bytes = sum(os.path.getsize(p.replace('<chrom>', c))
                       for p in kw['raw_npz_patterns'] + kw['bias_patterns']
                       for c in chroms)
        key = 'run_to_qvalues_%d' % trial if n_threads == -1 else \
            'run_to_qvalues_threads%d_%d' % (n_threads, trial)
        out[key] = dict(
            total=round(dt, 3), host_steps=host,
            prepare_host={k: round(v, 3) for k, v in h.timings.items()
                          if k.startswith('prepare/')}, union_pixels=n_px,
            input_bytes=in_bytes,
            output_bytes=nbytes, pixels_per_s=round(n_px / dt))
        del h, prepare
    out['generate_inputs_s'] = round(t_gen, 1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
