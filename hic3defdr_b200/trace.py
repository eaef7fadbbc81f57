"""
Stage tracing (off by default): ``H3D_TRACE=1`` makes every ``stage(name)``
block synchronise the device on both sides and accumulate its wall time, so a
run can print where a step's time goes (host glue, collectives and kernels
together).  With tracing off the context manager costs one attribute lookup
and never synchronises.  The reference prints progress per step to stderr
(hic3defdr/util/printing.py:19); this is the timing-aware equivalent.
"""
import collections
import contextlib
import os
import sys
import time

ENABLED = os.environ.get('H3D_TRACE', '0') not in ('', '0')
TIMES = collections.OrderedDict()
COUNTS = collections.Counter()
CALLS = collections.OrderedDict()      # name -> list of per-call ms
# event mode (bench.py): a pair of CUDA events per stage on the current stream,
# no synchronisation; ``event_ms()`` reads them after the caller synchronised
EVENTS = None                          # None: off; else list of (name, ev0, ev1)


def record_events(on=True):
    global EVENTS
    EVENTS = [] if on else None


def event_ms():
    """{stage: total ms} of the recorded event pairs (device time between the
    stage's first and last enqueued work on the stream that runs the stages)."""
    out = collections.OrderedDict()
    for name, e0, e1 in EVENTS or []:
        out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
    return out


@contextlib.contextmanager
def stage(name):
    if EVENTS is not None and not ENABLED:
        import torch
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        try:
            yield
        finally:
            e1.record()
            EVENTS.append((name, e0, e1))
        return
    if not ENABLED:
        yield
        return
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    try:
        yield
    finally:
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        TIMES[name] = TIMES.get(name, 0.0) + dt
        COUNTS[name] += 1
        CALLS.setdefault(name, []).append(round(1e3 * dt, 2))


def reset():
    TIMES.clear()
    COUNTS.clear()
    CALLS.clear()


def report(header='', divide_by=1, file=None):
    if not ENABLED or not TIMES:
        return
    file = file or sys.stderr
    tot = sum(v for k, v in TIMES.items() if '/' not in k)
    print('[h3d trace] %s (top-level total %.1f ms)'
          % (header, 1e3 * tot / divide_by), file=file)
    for k, v in TIMES.items():
        calls = CALLS.get(k, [])
        spread = '' if len(calls) < 2 else '  per call: %s' % calls[:12]
        print('[h3d trace]   %-28s %9.2f ms  (%d calls)%s'
              % (k, 1e3 * v / divide_by, COUNTS[k] // divide_by, spread),
              file=file)
