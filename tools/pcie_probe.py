"""Probe of the GPU box's host<->device copy rates (pinned memory), idle and
while kernels run, plus the PCIe link state nvidia-smi reports.  Diagnostic
for the e2e leg of bench.py; not part of the product."""
import subprocess
import time

import torch


def link():
    try:
        return subprocess.run(
            ['nvidia-smi', '--query-gpu=pcie.link.gen.current,pcie.link.gen.max,'
             'pcie.link.width.current,pstate,clocks.sm', '--format=csv,noheader',
             '-i', '0'], capture_output=True, text=True).stdout.strip()
    except Exception as e:
        return str(e)


def bw(fn, nbytes, reps=5):
    torch.cuda.synchronize()
    best = 0
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = max(best, nbytes / (time.perf_counter() - t0) / 1e9)
    return best


def main():
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device='cuda')
    print('link idle:', link())
    print('H2D idle  %.1f GB/s' % bw(lambda: d.copy_(h, non_blocking=True), n))
    print('D2H idle  %.1f GB/s' % bw(lambda: h.copy_(d, non_blocking=True), n))
    print('link after copies:', link())
    a = torch.randn(8192, 8192, device='cuda')
    side = torch.cuda.Stream()

    def busy_copy(dst, src):
        def f():
            for _ in range(40):
                a @ a
            with torch.cuda.stream(side):
                dst.copy_(src, non_blocking=True)
        return f
    # copy alone on the side stream, timed with events while matmuls run
    for name, dst, src in (('H2D', d, h), ('D2H', h, d)):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(60):
            a @ a
        with torch.cuda.stream(side):
            e0.record()
            dst.copy_(src, non_blocking=True)
            e1.record()
        for _ in range(60):
            a @ a
        torch.cuda.synchronize()
        print('%s under load %.1f GB/s' % (name, n / (e0.elapsed_time(e1) * 1e-3) / 1e9),
              'link:', link())
    # many small-ish copies (40 MB), like per-chromosome arrays
    m = 40 << 20
    hs = [torch.empty(m, dtype=torch.uint8).pin_memory() for _ in range(16)]
    ds = [torch.empty(m, dtype=torch.uint8, device='cuda') for _ in range(16)]

    def many_d2h():
        for x, y in zip(hs, ds):
            x.copy_(y, non_blocking=True)
    print('D2H 16 x 40 MB idle %.1f GB/s' % bw(many_d2h, 16 * m))
    # fresh pinned allocation cost
    t0 = time.perf_counter()
    x = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    print('pin 1 GiB: %.0f ms' % (1e3 * (time.perf_counter() - t0)))
    t0 = time.perf_counter()
    x = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
    print('alloc pinned 1 GiB directly: %.0f ms' % (1e3 * (time.perf_counter() - t0)))
    time.sleep(2.0)
    print('link after 2 s idle:', link())
    print('D2H after idle %.1f GB/s' % bw(lambda: h.copy_(d, non_blocking=True), n, reps=1))


if __name__ == '__main__':
    main()
