"""H2D / D2H bandwidth alone and concurrently (is the link's duplex rate the sum
of the two directions on this host?)."""
import torch
n = 1 << 28            # 1 GiB of float32
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, dtype=torch.float32, device='cuda')
d_out = torch.ones(n, dtype=torch.float32, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    gb = reps * n * 4 / 1e9
    return gb / (ms * 1e-3)


run(True, True)
print('H2D alone   %.1f GB/s' % run(True, False))
print('D2H alone   %.1f GB/s' % run(False, True))
print('both: each  %.1f GB/s (sum %.1f)' % (run(True, True), 2 * run(True, True)))
