"""CUDA-event timings of single ops at the sizes of the mouse-genome step
(A/B of library builds: H3D_LIB=<path> python tools/microbench.py)."""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from hic3defdr_b200 import ops  # noqa: E402


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 3)


def main():
    g = torch.Generator(device='cuda').manual_seed(1)
    out = {}
    n_big, n_chr = 38_700_000, 2_400_000
    p = torch.rand(n_big, generator=g, device='cuda', dtype=torch.float64) ** 2
    out['bh_38.7M_ms'] = timeit(lambda: ops.adjust_pvalues(p))
    row = torch.arange(n_chr // 201 + 1, device='cuda').repeat_interleave(201)[:n_chr]
    dist = (torch.arange(n_chr, device='cuda') % 201).to(torch.int32)
    out['stable_rank_2.4M_201keys_ms'] = timeit(lambda: ops.stable_rank(dist, 201), 20)
    dist_big = (torch.arange(n_big, device='cuda') % 201).to(torch.int32)
    out['stable_rank_38.7M_201keys_ms'] = timeit(lambda: ops.stable_rank(dist_big, 201))
    bal = torch.rand((n_chr, 4), generator=g, device='cuda', dtype=torch.float64) * 50
    out['size_factor_table_2.4M_ms'] = timeit(
        lambda: ops.size_factor_table(bal, dist, 200, 40, 'conditional_mor'), 20)
    # connected components (threshold()): every tested pixel of the genome as
    # ONE set, 55 % of a 200-bin band (the reference's find_clusters walks the
    # pixels in a Python loop: 31 k pixels/s on this container's CPU)
    n_rows = n_big // 110
    i = torch.arange(n_rows, device='cuda').repeat_interleave(201)
    j = i + torch.arange(201, device='cuda').repeat(n_rows)
    keep = torch.rand(i.numel(), generator=g, device='cuda') < 0.55
    i, j = i[keep].to(torch.int32), j[keep].to(torch.int32)
    ms = timeit(lambda: ops.connected_components(i, j), 3)
    out['connected_components_%.1fM_ms' % (i.numel() / 1e6)] = ms
    out['connected_components_Mpx_per_s'] = round(i.numel() / ms / 1e3, 1)
    out['connected_components_gbs_of_24B_per_px'] = round(
        24.0 * i.numel() / ms / 1e6, 1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
