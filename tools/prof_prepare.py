import sys, cProfile, pstats, io, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from hic3defdr_b200 import engine, staging
from hic3defdr_b200.synth import MM10_10KB
design = np.array([[1,0],[1,0],[0,1],[0,1]], dtype=bool)
inputs = []
for i,(c,n) in enumerate(MM10_10KB.items()):
    mats, bias = bench.gen_chrom_device(n, 4, 200, 20261018+1000+100*i, 300.0)
    inputs.append((staging.csr_to_device(mats, n), bias))
torch.cuda.synchronize()
for _ in range(3):
    st = engine.prepare_many(inputs, design, dist_max=200); torch.cuda.synchronize()
t=time.perf_counter(); st = engine.prepare_many(inputs, design, dist_max=200); torch.cuda.synchronize(); print('wall ms', 1e3*(time.perf_counter()-t))
pr = cProfile.Profile(); pr.enable()
st = engine.prepare_many(inputs, design, dist_max=200); torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(18); print(s.getvalue()[:3500])
