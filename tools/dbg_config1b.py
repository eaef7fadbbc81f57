import sys, os, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hic3defdr_b200 import HiC3DeFDR
from hic3defdr_b200.synth import MM10_10KB, write_dataset
root = tempfile.mkdtemp()
kw = write_dataset(root, {c: MM10_10KB[c] for c in ('chr18', 'chr19')}, n_reps=4, dist_max=200, config=1, amp=300.0)
kw.pop('loop_patterns')
h = HiC3DeFDR(outdir=os.path.join(root, 'out'), dist_thresh_max=200, **kw)
h.run_to_qvalues(n_threads=0)
got = np.load(os.path.join(root, 'out', 'disp_per_dist.npy'))
np.save('gpurun_out/config1_dpd.npy', got)
g = np.load('tests/golden/ref_config1.npz')
want = g['disp_per_dist']
rel = np.abs(got - want) / want
idx = np.argsort(-np.nan_to_num(rel.ravel()))[:12]
for i in idx:
    d, c = divmod(i, 2)
    print('d=%3d c=%d got %.12g want %.12g rel %.2e' % (d, c, got[d, c], want[d, c], rel[d, c]))
print('median rel', np.nanmedian(rel), 'stats', h.timings['qcml_stats'])
