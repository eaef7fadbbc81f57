import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hic3defdr_b200 import engine, ops
from hic3defdr_b200.synth import make_chrom
design = np.array([[1,0],[1,0],[0,1],[0,1]],dtype=bool)
ins = []
for i, n in enumerate((500, 400)):
    mats, bias, _ = make_chrom(n, 4, 40, seed=5+i, amp=200.0)
    ins.append((ops.DeviceCSR(mats), ops.dev(bias)))
states, dpd, fns, stats = engine.run_to_qvalues(ins, design, dist_max=40)
torch.cuda.synchronize()
print('ok', dpd[:6], stats)
