"""debug: per-bin device qCML vs the oracle on chr19 of BASELINE configs[0]"""
import sys, warnings, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
from oracle import pipeline as op
from hic3defdr_b200 import ops
from hic3defdr_b200.synth import make_chrom, BASE_SEED
design = np.array([[1,0],[1,0],[0,1],[0,1]],dtype=bool)
mats, bias, _ = make_chrom(6144, 4, 200, BASE_SEED + 1000*1 + 100*1, amp=300.0)
st = op.prepare_chrom(mats, bias, design, dist_max=200)
di = st['disp_idx']
row, col = st['row'][di], st['col'][di]
b = op.filter_bias(bias, 0.1)
f = op.combined_factor(b, row, col, st['size_factors'][di])
raw = st['raw'][di]; dist = col-row
for d in (5, 20, 60, 120, 190):
    sel = dist == d
    x, ff = raw[sel][:, :2], f[sel][:, :2]
    trace = []
    base = op.qcml(x, f=ff.copy(), trace=trace)
    got = ops.qcml(x, ff)
    pseudo = op.equalize(x, ff.copy(), base)
    dl = base / (1 + base)
    nl = [(ops.cml_nll(pseudo, t) - op.cml_nll(pseudo, t)) / abs(op.cml_nll(pseudo, t)) for t in (dl*0.5, dl, dl*(1+1e-5), dl*2)]
    pd = ops.equalize(x, ff, base).cpu().numpy()
    rel = np.abs(pd - pseudo) / np.maximum(pseudo, 1e-3)
    print('d=%3d n=%5d oracle %.12g device %.12g rel %.2e | nll rel diffs %s | pseudo max rel %.2e | oracle trace %s' % (
        d, sel.sum(), base, got, abs(got-base)/base, ['%.1e' % v for v in nl], rel.max(), trace), flush=True)
