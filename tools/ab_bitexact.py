"""python tools/ab_bitexact.py a.npz b.npz — exact comparison of two tools/dump_logits.py outputs."""
import sys
import numpy as np

a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
ok = True
for k in a.files:
    same = np.array_equal(a[k], b[k])
    ok &= same
    print(f"{k:12s} {'bit-identical' if same else 'DIFFERENT: max|d| = %g' % float(np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)).max())}")
sys.exit(0 if ok else 1)
