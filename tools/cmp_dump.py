import numpy as np, sys
a=np.load(sys.argv[1]); b=np.load(sys.argv[2])
for k in a.files:
    d=np.abs(a[k]-b[k]); print(k, "max", d.max(), "mean", d.mean(), "frac>1e-4", (d>1e-4).mean())
