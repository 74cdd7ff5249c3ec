import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import bench
bench.NB_TOTAL = 4096
for nb in (6, 300):
    eng, _ = bench.make_ensemble(0, nb, 0)
    st = eng.step(2)
    f = eng.get_f()
    print(nb, st['fmin'], st['negatives'], float(f.min()), int((f < 0).sum()), int((f == 0).sum()))
    eng.close()
