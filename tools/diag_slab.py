import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
import rgpot_b200
from rgpot_b200 import workloads
import bench

def percall(pot, pos, Z, box, n=30):
    F = np.zeros_like(pos)
    pot(pos, Z, box, out=F)
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); pot(pos, Z, box, out=F); ts.append((time.perf_counter() - t0) * 1e3)
    return [round(t, 2) for t in ts]

pos, Z, box = workloads.lj_fluid(1000000)
lj = rgpot_b200.LJPot(rgpot_b200.LJConfig(1.0, 2.5, 1.0), device=0)
print("lj pageable ms", percall(lj, pos, Z, box, 10))
if len(sys.argv) > 1:
    print("cpu baseline lj", bench.cpu_side_baseline("lj_fluid_1m")["value"])
    print("cpu linear", bench.cpu_lj_linear_comparator()["value"])
p2, Z2, b2 = workloads.cu_slab_triclinic(cells=40)
cu = rgpot_b200.CuH2Pot(device=0)
print("slab pageable ms", percall(cu, p2, Z2, b2, 40))
