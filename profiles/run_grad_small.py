import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
import snaq_jl_b200 as S
import bench
n, net, taxa, qt = bench.table100(S)
eng = S.Engine(0)
eng.set_data(S.DataCF(taxa, qt, np.full((len(qt), 3), 1 / 3)))
eng.set_network(net.flatten(taxa))
x0, _, _ = eng.get_params()
for _ in range(8):
    f, g = eng.eval_grad(x0)
print(f, g[:3])
