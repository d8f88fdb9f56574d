import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np, snaq_jl_b200 as S
n=100
net=S.simulate.random_level1_network(n,3,20261021)
taxa=[f"t{i+1}" for i in range(n)]
qt=S.simulate.all_quartets(n)
eng=S.Engine(0); eng.set_data(S.DataCF(taxa,qt,np.full((len(qt),3),1/3)))
fl=net.flatten(taxa)
for _ in range(3): eng.set_network(fl)
t0=time.perf_counter()
for _ in range(10): eng.set_network(fl)
print("set_network ms", (time.perf_counter()-t0)/10*1e3)
