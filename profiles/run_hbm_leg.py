import sys, json; sys.path.insert(0,'/root/repo')
import torch, numpy as np
import bench, snaq_jl_b200 as S
r = bench.hbm_leg(S, S.Engine, 0, torch)
print(json.dumps({k:r[k] for k in ("ms_per_call","frac","achieved","bytes_per_quartet","evals_per_s","descriptor_build_s")}))
