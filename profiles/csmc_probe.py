import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from aux_ssm_samplers_b200 import csmc
from aux_ssm_samplers_b200.draws import ExternalDraws
dt = torch.float64 if sys.argv[1]=='f64' else torch.float32
N=int(sys.argv[2]); T=int(sys.argv[3]); backward=sys.argv[4]=='1'; hist=sys.argv[5]=='1'; C=int(sys.argv[6])
f=lambda v: torch.tensor(v, dtype=dt, device='cuda')
M0=csmc.GaussianDistribution(f([0.0]), f([[1.0]])); Mt=csmc.LinearGaussianDynamics(f([[0.9]]), f([[0.19]]), f([0.0])); G=csmc.FlatPotential()
rng=np.random.default_rng(0)
to=lambda a: torch.as_tensor(a).to('cuda',dt)
_,k=csmc.get_kernel(M0,G,Mt,G,N,backward,return_history=hist)
out=k(ExternalDraws([to(rng.standard_normal((C,T,N,1)))],[to(rng.uniform(size=(C,T-1,N))),to(rng.uniform(size=(C,T)))]), csmc.CSMCState(to(rng.standard_normal((C,T,1))),None))
torch.cuda.synchronize(); print("ok", sys.argv[1:])
