import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from aux_ssm_samplers_b200 import csmc
from aux_ssm_samplers_b200.draws import ExternalDraws
from oracle import csmc as oc
dt = torch.float32; npd = np.float32
N, T, C, d = 300, 12, 3, 1
rng = np.random.default_rng(1312)
model = oc.LinearGaussianModel([0.2], [[0.8]], [[0.8]], [[0.4]], [0.1], "flat", dtype=npd)
x_star = rng.standard_normal((C, T, d)).astype(npd)
normals = rng.standard_normal((C, T, N, d)).astype(npd)
u_res = rng.uniform(size=(C, T - 1, N)).astype(npd)
u_bwd = rng.uniform(size=(C, T)).astype(npd)
to = lambda a: torch.as_tensor(a).to('cuda', dt)
M0 = csmc.GaussianDistribution(to(model.m0), to(model.P0)); Mt = csmc.LinearGaussianDynamics(to(model.F), to(model.Q), to(model.b)); G = csmc.FlatPotential()
_, k = csmc.get_kernel(M0, G, Mt, G, N, True, return_history=True)
st, h = k(ExternalDraws([normals], [u_res, u_bwd]), csmc.CSMCState(to(x_star), None))
torch.cuda.synchronize()
for c in range(C):
    x, upd, B, (w_T, xs, lws, As) = oc.csmc_kernel(oc.BootstrapFK(model), model, x_star[c], normals[c], u_res[c], u_bwd[c], True)
    gB = h["ancestors"][c].cpu().numpy()
    print("chain", c, "As equal", np.array_equal(h["As"][c].cpu().numpy(), As), "B oracle", B.tolist(), "B gpu", gB.tolist())
