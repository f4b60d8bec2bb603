import sys; sys.path[:0]=['/root/repo','/root/repo/dnn-mppi-mpc_b200','/root/repo/tests']
import numpy as np, torch
from golden_util import *
from gpu_util import *
from oracle import c_oracle as co, mppi_oracle as orc
def _dev(a): return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
g = Golden("racecar_default")
sp = g.spec(cost_mode="last", waypoint_mode="frozen")
eng = engine_from_spec(sp, g.path)
S = torch.zeros(sp.K, dtype=torch.float32, device="cuda")
for i in range(4):
    inp = g.tick_inputs(i)
    o = co.tick(sp, **inp)
    eng.set_nominal(inp["U"]); eng.set_waypoint_idx(inp["idx"])
    eng.rollout_costs(inp["x0"], S, _dev(inp["eps"]))
    Sg = S.cpu().numpy()
    frac, worst = cost_mismatch(Sg, o["S"])
    bad = np.nonzero(np.abs(Sg-o["S"]) > 1e-6+1e-5*np.abs(o["S"]))[0]
    print(i, frac, worst, bad[:5], Sg[bad[:5]], o["S"][bad[:5]])
    eng.set_waypoint_idx(inp["idx"])
    u0, useq = eng.step(inp["x0"], _dev(inp["eps"]))
    print("  U err", np.max(np.abs(useq - o["U_after"])), eng.get_waypoint_idx(), o["idx_after"])
# noise
sigma = np.array([[0.1, 0.02], [0.02, 0.05]])
sp = orc.diffdrive_spec(K=8192, T=31, sigma=sigma, cost_mode="sum", waypoint_mode="frozen")
g = Golden("diffdrive_pe0.05")
eng = engine_from_spec(sp, g.path)
out = torch.zeros(sp.K, sp.T, 2, dtype=torch.float32, device="cuda")
eng.generate_noise(out, seed=0x1234567890ABCDEF, tick=17)
e = out.cpu().numpy().astype(np.float64)
ref = orc.philox_noise(0x1234567890ABCDEF, 17, sp.K, sp.T, sigma)
d = np.abs(e-ref)
print("noise maxdiff", d.max(), np.unravel_index(d.argmax(), d.shape), e.reshape(-1,2).mean(0), np.cov(e.reshape(-1,2).T))
print(e[0,:3], ref[0,:3])
