import numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
import fetal_t2mapping_b200 as t2
t2.init(0)
te=np.array([114,132,150,176,202.])
k0,T0=483.08612060546875,158.7274932861328
clean=(k0*np.exp(-te/T0)).astype(np.float32)[None,:].repeat(64,0)
fp={"initial_guess":[650,165],"param_bounds":[(0,10000),(10,2000)]}
for tol in (2e-3,1e-4,1e-6):
  for init in ("loglinear","preset"):
    r=t2.fit_voxels_batch(torch.from_numpy(clean).cuda(), None, te, "gaussian", fp, prior=True, tol=tol, init_mode=init)
    torch.cuda.synchronize()
    print(tol, init, float(r.t2[0]), float(r.k[0]), int(r.nit[0]), int(r.status[0]), "rel", abs(float(r.t2[0])-T0)/T0)
