import sys, math, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, ccb200 as cc, orc
from helpers import rel_err, zd_of
mech, cfg = cc.examples.sawyer('/root/repo/tests/golden/sawyer_arm.json')
om = orc.OracleMech(mech.desc())
eqc_idx = [mech.eqc_index(i) for i in cfg["eqcids"]]
zd = zd_of(mech, cfg)
Qd = cc.blockdiag(cfg["Q"]) * mech.dt; Rd = cc.blockdiag(cfg["R"]) * mech.dt
N = 2000
out = cc.lqr_synthesize(mech, zd[None], np.zeros((1, 7)), eqc_idx, Qd, Rd, N)
K = out["Ku"][0]
print('K max', np.abs(K).max(), 'kexit', out['kexit'], 'iters', out['iters'])
lqr = cc.LQR.from_gains(mech, cfg["eqcids"], K, N, xd=zd[:, 0:3], qd=zd[:, 3:7])
for amp in (0.05, 0.01, 0.002):
    rng = np.random.default_rng(3)
    B = 32
    st = np.zeros((B, mech.nb, 13))
    for s in range(B):
        cc.urdf.zero_configuration(mech, {i: rng.uniform(-amp, amp) for i in range(7)})
        st[s] = mech.state_vector()
    co = om.make_ctrl(B, kind="lqr_finite", eqc=eqc_idx, N=N, K=K, zd=zd, Fd=np.zeros(7))
    steps = 300
    r = om.simulate(st, steps, ctrl=co, record=True, nthreads=8)
    sto, fin, status = cc.simulate(mech, steps * mech.dt, lqr, record=True, states=st, return_status=True)
    print('amp', amp, 'oracle fails', int((r['iters'] <= 0).sum()), 'gpu fails', int((status <= 0).sum()), 'mean iters', np.abs(r['iters']).mean()/steps, np.abs(status).mean()/steps)
    ok = (r['iters'] > 0) & (status > 0)
    if ok.any(): print('   parity on converged systems', rel_err(sto.data[:, ok], r['storage'][:, ok]), 'final dev from setpoint', np.abs(fin[ok][:, :, 0:3] - zd[:, 0:3]).max())
