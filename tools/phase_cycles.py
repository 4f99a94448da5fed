"""GPU: per-warp phase cycle counts from the debug record (pad slots 19..23 of sub-step 0)."""
import sys, os
import numpy as np, torch
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,"tests"))
import helpers
from ipcdnnwalk_b200 import SRBVecEnv
clips=helpers.fresh_clips()
prec=int(sys.argv[1]) if len(sys.argv)>1 else 32
N=4096
env=SRBVecEnv(clips,N,variant="test",precision=prec,lanes_per_env=8,auto_reset=True,seed=1)
env.attach(dbg=True); env.reset()
act=torch.zeros(N,22,device="cuda")
flush=torch.empty(256*1024*1024,dtype=torch.uint8,device="cuda")
for k in range(30):
    env.tracking_action(act,sigma=0.1,seed=1234,step=k)
    env.dbg.zero_(); flush.fill_(k)
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); env.step(act); e1.record(); torch.cuda.synchronize()
    if k>=25:
        d=env.dbg.cpu().numpy().reshape(N,4,64)
        tot=d[:,0,21]; pre=d[:,0,22]; subs=d[:,0,20]; entry=d[:,0,23]
        qp=d[:,:,19]; its=d[:,:,18]
        w=slice(0,N,4)   # one env per warp
        gexit=d[:,0,23]; gentry=d[:,1,23]
        t0=gentry[w].min()
        print(f"step {k}: event {e0.elapsed_time(e1)*1e3:.1f}us | warp cycles: load+decode {np.median(pre[w]):.0f}  4 substeps med {np.median(subs[w]):.0f} p99 {np.percentile(subs[w],99):.0f} max {subs[w].max():.0f} | total med {np.median(tot[w]):.0f} max {tot[w].max():.0f}"
              f" | post med {np.median((tot-subs)[w]):.0f} max {(tot-subs)[w].max():.0f} | globaltimer ns: entry spread {(gentry[w].max()-t0):.0f} exit min {(gexit[w].min()-t0):.0f} med {np.median(gexit[w]-t0):.0f} max {(gexit[w].max()-t0):.0f} | qp cyc/substep med {np.median(qp[qp>0]):.0f} iters/step max-in-warp mean {its.reshape(N//4,4,4).max(axis=1).sum(axis=1).mean():.1f}")
        rs = d[:, 1, 19:23]
        m = rs[:, 2] > 0                                   # environments that were reset in this step
        if m.any():
            print(f"   reset path (cycles after done is known, {int(m.sum())} envs): plan drawn {np.median(rs[m,0]):.0f}  state reset {np.median(rs[m,1]):.0f}  obs rewritten {np.median(rs[m,2]):.0f}  state stored {np.median(rs[m,3]):.0f} | no-reset envs: state stored {np.median(rs[~m,3]):.0f}")
env.close()
