import sys, time, numpy as np, torch
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import helpers_mmsto as hm
from ipcdnnwalk_b200 import ShortTrajOpt
skel=hm.skeleton()
ps=[hm.make_problem(s) for s in range(6)]
opt=ShortTrajOpt(skel,hm.NF,hm.MARKERS)
w=hm.ankle_dof_weights(); opt.set_dof_weights(w)
opt.set('weightEEcon',1000.0)
rng=np.random.default_rng(9)
xs=np.stack([p['x_original']+rng.normal(0,0.01,p['x_original'].shape) for p in ps])
args=[hm.stack(ps,k) for k in ('gpos_target','x_original','dx','ddx','com','momentum')]
obj,grad=opt.evaluate(xs,*args,con=hm.stack(ps,'con'),velcon=hm.stack(ps,'velcon'))
torch.cuda.synchronize()
for i,p in enumerate(ps[:3]):
    o=hm.oracle_opt(p,dof_weights=w,weightEEcon=1000.0)
    f,g=o.evaluate(xs[i].ravel())
    print(i,'obj',f,obj[i].item(),'rel',abs(f-obj[i].item())/abs(f),'grad err',np.abs(g.reshape(7,59)-grad[i].cpu().numpy()).max()/np.abs(g).max())
opt.set('weightEEcon',0.0)
for mi in (3,0):
    opt.set('max_iterations',mi)
    t=time.time()
    x,info=opt.solveEndEffectors_euler(hm.stack(ps,'euler_mot'),*args,con=hm.stack(ps,'con'),velcon=hm.stack(ps,'velcon'))
    torch.cuda.synchronize(); print('solve',mi,time.time()-t); print(info.cpu().numpy())
    for i,p in enumerate(ps[:2]):
        o=hm.oracle_opt(p,dof_weights=w)
        X,inf=o.solve(p['euler_mot'],max_iterations=mi)
        print(i,inf,'x err',np.abs(X-x[i].cpu().numpy()).max())
