import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'ilqr-admm_b200'); sys.path.insert(0,'tests')
from oracle import models as M, restated as R
import test_gpu_sls as T
from isls_b200 import SetConvexSOC
g=np.load('tests/golden/sls_admm.npz')
tag,pos_dim,N,dt='nb',1,100,0.01
n,m,A,B,tg,A_,b_=T._case(g,tag,pos_dim,N,dt)
s=T._make_sls(n,m,N,A,B,tg)
proj=SetConvexSOC(A_,b_,rho=1e1,max_iter=100,threshold=1e-3)
du,phi,logs=s.ADMM_SLS(project_u=proj,max_iter=50,rho_u=1e2,alpha=1.0,tol=1e-3,log=True)
print(logs[0,:6].cpu().numpy()); print(g['nb_logs'][0,:6])
print('inner', s.last.inner_total, 'iters', s.last.iters)
Qt=np.zeros((N,n)); Qt[-1]=1e6
xd=np.zeros((1,N,n)); xd[:,-1,:pos_dim]=tg
o=R.admm_sls(A,B,N,Qt,xd.reshape(1,-1),1e-2,A_,b_,1e2,max_iter=50,alpha=1.0,tol=1e-3,inner_rho=1e1,inner_max_iter=100,inner_threshold=1e-3)
print('oracle inner', o['inner_total'])
# one-iteration check with fixed budget
du1,phi1,l1=s.ADMM_SLS(project_u=proj,max_iter=1,rho_u=1e2,alpha=1.0,tol=1e-3,log=True,fixed_budget=True)
o1=R.admm_sls(A,B,N,Qt,xd.reshape(1,-1),1e-2,A_,b_,1e2,max_iter=1,alpha=1.0,tol=1e-3,inner_rho=1e1,inner_max_iter=100,inner_threshold=1e-3,fixed_budget=True)
print('x_u iter1 diff', np.abs(du1.cpu().numpy()-o1['du']).max(), np.abs(phi1.cpu().numpy()[:,:,:1]-o1['phi_u'][:,:,:1]).max(), 'inner', s.last.inner_total, o1['inner_total'], l1[0,0].cpu().numpy(), o1['logs'][0])
