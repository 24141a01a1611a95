import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from oracle import ransac
from test_gpu_refine import build, RIG

def rodrigues(d):
    th=np.linalg.norm(d)
    K=np.array([[0,-d[2],d[1]],[d[2],0,-d[0]],[-d[1],d[0],0]])
    if th<1e-8: return np.eye(3)+K+0.5*K@K
    return np.eye(3)+np.sin(th)/th*K+(1-np.cos(th))/th**2*K@K

def accumulate(R,t,p,f,cam,rig):
    q=(p-t)@R
    Rc=rig[cam,:,:3]; tc=rig[cam,:,3]
    y=np.einsum('nji,nj->ni',Rc,q-tc)
    ny=np.linalg.norm(y,axis=1,keepdims=True); n=y/ny
    cosv=(f*n).sum(1,keepdims=True); r=1-cosv[:,0]
    g=-(f-cosv*n)/ny
    h=np.einsum('nij,nj->ni',Rc,g)
    Jt=-(h@R.T)
    Jr=np.cross(h,q)
    J=np.hstack([Jt,Jr])
    return r@r, J.T@J, J.T@r

def lm(R,t,p,f,cam,rig,max_iters=80,mode='x10',verbose=False):
    lam=1e-3; cur=None; nu=2.0
    Rt,tt=R.copy(),t.copy()
    for it in range(max_iters):
        cost,A,g=accumulate(Rt,tt,p,f,cam,rig)
        stop=False
        if cur is None:
            acc=True
        elif cost<cur[0]:
            acc=True
            if cur[0]-cost<=1e-15*cur[0]: stop=True
            if mode=='x10': lam=max(lam*0.1,1e-15)
            else:
                rho=(cur[0]-cost)/pred
                lam=lam*max(1/3,1-(2*rho-1)**3); nu=2.0
        else:
            acc=False
            if mode=='x10':
                if cost-cur[0]<=1e-14*cur[0]: stop=True
                lam*=10
            else:
                lam*=nu; nu*=2
            if lam>1e12: stop=True
        if acc: cur=(cost,A,g,Rt.copy(),tt.copy())
        if verbose: print(it,cost,acc,lam)
        if stop: break
        c,A0,g0,Rc_,tc_=cur
        dx=np.linalg.solve(A0+lam*np.diag(np.diag(A0)),-g0)
        pred=-(dx@(2*g0)+dx@A0@dx)  # predicted decrease of sum r^2 (GN model)
        if np.abs(dx).max()<1e-14: break
        Rt=Rc_@rodrigues(dx[3:]); tt=tc_+dx[:3]
    return cur[3],cur[4],cur[0],it+1

rng=np.random.default_rng(5)
P,F,C,M,pose0=build(rng,[300,300],512,RIG)
p=P[0,:300].astype(float); f=F[0,:300].astype(float); cam=C[0,:300].astype(int)
R0=pose0[0,:,:3].astype(float); t0=pose0[0,:,3].astype(float)
U,_,Vt=np.linalg.svd(R0); R0=U@Vt
want,c0,c1=ransac.refine_pose_lm(p,f,pose0[0],cam,RIG,None)
print('minpack',c0,c1)
for mode in ['x10','nielsen']:
    R,t,c,it=lm(R0,t0,p,f,cam,RIG,mode=mode,verbose=(mode=='x10'))
    print(mode,c,it,np.abs(np.hstack([R,t[:,None]])-want).max())
print('--- masked problems')
for rigv,seed in ((RIG,12),(None,11)):
    rng=np.random.default_rng(seed)
    sizes=[40,700,5000]
    P,F,C,M,pose0=build(rng,sizes,5120,rigv)
    rg=RIG if rigv is not None else np.tile(np.hstack([np.eye(3),np.zeros((3,1))]),(2,1,1))
    for b,nb in enumerate(sizes):
        m=M[b,:nb].astype(bool)
        p=P[b,:nb][m].astype(float); f=F[b,:nb][m].astype(float); cam=C[b,:nb][m].astype(int)
        R0=pose0[b,:,:3].astype(float); t0=pose0[b,:,3].astype(float)
        U,_,Vt=np.linalg.svd(R0); R0=U@Vt
        want,c0,c1=ransac.refine_pose_lm(p,f,pose0[b],cam,rigv,None)
        R,t,c,it=lm(R0,t0,p,f,cam,rg,mode='nielsen',max_iters=100)
        print(b,nb,'minpack',c0,c1,'ours',c,it,np.abs(np.hstack([R,t[:,None]])-want).max())
print('--- trace 5000 central')
def lm_trace(R,t,p,f,cam,rig,want,max_iters=60):
    lam=1e-3; cur=None; nu=2.0
    Rt,tt=R.copy(),t.copy()
    for it in range(max_iters):
        cost,A,g=accumulate(Rt,tt,p,f,cam,rig)
        if cur is None: acc=True
        elif cost<cur[0]:
            acc=True; rho=(cur[0]-cost)/pred; lam=lam*max(1/3,1-(2*rho-1)**3); nu=2.0
        else:
            acc=False; lam*=nu; nu*=2
        if acc: cur=(cost,A,g,Rt.copy(),tt.copy())
        print(it,'%.17g'%cost,acc,'%.3g'%lam,'err %.3g'%np.abs(np.hstack([cur[3],cur[4][:,None]])-want).max())
        c,A0,g0,Rc_,tc_=cur
        dx=np.linalg.solve(A0+lam*np.diag(np.diag(A0)),-g0)
        pred=-(dx@(2*g0)+dx@A0@dx)
        Rt=Rc_@rodrigues(dx[3:]); tt=tc_+dx[:3]
lm_trace(R0,t0,p,f,cam,rg,want,40)
print('--- kappa 1.5')
def lm2(R,t,p,f,cam,rig,want=None,max_iters=60,kappa=1.5,verbose=False):
    lam=1e-3; cur=None; nu=2.0
    Rt,tt=R.copy(),t.copy()
    for it in range(max_iters):
        cost,A,g=accumulate(Rt,tt,p,f,cam,rig)
        stop=False
        if cur is None: acc=True
        elif cost<cur[0]:
            acc=True; rho=(cur[0]-cost)/pred; lam=max(lam*max(1/3,1-(2*rho-1)**3),1e-12); nu=2.0
            if cur[0]-cost<=1e-14*cur[0]: stop=True
        else:
            acc=False; lam*=nu; nu*=2
            if cost-cur[0]<=1e-14*cur[0] and dxmax<1e-9: stop=True
        if acc: cur=(cost,A,g,Rt.copy(),tt.copy())
        if verbose: print(it,'%.17g'%cost,acc,'%.3g'%lam,'err %.3g'%np.abs(np.hstack([cur[3],cur[4][:,None]])-want).max())
        if stop: break
        c,A0,g0,Rc_,tc_=cur
        dx=np.linalg.solve(kappa*A0+lam*np.diag(np.diag(A0)),-g0)
        dxmax=np.abs(dx).max()
        pred=-(2*dx@g0+kappa*dx@A0@dx)
        Rt=Rc_@rodrigues(dx[3:]); tt=tc_+dx[:3]
    return cur[3],cur[4],cur[0],it+1
lm2(R0,t0,p,f,cam,rg,want,40,verbose=True)
for rigv,seed in ((RIG,12),(None,11)):
    rng=np.random.default_rng(seed)
    sizes=[40,700,5000]
    P,F,C,M,pose0=build(rng,sizes,5120,rigv)
    rg=RIG if rigv is not None else np.tile(np.hstack([np.eye(3),np.zeros((3,1))]),(2,1,1))
    for b,nb in enumerate(sizes):
        m=M[b,:nb].astype(bool)
        p=P[b,:nb][m].astype(float); f=F[b,:nb][m].astype(float); cam=C[b,:nb][m].astype(int)
        R0=pose0[b,:,:3].astype(float); t0=pose0[b,:,3].astype(float)
        U,_,Vt=np.linalg.svd(R0); R0=U@Vt
        want,c0,c1=ransac.refine_pose_lm(p,f,pose0[b],cam,rigv,None)
        R,t,c,it=lm2(R0,t0,p,f,cam,rg,max_iters=100)
        print(b,nb,'minpack',c0,c1,'ours',c,it,np.abs(np.hstack([R,t[:,None]])-want).max())
rng=np.random.default_rng(5)
P,F,C,M,pose0=build(rng,[300,300],512,RIG)
p=P[0,:300].astype(float); f=F[0,:300].astype(float); cam=C[0,:300].astype(int)
R0=pose0[0,:,:3].astype(float); t0=pose0[0,:,3].astype(float)
U,_,Vt=np.linalg.svd(R0); R0=U@Vt
want,c0,c1=ransac.refine_pose_lm(p,f,pose0[0],cam,RIG,None)
R,t,c,it=lm2(R0,t0,p,f,cam,RIG,max_iters=100)
print('nomask minpack',c1,'ours',c,it,np.abs(np.hstack([R,t[:,None]])-want).max())
print('--- full hessian')
def skew(q):
    z=np.zeros(len(q))
    return np.stack([np.stack([z,-q[:,2],q[:,1]],1),np.stack([q[:,2],z,-q[:,0]],1),np.stack([-q[:,1],q[:,0],z],1)],1)
def accumulate2(R,t,p,f,cam,rig):
    q=(p-t)@R
    Rc=rig[cam,:,:3]; tc=rig[cam,:,3]
    y=np.einsum('nji,nj->ni',Rc,q-tc)
    ny=np.linalg.norm(y,axis=1,keepdims=True); n=y/ny
    cosv=(f*n).sum(1,keepdims=True); r=1-cosv[:,0]
    Dq=np.concatenate([np.broadcast_to(-R.T,(len(p),3,3)),skew(q)],2)  # n,3,6
    Dy=np.einsum('nji,njk->nik',Rc,Dq)
    Pn=(np.eye(3)[None]-n[:,:,None]*n[:,None,:])/ny[:,:,None]
    N=np.einsum('nij,njk->nik',Pn,Dy)   # dn/dx
    J=-np.einsum('ni,nik->nk',f,N)
    A=J.T@J+np.einsum('n,nik,nil->kl',r,N,N)
    return r@r,A,J.T@r
def lm3(R,t,p,f,cam,rig,want=None,max_iters=60,verbose=False):
    lam=1e-4; cur=None; nu=2.0
    Rt,tt=R.copy(),t.copy()
    for it in range(max_iters):
        cost,A,g=accumulate2(Rt,tt,p,f,cam,rig)
        stop=False
        if cur is None: acc=True
        elif cost<cur[0]:
            acc=True; rho=(cur[0]-cost)/pred; lam=max(lam*max(1/3,1-(2*rho-1)**3),1e-12); nu=2.0
            if cur[0]-cost<=1e-14*cur[0]: stop=True
        else:
            acc=False; lam*=nu; nu*=2
            if dxmax<1e-9: stop=True
        if acc: cur=(cost,A,g,Rt.copy(),tt.copy())
        if verbose: print(it,'%.17g'%cost,acc,'%.3g'%lam,'err %.3g'%np.abs(np.hstack([cur[3],cur[4][:,None]])-want).max())
        if stop: break
        c,A0,g0,Rc_,tc_=cur
        dx=np.linalg.solve(A0+lam*np.diag(np.diag(A0)),-g0)
        dxmax=np.abs(dx).max()
        pred=-(2*dx@g0+dx@A0@dx)
        Rt=Rc_@rodrigues(dx[3:]); tt=tc_+dx[:3]
    return cur[3],cur[4],cur[0],it+1
lm3(R0,t0,p,f,cam,RIG,want,40,verbose=True)
for rigv,seed in ((RIG,12),(None,11)):
    rng=np.random.default_rng(seed)
    sizes=[40,700,5000]
    P,F,C,M,pose0=build(rng,sizes,5120,rigv)
    rg=RIG if rigv is not None else np.tile(np.hstack([np.eye(3),np.zeros((3,1))]),(2,1,1))
    for b,nb in enumerate(sizes):
        m=M[b,:nb].astype(bool)
        p=P[b,:nb][m].astype(float); f=F[b,:nb][m].astype(float); cam=C[b,:nb][m].astype(int)
        R0=pose0[b,:,:3].astype(float); t0=pose0[b,:,3].astype(float)
        U,_,Vt=np.linalg.svd(R0); R0=U@Vt
        want,c0,c1=ransac.refine_pose_lm(p,f,pose0[b],cam,rigv,None)
        R,t,c,it=lm3(R0,t0,p,f,cam,rg,max_iters=100)
        print(b,nb,'minpack',c0,c1,'ours',c,it,np.abs(np.hstack([R,t[:,None]])-want).max())
