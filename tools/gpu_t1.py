import sys, math, torch
sys.path.insert(0, '/root/repo')
from conformal_bayesopt_b200 import ops, GPModelState
from oracle import closed_form, gp as ogp
from tests.problems import make_problem, make_inputs, oracle_state
torch.manual_seed(0)
dev='cuda:0'
# gemm test
for M in (100, 300):
    K=256
    A=torch.randn(M,K,dtype=torch.float64,device=dev); B=torch.randn(K,K,dtype=torch.float64,device=dev)
    Bl=torch.tril(B)
    C1,rn=ops.gemm_tn(A,Bl,1,True); ref=A@Bl.T
    print('gemm lower', (C1-ref).abs().max().item(), (rn.sum(1)-(ref**2).sum(1)).abs().max().item())
    Bu=torch.triu(B); C2=ops.gemm_tn(A,Bu,2); print('gemm upper',(C2-A@Bu.T).abs().max().item())
    C3=ops.gemm_tn(A,B,0); print('gemm full',(C3-A@B.T).abs().max().item())
for (N,d,b,G,temp,rand,ratio_on,refine,kind,S) in [(40,3,6,16,3e-2,False,False,0,'ucb',1),(200,8,37,64,3e-2,True,True,0,'ucb',1),(130,4,9,32,0.0,True,True,0,'ei',2),(64,2,20,64,1e-6,True,False,2,'ei',1),(300,8,50,64,3e-2,True,True,3,'ucb',3)]:
    p=make_problem(N,d,seed=N); gp=oracle_state(p); inp=make_inputs(N,d,b,1,G,S=S,seed=N)
    ratio=ogp.make_ratio_mlp(d) if ratio_on else None
    kw=dict(kind=kind,alpha=0.1,temp=temp,grid_res=G,max_grid_refinements=refine,randomized=rand,optimistic=True,beta_prime=math.sqrt(0.2*math.pi/2),best_f=0.3)
    o=closed_form.conformal_acquisition(gp,inp['Xq'],eps_grid=inp['eps_grid'],z=inp['z'],eps_mc=inp['eps_mc'],ratio=ratio,u_rand=inp['u_rand'],**kw)
    st=GPModelState.fit_fixed(p['X'],p['y'],p['mean'],p['lengthscale'],p['outputscale'],p['noise'],device=dev)
    rw=None
    if ratio is not None:
        rw=ops.RatioWeights(*[t.to(dev) for t in (ratio.W1,ratio.b1,ratio.W2,ratio.b2,ratio.W3,ratio.b3)])
    g=ops.conformal_acq(st,inp['Xq'].to(dev),eps_grid=inp['eps_grid'][:,0].to(dev),z=inp['z'].to(dev),eps_mc=inp['eps_mc'][:,:,0].to(dev),u_rand=inp['u_rand'][:,:,0].to(dev),ratio=rw,outputs=('target_grid','grid_logp','conf_pred_mask','ood_mask','post_mean','post_var'),**kw)
    torch.cuda.synchronize()
    def rel(a,b): return ((a.cpu()-b).abs()/(b.abs()+1e-12)).max().item()
    print(N,kind,temp,'acq',rel(g['acq'],o['acq']),'grid',rel(g['target_grid'],o['target_grid'][:,:,0]),'logp',rel(g['grid_logp'],o['logp']),'mask',(g['conf_pred_mask'].cpu()-o['mask'][:,:,0]).abs().max().item(),'mu',rel(g['post_mean'],o['mu'][:,0]),'s',rel(g['post_var'],o['S'][:,0,0]))
