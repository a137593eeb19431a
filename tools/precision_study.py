"""CPU study behind the operand format of the tensor-core MPNN kernels: the reference's forward re-evaluated with every
activation (am) and every weight of the 64-wide linears (wm) rounded to a candidate operand format at exactly the points where
the kernels round (S/D, g, e, h0, agg, m, h), fp32 accumulation, on the states / Q-values recorded from the reference
(tests/golden/*.npz).  Printed per case: the worst |dq| in units of the parity tolerance (1e-3 |q| + 1e-4 max_row|q|), the
worst error relative to max|Q|, and how often the argmax survives.  Result (round 2): bf16 hi+lo stays below 0.3 of the
tolerance; one fp16 term (= tf32's 10-bit mantissa) exceeds it 10-20x on the 200-vertex cases; one bf16 term 70-110x.
Test infrastructure (imports oracle/): not used by the product."""
import os, sys, glob, numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.mpnn import KEYS, weights_from_npz, as_torch_weights, mpnn_forward

def rnd(x, mode):
    if mode=='fp32': return x
    if mode=='fp16': return x.half().float()
    if mode=='bf16': return x.bfloat16().float()
    if mode=='bf16x2':
        hi=x.bfloat16().float(); lo=(x-hi).bfloat16().float(); return hi+lo
    raise ValueError

def fwd(w, x, adj, am, wm, dmax=None):
    # x [B,N,7], adj [B,N,N]; am: activation rounding, wm: weight rounding (big linears only)
    B,N,_=adj.shape
    deg=(adj!=0).sum(2,keepdim=True).float().clamp(min=1)
    dm = deg.max() if dmax is None else dmax
    We=w[KEYS[1]]; w0=We[:,0]; Wx=We[:,1:]
    P=x@Wx.T
    Rp=F.relu(P+w0); Rm=F.relu(P-w0)
    S=rnd(Rp+Rm,am); D=rnd(Rp-Rm,am)
    g=(adj.abs()@S + adj@D)*0.5/deg
    g=torch.cat([g, deg/dm],-1)
    g=rnd(g,am)
    e=rnd(F.relu(g@rnd(w[KEYS[2]],wm).T),am)
    h=rnd(F.relu(x@w[KEYS[0]].T),am)
    for l in range(3):
        agg=rnd((adj@h)/deg,am)
        m=rnd(F.relu(torch.cat([agg,e],-1)@rnd(w[KEYS[3+2*l]],wm).T),am)
        h=F.relu(torch.cat([h,m],-1)@rnd(w[KEYS[4+2*l]],wm).T)
        if l<2: h=rnd(h,am)
    pooled=(h.sum(1)/N)@w[KEYS[9]].T
    f=F.relu(torch.cat([pooled.unsqueeze(1).expand_as(h),h],-1))
    return (f@w[KEYS[10]].T+w[KEYS[11]]).squeeze(-1)

torch.set_num_threads(8)
for fn in sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', '*.npz'))):
    z=np.load(fn)
    if 'obs' not in z.files or 'q' not in z.files or 'J' not in z.files: continue
    if fn.split('/')[-1].startswith(('s2v','dqn','graphsets','multi','generators')): continue
    w=as_torch_weights(weights_from_npz(z))
    obs=torch.tensor(z['obs']).reshape(-1,7,z['obs'].shape[-1]); q=torch.tensor(z['q']).reshape(-1,z['q'].shape[-1])
    J=torch.tensor(z['J']).float()
    x=obs.transpose(1,2); adj=J.unsqueeze(0).expand(x.shape[0],-1,-1)
    out=[]
    for am,wm in [('fp32','fp32'),('bf16x2','bf16x2'),('fp16','fp16'),('fp16','fp32'),('bf16','bf16')]:
        qq=fwd(w,x,adj,am,wm)
        tol=1e-3*q.abs()+1e-4*q.abs().max(1,keepdim=True).values
        r=((qq-q).abs()/tol).max().item()
        rel=((qq-q).abs().max(1).values/q.abs().max(1).values).max().item()
        # argmax agreement
        agree=(qq.argmax(1)==q.argmax(1)).float().mean().item()
        out.append('%s/%s: tol-ratio %.3f relmax %.1e argmax %.2f'%(am,wm,r,rel,agree))
    print(fn.split('/')[-1], q.abs().max().item()); [print('   ',o) for o in out]
