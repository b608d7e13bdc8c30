#!/usr/bin/env python
"""CPU emulation (test infrastructure: uses the oracle) of the dither's per-step search work for two palettes of the same
noise image - unweighted and saliency-weighted: fraction of queries (pixel + diffused error) that leave the candidate
grid of pb_nngrid.cu, candidates per step by k_nn_cells' criterion, size of the diffused error.  Written to test one
explanation of profiles/r02_saliency.md's open question (it refutes it: 8.3 vs 9.1 candidates per step, 0.006 % vs
0.004 % of the queries outside the grid).      python tools/dither_query_stats.py"""
import sys, numpy as np, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle.reflib import OracleLib
from oracle.ref_build import build_ref_pyx
from scipy.spatial import cKDTree
ref = build_ref_pyx.load()
side=512
rng=np.random.default_rng(3)
colors=rng.integers(0,256,(side*side,3),dtype=np.uint8)/255.0
w=np.asarray(ref.get_weights(colors.reshape(side,side,3), side/32.0))
o=OracleLib()
M=np.array([[0.6274,0.3293,0.0433],[0.0691,0.9195,0.0114],[0.0164,0.0880,0.8956]])
cw=np.array([0.51254268114958,0.8234075540095561,0.2435159132377184])
def lin(c): return np.where(c<=0.04045,c/12.92,((c+0.055)/1.055)**2.4)
pix=lin(colors)@M.T
# hilbert order
def d2xy(n,d):
    x=y=0; t=d; s=1
    while s<n:
        rx=1&(t//2); ry=1&(t^rx)
        if ry==0:
            if rx==1: x=s-1-x; y=s-1-y
            x,y=y,x
        x+=s*rx; y+=s*ry; t//=4; s*=2
    return x,y
order=[ (lambda xy: xy[1]*side+xy[0])(d2xy(side,d)) for d in range(side*side)]
m=16**(1/15); qw=np.array([m**i/16 for i in range(16)])
pw=pix*cw; lo=pw.min(0); hi=pw.max(0); r=hi-lo; glo=lo-0.25*r; ghi=hi+0.25*r; NG=32; wd=(ghi-glo)/NG
for name,ww in (('unweighted',None),('saliency',w)):
    code,pal,pmap=o.quantize(side,side,colors,256,dither=False,color_space=2,kmeans_niter=10,weights=ww)
    pal=pal[pal[:,0]>=0]
    P=lin(np.clip(pal,0,1))@M.T; PW=P*cw
    # per-cell candidate counts
    idx=np.stack(np.meshgrid(np.arange(NG),np.arange(NG),np.arange(NG),indexing='ij'),-1).reshape(-1,3)
    clo=glo+idx*wd; chi=clo+wd; cnts=np.empty(len(idx),int)
    for s in range(0,len(idx),2048):
        a=clo[s:s+2048,None,:]; b=chi[s:s+2048,None,:]; p=PW[None,:,:]
        out=np.maximum(np.maximum(a-p,p-b),0.0); mind=(out**2).sum(-1)
        far=np.maximum(np.abs(p-a),np.abs(p-b)); maxd=(far**2).sum(-1)
        cnts[s:s+2048]=(mind<=maxd.min(1,keepdims=True)).sum(1)
    tree=cKDTree(PW)
    q=np.zeros((16,3)); outside=0; cand=0; errs=[]
    t0=time.time()
    for p_ in order:
        e=(q*qw[:,None]).sum(0)
        c=pix[p_]+e
        cq=c*cw
        cell=np.floor((cq-glo)/wd).astype(int)
        if (cell<0).any() or (cell>=NG).any(): outside+=1; cand+=256
        else: cand+=cnts[cell[0]*NG*NG+cell[1]*NG+cell[2]]
        _,j=tree.query(cq)
        q[:-1]=q[1:]; q[-1]=pix[p_]-P[j]
        errs.append(np.abs(e).max())
    n=len(order)
    print(name,'out-of-grid queries %.3f %%, mean candidates per step %.1f, |diffused error| mean %.3f p99 %.3f max %.3f (%.0fs)'%(100*outside/n,cand/n,np.mean(errs),np.percentile(errs,99),np.max(errs),time.time()-t0))
