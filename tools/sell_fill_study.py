"""Offline study of SELL layouts (rejected variants): places per live entry of one cohort sample for the paired even/odd slots the kernel uses and for the alternatives (even-then-odd slots, single stream with a parity bit, pairs in index order, half-lines, slices of 16 / 8 lines).  Pure numpy, no GPU."""
import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bambu_b200 import synthetic as S
from bambu_b200.wire import Wire
ann=S.annotation_for("config4"); a=S.cohort_sample(ann,0)
w=Wire.from_arena(a)
G=w.n_groups
M=np.diff(w.grp_row_ptr); NL=np.diff(w.grp_col_ptr); E=np.diff(w.grp_ent_ptr)
idx16=None
# decode indices per group (all u16 here)
tot={}
def add(k,v): tot[k]=tot.get(k,0)+v
rng=np.random.default_rng(0)
sel=rng.choice(G,600,replace=False)
for g in sel:
    m,nl,e=int(M[g]),int(NL[g]),int(E[g])
    if m<2 or e==0: continue
    # find idx offset: recompute
    pass
# compute idx offsets
off=np.zeros(G+1,np.int64); o=0
for g in range(G):
    o=(o+3)//4*4; off[g]=o; o+=E[g]*(2 if NL[g]<=32767 else 4)
for g in sel:
    m,nl,e=int(M[g]),int(NL[g]),int(E[g])
    if m<2 or e==0: continue
    c=w.cidx[off[g]:off[g]+2*e].view(np.uint16).astype(np.int64)
    lj=c>>1; par=c&1
    re=w.row_end[w.grp_row_ptr[g]:w.grp_row_ptr[g+1]].astype(np.int64)
    rows=np.repeat(np.arange(m),np.diff(np.concatenate(([0],re))))
    for side,(line,parity,n) in {"R":(rows,par,m),"C":(lj,rows&1,nl)}.items():
        ne=np.bincount(line[parity==0],minlength=n); no=np.bincount(line[parity==1],minlength=n)
        add(side+"_live",e)
        # current: L=max(ne,no), sort desc, slices of 32, slots=sum of slice max; places=2*32*slots
        L=np.maximum(ne,no); Ls=np.sort(L)[::-1]
        pad=(-len(Ls))%32; Ls=np.concatenate((Ls,np.zeros(pad,int))).reshape(-1,32)
        add(side+"_cur_places",2*32*Ls.max(1).sum())
        # split-phase, sort by (ne,no) lexicographic desc
        o2=np.lexsort((-no,-ne)); 
        for name,order in (("lex",o2),("tot",np.lexsort((-ne,-(ne+no)))),("max",np.lexsort((-np.minimum(ne,no),-L)))):
            e2=np.concatenate((ne[order],np.zeros(pad,int))).reshape(-1,32); o3=np.concatenate((no[order],np.zeros(pad,int))).reshape(-1,32)
            add(side+"_split_"+name,32*(e2.max(1)+o3.max(1)).sum())
        # single-stream with parity bit: slots = max total per slice sorted by total
        t=np.sort(ne+no)[::-1]; t=np.concatenate((t,np.zeros(pad,int))).reshape(-1,32)
        add(side+"_single",32*t.max(1).sum())
        # pairs in order (2 per slot)
        t2=np.sort((ne+no+1)//2)[::-1]; t2=np.concatenate((t2,np.zeros(pad,int))).reshape(-1,32)
        add(side+"_pairs",2*32*t2.max(1).sum())
for k in sorted(tot): print(k, tot[k], "fill %.3f"%(tot[k[0]+"_live"]/tot[k]) if not k.endswith("live") else "")
print("---- half-lines / slices of 16")
tot={}
for g in sel:
    m,nl,e=int(M[g]),int(NL[g]),int(E[g])
    if m<2 or e==0: continue
    c=w.cidx[off[g]:off[g]+2*e].view(np.uint16).astype(np.int64)
    lj=c>>1; par=c&1
    re=w.row_end[w.grp_row_ptr[g]:w.grp_row_ptr[g+1]].astype(np.int64)
    rows=np.repeat(np.arange(m),np.diff(np.concatenate(([0],re))))
    for side,(line,parity,n) in {"R":(rows,par,m),"C":(lj,rows&1,nl)}.items():
        ne=np.bincount(line[parity==0],minlength=n); no=np.bincount(line[parity==1],minlength=n)
        add(side+"_live",e)
        h=np.sort(np.concatenate((ne,no)))[::-1]; h=h[h>0] if (h>0).any() else h[:1]
        pad=(-len(h))%32; hh=np.concatenate((h,np.zeros(pad,int))).reshape(-1,32)
        add(side+"_half32",32*hh.max(1).sum())
        for W in (16,8):
            L=np.maximum(ne,no); Ls=np.sort(L)[::-1]; pad=(-len(Ls))%W; Ls=np.concatenate((Ls,np.zeros(pad,int))).reshape(-1,W)
            add(side+"_cur%d"%W,2*W*Ls.max(1).sum())
            t=np.sort(ne+no)[::-1]; pad=(-len(t))%W; t=np.concatenate((t,np.zeros(pad,int))).reshape(-1,W)
            add(side+"_single%d"%W,W*t.max(1).sum())
for k in sorted(tot): print(k, tot[k], "fill %.3f"%(tot[k[0]+"_live"]/tot[k]) if not k.endswith("live") else "")
