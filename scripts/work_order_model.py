"""Work-order model of the thread-per-column CAPE kernel from the CPU oracle Brent trace: lock-step cost of 32-column
warps in model order vs bucketed by parcel level count (global / block-local sort), the longest warp chain, and an
upper bound for a secondary key.  CPU only (the oracle is test infrastructure)."""
import sys, os, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cam_nor_physics_b200 import soundings as S
from helpers import get_oracle
ncols = 4096
o, p, rc = get_oracle("pm", 16, 32)
ch = S.make_chunks(ncols, 32, 16, p_conv=0.35)
cap = 4 * 40_000_000 // 10
buf = np.zeros(cap, np.int32)
o.lib.zmo_trace_set(buf.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(cap))
ref = o.convr_batch(ch, nthreads=1)
n = o.lib.zmo_trace_count()
o.lib.zmo_trace_set(None, 0)
tr = buf[:4 * n].reshape(n, 4)
from collections import defaultdict
passes = {1: defaultdict(lambda: {2: [], 3: [], 4: []}), 2: defaultdict(lambda: {2: [], 3: [], 4: []})}
cur_pass = {}; seen4 = {}
for rcall, icol, lchnk, ev in tr:
    pz = cur_pass.get(lchnk, 1)
    if rcall == 2 and seen4.get(lchnk, False):
        pz = 2; cur_pass[lchnk] = 2; seen4[lchnk] = False
    if rcall == 4: seen4[lchnk] = True
    passes[pz][(lchnk, icol)][rcall].append(ev)
for pz in (1,2):
    cols = sorted(passes[pz])
    G = [passes[pz][c] for c in cols]
    nA = np.array([len(g[2]) for g in G]); nB = np.array([len(g[4])//2 for g in G]); nL=np.array([len(g[3]) for g in G])
    evA = np.array([sum(g[2]) for g in G]); evB = np.array([sum(g[4]) for g in G]); evL=np.array([sum(g[3]) for g in G])
    print("pass",pz,"cols",len(G))
    print(" levels A: mean %.1f min %d max %d ; levels B: mean %.1f min %d max %d; lcl inv mean %.2f"%(nA.mean(),nA.min(),nA.max(),nB.mean(),nB.min(),nB.max(),nL.mean()))
    print(" evals A mean %.1f  B mean %.1f  L mean %.1f"%(evA.mean(),evB.mean(),evL.mean()))
    print(" hist nB:", np.bincount(nB))
    def lock(groups, key):
        tot=0; mean=0
        for grp in groups:
            nlev = max(len(g[key])//(2 if key==4 else 1) for g in grp)
            for j in range(nlev):
                if key==2:
                    a=np.array([g[2][j] if j<len(g[2]) else 0 for g in grp]); tot+=a.max()+ (1); mean+=a.sum()/32+ (a>0).sum()/32
                else:
                    b=np.array([g[4][2*j] if 2*j<len(g[4]) else 0 for g in grp]); c=np.array([g[4][2*j+1] if 2*j+1<len(g[4]) else 0 for g in grp])
                    tot+=b.max()+c.max()+1; mean+=(b.sum()+c.sum())/32+(b>0).sum()/32
        return tot, mean
    warps=[G[i:i+32] for i in range(0,len(G),32)]
    tA,mA=lock(warps,2); tB,mB=lock(warps,4)
    print(" unsorted: A lock %.0f eff %.2f ; B lock %.0f eff %.2f ; total eff %.2f"%(tA,mA/tA,tB,mB/tB,(mA+mB)/(tA+tB)))
    order=np.argsort(nB,kind='stable'); Gs=[G[i] for i in order]
    warps=[Gs[i:i+32] for i in range(0,len(Gs),32)]
    tB2,mB2=lock(warps,4)
    print(" B sorted by nB: lock %.0f eff %.2f ; total eff %.2f ; time ratio %.3f"%(tB2,mB2/tB2,(mA+mB2)/(tA+tB2),(tA+tB2)/(tA+tB)))
    # also sort by evB (oracle knowledge; upper bound)
    order=np.argsort(evB,kind='stable'); Gs=[G[i] for i in order]
    warps=[Gs[i:i+32] for i in range(0,len(Gs),32)]
    tB3,mB3=lock(warps,4)
    print(" B sorted by evB (bound): lock %.0f eff %.2f ; time ratio %.3f"%(tB3,mB3/tB3,(tA+tB3)/(tA+tB)))
print("==== sort by level count (known before the sweep)")
cols = sorted(passes[1]); G=[passes[1][c] for c in cols]
nA = np.array([len(g[2]) for g in G])
def total(Gx):
    warps=[Gx[i:i+32] for i in range(0,len(Gx),32)]
    tA,mA=lock(warps,2); tB,mB=lock(warps,4)
    return tA+tB, mA+mB, tA, tB
base=total(G)
print("unsorted total %.0f eff %.3f"%(base[0],base[1]/base[0]))
order=np.argsort(nA,kind='stable'); t=total([G[i] for i in order]); print("global sort: total %.0f eff %.3f ratio %.3f (A %.3f B %.3f)"%(t[0],t[1]/t[0],t[0]/base[0],t[2]/base[2],t[3]/base[3]))
for blk in (128,256,512,1024):
    o2=[]
    for b0 in range(0,len(G),blk):
        idx=np.arange(b0,min(b0+blk,len(G))); o2+=list(idx[np.argsort(nA[idx],kind='stable')])
    t=total([G[i] for i in o2]); print("block-local sort %d: total %.0f eff %.3f ratio %.3f"%(blk,t[0],t[1]/t[0],t[0]/base[0]))
print("==== align by absolute level k (lanes with a lower launch level idle first), no sort")
def lock_k(groups):
    tot=0; mean=0
    for grp in groups:
        nlev=max(len(g[2]) for g in grp)
        for j in range(1,nlev+1):   # j-th level from the top end
            a=np.array([g[2][-j] if j<=len(g[2]) else 0 for g in grp])
            b=np.array([g[4][-2*j] if 2*j<=len(g[4]) else 0 for g in grp]); c=np.array([g[4][-2*j+1] if 2*j<=len(g[4]) else 0 for g in grp])
            tot+=a.max()+b.max()+c.max()+2; mean+=(a.sum()+b.sum()+c.sum())/32+2*(a>0).sum()/32
    return tot,mean
warps=[G[i:i+32] for i in range(0,len(G),32)]
t,m=lock_k(warps); print("k-aligned unsorted: total %.0f eff %.3f ratio vs base %.3f"%(t,m/t,t/base[0]))
order=np.argsort(nA,kind='stable'); Gs=[G[i] for i in order]
warps=[Gs[i:i+32] for i in range(0,len(Gs),32)]
t,m=lock_k(warps); print("k-aligned global sort: total %.0f eff %.3f ratio vs base %.3f"%(t,m/t,t/base[0]))
print("==== per-warp lockstep chain lengths (evaluations): the single wave ends with the longest warp")
def chains(Gx):
    out=[]
    for i in range(0,len(Gx),32):
        grp=Gx[i:i+32]; tA,_=lock([grp],2); tB,_=lock([grp],4); out.append(tA+tB)
    return np.array(out)
c0=chains(G); order=np.argsort(nA,kind='stable'); c1=chains([G[i] for i in order])
print("unsorted: mean %.0f max %.0f p90 %.0f"%(c0.mean(),c0.max(),np.percentile(c0,90)))
print("sorted  : mean %.0f max %.0f p90 %.0f min %.0f"%(c1.mean(),c1.max(),np.percentile(c1,90),c1.min()))
# per-lane totals (free-running bound)
lt=np.array([sum(g[2])+sum(g[4])+2*len(g[2]) for g in G]); print("per-lane totals: mean %.0f max %.0f p99 %.0f"%(lt.mean(),lt.max(),np.percentile(lt,99)))
print("==== secondary keys")
tot_ev=np.array([sum(g[2])+sum(g[4]) for g in G])
def report(name, order):
    Gx=[G[i] for i in order]; c=chains(Gx); t=total(Gx)
    print("%-28s total %.0f ratio %.3f ; chain max %.0f mean %.0f"%(name,t[0],t[0]/base[0],c.max(),c.mean()))
report("primary only", np.argsort(-nA,kind='stable'))
report("primary + total evals (bound)", np.lexsort((tot_ev,-nA)))
