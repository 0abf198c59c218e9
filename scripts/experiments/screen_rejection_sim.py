"""Monte-Carlo sizing of the popcount screen (segalign_b200/csrc/screen_bound.h): fraction of RANDOM seed hits
the screen decides, for different window lengths / block sizes / class-score sets, under the default matrix,
xdrop 910, hspthresh 3000, seed 12of19 with transitions (12 of 13 hits carry one transition inside the seed).
Pure numpy; run: python scripts/experiments/screen_rejection_sim.py"""
import numpy as np
rng = np.random.default_rng(1)
N = 400000
W = 256  # cells per side simulated
M = np.array([[91,-114,-31,-123],[-114,100,-125,-31],[-31,-125,100,-114],[-123,-31,-114,91]])
X = 910; THR = 3000
shape = "TTT0T00TT00T0T0TTTT"  # seed cells, leftmost first; left walk goes from the last char backwards
care = [i for i,c in enumerate(shape) if c=='T']
# right side
rr = rng.integers(0,4,(N,W)); qr = rng.integers(0,4,(N,W))
# left side: cell k=1.. : k=1 is the seed's last base
rl = rng.integers(0,4,(N,W)); ql = rng.integers(0,4,(N,W))
# seed: positions 0..18 in seed coords correspond to left cell k = 19 - pos
var = rng.integers(0,13,N)  # 0 = exact, 1..12 = transition at care index var-1
for ci,pos in enumerate(care):
    k = 19 - pos - 1  # index into left arrays (k-1)
    ql[:,k] = rl[:,k]
    sel = var == ci+1
    ql[sel,k] = rl[sel,k] ^ 2
def walk(r,q, init_pos_zero):
    sc = M[r,q]
    ps = np.cumsum(sc,1)
    rm = np.maximum.accumulate(np.maximum(ps,0),1)
    # M before cell k (strict): max(0, ps[:k])  ; drop when M_k - s_k > X where M_k includes cell k (s_k<=M_k)
    drop = (rm - ps) > X
    first = np.where(drop.any(1), drop.argmax(1), W)
    idx = np.arange(W)[None,:]
    valid = idx < first[:,None]
    best = np.where(valid, ps, -10**9).max(1)
    return np.maximum(best,0), first, sc
Rs, Rf, scR = walk(rr,qr,False)
Ls, Lf, scL = walk(rl,ql,True)
tot = Rs+Ls
print("true: mean total", tot.mean(), "P(tot>=3000)", (tot>=THR).mean(), "mean term R", Rf.mean(), "L", Lf.mean())
print("P(Rf<=32)",(Rf<32).mean(),"P(Rf<64)",(Rf<64).mean(),"P(Lf<64)",(Lf<64).mean(),"P(Lf<96)",(Lf<96).mean(), "both<64", ((Rf<64)&(Lf<64)).mean(), "R<64&L<96", ((Rf<64)&(Lf<96)).mean())

def screen(sc, r, q, win, blk, up):
    # class bounds
    x = r ^ q
    m = (x==0); ts = (x==2); tv = (x==1)|(x==3)
    if up=='class3':
        hi = np.where(m,100,np.where(ts,-31,-114)); lo = np.where(m,91,np.where(ts,-31,-125))
    else:
        hi = sc; lo = sc
    nb = win//blk
    hi = hi[:,:win].reshape(N,nb,blk); lo = lo[:,:win].reshape(N,nb,blk)
    mm = m[:,:win].reshape(N,nb,blk).sum(2)
    Bhi = hi.sum(2); Blo = lo.sum(2)
    Phi = np.cumsum(Bhi,1); Plo = np.cumsum(Blo,1)
    Pprev = Phi - Bhi
    Mhat = np.maximum((Pprev + 100*mm).max(1),0)
    LM = np.maximum.accumulate(np.maximum(Plo,0),1)
    proven = ((LM - Phi) > X)
    # bound: only blocks up to first proven
    first = np.where(proven.any(1), proven.argmax(1), nb)
    idx = np.arange(nb)[None,:]
    Mhat2 = np.maximum(np.where(idx<=first[:,None], Pprev+100*mm, -10**9).max(1),0)
    return proven.any(1), Mhat2
for (wr,wl,blk,up) in [(64,64,8,'exact'),(64,64,16,'exact'),(64,64,8,'class3'),(64,64,16,'class3'),(64,96,8,'class3'),(64,96,16,'class3'),(96,96,16,'class3'),(64,128,16,'class3'),(32,64,8,'class3'),(64,96,32,'class3')]:
    pr, ur = screen(scR, rr, qr, wr, blk, up)
    pl, ul = screen(scL, rl, ql, wl, blk, up)
    assert (ur[pr] >= Rs[pr]).all() and (ul[pl]>=Ls[pl]).all()
    rej = pr & pl & (ur+ul < THR)
    print(wr,wl,blk,up,"provenR",pr.mean(),"provenL",pl.mean(),"reject",rej.mean(), "bound>=thr given proven", ((ur+ul>=THR)&pr&pl).mean())

def screen_mixed(r, q, sizes):
    x = r ^ q
    m = (x==0); ts = (x==2)
    hi = np.where(m,100,np.where(ts,-31,-114)); lo = np.where(m,91,np.where(ts,-31,-125))
    pos=0; Phi=np.zeros(N,dtype=np.int64); Plo=np.zeros(N,dtype=np.int64); LM=np.zeros(N,dtype=np.int64); Mhat=np.zeros(N,dtype=np.int64); proven=np.zeros(N,bool)
    for b in sizes:
        mm = m[:,pos:pos+b].sum(1)
        Mhat=np.maximum(Mhat, Phi+100*mm)
        Phi=Phi+hi[:,pos:pos+b].sum(1); Plo=Plo+lo[:,pos:pos+b].sum(1)
        LM=np.maximum(LM,Plo)
        proven|=(LM-Phi)>X
        pos+=b
    return proven, Mhat
for (rs,ls) in [((16,16,16,16),(16,16,16,16,16,16)),((16,16,32),(16,16,32,32)),((16,16,32),(16,16,16,16,32)),((16,16,16,16),(16,16,16,16,32)),((16,16,32),(16,16,16,16,16,16)),((16,16,32),(32,16,16,32)), ((8,8,16,32),(16,16,16,16,32)),((16,16,32,32),(16,16,16,16,32,32))]:
    pr,ur=screen_mixed(rr,qr,rs); pl,ul=screen_mixed(rl,ql,ls)
    assert (ur[pr] >= Rs[pr]).all() and (ul[pl]>=Ls[pl]).all()
    rej = pr & pl & (ur+ul < THR)
    print(rs,ls,"steps",len(rs)+len(ls),"reject",round(rej.mean(),4),"provenR",round(pr.mean(),4),"provenL",round(pl.mean(),4))
