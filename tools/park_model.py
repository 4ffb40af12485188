"""Offline model (CPU only, test infrastructure: uses the oracle) of `dataflow_park_kernel`: how many of a warp's 32 lanes are
busy when the last `K` unfinished secant solves of a tile are parked after trip >= `kmin` and finished later 32 at a time?

    python tools/park_model.py [n_segments] [nsteps]

Trip counts of every (segment, step) come from the oracle (pinned-pow build); a tile is 32 consecutive positions of a level in
caller order; cost of a trip = 1, of the extra phase A of a first trip = 0.7.  Output of `100000 288` is kept in
profiles/r02_v7_park/park_model_100000x288.txt: 21.8 busy lanes without parking, 27.1 with K = 8 (10 % of the lanes parked)."""
import os, sys, time
import numpy as np
ROOT="/root/repo"
for p in (ROOT, os.path.join(ROOT,"t-route_b200"), os.path.join(ROOT,"tests"), os.path.join(ROOT,"tools")):
    sys.path.insert(0,p)
import trip_order_study as S
import helpers as H
from oracle import oracle as o
from troute_b200 import synth, hostgraph
n=int(sys.argv[1]) if len(sys.argv)>1 else 100000; T=int(sys.argv[2]) if len(sys.argv)>2 else 288
o.build()
down = synth.conus_like(n_total=n, n_basins=max(4, n // 185), seed=16, style="nhd")
case = H.make_case(down, nsteps=T)
fvd,_,ex = H.oracle_route(o, case, False)
trips = S.trip_matrix(o, case, fvd)
level = hostgraph.levels(down, case["up_ptr"]).astype(np.int64)
print("trip histogram", np.bincount(trips.ravel())[:10])

trips=trips.astype(np.int32)
n,T=trips.shape
order=np.lexsort((np.arange(n),level))
lv=level[order]; tr=trips[order]
# tiles: groups of 32 within level (levels >= 64 wide)
starts=np.flatnonzero(np.r_[True,lv[1:]!=lv[:-1]]); ends=np.r_[starts[1:],lv.size]
tiles=[]
for a,b in zip(starts,ends):
    if b-a<64: continue
    x=tr[a:b]; pad=(-x.shape[0])%32
    if pad: x=np.concatenate([x,np.zeros((pad,T),dtype=x.dtype)])
    tiles.append(x.reshape(-1,32,T))
G=np.concatenate(tiles,axis=0)          # [ntiles,32,T]
G=G.transpose(0,2,1).reshape(-1,32)      # tile-steps
G=G[G.max(axis=1)>0]
A=0.7
useful=(G[G>0]+A).sum()
base=((G.max(axis=1)+A)*32).sum()
print("tile-steps",G.shape[0],"baseline eff %.4f lanes %.2f"%(useful/base,32*useful/base))
rng=np.random.default_rng(0)
def park(K,kmin,nwarps=64,repark=False, over=0.08):
    # warps take tile-steps round robin; pool per warp
    cost=0.0
    idx=np.arange(G.shape[0]); 
    mx=G.max(axis=1)
    # trips executed in main loop: smallest k>=kmin with count(trips>k)<=K
    cnt=np.stack([(G>k).sum(axis=1) for k in range(0,12)],axis=1)   # cnt[:,k] = lanes needing more than k trips
    kstop=np.full(G.shape[0],-1)
    for k in range(kmin,12):
        m=(kstop<0)&(cnt[:,k]<=K)
        kstop[m]=k
    kstop[kstop<0]=mx[kstop<0]
    kstop=np.minimum(kstop,np.maximum(mx,1))
    kstop=np.where(mx<kmin,mx,kstop)
    cost_main=((kstop+A)*32).sum()
    rem=np.maximum(G-kstop[:,None],0)        # remaining trips of parked lanes
    nparked=(rem>0).sum()
    # pools: assign tile-steps to warps in chunks (a warp sees a random subsequence): shuffle then split
    perm=rng.permutation(G.shape[0])
    r=rem[perm].ravel(); r=r[r>0]
    # batches of 32 in arrival order
    m=(r.size//32)*32
    b=r[:m].reshape(-1,32)
    cost_batch=(b.max(axis=1)*32).sum()+ (r[m:].max()*32 if r.size>m else 0)
    ovh=over*32*( (rem>0).any(axis=1).sum()*0.5 + b.shape[0])   # park code per tile w/ stragglers + restore per batch, in trip units
    tot=cost_main+cost_batch+ovh
    return useful/tot, nparked/ (G>0).sum(), cost_main/tot, cost_batch/tot
for kmin in (2,3):
    for K in (0,2,4,6,8,12,16):
        e,fp,cm,cb=park(K,kmin)
        print(f"kmin {kmin} K {K:2d}: eff {e:.4f} lanes {32*e:.2f} parked {100*fp:.1f}% main {cm:.2f} batch {cb:.2f}")
