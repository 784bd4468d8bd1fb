"""One-off campaign: the three oracles against the reference's own compiled sources (oracle/_ref) on a few thousand
cases beyond what tests/test_reference_pins.py runs.  Last run: solver 1536/1536 solves bit-identical (horizons 30-200,
both roads, 2 lambda-overflow exits), corridor 5592/5592 knots, DP planner 60/60 scenes (8 failed plans included)."""
import sys, time, numpy as np, collections
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cilqr_b200 import scenarios
from oracle import binding as orc, ref_binding as ref, dp_binding as dp, corridor_binding as cb
t0=time.time()
# solver: 1536 scenarios
c=collections.Counter()
for (seed,B,N,road) in ((101,512,30,'shipped'),(102,384,50,'gentle'),(103,256,80,'shipped'),(104,256,100,'gentle'),(105,128,200,'gentle')):
    batch=scenarios.generate(seed,0,B,N=N,road_name=road)
    X,U,S,conv=orc.solve_batch(batch,nthreads=8)
    for b in range(B):
        r=ref.ilqr_solve(batch,b)
        same=np.array_equal(r['states'],X[b]) and np.array_equal(r['controls'],U[b])
        c[(int(S[b,0]),same)]+=1
print('solver',sorted(c.items()),'%.0fs'%(time.time()-t0))
# corridor: ~6000 knots
t0=time.time(); bad=0; tot=0
for (seed,B,N,n_obs) in ((201,24,100,20),(202,24,80,11),(203,24,50,3)):
    _,ci=scenarios.generate_with_obstacles(seed,0,B,N=N,n_obs=n_obs)
    cor,cnt,poly,code=cb.plan_batch(ci.traj,ci.obs_points,ci.obs_cnt,64)
    for b in range(ci.B):
        for k in range(ci.K):
            n=int(ci.obs_cnt[b,k]); m,cons,pl=ref.build_corridor(*ci.traj[b,k],ci.obs_points[b,k,:n],cap=64)
            tot+=1; bad+= not (m==cnt[b,k] and np.array_equal(cons,cor[b,k,:m]) and np.array_equal(pl,poly[b,k,:m]))
print('corridor knots',tot,'mismatches',bad,'%.0fs'%(time.time()-t0))
# dp: 60 scenes
t0=time.time(); bad=0; oks=0
for (seed,n_obs) in ((301,11),(302,5),(303,16)):
    db=scenarios.generate_dp(seed,20,n_obs=n_obs); bar=dp.build_barrier(db.ref)
    for b in range(db.B):
        okr,tr=ref.dp_plan(db.ref,db.static_poly[b],db.static_nv[b],db.dyn_time[b],db.dyn_samples[b],db.dyn_poly[b],db.dyn_nv[b],*db.start[b])
        ok,traj,cost,wp=dp.plan(dp.Scene(db.ref,bar,db.static_poly[b],db.static_nv[b],db.dyn_time[b],db.dyn_samples[b],db.dyn_poly[b],db.dyn_nv[b]),*db.start[b])
        bad+= not (ok==okr and np.array_equal(tr,traj[:,:11],equal_nan=True)); oks+=ok
print('dp scenes 60 mismatches',bad,'planned ok',oks,'%.0fs'%(time.time()-t0))
