"""Differential fuzzing of the ORACLE against the unmodified reference (build container only: imports /root/reference
through oracle/refstubs, like tests/golden/make_golden.py; nothing here runs on the GPU box or in the test suite).

    python tools/fuzz_oracle_vs_reference.py [seed ...]

Three sweeps over adversarial inputs (exact and near ties, single walkers, mirror-image groups, massive events):
  distit      oracle.distit vs DistIt.run (distance_descriptors.py:177-213) -- random molecules of 3-16 atoms, random sorted
              sub-lists / groups / method / matrix form, near-symmetric structures with noise 0, 1e-12, 1e-6, 0.1.
              Mismatches are expected ONLY for an exact tie of column norms inside sorted_atoms: the reference's order is
              then whatever NumPy's SIMD argsort does on this CPU (known gaps); they are counted separately.
  continuous  oracle.branch_continuous vs DMC_Sim.birth_or_death (pyvibdmc.py:432-454, :340-356)
  discrete    oracle.birth_or_death_discrete vs DMC_Sim.birth_or_death (pyvibdmc.py:391-431), error text included
This sweep is what found the summation-order dependence of sort_groups (:146) that tests/test_descriptors.py now pins."""
import os
import shutil
import sys
import warnings

import numpy as np

warnings.simplefilter('ignore')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
import make_golden as G                                            # noqa: E402  (puts the reference on sys.path)
from oracle import dmc_oracle as O                                 # noqa: E402
from pyvibdmc.simulation_utilities.tensorflow_descriptors.distance_descriptors import DistIt   # noqa: E402


def fuzz_distit(seed):
    rng = np.random.default_rng(seed)
    bad=0; tot=0; ties=0
    for it in range(150):
        na=int(rng.integers(3,17))
        # group layout
        sg=None; sa=None
        if rng.random()<0.7:
            ng=int(rng.integers(2,min(5,na)+0)); gs=int(rng.integers(1,max(na//ng,1)+1))
            if ng*gs<=na and gs>=1:
                perm=rng.permutation(na)[:ng*gs]
                sg=[sorted(perm[i*gs:(i+1)*gs].tolist()) for i in range(ng)] if rng.random()<0.5 else [perm[i*gs:(i+1)*gs].tolist() for i in range(ng)]
        if rng.random()<0.6:
            perm=rng.permutation(na).tolist(); sa=[]; i=0
            while i<na:
                l=int(rng.integers(1,5)); sa.append(sorted(perm[i:i+l])); i+=l
        method=['distance','coulomb','spf'][int(rng.integers(3))]
        full=bool(rng.random()<0.5)
        n=int(rng.choice([1,2,5,40]))
        # near-symmetric structure: mirror pairs + tiny noise
        base=rng.normal(0,2.0,size=(na,3))
        if rng.random()<0.6:
            h=na//2; base[h:2*h]=base[:h][rng.permutation(h)]*np.array([-1,1,1])
            base[:,0]+=np.where(np.arange(na)<h,0.7,0); base[h:2*h,0]-=0.7
        noise=rng.choice([0,1e-12,1e-6,0.1])
        cds=base[None]+rng.normal(0,1,size=(n,na,3))*noise
        # avoid coincident atoms
        zs=rng.choice([1,6,8],size=na).tolist()
        eqn=rng.choice([0,1e-3])                     # 0: the symmetric structure itself is the equilibrium structure
        eq=base+rng.normal(0,1,size=base.shape)*eqn if method=='spf' else None
        kw=dict(method=method,sorted_atoms=sa,sorted_groups=sg,full_mat=full)
        try:
            r=np.asarray(DistIt(zs,force_numpy=True,eq_xyz=eq,**kw).run(cds))
        except Exception as e:
            continue
        try:
            a=O.distit(cds,zs,eq_xyz=eq,**kw)
        except Exception as e:
            print("oracle raised",type(e).__name__,e,na,kw,n); bad+=1; continue
        tot+=1
        if not (a.shape==r.shape and np.array_equal(a,r,equal_nan=True)):
            if sa is not None and (noise == 0 or (method == 'spf' and eqn == 0)):      # exact tie of column norms inside sorted_atoms (walkers or eq_xyz): platform dependent
                ties += 1
                continue
            bad+=1
            print("MISMATCH",na,n,noise,kw, int((a!=r).sum()))
    print(f"  ({ties} exact-tie atom sorts set aside)")
    return tot, bad


def fuzz_continuous(seed):
    rng = np.random.default_rng(seed)
    bad=0; tot=0
    for it in range(120):
        n=int(rng.choice([1,2,3,7,64,500,1500]))
        mode=int(rng.integers(5))
        thresh=[None,[0.3],[0.05,3.0],[0.5,1.5],[0.9]][int(rng.integers(5))]
        sim=G.make_sim('continuous', n, ["H","H","O"], np.tile(G.EQ,(n,1,1)), dt=5.0, cont_wt_thresh=thresh)
        if mode==0: w0=np.exp(rng.normal(0,2.5,size=n))
        elif mode==1: w0=np.round(np.exp(rng.normal(0,1,size=n))*4)/4+0.25
        elif mode==2: w0=np.full(n,1.0); w0[rng.random(n)<0.6]=1e-6   # more kills than big donors
        elif mode==3: w0=rng.choice([0.01,0.5,1.0,2.0,8.0],size=n)
        else: w0=np.exp(rng.normal(0,4,size=n))
        v=np.full(n,0.021) if mode in (1,2,3) and rng.random()<0.7 else 0.021+0.004*rng.standard_normal(n)
        sim._cont_wts=w0.copy(); sim._walker_pots=v.copy(); sim._vref=0.021
        sim._walker_coords=np.arange(n*9,dtype=float).reshape(n,3,3); sim._who_from=np.arange(n); sim._desc_wt=True
        try:
            nb,mx,mn=sim.birth_or_death(); rerr=None
        except Exception as e:
            rerr=e
        lower=sim._thresh_lower; upper=sim._thresh_upper
        try:
            w,src,nb2,mx2,mn2=O.branch_continuous(w0,v,0.021,5.0,lower,upper); oerr=None
        except Exception as e:
            oerr=e
        tot+=1
        if (rerr is None)!=(oerr is None):
            bad+=1; print("ERR MISMATCH",n,mode,thresh,repr(rerr),repr(oerr)); continue
        if rerr is not None: 
            if type(rerr)!=type(oerr): print("different exception types",repr(rerr),repr(oerr)); bad+=1
            continue
        if upper is None:
            ok=np.array_equal(w,sim._cont_wts) and np.array_equal(src,sim._who_from) and [nb,mx,mn]==[nb2,mx2,mn2]
        else:
            ok=np.array_equal(np.sort(w),np.sort(sim._cont_wts)) and nb==nb2
        if not ok:
            bad+=1; print("MISMATCH",n,mode,thresh,(nb,mx,mn),(nb2,mx2,mn2), int((w!=sim._cont_wts).sum()), int((src!=sim._who_from).sum()))
    return tot, bad


def fuzz_discrete(seed):
    rng = np.random.default_rng(seed)
    bad=0; tot=0
    for it in range(150):
        n=int(rng.choice([1,2,3,50,700]))
        dt=float(rng.choice([1.0,5.0,60.0]))
        spread=float(rng.choice([0.0,0.004,0.02,0.1,0.5]))
        sim=G.make_sim('discrete', n, ["H","H","O"], np.tile(G.EQ,(n,1,1)), dt=dt)
        sim._walker_coords=np.arange(n*9,dtype=float).reshape(n,3,3)
        v=0.021+spread*rng.standard_gamma(2.0,size=n)-spread*2
        if rng.random()<0.2: v[rng.integers(n)]-=rng.choice([0.2,1.0,10.0])
        sim._walker_pots=v.copy(); sim._vref=float(np.mean(v))+float(rng.choice([0,0.001,-0.001,0.05,-0.05]))
        sim._who_from=np.arange(n); sim._desc_wt=True
        seed=int(rng.integers(1<<30)); np.random.seed(seed); u=np.random.random(n); np.random.seed(seed)
        if rng.random()<0.2: pass
        try:
            b,d,p=sim.birth_or_death(); rerr=None
        except Exception as e: rerr=e
        try:
            out=O.birth_or_death_discrete(v,sim._vref,dt,u,n); oerr=None
        except Exception as e: oerr=e
        tot+=1
        if (rerr is None)!=(oerr is None) or (rerr is not None and (type(rerr)!=type(oerr) or str(rerr)!=str(oerr))):
            bad+=1; print("ERR MISMATCH",n,dt,spread,repr(rerr),repr(oerr)); continue
        if rerr is not None: continue
        _,idx,b2,d2,p2=out
        if not (np.array_equal(idx,sim._who_from) and (b,d)==(b2,d2)):
            bad+=1; print("MISMATCH",n,dt,spread,(b,d,p),(b2,d2,p2))
    return tot, bad


if __name__ == "__main__":
    try:
        for seed in [int(a) for a in sys.argv[1:]] or [1]:
            for name, fn in (("distit", fuzz_distit), ("continuous", fuzz_continuous), ("discrete", fuzz_discrete)):
                tot, bad = fn(seed)
                print(f"seed {seed} {name}: {tot} cases, {bad} mismatches")
    finally:
        shutil.rmtree(G.TMP, ignore_errors=True)
