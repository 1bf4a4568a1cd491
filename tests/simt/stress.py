"""Randomised runs of the deposit kernels' source on CPU fibers against the oracle (not collected by pytest).

usage: python tests/simt/stress.py <seed> <seconds>.  Mesh 4-48, particles inside / outside / far outside the box, on exact
cell and half-cell faces, clumped, float32 / float64, AoS / SoA, masses, interlaced pair, slab plans.  This is how the
single-brick twin-wrap defect was found (profiles/r01_session2_measurements.md)."""
import sys, time, ctypes as ct
import os
HERE=os.path.dirname(os.path.abspath(__file__)); ROOT=os.path.dirname(os.path.dirname(HERE))
for p_ in (ROOT, os.path.join(ROOT,'tests'), HERE): sys.path.insert(0,p_)
import numpy as np
import build_simt
from oracle import pk_oracle_fast as f
from test_simt_deposit import load, deposit
lib=load(build_simt.build())
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
nfail=0
t0=time.time(); it=0
while time.time()-t0 < float(sys.argv[2] if len(sys.argv)>2 else 120):
    it+=1
    N=int(rng.integers(4,49))
    L=float(rng.choice([1.0, 250.0, 1000.0, 7.3]))
    npart=int(rng.integers(1,4000))
    kind=rng.integers(0,5)
    if kind==0: pos=rng.random((npart,3))*L
    elif kind==1: pos=rng.random((npart,3))*3*L-L          # up to one box outside
    elif kind==2: pos=(rng.integers(0,2*N+1,(npart,3))*0.5)*L/N   # exactly on cell / half-cell boundaries
    elif kind==3: pos=rng.normal(0.37*L,0.02*L,(npart,3))        # one dense clump
    else: pos=rng.random((npart,3))*9*L-4*L                # far outside (float64 fallback path)
    f64=bool(rng.integers(0,2)) and kind!=2
    pos=pos.astype(np.float64 if f64 else np.float32)
    mass=rng.random(npart).astype(np.float32) if rng.integers(0,2) else None
    res=str(rng.choice(['cic','tsc']))
    pair=bool(rng.integers(0,2))
    shift=0.0 if pair else float(rng.choice([0.0,0.5]))
    soa=bool(rng.integers(0,2))
    slab=bool(rng.integers(0,3)==0) and N>=8 and kind!=2   # exact cell boundaries: ownership is decided by the exact product, not by the float64-rounded one
    x0,n0=(0,N)
    if slab:
        P=int(rng.choice([2,4])); n0=N//P
        if n0*P!=N or n0<2: slab=False; n0=N
        else: x0=int(rng.integers(0,P))*n0
    try:
        a,b=deposit(lib,pos,mass,N,L,res,pair=pair,shift=shift,soa=soa,x0=x0,n0=n0)
    except Exception as e:
        print('EXC',e); nfail+=1; continue
    p=pos
    if slab:
        cell=np.floor(pos[:,0].astype(np.float64)*(N/L)).astype(np.int64)%N   # ownership: unshifted floor
        # use the same float64 expression as the kernel: x*scale with scale=pos_scale*N = N/L ... (1/L)*N
        cell=np.floor(pos[:,0].astype(np.float64)*((1.0/L)*N)).astype(np.int64)%N
        sel=(cell>=x0)&(cell<x0+n0); p=pos[sel]; m=None if mass is None else mass[sel]
    else: m=mass
    for got,sh in ((a,shift),(b,0.5)) if pair else ((a,shift),):
        want=f.paint(p,m,N,L,res,sh) if len(p) else np.zeros((N,N,N))
        if slab:
            idx=[(x0-1+i)%N for i in range(n0+3)]
            # ghost planes may alias owned planes when n0+3 > N: skip those configs
            if n0+3>N: continue
            want=want[idx]
        tol=6e-6*max(want.max(),1e-30) + 1e-30
        err=np.abs(got-want).max()
        if not (err<=tol) or np.isnan(got).any():
            nfail+=1
            print('FAIL',dict(N=N,L=L,npart=npart,kind=int(kind),f64=f64,mass=mass is not None,res=res,pair=pair,shift=sh,soa=soa,slab=slab,x0=x0,n0=n0),'err',err,'tol',tol,'sum',got.sum(),want.sum())
print('iterations',it,'failures',nfail)
