"""Randomised runs of the slab routing kernels' source on CPU fibers against NumPy (not collected by pytest).

usage: python tests/simt/stress_route.py <seed> <seconds>."""
import sys, time
import os
HERE=os.path.dirname(os.path.abspath(__file__)); ROOT=os.path.dirname(os.path.dirname(HERE))
for p_ in (ROOT, os.path.join(ROOT,'tests'), HERE): sys.path.insert(0,p_)
import numpy as np, ctypes as ct
import build_simt
vp,i32,i64,f64=ct.c_void_p,ct.c_int,ct.c_longlong,ct.c_double
lib=ct.CDLL(build_simt.build_misc())
lib.simt_route.argtypes=[vp,vp,vp,vp,i64,i32,f64,i32,i32,vp,i64,vp,vp,i32]
ptr=lambda a: None if a is None else a.ctypes.data_as(vp)
rng=np.random.default_rng(int(sys.argv[1])); t0=time.time(); it=0; nfail=0
while time.time()-t0<float(sys.argv[2]):
    it+=1
    P=int(rng.choice([2,4,8])); n0=int(rng.integers(2,9)); N=P*n0; rank=int(rng.integers(0,P)); L=float(rng.choice([1.0,1000.0,7.3]))
    n=int(rng.integers(1,6000)); kind=rng.integers(0,4)
    if kind==0: x=rng.random(n)*L
    elif kind==1: x=rng.random(n)*3*L-L
    elif kind==2: x=rng.normal((rank+0.5)/P*L,0.1*L,n)
    else: x=rng.random(n)*9*L-4*L
    x=x.astype(np.float32); y=rng.random(n).astype(np.float32); z=rng.random(n).astype(np.float32)
    scale=(1.0/L)*N
    g=x.astype(np.float64)*scale
    # skip particles whose product is within 1e-9 of an integer (exact vs rounded product may disagree there)
    amb=np.abs(g-np.round(g))<1e-9
    if amb.any(): continue
    cell=np.floor(g).astype(np.int64)%N; dest=cell//n0; leaving=dest!=rank
    cap=int(leaving.sum())+7
    counts=np.zeros(2*P,np.uint64); out=np.full((cap,3),np.nan,np.float32)
    lib.simt_route(ptr(x),ptr(y),ptr(z),None,n,N,1.0/L,P,rank*n0,ptr(counts),cap,ptr(out),None,2)
    want=np.bincount(dest[leaving],minlength=P)
    ok=np.array_equal(counts[:P].astype(np.int64),want)
    if ok:
        s=0
        for d in range(P):
            k=int(want[d]); got=out[s:s+k]; sel=leaving&(dest==d); w=np.stack([x[sel],y[sel],z[sel]],1)
            ok=ok and np.array_equal(got[np.lexsort(got.T[::-1])], w[np.lexsort(w.T[::-1])]); s+=k
    if not ok: nfail+=1; print('FAIL',dict(P=P,n0=n0,rank=rank,L=L,n=n,kind=int(kind)))
print('iterations',it,'failures',nfail)
