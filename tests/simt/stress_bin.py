"""Randomised runs of the fused binning kernel's source on CPU fibers against the oracle (not collected by pytest).

usage: python tests/simt/stress_bin.py <seed> <seconds>.  Random mesh size, box, kmin / dk / kmax, auto / cross,
interlaced, compensated, number of CTAs; mode counts and edges must be bit-identical."""
import sys, time
import os
HERE=os.path.dirname(os.path.abspath(__file__)); ROOT=os.path.dirname(os.path.dirname(HERE))
for p_ in (ROOT, os.path.join(ROOT,'tests'), HERE): sys.path.insert(0,p_)
import numpy as np, ctypes as ct
import build_simt
from oracle import pk_oracle as o
import test_simt_bin_power as T
from astrild_b200 import tables
lib=ct.CDLL(build_simt.build_bin())
lib.simt_bin_power.restype = ct.c_int
lib.simt_bin_power.argtypes = [ct.c_void_p] * 4 + [ct.c_int] * 3 + [ct.c_void_p] * 5 + [ct.c_int] + [ct.c_void_p] * 6 + [ct.c_int] * 3 + [ct.c_void_p] * 4
rng=np.random.default_rng(int(sys.argv[1])); t0=time.time(); it=0; nfail=0
orig_edges=tables.k_edges
while time.time()-t0 < float(sys.argv[2]):
    it+=1
    N=int(rng.integers(4,41)); L=float(rng.choice([1.0,300.0,1000.0,2*np.pi, 123.456]))
    kf=2*np.pi/L
    kmin=float(rng.choice([0.0,kf,0.5*kf,2.3*kf]))
    dk=None if rng.integers(0,2) else float(rng.choice([kf,0.5*kf,1.7*kf,3*kf]))
    kmax=None if rng.integers(0,2) else float(rng.uniform(2,N)*kf)
    edges=orig_edges(N,L,kmin,dk,kmax)
    if len(edges)<2: continue
    tables.k_edges=lambda N_,L_,kmin_=0.0,dk_=None,kmax_=None,_e=edges:_e   # the harness helper calls tables.k_edges(N, L, kmin)
    shape=(N,N,N//2+1)
    mk=lambda:(rng.normal(size=shape)+1j*rng.normal(size=shape)).astype(np.complex64)
    inter=bool(rng.integers(0,2)); cross=bool(rng.integers(0,2)); comp=bool(rng.integers(0,2))
    res=str(rng.choice(['cic','tsc']))
    c1=mk(); c1s=mk() if inter else None; c2=mk() if cross else None; c2s=mk() if (cross and inter) else None
    a=c1.astype(np.complex128); b=None if c2 is None else c2.astype(np.complex128)
    if inter:
        a=o.interlace_combine(a,c1s.astype(np.complex128),N,L)
        if cross: b=o.interlace_combine(b,c2s.astype(np.complex128),N,L)
    if comp:
        a=o.compensate(a,res,inter,N)
        if cross: b=o.compensate(b,res,inter,N)
    want=o.fftpower_1d(a,b,N,L,kmin=kmin,dk=dk,kmax=kmax)
    got=T.bin_power(lib,N,L,c1,c1s,c2,c2s,kmin=kmin,compensation=(res,inter) if comp else None,ctas=int(rng.integers(1,5)))
    ok=np.array_equal(got['Nsum'],want['Nsum']) and np.array_equal(got['edges'],want['edges'])
    sc=np.abs(want['power'][np.isfinite(want['power'])]).max() if np.isfinite(want['power']).any() else 1.0
    m=np.isfinite(want['power'])
    ok=ok and np.allclose(got['power'][m],want['power'][m],rtol=0,atol=3e-5*sc) and np.allclose(got['k'][m],want['k'][m],rtol=1e-11)
    if not ok:
        nfail+=1; print('FAIL',dict(N=N,L=L,kmin=kmin,dk=dk,kmax=kmax,inter=inter,cross=cross,comp=comp,res=res), 'modes eq',np.array_equal(got['Nsum'],want['Nsum']))
tables.k_edges=orig_edges
print('iterations',it,'failures',nfail)
