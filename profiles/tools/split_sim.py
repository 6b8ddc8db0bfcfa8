# how do the lanes of one 32-ray bundle spread over RK4 stage slots when they cross z faces?
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from synthpy_b200 import field_generator as fg
from scipy.interpolate import RegularGridInterpolator as RGI
n=int(sys.argv[1]) if len(sys.argv)>1 else 128
ne = fg.turbulent_ne(n // 2, device="cpu", seed=1, noise="numpy").numpy()
print(ne.shape, ne.min(), ne.max())
c=299792458.0; lwl=1064e-9; om=2*np.pi*c/lwl; nc=3.14207787e-4*om**2
L=np.array([5e-3,5e-3,10e-3]); ax=[np.float32(np.linspace(-l,l,n)).astype(np.float64) for l in L]
nn=ne/nc
g=[ -0.5*c*c*np.gradient(nn,ax[a],axis=a) for a in range(3)]
I=[RGI(tuple(ax),gg,bounds_error=False,fill_value=0.0) for gg in g]
def acc(p): return np.stack([i(p) for i in I],1)
dz=2*L[2]/(n-1); h=0.5*dz/c
rng=np.random.default_rng(0)
dx=2*L[0]/(n-1)
# 32 rays within one cell column
ix,iy=n//2+3,n//2-5
p=np.zeros((32,3)); p[:,0]=ax[0][ix]+dx*rng.random(32); p[:,1]=ax[1][iy]+dx*rng.random(32); p[:,2]=-L[2]
v=np.zeros((32,3)); v[:,2]=c
slots=[]  # per (step,stage) -> set of lanes that entered a new z cell
cell=np.searchsorted(ax[2],p[:,2],side='right')-1
events={}
off=float(sys.argv[2]) if len(sys.argv)>2 else 0.0
H=h
for s in range(2*(n-1)+1):
    h=H*off*2 if s==0 else H
    if h==0: continue
    k1=acc(p); p2=p+0.5*h*v; 
    k2=acc(p2); p3=p+0.5*h*v+0.25*h*h*k1
    k3=acc(p3); p4=p+h*v+0.5*h*h*k2
    k4=acc(p4)
    for st,pp in enumerate((p,p2,p3,p4)):
        cz=np.searchsorted(ax[2],pp[:,2],side='right')-1
        ch=np.nonzero(cz!=cell)[0]
        if len(ch): events[(s,st)]=len(ch)
        cell=cz
    p=p+h*v+h*h/6*(k1+k2+k3); v=v+h/6*(k1+2*k2+2*k3+k4)
ev=np.array(list(events.values()))
print('reload slots',len(ev),'per cell',len(ev)/(n-1),'avg lanes',ev.mean())
print('theta', np.degrees(np.arctan(v[:,0]/v[:,2])).std()*17.45,'mrad rms x;  |v|/c-1 range',(np.linalg.norm(v,axis=1)/c-1).min(),(np.linalg.norm(v,axis=1)/c-1).max())
import collections
print(collections.Counter(k[1] for k in events))
print(sorted(events.items())[:40])
it=sorted(events.items())
print(it[300:340]); print(it[700:740])
print('lead (cells) per lane at end:', np.round((p[:,2]-(-L[2]+2*(n-1)*0.5*dz))/dz,4))
