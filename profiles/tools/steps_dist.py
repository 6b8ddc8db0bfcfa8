import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench
from synthpy_b200 import engine as E, _lib as L
grid=512
ne=bench.build_ne(bench.parse(['--grid',str(grid)]),'cuda')
c=299792458.0; lwl=1064e-9; om=2*np.pi*c/lwl
axes=[np.float32(np.linspace(-l/2,l/2,grid)) for l in bench.LENGTHS]
fld=E.DeviceField.from_ne(ne, axes[0], axes[1], axes[2], om, march_axis=2)
n=1<<20
beam=E.make_beam('circular', bench.BEAM_R, bench.BEAM_DIV, bench.EXTENT, 'z', seed=2)
s0=E.beam_generate(beam,n)
for rtol,atol in ((1e-3,1e-6),(1e-7,1e-9)):
    P=E.make_params('rk45', extent=bench.EXTENT, omega=om, rtol=rtol, atol=atol)
    out=E.propagate(fld,P,s0=s0,want_steps=True,want_rf=False)
    torch.cuda.synchronize()
    st=out['steps'].cpu().numpy().astype(np.int64)
    print('rtol',rtol,'steps/nfev per ray: mean',st.mean(),'median',np.median(st),'p90',np.percentile(st,90),'p99',np.percentile(st,99),'max',st.max())
    x=s0[0].cpu().numpy(); y=s0[1].cpu().numpy()
    dx=bench.LENGTHS[0]/(grid-1)
    ix=np.floor((x+bench.LENGTHS[0]/2)/dx).astype(np.int64); iy=np.floor((y+bench.LENGTHS[1]/2)/dx).astype(np.int64)
    o=np.lexsort((iy,ix)); s=st[o][:(n//32)*32].reshape(-1,32)
    print('  bundle mean/max utilisation', (s.mean(1)/s.max(1)).mean(), ' global mean/ mean-of-max', s.mean()/s.max(1).mean())
    print('  hist', np.histogram(st,bins=10)[0], np.histogram(st,bins=10)[1].astype(int))
