import sys, numpy as np
sys.path.insert(0, '.')
from oracle import enmap_np as oenmap, maps_np as omaps, stats_np as ostats, theory as otheory
from orphics_b200 import maps, stats
th = otheory.load_theory()
for npix,res in ((512,2.0),(96,3.0),(256,1.0)):
    w=npix*res
    shape,wcs = maps.rect_geometry(width_arcmin=w, px_res_arcmin=res)
    so,wo = omaps.rect_geometry(width_arcmin=w, px_res_arcmin=res)
    modl = np.asarray(oenmap.modlmap(so,wo)); ells=np.arange(0,modl.max()+1,1.)
    ps = otheory.power_from_theory(ells, th, lensed=True, pol=False)
    mg = maps.MapGen(shape,wcs,ps); og = omaps.MapGen(so,wo,ps)
    fc, ofc = maps.FourierCalc(shape,wcs), omaps.FourierCalc(so,wo)
    m = mg.get_map(seed=1000); mo = og.get_map(seed=1000)
    rel = lambda a,b: np.max(np.abs(np.asarray(a)-np.asarray(b)))/np.max(np.abs(b))
    print(npix, 'covsqrt', rel(mg.covsqrt, og.covsqrt), 'map', rel(m,mo))
    p2d,k1,_ = fc.power2d(m); p2o,ko,_ = ofc.power2d(mo)
    print('  k (own maps)', rel(k1,ko), 'p2d', rel(p2d,p2o))
    p2d,k1,_ = fc.power2d(mo)
    print('  k (same map)', rel(k1,ko), 'p2d', rel(p2d,p2o))
    d = np.abs(np.asarray(k1)-np.asarray(ko)); i = np.unravel_index(d.argmax(), d.shape); print('  argmax', i, np.asarray(k1)[i], np.asarray(ko)[i])
