import ctypes as C, numpy as np, sys
sys.path.insert(0,'.')
import oracle
from gslam_b200 import capi
from gslam_b200.api import Context
ctx = Context(0); L = capi.lib()
L.gb_dbg_orb_level.restype = C.c_int
L.gb_dbg_orb_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
g=np.load('tests/golden/orb_320x240_n300.npz'); img=g['image']
kps, desc = ctx.orb_extract(img, 300)
for l in range(8):
    w=C.c_int(); h=C.c_int(); buf=np.zeros(320*240,np.uint8)
    rc=L.gb_dbg_orb_level(ctx.handle,l,buf.ctypes.data,buf.size,C.byref(w),C.byref(h))
    got=buf[:w.value*h.value].reshape(h.value,w.value)
    want=oracle.orb_pyramid_level(img,l)
    d=(got!=want)
    print('level',l,(w.value,h.value),want.shape,'mismatching px',int(d.sum()), 'rows with mismatch', np.flatnonzero(d.any(axis=1))[:10], 'cols', np.flatnonzero(d.any(axis=0))[:10])
