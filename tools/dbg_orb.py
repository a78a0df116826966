import ctypes as C, numpy as np, sys
sys.path.insert(0,'.')
import oracle
from gslam_b200 import capi
from gslam_b200.api import Context
ctx = Context(0); L = capi.lib()
L.gb_dbg_orb_candidates.restype = C.c_int
g=np.load('tests/golden/orb_320x240_n300.npz'); img=g['image']
kps, desc = ctx.orb_extract(img, 300)
q = oracle.orb_quotas(300)
for l in range(8):
    cap = 100000
    pos=np.zeros(cap,np.uint32); sc=np.zeros(cap,np.uint8); rs=np.zeros(cap,np.float32); key=np.zeros(cap,np.uint32); kp=np.zeros(4096,np.uint32)
    n=C.c_int(); nk=C.c_int()
    rc = L.gb_dbg_orb_candidates(ctx.handle, l, pos.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p), rs.ctypes.data_as(C.c_void_p), key.ctypes.data_as(C.c_void_p), cap, C.byref(n), kp.ctypes.data_as(C.c_void_p), 4096, C.byref(nk))
    n=n.value; nk=nk.value
    lv = oracle.orb_pyramid_level(img, l); h,w = lv.shape
    xs,ys,s0 = oracle.fast_detect(lv,20,True)
    m=(xs>=31)&(xs<w-31)&(ys>=31)&(ys<h-31); xs,ys,s0=xs[m],ys[m],s0[m]
    G = {(int(p&0xffff),int(p>>16)):int(s) for p,s in zip(pos[:n],sc[:n])}
    O = {(int(x),int(y)):int(s) for x,y,s in zip(xs,ys,s0)}
    print('level',l,'gpu cands',n,'dups',n-len(G),'oracle',len(O),'extra',sorted(set(G)-set(O))[:5],'missing',sorted(set(O)-set(G))[:5],'score mism',sum(1 for k in G if k in O and G[k]!=O[k]),'kept',nk,'quota',q[l])
    Gk = {(int(p&0xffff),int(p>>16)) for p in kp[:nk]}
    Ok = {(int(round(k['x']/np.float32(np.float64(np.float32(1.2))**l))), int(round(k['y']/np.float32(np.float64(np.float32(1.2))**l)))) for k in g['kps'] if k['octave']==l}
    print('     kept extra', sorted(Gk-Ok)[:5], 'missing', sorted(Ok-Gk)[:5])
    if l==0:
        for (x,y) in sorted(Gk-Ok)[:3]:
            i=[k for k,p in enumerate(pos[:n]) if (int(p&0xffff),int(p>>16))==(x,y)]
            print('     extra cand entries', [(int(sc[k]), float(rs[k]), hex(int(key[k]))) for k in i], 'oracle harris', oracle.harris_response(lv,x,y))
