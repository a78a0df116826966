import ctypes as C, numpy as np, sys
sys.path.insert(0,'.')
import oracle
from gslam_b200 import capi
from gslam_b200.api import Context
ctx = Context(0); L = capi.lib()
L.gb_dbg_orb_candidates.restype = C.c_int
g=np.load('tests/golden/orb_320x240_n300.npz'); img=g['image']
def run():
    kps, desc = ctx.orb_extract(img, 300)
    cap=100000
    pos=np.zeros(cap,np.uint32); sc=np.zeros(cap,np.uint8); n=C.c_int(); nk=C.c_int()
    L.gb_dbg_orb_candidates(ctx.handle, 0, pos.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p), None, None, cap, C.byref(n), None, 0, C.byref(nk))
    n=n.value
    return {(int(p&0xffff),int(p>>16)):int(s) for p,s in zip(pos[:n],sc[:n])}
G1=run(); G2=run()
print('deterministic across runs:', G1==G2, len(G1), len(G2))
sm = oracle.fast_score_map(img,20)
xs,ys,s0 = oracle.fast_detect(img,20,True)
O={(int(x),int(y)):int(s) for x,y,s in zip(xs,ys,s0) if 31<=x<320-31 and 31<=y<240-31}
mism=[(k,G1[k],O[k]) for k in G1 if k in O and G1[k]!=O[k]]
print('score mismatches', len(mism), mism[:12])
extra=sorted(set(G1)-set(O))[:12]
print('extra (pos, gpu score, oracle raw score map)', [(k,G1[k],int(sm[k[1],k[0]])) for k in extra])
miss=sorted(set(O)-set(G1))[:12]
print('missing (pos, oracle score)', [(k,O[k]) for k in miss])
# distribution of mismatch positions within tiles
print('mism lx%64', sorted(set(k[0]%64 for k,_,_ in mism))[:40])
print('mism ly%16', sorted(set(k[1]%16 for k,_,_ in mism)))
