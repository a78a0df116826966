"""CPU-only fuzz of the ORB oracle against the installed cv2 (random sizes, scale factors, thresholds, feature counts)."""
import sys, numpy as np, time
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests','golden'))
import oracle
from gslam_b200 import synth
from make_golden_orb import cv2_orb_canonical
rng=np.random.default_rng(4242)
bad=0
t0=time.time()
for it in range(300):
    w=int(rng.integers(70,1500)); h=int(rng.integers(70,1200)); seed=int(rng.integers(0,10000))
    n=int(rng.choice([5,37,150,500,1500,4000]))
    nl=int(rng.integers(1,9)); sf=float(rng.choice([1.05,1.1,1.15,1.2,1.25,1.3,1.33,1.4,1.5,1.7,2.0])); ft=int(rng.choice([5,10,20,30,45]))
    img=synth.synth_frame(w,h,seed)
    try:
        want,wdesc,_=cv2_orb_canonical(img,n,nlevels=nl,scaleFactor=sf,fastThreshold=ft)
        kps,desc=oracle.orb_extract(img,n,nlevels=nl,scale_factor=sf,fast_threshold=ft)
    except Exception as e:
        print("EXC",w,h,seed,n,nl,sf,ft,repr(e)[:200]); bad+=1; continue
    ok=len(kps)==len(want) and all(np.array_equal(kps[f],want[f]) for f in ("octave","x","y","size","angle","response")) and np.array_equal(desc,wdesc)
    if not ok:
        bad+=1
        msg=[f for f in ("octave","x","y","size","angle","response") if len(kps)==len(want) and not np.array_equal(kps[f],want[f])]
        print("MISMATCH",w,h,seed,n,nl,sf,ft,len(kps),len(want),msg, int((desc!=wdesc).sum()) if len(kps)==len(want) else -1)
print("done",bad,"bad of N in",round(time.time()-t0,1),"s")
