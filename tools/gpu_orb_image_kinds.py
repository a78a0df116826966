"""GPU-vs-oracle ORB parity on the image kinds of tests/_images.py (noise, binary, checkerboard ... : massive ties, saturated
regions).  The oracle is green against cv2 on all of them (tests/test_oracle_orb.py::test_live_cv2_image_kinds); this script is the
next step: run it once under `compute-sanitizer --tool memcheck` on a B200, then promote the cases into tests/test_orb_gpu.py."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import oracle
from _images import KIND_CASES, image_of_kind
from gslam_b200.api import Context

ctx = Context(0)
bad = 0
for kind, w, h, seed, n, nl, sf, ft in KIND_CASES:
    img = np.ascontiguousarray(image_of_kind(kind, w, h, seed))
    kw = dict(nlevels=nl, scale_factor=sf, fast_threshold=ft)
    wk, wd = oracle.orb_extract(img, n, **kw)
    try:
        kps, desc = ctx.orb_extract(img, n, **kw)
        ok = len(kps) == len(wk) and all(np.array_equal(kps[f], wk[f]) for f in ("octave", "x", "y", "size", "angle", "response")) \
            and np.array_equal(desc, wd)
    except Exception as e:  # capacity errors are reported, not fatal
        ok = False; print("  error:", e)
    bad += not ok
    print(("ok      " if ok else "MISMATCH"), kind, w, h, n, nl, sf, ft, "oracle", len(wk))
print("mismatches:", bad)
