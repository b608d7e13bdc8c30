import os, sys, numpy as np
sys.path.insert(0,'/root/repo')
import patolette_b200 as pb
side=4096; n=side*side
rgb8=np.random.default_rng(3).integers(0,256,(n,3),dtype=np.uint8)
kw=dict(dither=True,color_space=2,kmeans_niter=10)
for keep in ("1","0"):
    os.environ["PB_SAL_KEEP_CARVEOUT"]=keep
    for _ in range(2): pb.quantize_u8(side,side,rgb8,256,tile_size=512,**kw)
    print("keep" if keep=="1" else "reset", {k:round(v,2) for k,v in pb.last_timings().items() if k in("dither","lq","saliency","total")}, flush=True)
pb.quantize_u8(side,side,rgb8,256,**kw); print("unweighted", round(pb.last_timings()["dither"],2))
