import os, sys, torch
sys.path.insert(0, "/root/repo")
import rpst
from oracle import restate as R
for shape in ((1,200,31,5),(1,200,40,40),(1,256,10,20),(2,64,100,36),(1,128,33,4),(3,16,8,8),(1,64,9,8)):
    c, s = R.synth_features(shape, cfg=3, device="cuda")
    for m in ("closed-form","original"):
        a=rpst.get_tuning("wct_ns_flagged")
        out=rpst.wct_fuse(c,s,m)
        b=rpst.get_tuning("wct_ns_flagged")
        print(shape,m,"flagged",b-a, flush=True)
