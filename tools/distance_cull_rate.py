#!/usr/bin/env python
"""CPU estimate for DESIGN.md section 9: how many candidate tests of primary rays a distance cull in the kernel's candidate walk
would skip.  Replays traceScene's phase A / phase B on a grid of primary rays with the real item bounds (lower.cpp through the
probe of tests/test_lowering_cpu.py) and the oracle's hit lists: an item tested after a hit at t is known could be skipped when its
bounding sphere begins beyond t.  usage: python tools/distance_cull_rate.py"""
import sys, ctypes as C, numpy as np, subprocess, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from functracer_b200 import scenes, frontend, abi
from oracle import ftb_oracle as orc
import test_lowering_cpu as T
import tempfile
d=tempfile.mkdtemp(prefix='ftb_probe_')
src,so=d+'/probe.cpp',d+'/libprobe.so'
open(src,'w').write(T.PROBE % dict(csrc=T.CSRC))
subprocess.check_call(["g++","-O1","-std=c++17","-fPIC","-shared","-o",so,src,os.path.join(T.CSRC,"cuda","lower.cpp")])
lib=C.CDLL(so)
lib.ftb_probe_bounds.argtypes=[C.POINTER(abi.SceneDesc),C.c_int,C.POINTER(C.c_double),C.c_int,C.POINTER(C.c_int),C.POINTER(C.c_int),C.POINTER(C.c_int)]
for name in ["cfg5-repeat","cfg3-house","cfg3-night-house","cfg2-hollow-sphere"]:
    W,H=96,54
    sc=frontend.ParsedScene(scenes.config_text(name,res=(W,H),spp=1),scenes.asset_dir())
    bounds,prim_item=T._item_bounds(lib,sc)
    tot_c=0; after=0; culled=0; rays=0
    for py in range(0,H,2):
        for px in range(0,W,2):
            ray=orc.primary_ray(sc.camera,W,H,px,py); o,dd=ray[:3],ray[3:]
            o=np.array(o); dd=np.array(dd); o=o+1e-4*dd
            du=dd/np.linalg.norm(dd); L=np.linalg.norm(dd)
            hits=orc.node_hits(sc,o,dd,max_hits=4096)
            # nearest t>=0 per item
            best={}
            for h in hits:
                if h["t"]>=0:
                    it=prim_item[h["prim"]]; best[it]=min(best.get(it,np.inf),h["t"])
            # candidates in enumeration order
            cur=np.inf
            rays+=1
            for it in range(bounds.shape[0]):
                c,r=bounds[it,:3],bounds[it,3]
                if r>=0:
                    oc=c-o; b=oc@du; 
                    if oc@oc-b*b>r*r*1.004 or (b<0 and oc@oc>r*r*1.004): continue
                tot_c+=1
                if cur<np.inf:
                    after+=1
                    if r>=0 and (b-r*1.002)/L>cur: culled+=1
                t=best.get(it,np.inf)
                if t<cur: cur=t
    print("%-20s rays %d: candidates/ray %.2f, tested after a hit is known %.2f/ray, of which a distance cull would skip %.2f/ray (%.0f %% of all candidate tests)"%(name,rays,tot_c/rays,after/rays,culled/rays,100*culled/max(tot_c,1)))
