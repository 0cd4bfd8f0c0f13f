"""Host-side cost of the expansion driver without a GPU: MVS::expansionPatches with the plane stand-in for refine() (test hook
library) on a 5 x 1600x1200 camera rig, 64 seeds -> ~120 k patches; prints the driver's pop / generate / commit seconds.
usage: python tools/host_expansion_bench.py [cameras cellSize]"""
import ctypes as C
import os
import sys
import time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path[:0] = [os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'pais-mvs_b200', 'python')]
import numpy as np
from pmvs_b200 import abi
import orc_host as oh
import test_host_parity_cpu as T
L=C.CDLL(T.HOOKS)
L.tmvs_hook_create.restype=C.c_void_p; L.tmvs_hook_create.argtypes=[C.POINTER(abi.PmvsConfig)]
L.tmvs_hook_add_camera.argtypes=[C.c_void_p,C.c_double,C.c_void_p,C.c_void_p,C.c_int,C.c_int,C.c_void_p]
L.tmvs_hook_put_patch.argtypes=[C.c_void_p,C.c_int,C.c_void_p,C.c_void_p,C.c_double,C.c_double,C.c_double,C.c_int,C.c_void_p,C.c_void_p,C.c_int]; L.tmvs_hook_put_patch.restype=None
L.tmvs_hook_expand_plane.restype=C.c_long; L.tmvs_hook_expand_plane.argtypes=[C.c_void_p,C.c_double,C.c_int,C.c_int,C.c_void_p]
L.tmvs_hook_patch_count.argtypes=[C.c_void_p]
NC=int(sys.argv[1]) if len(sys.argv)>1 else 5
cfg=abi.readme_config(); cfg.cellSize=int(sys.argv[2]) if len(sys.argv)>2 else 4; cfg.maxCellPatchNum=3; cfg.minCamNum=3; cfg.maxFitness=10.0; cfg.minCorrelation=0.9; cfg.neighborRadiusScalar=0.01
P=T.Pair(L,cfg,n_cams=NC,cols=1600,rows=1200,seed=3,with_grey=False)
rng=np.random.RandomState(1)
for pid in range(64):
    c=[2.0*(2*rng.rand()-1),1.5*(2*rng.rand()-1),0.0]
    pts=[cam.project(c,0,cfg.lodRatio)[0] for cam in P.o.cameras]
    P.put(oh.Patch(pid,c,[0.0,0.0,-1.0],1.0,1.0+0.1*pid,0.95,range(NC),pts))
ref=C.c_long(0); t=time.time()
calls=L.tmvs_hook_expand_plane(P.h,0.0,1024,1,C.byref(ref))
print("calls",calls,"refined",ref.value,"patches",L.tmvs_hook_patch_count(P.h),"wall %.2f s"%(time.time()-t))
