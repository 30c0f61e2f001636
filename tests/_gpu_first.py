import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from conftest import render, ORACLE_LIB, CUDA_LIB
from ray_tracing_b200 import scenes
sc = scenes.cornell_spheres(256,256,4,1)
fo, ao, so = render(ORACLE_LIB, sc, frames=2, want_stats=True)
for k in (0,1):
    t=time.time()
    fg, ag, sg = render(CUDA_LIB, sc, frames=2, options={'kernel':k,'countStats':1}, want_stats=True)
    print('kernel',k,'time',time.time()-t,'bitwise frame',np.array_equal(fo.view(np.uint32),fg.view(np.uint32)),'accum',np.array_equal(ao.view(np.uint32),ag.view(np.uint32)),
          'maxdiff',np.nanmax(np.abs(ao-ag)), 'stats',sg, so)
