import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, images
from cases import gpu_encode, oracle_encode
ok = True
for (w,h) in [(1920,1080),(258,128),(1000,37),(17,400)]:
    img = images.synth_frame(w,h,3,seed=3)
    cfg = dict(quality=90, sampling=(2,2))
    ok &= gpu_encode(img,w,h,'rgb',cfg) == oracle_encode(img,w,h,'rgb',cfg)
print('variant', os.environ.get('JPGB_STAGE_A_MINB'), 'parity', ok)
