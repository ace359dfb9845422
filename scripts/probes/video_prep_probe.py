import sys, random, numpy as np, torch
sys.path.insert(0, '.')
from avid_cma_b200.datasets.gpu_preprocessing import VideoPrep_MSC_CJ
B,T,H,W=64,8,256,340
clips=torch.from_numpy(np.random.default_rng(0).integers(0,256,(B,T,H,W,3),dtype=np.uint8)).cuda()
prep=VideoPrep_MSC_CJ(crop=(224,224),num_frames=T)
random.seed(0)
params=[prep.draw(W,H) for _ in range(B)]
run,_=prep.plan_batch(list(clips),params)
for _ in range(3): run()
torch.cuda.synchronize()
