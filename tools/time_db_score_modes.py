import torch, numpy as np, time
from ocr_system_b200 import ops
from oracle import db_post as D
preds=np.stack([D.synth_prob_map(960,960,k,n_boxes=500) for k in range(8)])
p=torch.from_numpy(np.tile(preds,(8,1,1))).cuda()
hw=np.tile(np.array([[960,960]],np.int32),(p.shape[0],1))
for mode in ("fast","slow"):
    for _ in range(3): ops.db_postprocess(p,hw,0.3,0.6,1.5,1000,3,False,mode)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.db_postprocess(p,hw,0.3,0.6,1.5,1000,3,False,mode)
    e1.record(); torch.cuda.synchronize(); print(mode, e0.elapsed_time(e1)/10, "ms per 64 maps")
