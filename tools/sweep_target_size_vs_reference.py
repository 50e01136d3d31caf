"""get_optimal_size of the UNMODIFIED reference (/root/reference, build container only) against lumina_target_size (the
C-ABI host entry) and the oracle on random and adversarial (near-integer quotient) sizes: int(h * max_dim / w) is a float
division in the reference, so the C restatement has to round the same way everywhere.  Result: profiles/r2_sweep_reference_vs_oracle.txt.

    python tools/sweep_target_size_vs_reference.py
"""
import sys, logging, numpy as np
logging.disable(logging.CRITICAL)
sys.path.insert(0,"/root/reference/backend"); sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from utils.image_preprocessing import ImagePreprocessor as Ref
from ocr_system_b200 import ops
import oracle as O
ref=Ref()
rng=np.random.default_rng(0)
bad=0
N=300000
for i in range(N):
    if i%3==0: w,h,md=int(rng.integers(1,20000)),int(rng.integers(1,20000)),int(rng.integers(1,5000))
    elif i%3==1: w,h,md=int(rng.integers(1,70000)),int(rng.integers(1,70000)),int(rng.choice([960,2000,4096,1234,65500]))
    else:
        md=int(rng.integers(2,4000)); w=int(rng.integers(md,md*40)); k=int(rng.integers(1,md)); h=max(1,(k*w)//md + int(rng.integers(-1,2)))  # near-integer quotients
    r=ref.get_optimal_size(w,h,md)
    if tuple(r)!=ops.target_size(w,h,md) or tuple(r)!=O.target_size(w,h,md):
        bad+=1
        if bad<5: print("MISMATCH",w,h,md,r,ops.target_size(w,h,md),O.target_size(w,h,md))
print("cases",N,"mismatches",bad)
