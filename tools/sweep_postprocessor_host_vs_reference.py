"""Host glue of ocr_postprocessor against the UNMODIFIED reference module (/root/reference, build container only):
parse_rapidocr_output on random well- and ill-formed RapidOCR results (None, empty, short items, string / numpy
confidences, malformed boxes, non-string text), format_merged_output, and the derived TextBlock properties the device kernel
also computes.  Results (dataclass tuples, exception types, strings, repr of floats) must be identical.

    python tools/sweep_postprocessor_host_vs_reference.py
"""
import sys, logging, dataclasses, numpy as np
logging.disable(logging.CRITICAL)
sys.path.insert(0,"/root/reference/backend"); sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import utils.ocr_postprocessor as R
import ocr_system_b200.ocr_postprocessor as O
rng=np.random.default_rng(0)
def rand_item():
    k=rng.integers(0,10)
    x,y=rng.uniform(0,1000,2); w,h=rng.uniform(1,200,2)
    box=[[x,y],[x+w,y],[x+w,y+h],[x,y+h]]
    if k==0: return None
    if k==1: return [box]                       # too short
    if k==2: return [box,"txt"]                 # no confidence
    if k==3: return [box,"txt","0.5"]           # str conf
    if k==4: return (np.array(box),"t",np.float32(0.25))
    if k==5: return [[[1,2],[3,4]],"bad box",0.9]
    if k==6: return [box,"",0.1]
    if k==7: return [box, 123, 0.7]
    return [box,"word%d"%rng.integers(0,99),float(rng.random())]
def as_tuple(b): return tuple(dataclasses.astuple(b))
bad=0
for it in range(3000):
    res=[rand_item() for _ in range(rng.integers(0,12))]
    if it%50==0: res=None
    if it%50==1: res=[]
    try: r=[as_tuple(b) for b in R.parse_rapidocr_output(res)]; re_=None
    except Exception as e: r=None; re_=type(e)
    try: o=[as_tuple(b) for b in O.parse_rapidocr_output(res)]; oe=None
    except Exception as e: o=None; oe=type(e)
    if repr(r)!=repr(o) or re_!=oe:
        bad+=1
        if bad<4: print("MISMATCH",res,r,o,re_,oe)
# format_merged_output
for it in range(500):
    lines=[R.MergedLine(text="t%d"%i, confidence=float(rng.random()), y_position=float(rng.random()*1000), blocks=[]) for i in range(rng.integers(0,6))]
    lo=[O.MergedLine(text=l.text, confidence=l.confidence, y_position=l.y_position, blocks=[]) for l in lines]
    for sc in (False,True):
        if R.format_merged_output(lines,sc)!=O.format_merged_output(lo,sc): bad+=1; print("FMT MISMATCH")
print("mismatches",bad)
# TextBlock derived properties (the quantities the device kernel also computes)
props=[n for n in dir(R.TextBlock) if isinstance(getattr(R.TextBlock,n),property)]
pbad=0
for it in range(3000):
    box=[[float(v) for v in rng.uniform(-50,2000,2)] for _ in range(4)]
    if it%7==0: box=[[int(a),int(b)] for a,b in box]
    rb=R.TextBlock(text="x",confidence=0.5,box=box); ob=O.TextBlock(text="x",confidence=0.5,box=box)
    for n in props:
        a,b=getattr(rb,n),getattr(ob,n)
        if repr(a)!=repr(b): pbad+=1; print("PROP MISMATCH",n,a,b)
print("properties",props,"mismatches",pbad)
