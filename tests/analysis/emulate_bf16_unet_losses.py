"""Analysis (CPU, oracle only — test infrastructure): does running the LAST U-Net blocks exactly remove the bf16 error of the
BCE / CCE losses?  Emulates bf16 rounding of weights / ReLU outputs / BatchNorm outputs per block on
the fp32 oracle (B=1 @32^3).  Result: BCE error all-bf16 4.6e-4; c18+heads exact 1.9e-4; c17,c18+heads exact 5.1e-4;
c13..c18+heads exact 7.6e-4 -> zero-mean noise, no layer whose precision removes it (it averages out over voxels: 3e-5 at
B=8 on the GPU).  usage: python tests/analysis/emulate_bf16_unet_losses.py"""
import sys, torch, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import nets, keras_ops as K
from tests.util import synthetic_batch
from icsg3d_b200.params import ParamStore, unet_specs
B,d=1,32
ps=ParamStore(unet_specs(4,95),'cpu').init(5)
p={k:torch.from_numpy(v) for k,v in ps.to_dict().items()}
M,_,S=synthetic_batch(B,d=d,seed=0)
bf=lambda t: t.to(torch.bfloat16).float()
ORDER=["c1","c2","c3","c4","c5","c6","c9","c10","c13","c14","c15","c16","c17","c18"]
def fwd(exact=set(), exact_heads=False, exact_in=False):
    """bf16 emulation: per block: weights bf16, conv+relu output a -> bf16, BN(train) -> y bf16; blocks in `exact` keep w, a, y fp32"""
    def blk(name,x):
        e = name in exact
        w=p[name+"/kernel"]; w = w if e else bf(w)
        a=K.relu(K.conv3d_same(x,w,p[name+"/bias"]))
        if not e: a=bf(a)
        y,_,_=K.batchnorm(a,p[f"bn_{name}/gamma"],p[f"bn_{name}/beta"],None,None,True)
        if not e: y=bf(y)
        return y
    x = M if exact_in else bf(M)
    c1=blk("c1",x); c2=blk("c2",c1); c3=blk("c3",K.maxpool2(c2)); c4=blk("c4",c3); c5=blk("c5",K.maxpool2(c4)); c6=blk("c6",c5)
    c9=blk("c9",K.maxpool2(c6)); c10=blk("c10",c9)
    c13=blk("c13",torch.cat([c6,K.upsample2(c10)],-1)); c14=blk("c14",c13)
    c15=blk("c15",torch.cat([c4,K.upsample2(c14)],-1)); c16=blk("c16",c15)
    c17=blk("c17",torch.cat([c2,K.upsample2(c16)],-1)); c18=blk("c18",c17)
    ws,wg=p["soft/kernel"],p["sig/kernel"]
    if not exact_heads: ws,wg=bf(ws),bf(wg)
    soft=K.conv3d_same(c18,ws,p["soft/bias"]); sig=K.conv3d_same(c18,wg,p["sig/bias"])
    return float(nets.weighted_cce(soft,S.long(),95.0)), float(nets.sigmoid_bce(sig,S!=0))
with torch.no_grad():
    ref=fwd(set(ORDER),True,True); print("ref",ref)
    for name,args in [("all bf16",(set(),False,False)),("c18+heads exact",({"c18"},True,False)),("c17,c18+heads exact",({"c17","c18"},True,False)),
                      ("c13..c18+heads exact",(set(ORDER[8:]),True,False)), ("c1,c2 exact + input",({"c1","c2"},False,True))]:
        r=fwd(*args); print(f"{name:28s} cce {r[0]:.4f} d {abs(r[0]-ref[0]):.2e}  bce {r[1]:.6f} d {abs(r[1]-ref[1]):.2e}")
