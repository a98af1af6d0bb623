"""Analysis (CPU, oracle only — test infrastructure): which bf16 rounding points of the VAE encoder move the KLD metric.
Emulates the engine's rounding points (input, weights, conv outputs, pooled outputs) on the fp32 oracle, one at a time and
layer by layer.  Result (B=8, seed 52): every source alone shifts the KLD by 3.6e-4 ... 7.4e-4 relative, all bf16 1.1e-3;
keeping only layers 3-5 exact does not help (1.3e-3) -> the whole encoder forward runs on split operands (engine.py).
usage: python tests/analysis/emulate_bf16_kld.py"""
import sys, torch, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import nets, keras_ops as K
from tests.util import synthetic_batch
from icsg3d_b200.params import ParamStore, vae_specs
B=8
ps=ParamStore(vae_specs(), 'cpu').init(3)
pv={k:torch.from_numpy(v) for k,v in ps.to_dict().items()}
M,cond,_=synthetic_batch(B,d=32,seed=52)
eps=torch.randn(B,256,generator=torch.Generator().manual_seed(7))
bf=lambda t: t.to(torch.bfloat16).float()
def enc(p, M, cond, rw, rc, ry, rx=True):
    """rw[i]: round weights of conv i (1..5) to bf16; rc[i]: round conv output; ry[i]: round pooled output; rx: round input"""
    x=torch.cat([M, nets.tile_cond(cond, M.shape[1:4], reps=4)],dim=-1)
    if rx: x=bf(x)
    for i in range(1,5):
        w=p[f"enc_conv{i}/kernel"]; w=bf(w) if rw[i] else w
        x=K.conv3d_same(x,w,p[f"enc_conv{i}/bias"])
        if rc[i]: x=bf(x)
        x,_,_=K.batchnorm(x,p[f"enc_bn{i}/gamma"],p[f"enc_bn{i}/beta"],None,None,True)
        x=K.maxpool2(K.leaky_relu(x))
        if ry[i]: x=bf(x)
    w=p["enc_conv5/kernel"]; w=bf(w) if rw[5] else w
    x=K.leaky_relu(K.conv3d_same(x,w,p["enc_conv5/bias"]))
    h=K.relu(K.dense(x.reshape(B,-1),p["enc_dense/kernel"],p["enc_dense/bias"]))
    zm=K.dense(h,p["z_mean/kernel"],p["z_mean/bias"]); lv=K.dense(h,p["z_log_var/kernel"],p["z_log_var/bias"])
    kl=(-0.5*(1+lv-zm**2-torch.exp(lv)).sum(1)).mean()
    return float(kl), zm
N={i:False for i in range(1,6)}; A={i:True for i in range(1,6)}
with torch.no_grad():
    ref,zr=enc(pv,M,cond,N,N,N,rx=False)
    def rep(name,*a,**k):
        v,z=enc(pv,M,cond,*a,**k); print(f"{name:50s} kld {v:.5f} rel err {abs(v-ref)/ref:.2e}  zmean relL2 {float((z-zr).norm()/zr.norm()):.2e}")
    print("ref",ref)
    rep("all bf16",A,A,A)
    rep("only input rounding",N,N,N)
    rep("only weights",A,N,N,rx=False)
    rep("only conv-out rounding",N,A,N,rx=False)
    rep("only pooled rounding",N,N,A,rx=False)
    sel=lambda s:{i:(i in s) for i in range(1,6)}
    rep("layers 1,2 bf16; 3-5 exact",sel({1,2}),sel({1,2}),sel({1,2}))
    rep("layer 1 bf16; 2-5 exact",sel({1}),sel({1}),sel({1}))
    rep("layer 1 bf16 w/ fp32 c; 2-5 exact",sel({1}),N,sel({1}))
    rep("layers 1,2 bf16 w/ fp32 c; 3-5 exact",sel({1,2}),N,sel({1,2}))
    rep("all bf16 but fp32 c",A,N,A)
    rep("exact except input",N,N,N,rx=True)
    rep("exact layer1 (and input), rest bf16",sel({2,3,4,5}),sel({2,3,4,5}),sel({2,3,4,5}),rx=False)
