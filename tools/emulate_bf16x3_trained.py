"""Where does the tensor-core path's logit error come from on the reference's shipped 3x1024 model?  (CPU emulation.)

    python tools/emulate_bf16x3_trained.py        # needs tests/golden/_local/trained_3x1024.npz

Runs the oracle's forward pass on the speech-like input of tests/test_gpu_trained.py three ways -- float64, plain fp32
(what TensorFlow computes) and with every matrix product replaced by its bf16x3 emulation (operands split in two bf16
pieces, three products, as the kernels do; fp32 state between kernels) -- then with bf16x3 in ONE product family at a
time, and finally as the product computes it (six-product input dense, bf16x3 elsewhere).  Test infrastructure only.
"""
import sys, numpy as np, torch
ROOT=__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,ROOT+'/tests')
src = open(ROOT+"/tests/test_gpu_trained.py").read().split("def test_shipped")[0].replace("from conftest import GOLDEN","GOLDEN=ROOT+'/tests/golden'").replace("pytestmark = pytest.mark.gpu","")
ns = {"ROOT": ROOT}; exec(src, ns)
from oracle import features, model
w = np.load(ns["FIXTURE"])
L,H,F,C=3,1024,120,80
p={"input_w":w["Input_Layer/input_w"],"input_b":w["Input_Layer/input_b"],"output_w":w["Output_layer/output_w"],"output_b":w["Output_layer/output_b"]}
for l in range(L):
    p["kernel_%d"%l]=w["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel"%l]; p["bias_%d"%l]=w["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/bias"%l]
rng=np.random.default_rng(2024)
sigs=[ns["_speechlike"](rng,s,22050) for s in (3.0,2.4,2.8,1.7)]
fs=[features.fbank(s,22050,320) for s in sigs]
lens=np.array([n for _,n in fs]); T=lens.max()
x=np.zeros((T,4,F),np.float32)
for b,(f,n) in enumerate(fs): x[:n,b]=f
print("feature magnitude: max |x| %.1f rms %.2f; |w_i| max %.2f; kernel max %.2f" % (np.abs(x).max(), np.sqrt((x**2).mean()), np.abs(p["input_w"]).max(), np.abs(p["kernel_0"]).max()))
def bf(a): return torch.tensor(np.asarray(a,np.float32)).to(torch.bfloat16).to(torch.float64).numpy()
def mm3(a,b):
    a=np.asarray(a,np.float32).astype(np.float64); b=np.asarray(b,np.float32).astype(np.float64)   # fp32 storage between kernels
    ah=bf(a); al=bf(a-ah); bh=bf(b); bl=bf(b-bh)
    return ah@bh+ah@bl+al@bh
def forward(p,x,lens,mm,f32=False):
    T,B,F=x.shape
    rd=(lambda v: np.asarray(v,np.float32).astype(np.float64)) if f32 else (lambda v:v)
    valid=(np.arange(T)[:,None]<lens[None,:])
    cur=rd(mm(x.reshape(T*B,F),p["input_w"])+p["input_b"]).reshape(T,B,H)
    for l in range(L):
        K,bb=p["kernel_%d"%l].astype(np.float64),p["bias_%d"%l].astype(np.float64)
        gx=rd(mm(cur.reshape(T*B,H),K[:H])+bb).reshape(T,B,4*H)
        c=np.zeros((B,H)); h=np.zeros((B,H)); out=np.zeros((T,B,H))
        for t in range(T):
            g=rd(gx[t]+mm(h,K[H:]))
            i=1/(1+np.exp(-g[:,:H])); j=np.tanh(g[:,H:2*H]); f=1/(1+np.exp(-(g[:,2*H:3*H]+1))); o=1/(1+np.exp(-g[:,3*H:]))
            cn=rd(c*f+i*j); hn=rd(np.tanh(cn)*o); v=valid[t][:,None]
            out[t]=np.where(v,hn,0); c=np.where(v,cn,c); h=np.where(v,hn,h)
        cur=out
    return mm(cur.reshape(T*B,H),p["output_w"]).reshape(T,B,C)+p["output_b"]
import warnings; warnings.simplefilter("ignore")
exact=forward(p,x.astype(np.float64),lens,lambda a,b:np.asarray(a,np.float64)@np.asarray(b,np.float64))
x3=forward(p,x.astype(np.float64),lens,mm3,f32=True)
f32=forward(p,x.astype(np.float64),lens,lambda a,b:(np.asarray(a,np.float32)@np.asarray(b,np.float32)).astype(np.float64),f32=True)
valid=(np.arange(T)[:,None]<lens[None,:])
print("logit range", exact[valid].min(), exact[valid].max())
print("bf16x3 emulation (fp32 state) vs float64: max |logit err| %.3e" % np.abs(x3-exact)[valid].max())
print("plain fp32 (numpy sgemm, fp32 state) vs float64: max |logit err| %.3e" % np.abs(f32-exact)[valid].max())

def mm6(a,b):
    a=np.asarray(a,np.float32).astype(np.float64); b=np.asarray(b,np.float32).astype(np.float64)
    ah=bf(a); am=bf(a-ah); al=bf(a-ah-am); bh=bf(b); bm=bf(b-bh); bl=bf(b-bh-bm)
    return ah@bh+am@bh+al@bh+ah@bm+ah@bl+am@bm
# which product family carries the error?  exact (float64) everywhere except one family in bf16x3
ex=lambda a,b:np.asarray(a,np.float64)@np.asarray(b,np.float64)
def forward_sel(p,x,lens,which):
    T,B,F=x.shape
    rd=lambda v: np.asarray(v,np.float32).astype(np.float64)
    valid=(np.arange(T)[:,None]<lens[None,:])
    cur=rd((mm3 if which=="input" else mm6 if which=="product" else ex)(x.reshape(T*B,F),p["input_w"])+p["input_b"]).reshape(T,B,H)
    for l in range(L):
        K,bb=p["kernel_%d"%l].astype(np.float64),p["bias_%d"%l].astype(np.float64)
        gx=rd((mm3 if which in ("gx","product") else ex)(cur.reshape(T*B,H),K[:H])+bb).reshape(T,B,4*H)
        c=np.zeros((B,H)); h=np.zeros((B,H)); out=np.zeros((T,B,H))
        for t in range(T):
            g=rd(gx[t]+(mm3 if which in ("rec","product") else ex)(h,K[H:]))
            i=1/(1+np.exp(-g[:,:H])); j=np.tanh(g[:,H:2*H]); f=1/(1+np.exp(-(g[:,2*H:3*H]+1))); o=1/(1+np.exp(-g[:,3*H:]))
            cn=rd(c*f+i*j); hn=rd(np.tanh(cn)*o); v=valid[t][:,None]
            out[t]=np.where(v,hn,0); c=np.where(v,cn,c); h=np.where(v,hn,h)
        cur=out
    return (mm3 if which in ("output","product") else ex)(cur.reshape(T*B,H),p["output_w"]).reshape(T,B,C)+p["output_b"]
for which in ("none","input","gx","rec","output","product"):
    y=forward_sel(p,x.astype(np.float64),lens,which)
    print("bf16x3 only in %-7s: max |logit err| %.3e%s" % (which, np.abs(y-exact)[valid].max(), "   (= the product: six-product input dense, bf16x3 elsewhere)" if which == "product" else ""))
