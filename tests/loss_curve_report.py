"""Report (not a test): 24 SGD steps through the public API on the emulated ABI next to the oracle with bf16 storage
emulation and the plain fp32 oracle (= the reference's arithmetic); the table in profiles/r1_summary.md comes from here.
    python tests/loss_curve_report.py [--sim] [--steps N]
--sim: the library is libsseg_sim.so, i.e. the REAL csrc/*.cu kernel sources on the CPU simulator (tests/cusim), not the
Python restatement of the ABI (minutes per step).
"""
import sys, os, torch, torch.nn as nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'semantic-segmentation-pytorch_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, _p)
from abi_emulator import EmuLib
from oracle import segnet_oracle as O
from test_program_dry import _seg
from test_program_emulated import _load
from mit_semseg.engine import _C, ops, functional as EF, program as PR
SIM="--sim" in sys.argv
STEPS=int(sys.argv[sys.argv.index("--steps")+1]) if "--steps" in sys.argv else 24
if SIM:
    os.environ.setdefault("CUSIM_SMS","16")
    import conftest
    lib=conftest.sim_lib()
else:
    lib=EmuLib()
_C.lib=lambda: lib; ops._stream=lambda: None
real=PR.SegProgram
def factory(*a,**k):
    k["dry_run"]=True; p=real(*a,**k); p.dry_run=False; p.serial=True; return p
EF.SegProgram=factory
real.capture=lambda self, warm=True: None
torch.Tensor.is_cuda=property(lambda self: True)
torch.manual_seed(0)
enc,dec,fc="resnet18dilated","ppm_deepsup",512
seg=_seg(enc,dec,fc); esd,dsd=_load(seg,enc,dec,fc); seg.train()
for m in seg.modules():
    if isinstance(m,nn.Dropout2d): m.p=0.0
opt=torch.optim.SGD(seg.parameters(),lr=0.02,momentum=0.9,weight_decay=1e-4)
e={k:v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k,v in esd.items()}
d={k:v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k,v in dsd.items()}
optr=torch.optim.SGD([v for v in list(e.values())+list(d.values()) if v.requires_grad],lr=0.02,momentum=0.9,weight_decay=1e-4)
e32={k:v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k,v in esd.items()}
d32={k:v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k,v in dsd.items()}
opt32=torch.optim.SGD([v for v in list(e32.values())+list(d32.values()) if v.requires_grad],lr=0.02,momentum=0.9,weight_decay=1e-4)
rows=[]
for step in range(STEPS):
    feed=O.synth_batch(2,64,64,8,200+step%4)
    opt.zero_grad(); loss,acc=seg(feed); loss.backward(); opt.step()
    optr.zero_grad(); lr_,_=O.segmentation_forward(feed,e,d,enc,dec,O.BNState(True,emulate="bf16",update_running=True),0.4,dropout_p=0.0); lr_.backward(); optr.step()
    opt32.zero_grad(); l32,_=O.segmentation_forward(feed,e32,d32,enc,dec,O.BNState(True,update_running=True),0.4,dropout_p=0.0); l32.backward(); opt32.step()
    rows.append((step,loss.item(),lr_.item(),l32.item()))
    print("%2d engine-schedule %.4f  oracle(bf16 storage) %.4f  oracle(fp32 = the reference) %.4f" % rows[-1], flush=True)
