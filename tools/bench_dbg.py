import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, json, torch
sys.path.insert(0, %r)
from bin_b200 import ops
dev="cuda"; B,h,w=5,360,640
def run(cin, cout, k, split, relu=True, res=False):
    cin_pad=(cin+31)//32*32
    wt=torch.randn(cout,cin,k,k,device=dev)/(cin*k*k)**0.5
    wp,bp=ops.pack_conv_weight(wt,cout,cin_pad),ops.pad_bias(torch.zeros(cout,device=dev),cout)
    x=torch.randn(B,12,h,w,8,device=dev).half(); g=torch.randn(B,16,h,w,8,device=dev).half()
    out=ops.empty_p8(B,cout//8,h,w,dev)
    kw=dict(in0_planes=12,in1=g,in1_planes=(cin-96)//8,relu=relu,out=out, res=x if res else None)
    for _ in range(3): ops.conv_fwd(x,wp,bp,k,cout,**kw)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.conv_fwd(x,wp,bp,k,cout,**kw)
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1)/10*1e3,1)
print(json.dumps({"conv0_us":run(96,32,3,None),"conv3_us":run(192,32,3,96),"lff_us":run(224,96,1,96,relu=False,res=True)}))
''' % ROOT
for dbg in (0, 1, 2, 3, 4, 5, 6, 7):
    env = dict(os.environ, BIN_B200_DEBUG=str(dbg))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    print("debug", dbg, "(noTMA=%d noEPI=%d noMMA=%d)" % (dbg & 1, (dbg >> 1) & 1, (dbg >> 2) & 1), r.stdout.strip(), r.stderr.strip()[-300:], flush=True)
