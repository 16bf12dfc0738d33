import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.nn.functional as F
from bin_b200 import ops
from bin_b200._lib import Act, check, lib
cin, cout, k = 36, 96, 5
B, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (2, 27, 41)
torch.manual_seed(1); dev = "cuda"
x = torch.randn(B, cin, H, W, device=dev).half().float().requires_grad_(True)
w = (torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5).half().float().requires_grad_(True)
dy = torch.randn(B, cout, H, W, device=dev).half().float()
if os.environ.get("INNER"):
    m = int(os.environ["INNER"]); mask = torch.zeros_like(dy); mask[:, :, m:H - m, m:W - m] = 1; dy = dy * mask
torch.backends.cudnn.allow_tf32 = False
gw_ref, = torch.autograd.grad((F.conv2d(x, w, None, padding=k // 2) * dy).sum(), [w])
w64 = w.detach().double().cpu().requires_grad_(True)
gw64, = torch.autograd.grad((F.conv2d(x.detach().double().cpu(), w64, None, padding=k // 2) * dy.double().cpu()).sum(), [w64])
print("cudnn fp32 vs cpu fp64 ref err:", (gw_ref.cpu().double() - gw64).abs().max().item())
gw_ref = gw64.float().cuda()
scale = torch.full((1,), 4.0, device=dev)
dys = ops.nchw_to_p8(dy * 4.0, pad_to=8)
x0 = ops.nchw_to_p8(x.detach(), pad_to=32)
dw = torch.zeros_like(w)
st = torch.cuda.current_stream().cuda_stream
check(lib().bin_conv_wgrad(ops.act_view(x0), 0, x0.shape[1], Act(None, 0, 0, 0, 0), 0, 0, ops.act_view(dys), 0, cout, cin, k, scale.data_ptr(), dw.data_ptr(), st))
torch.cuda.synchronize()
err = (dw - gw_ref).abs()
print("max ref", gw_ref.abs().max().item(), "max err", err.max().item())
print("err per tap (ky,kx):")
print((err.amax(dim=(0, 1)) / gw_ref.abs().max()).cpu())
print("err per ci block of 8:", [round((err[:, i:i + 8].max() / gw_ref.abs().max()).item(), 4) for i in range(0, cin, 8)])
