import os
os.environ.setdefault("BIN_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bin_b200", "libbin_b200_tools.so"))  # tools build: timeline hooks + microbenchmarks


def microbench_mma(n, iters=4096, mode=0):
    import ctypes as C
    from bin_b200 import _lib
    v = C.c_float(0)
    _lib.check(_lib.lib().bin_tools_microbench_mma(n, iters, mode, C.byref(v)))
    return v.value

"""First-contact GPU probe: tcgen05 issue rates and per-instantiation conv parity.
Each case runs in its own subprocess (a device trap must not take the others down).
Usage: python tools/gpu_probe.py [case ...]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def case_mma():
    from bin_b200 import ops
    res = {}
    names = {0: "ns", 1: "sw128", 2: "sw64", 3: "sw32"}
    for la in range(4):
        for lb in (0, 1):
            for n in (16, 32, 64, 96, 128, 256):
                mode = 2 | (la << 4) | (lb << 6)
                res[f"A{names[la]}_B{names[lb]}_N{n}"] = round(microbench_mma(n, 16384, mode), 2)
    for la in (0, 1):
        mode = (la << 4) | (la << 6)
        res[f"A{names[la]}_1acc_N32"] = round(microbench_mma(32, 16384, mode | 1), 2)
        res[f"A{names[la]}_shift_N32"] = round(microbench_mma(32, 16384, mode | 2 | 0x100), 2)
        res[f"A{names[la]}_M64_N32"] = round(microbench_mma(32, 16384, mode | 2 | 0x200), 2)
        res[f"A{names[la]}_M64_N128"] = round(microbench_mma(128, 16384, mode | 2 | 0x200), 2)
    print(json.dumps(res))


def _conv_case(cin, cout, k, H, W, B=2, relu=False, res=False, epilogue=0, cin_split=None, seed=0, variant=0):
    import torch
    import torch.nn.functional as F
    from bin_b200 import ops
    torch.manual_seed(seed)
    dev = "cuda"
    cin_pad = (cin + 31) // 32 * 32
    cout_pad = {3: 16}.get(cout, cout)
    x = torch.randn(B, cin, H, W, device=dev)
    w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device=dev) * 0.1
    xh = x.half().float()
    wh = w.half().float()
    ref = F.conv2d(xh, wh, b, padding=k // 2)
    if relu:
        ref = ref.relu()
    wp = ops.pack_conv_weight(w, cout_pad, cin_pad, variant)
    bp = ops.pad_bias(b, cout_pad)
    if cin_split is None:
        in0 = ops.nchw_to_p8(x)
        in1, in1_planes = None, 0
        in0_planes = cin_pad // 8
    else:
        in0 = ops.nchw_to_p8(x[:, :cin_split].contiguous())
        in1 = ops.nchw_to_p8(x[:, cin_split:].contiguous())
        in0_planes, in1_planes = cin_split // 8, (cin - cin_split) // 8
    if epilogue == 0:
        out = torch.full((B, cout_pad // 8 + 1, H, W, 8), 7.0, dtype=torch.float16, device=dev)
        r = None
        if res:
            rx = torch.randn(B, cout, H, W, device=dev)
            r = ops.nchw_to_p8(rx)
            ref = ref + rx.half().float()
        ops.conv_fwd(in0, wp, bp, k, cout_pad, in0_planes=in0_planes, in1=in1, in1_planes=in1_planes, relu=relu,
                     out=out, out_plane0=1, res=r, variant=variant)
        torch.cuda.synchronize()
        got = ops.p8_to_nchw(out, cout, plane0=1)
        guard = out[:, 0].float()
        assert (guard == 7.0).all(), "guard plane overwritten"
    elif epilogue == 1:
        out = torch.zeros((B, 8, 2 * H, 2 * W, 8), dtype=torch.float16, device=dev)
        ops.conv_fwd(in0, wp, bp, k, cout_pad, epilogue=1, out=out)
        torch.cuda.synchronize()
        got = ops.p8_to_nchw(out, 64)
        ref = F.pixel_shuffle(ref, 2)
    else:
        frames = [[torch.rand(1, 3, H, W, device=dev) for _ in range(3)] for _ in range(B)]
        outs = [torch.zeros(1, 3, H, W, device=dev) for _ in range(B)]
        fr = ops.make_frames(frames, outs)
        ops.conv_fwd(in0, wp, bp, k, cout_pad, epilogue=2, frames=fr)
        torch.cuda.synchronize()
        got = torch.cat(outs, 0)
        ref = ref + torch.cat([sum(f) / 3.0 for f in frames], 0)
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(json.dumps({"max_abs_err": err, "ref_max": scale, "ok": bool(err <= 4e-3 * max(1.0, scale))}))


def case_mmapat():
    from bin_b200 import ops
    res = {}
    for vary in range(8):
        res[f"A{'var' if vary & 1 else 'fix'}_B{'var' if vary & 2 else 'fix'}_{'data' if vary & 4 else 'zero'}"] = round(
            microbench_mma(96, 200, 0x1000 | vary), 2)
    print(json.dumps(res))


def case_sub(cin=128, k=3, cout=32):
    import torch
    import torch.nn.functional as F
    from bin_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    B, H, W = 3, 56, 64
    x = torch.randn(B, cin, H, W, device=dev)
    w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device=dev) * 0.1
    ref = F.conv2d(x.half().float(), w.half().float(), b, padding=k // 2)
    wp, bp = ops.pack_conv_weight(w, cout, cin), ops.pad_bias(b, cout)
    in0 = ops.nchw_to_p8(x)
    res = {}
    for sub in [(0, 0, 0, 0), (1, 1, 0, 0), (0, 0, 11, 23), (2, 1, 30, 26), (1, 2, 5, 9)]:
        out = torch.zeros((B, cout // 8, H, W, 8), dtype=torch.float16, device=dev)
        for rep in range(20):
            ops.conv_fwd(in0, wp, bp, k, cout, out=out, sub=sub)
        torch.cuda.synchronize()
        got = ops.p8_to_nchw(out, cout)
        b0, nb, y0, ny = sub
        nb = nb or B - b0
        ny = ny or H - y0
        mask = torch.zeros_like(ref)
        mask[b0:b0 + nb, :, y0:y0 + ny] = 1
        err = ((got - ref) * mask).abs().max().item()
        outside = (got * (1 - mask)).abs().max().item()
        res[str(sub)] = (round(err, 5), outside)
    print(json.dumps(res))


def case_mma2():
    from bin_b200 import ops
    res = {}
    for n in (32, 64, 96, 128, 192, 256):
        res[f"pair_M256_N{n}"] = round(microbench_mma(n, 8192, 0x2000), 2)
    for n in (96, 128, 256):
        res[f"single_M128_N{n}"] = round(microbench_mma(n, 8192, 2), 2)
    print(json.dumps(res))


CASES = {
    "mma2": case_mma2,
    "mmapat": case_mmapat,
    "sub3": lambda: case_sub(128, 3, 32),
    "sub1": lambda: case_sub(224, 1, 96),
    "mma": case_mma,
    "c3_32_small": lambda: _conv_case(96, 32, 3, 8, 30, B=1),
    "c3_32": lambda: _conv_case(96, 32, 3, 20, 37, relu=True),
    "c3_32_split": lambda: _conv_case(192, 32, 3, 24, 70, relu=True, cin_split=96),
    "c3_96_res": lambda: _conv_case(96, 96, 3, 19, 45, res=True),
    "c5_96": lambda: _conv_case(24, 96, 5, 18, 33),
    "c5_96_60": lambda: _conv_case(60, 96, 5, 30, 61),
    "c1_96": lambda: _conv_case(224, 96, 1, 21, 50, res=True, cin_split=96),
    "c1_96_1152": lambda: _conv_case(1152, 96, 1, 16, 40),
    "c3_256_ps": lambda: _conv_case(96, 256, 3, 17, 35, epilogue=1),
    "c3_final": lambda: _conv_case(64, 3, 3, 26, 44, epilogue=2),
    "c3_32_big": lambda: _conv_case(160, 32, 3, 180, 320, B=2, relu=True, cin_split=96),
    "c3_32_plain": lambda: _conv_case(96, 32, 3, 20, 37, relu=True, variant=1),
    "c3_32_plain_split": lambda: _conv_case(192, 32, 3, 24, 70, relu=True, cin_split=96, variant=1),
}

if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--run":
        CASES[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(CASES)
    summary = {}
    for n in names:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", n], capture_output=True, text=True,
                               timeout=180)
            tail = (r.stdout.strip().splitlines() or [""])[-1]
            summary[n] = {"rc": r.returncode, "out": tail, "err": r.stderr.strip()[-600:] if r.returncode else ""}
        except subprocess.TimeoutExpired:
            summary[n] = {"rc": "timeout"}
        print(n, summary[n], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as fh:
        json.dump(summary, fh, indent=1)
