"""A/B of the conv kernel's MMA-warp schemes at the bench shapes, one configuration per child process (the library reads
its switches once): stage alternation vs M-split, 12 vs 24 MMAs per stage.  Prints the x-stacked RDB conv chain (convs
0..2, 5 x 360 x 640), the 96->96 3x3, the 1x1 GFF.0 and a whole graphed window.  usage: python tools/ab_conv.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from bin_b200 import ops, rdn
    from oracle import bin_oracle as O
    dev = "cuda"
    B, h, w = 5, 360, 640

    def t(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    x = torch.randn(B, 12, h, w, 8, device=dev).half()
    g = torch.randn(B, 16, h, w, 8, device=dev).half()
    res = {"cfg": {k: os.environ.get(k, "") for k in ("BIN_B200_QUAD", "BIN_B200_MSPLIT", "BIN_B200_STAGE_MMAS", "BIN_B200_PAIR", "BIN_B200_SPREAD", "BIN_B200_DEBUG", "BIN_B200_LIB")}}
    tot, fl = 0.0, 0.0
    for c in range(3):
        cin = 96 + 32 * c
        wp = ops.pack_conv_weight(torch.randn(32, cin, 3, 3, device=dev) / (cin * 9) ** 0.5, 32, cin)
        bp = ops.pad_bias(torch.zeros(32, device=dev), 32)
        kw = dict(in0_planes=12, in1=g, in1_planes=4 * c, relu=True, out=g, out_plane0=4 * c)
        tot += t(lambda: ops.conv_fwd(x, wp, bp, 3, 32, **kw))
        fl += 2.0 * B * h * w * cin * 32 * 9
    res["rdb_convs_ms"] = round(tot, 4); res["rdb_convs_tflops"] = round(fl / tot / 1e9, 1)
    w96 = ops.pack_conv_weight(torch.randn(96, 96, 3, 3, device=dev) / 864 ** 0.5, 96, 96)
    b96 = ops.pad_bias(torch.zeros(96, device=dev), 96)
    o96 = torch.empty(B, 12, h, w, 8, device=dev).half()
    ms = t(lambda: ops.conv_fwd(x, w96, b96, 3, 96, in0_planes=12, out=o96))
    res["conv3x3_96_ms"] = round(ms, 4); res["conv3x3_96_tflops"] = round(2.0 * B * h * w * 96 * 96 * 9 / ms / 1e9, 1)
    w256 = ops.pack_conv_weight(torch.randn(256, 96, 3, 3, device=dev) / 864 ** 0.5, 256, 96)
    b256 = ops.pad_bias(torch.zeros(256, device=dev), 256)
    u = torch.empty(B, 8, 2 * h, 2 * w, 8, device=dev).half()
    ms = t(lambda: ops.conv_fwd(x, w256, b256, 3, 256, in0_planes=12, epilogue=1, out=u))
    res["upnet0_ms"] = round(ms, 4); res["upnet0_tflops"] = round(2.0 * B * h * w * 96 * 256 * 9 / ms / 1e9, 1)
    x0 = torch.randn(B, 4, h, w, 8, device=dev).half()
    w5 = ops.pack_conv_weight(torch.randn(96, 24, 5, 5, device=dev) / 600 ** 0.5, 96, 32)
    ms = t(lambda: ops.conv_fwd(x0, w5, b96, 5, 96, in0_planes=4, out=o96))
    res["sfenet1_ms"] = round(ms, 4)
    net = rdn.bin_stage4_lstm(); net.load_state_dict(O.synth_state_dict(0), strict=True); net = net.cuda().eval()
    fr = [f.cuda() for f in O.synth_frames(6, 1, 720, 1280, seed=1234, smooth=True)]
    with torch.no_grad():
        res["window_ms"] = round(t(lambda: net(*fr), reps=10), 3)
    print(json.dumps(res))
else:
    cfgs = [{"BIN_B200_QUAD": "0"}, {"BIN_B200_QUAD": "1"}, {"BIN_B200_QUAD": "0"}, {"BIN_B200_QUAD": "1"}]
    if len(sys.argv) > 1 and sys.argv[1] == "noload":
        # tools library: BIN_B200_DEBUG bit 10 = the x-stacked conv's producers skip every input TMA load, bit 8 = no output
        # stores: the kernel with NO HBM traffic is the upper bound of what fusing convs 0..2 (growth maps kept on chip) could save
        tl = os.path.join(ROOT, "bin_b200", "libbin_b200_tools.so")
        cfgs = [{"BIN_B200_LIB": tl, "BIN_B200_DEBUG": d} for d in ("0", "1024", "1280", "0", "1024", "1280")]
    if len(sys.argv) > 1 and sys.argv[1] == "prevlib":
        # A/B against a library built from the previous sources (bin_b200/libbin_b200_prev.so, built by hand with the same flags)
        pl = os.path.join(ROOT, "bin_b200", "libbin_b200_prev.so")
        cfgs = [{"BIN_B200_LIB": pl}, {}, {"BIN_B200_LIB": pl}, {}]
    if len(sys.argv) > 1 and sys.argv[1] == "spread":
        cfgs = [{"BIN_B200_SPREAD": "0"}, {"BIN_B200_SPREAD": "1"}, {"BIN_B200_SPREAD": "0"}, {"BIN_B200_SPREAD": "1"}]
    for cfg in cfgs:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, **cfg), capture_output=True,
                           text=True, timeout=600)
        print(r.stdout.strip() or r.stderr[-1500:], flush=True)
