// bin_b200 -- extern "C" entry points (include/bin_b200.h) and the host-side orchestration of
// one backbone / one 6-frame window.  Host code only: every arithmetic step is a kernel in
// conv_igemm.cu / aux_kernels.cu.
#include <string.h>

#include <string>
#include <vector>

#include "internal.h"

namespace binb {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

int launch_nchw_to_p8(const float* x, int C, const bin_act_t& dst, int plane0, cudaStream_t s);
int launch_p8_to_nchw(const bin_act_t& src, int plane0, int C, float* y, cudaStream_t s);
int launch_pack_frames(const bin_frames_t& fr, int H, int W, const bin_act_t& dst, cudaStream_t s);
int launch_pack_weight(const float* w, int cout, int cin, int ks, int cout_pad, int cin_pad, int variant,
                       void* packed, cudaStream_t s);
int launch_pack_bias(const float* b, int cout, int cout_pad, float* dst, cudaStream_t s);
int launch_convlstm(const float* x, const float* c_prev, const float* h_prev, const float* w, const float* b,
                    float* h_out, float* c_out, int B, int H, int W, cudaStream_t s);
int run_mma_bench(int n, int iters, int mode, float* cycles_host);

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------ backbone conv table
// Order = nn.Module registration order of the reference backbones (RDN.py:187-208): SFENet1,
// SFENet2, RDBs.{i}.convs.{0..3}, RDBs.{i}.LFF, GFF.0, GFF.1, UPNet.0, UPNet.2.
struct ConvSpec {
  int cin, cout, ks, cin_pad, cout_pad;
  size_t w_off, b_off;
};
constexpr int kG0 = 96, kD = 12, kCgrow = 4, kG = 32;

struct BackboneLayout {
  ConvSpec conv[BIN_BACKBONE_NCONV];
  size_t bytes;
};

static BackboneLayout backbone_layout(int nframes) {
  BackboneLayout L;
  int k = 0;
  auto add = [&](int cin, int cout, int ks, int cout_pad) {
    ConvSpec c;
    c.cin = cin; c.cout = cout; c.ks = ks;
    c.cin_pad = (int)align_up(cin, kKC);
    c.cout_pad = cout_pad;
    c.w_off = c.b_off = 0;
    L.conv[k++] = c;
  };
  add(12 * nframes, kG0, 5, 96);
  add(kG0, kG0, 3, 96);
  for (int i = 0; i < kD; ++i) {
    for (int c = 0; c < kCgrow; ++c) add(kG0 + c * kG, kG, 3, 32);
    add(kG0 + kCgrow * kG, kG0, 1, 96);
  }
  add(kD * kG0, kG0, 1, 96);
  add(kG0, kG0, 3, 96);
  add(kG0, 256, 3, 256);
  add(64, 3, 3, 16);
  size_t off = 0;
  for (int i = 0; i < BIN_BACKBONE_NCONV; ++i) {
    ConvSpec& c = L.conv[i];
    c.w_off = off;
    off = align_up(off + (size_t)c.cout_pad * c.cin_pad * c.ks * c.ks * sizeof(__half), 256);
    c.b_off = off;
    off = align_up(off + (size_t)c.cout_pad * sizeof(float), 256);
  }
  L.bytes = off;
  return L;
}

static bool valid_nframes(int n) { return n == 2 || n == 3 || n == 5; }

// ------------------------------------------------------------------ backbone workspace
struct BackboneWs {
  bin_act_t x0, f1, f2, cat, g, t1, t2, u;
  size_t bytes;
};
static BackboneWs backbone_ws(int nframes, int Btot, int H, int W, void* base) {
  BackboneWs w;
  const int h = H / 2, wd = W / 2;
  size_t off = 0;
  auto carve = [&](int planes, int hh, int ww) {
    bin_act_t t;
    t.ptr = base ? (void*)((uint8_t*)base + off) : nullptr;
    t.B = Btot; t.planes = planes; t.H = hh; t.W = ww;
    off = align_up(off + (size_t)Btot * planes * hh * ww * 16, 256);
    return t;
  };
  w.x0 = carve((int)align_up(12 * nframes, kKC) / 8, h, wd);
  w.f1 = carve(12, h, wd);
  w.f2 = carve(12, h, wd);
  w.cat = carve(12 * kD, h, wd);
  w.g = carve(16, h, wd);
  w.t1 = carve(12, h, wd);
  w.t2 = carve(12, h, wd);
  w.u = carve(8, H, W);
  w.bytes = off;
  return w;
}

static bin_conv_args_t conv_args(const void* blob, const ConvSpec& c) {
  bin_conv_args_t a;
  memset(&a, 0, sizeof(a));
  a.w_packed = (const uint8_t*)blob + c.w_off;
  a.bias = (const float*)((const uint8_t*)blob + c.b_off);
  a.ksize = c.ks;
  a.cout_pad = c.cout_pad;
  return a;
}

// One RDB: 4 x (conv3x3+ReLU -> growth planes) + LFF 1x1 + residual (RDN.py:149-165).
static int run_rdb(const void* blob, const BackboneLayout& L, int i, const bin_act_t& xin, int x_plane0,
                   const bin_act_t& g, const bin_act_t& out, int out_plane0, cudaStream_t s) {
  const int base = 2 + i * (kCgrow + 1);
  for (int c = 0; c < kCgrow; ++c) {
    bin_conv_args_t a = conv_args(blob, L.conv[base + c]);
    a.in0 = xin; a.in0_plane0 = x_plane0; a.in0_planes = 12;
    a.in1 = g; a.in1_plane0 = 0; a.in1_planes = 4 * c;
    a.relu = 1; a.epilogue = BIN_EPI_P8;
    a.out = g; a.out_plane0 = 4 * c;
    BIN_TRY(launch_conv(a, s));
  }
  bin_conv_args_t a = conv_args(blob, L.conv[base + kCgrow]);
  a.in0 = xin; a.in0_plane0 = x_plane0; a.in0_planes = 12;
  a.in1 = g; a.in1_plane0 = 0; a.in1_planes = 16;
  a.epilogue = BIN_EPI_P8;
  a.out = out; a.out_plane0 = out_plane0;
  a.res = xin; a.res_plane0 = x_plane0;
  return launch_conv(a, s);
}

static int run_backbone(int nframes, const void* blob, const bin_frames_t& fr, int H, int W, void* workspace,
                        size_t workspace_bytes, cudaStream_t s) {
  if (!valid_nframes(nframes) || fr.nframes != nframes) return fail(BIN_ERR_ARG, "backbone: nframes must be 2, 3 or 5");
  if (fr.ncalls < 1 || fr.ncalls > BIN_MAX_CALLS || fr.Bc < 1) return fail(BIN_ERR_ARG, "backbone: bad call table");
  if ((H & 1) || (W & 1) || H < 2 || W < 2) return fail(BIN_ERR_ARG, "backbone: H and W must be even (RDN.py:123-128)");
  const int Btot = fr.ncalls * fr.Bc;
  const BackboneLayout L = backbone_layout(nframes);
  const BackboneWs ws = backbone_ws(nframes, Btot, H, W, workspace);
  if (ws.bytes > workspace_bytes) return fail(BIN_ERR_WORKSPACE, "backbone: workspace too small");
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail(BIN_ERR_ARG, "backbone: workspace must be 256-byte aligned");

  BIN_TRY(launch_pack_frames(fr, H, W, ws.x0, s));                               // RDN.py:211
  {
    bin_conv_args_t a = conv_args(blob, L.conv[0]);                              // SFENet1 (RDN.py:212)
    a.in0 = ws.x0; a.in0_planes = ws.x0.planes; a.epilogue = BIN_EPI_P8; a.out = ws.f1;
    BIN_TRY(launch_conv(a, s));
  }
  {
    bin_conv_args_t a = conv_args(blob, L.conv[1]);                              // SFENet2 (RDN.py:213)
    a.in0 = ws.f1; a.in0_planes = 12; a.epilogue = BIN_EPI_P8; a.out = ws.f2;
    BIN_TRY(launch_conv(a, s));
  }
  for (int i = 0; i < kD; ++i) {                                                 // RDN.py:215-217
    if (i == 0) BIN_TRY(run_rdb(blob, L, i, ws.f2, 0, ws.g, ws.cat, 0, s));
    else BIN_TRY(run_rdb(blob, L, i, ws.cat, 12 * (i - 1), ws.g, ws.cat, 12 * i, s));
  }
  {
    bin_conv_args_t a = conv_args(blob, L.conv[62]);                             // GFF.0 on the 1152-ch concat (RDN.py:218)
    a.in0 = ws.cat; a.in0_planes = 12 * kD; a.epilogue = BIN_EPI_P8; a.out = ws.t1;
    BIN_TRY(launch_conv(a, s));
  }
  {
    bin_conv_args_t a = conv_args(blob, L.conv[63]);                             // GFF.1, x += f__1 (RDN.py:219)
    a.in0 = ws.t1; a.in0_planes = 12; a.epilogue = BIN_EPI_P8; a.out = ws.t2; a.res = ws.f1;
    BIN_TRY(launch_conv(a, s));
  }
  {
    bin_conv_args_t a = conv_args(blob, L.conv[64]);                             // UPNet.0 + PixelShuffle (RDN.py:205-206)
    a.in0 = ws.t2; a.in0_planes = 12; a.epilogue = BIN_EPI_PIXSHUF; a.out = ws.u;
    BIN_TRY(launch_conv(a, s));
  }
  {
    bin_conv_args_t a = conv_args(blob, L.conv[65]);                             // UPNet.2 + mean(frames) (RDN.py:207,221)
    a.in0 = ws.u; a.in0_planes = 8; a.epilogue = BIN_EPI_FINAL; a.fr = fr;
    BIN_TRY(launch_conv(a, s));
  }
  return BIN_OK;
}

// ------------------------------------------------------------------ window orchestration
struct Call {
  const float* in[BIN_MAX_FRAMES];
  float* out;
};
static int run_stage(const bin_net_t* net, int which, int nframes, const std::vector<Call>& calls, int B, int H, int W,
                     void* ws, size_t ws_bytes, cudaStream_t s) {
  bin_frames_t fr;
  memset(&fr, 0, sizeof(fr));
  fr.ncalls = (int)calls.size(); fr.nframes = nframes; fr.Bc = B;
  for (int k = 0; k < fr.ncalls; ++k) {
    for (int f = 0; f < nframes; ++f) fr.frame[k][f] = calls[k].in[f];
    fr.out[k] = calls[k].out;
  }
  return run_backbone(nframes, net->blob[which], fr, H, W, ws, ws_bytes, s);
}

static size_t window_ws_bytes(int B, int H, int W, int max_calls, int ntemp) {
  size_t bb = 0;
  const int nf[3] = {2, 3, 5};
  for (int i = 0; i < 3; ++i) {
    size_t v = backbone_ws(nf[i], max_calls * B, H, W, nullptr).bytes;
    bb = v > bb ? v : bb;
  }
  return bb + (size_t)ntemp * align_up((size_t)B * 3 * H * W * sizeof(float), 256);
}

}  // namespace binb

using namespace binb;

extern "C" {

int bin_abi_version(void) { return BIN_ABI_VERSION; }
const char* bin_last_error(void) { return g_err.c_str(); }

int bin_check_device(void) {
  int dev = 0;
  BIN_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  BIN_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(BIN_ERR_UNSUPPORTED, std::string("bin_b200 needs an sm_100 device, found sm_") +
                                                          std::to_string(prop.major) + std::to_string(prop.minor));
  return BIN_OK;
}

int bin_nchw_to_p8(const float* x, int C, bin_act_t dst, int plane0, bin_stream_t s) {
  return launch_nchw_to_p8(x, C, dst, plane0, (cudaStream_t)s);
}
int bin_p8_to_nchw(bin_act_t src, int plane0, int C, float* y, bin_stream_t s) {
  return launch_p8_to_nchw(src, plane0, C, y, (cudaStream_t)s);
}
int bin_pack_frames(const bin_frames_t* fr, int H, int W, bin_act_t dst, bin_stream_t s) {
  if (!fr) return fail(BIN_ERR_ARG, "pack_frames: null frame table");
  return launch_pack_frames(*fr, H, W, dst, (cudaStream_t)s);
}
size_t bin_packed_weight_bytes(int cout_pad, int cin_pad, int ksize) {
  return (size_t)cout_pad * cin_pad * ksize * ksize * sizeof(__half);
}
int bin_pack_conv_weight(const float* w_oihw, int cout, int cin, int ksize, int cout_pad, int cin_pad, int variant,
                         void* packed, bin_stream_t s) {
  return launch_pack_weight(w_oihw, cout, cin, ksize, cout_pad, cin_pad, variant, packed, (cudaStream_t)s);
}
int bin_conv_fwd(const bin_conv_args_t* a, bin_stream_t s) {
  if (!a) return fail(BIN_ERR_ARG, "conv: null args");
  return launch_conv(*a, (cudaStream_t)s);
}
int bin_convlstm_fwd(const float* x, const float* c_prev, const float* h_prev, const float* w, const float* b,
                     float* h_out, float* c_out, int B, int H, int W, bin_stream_t s) {
  return launch_convlstm(x, c_prev, h_prev, w, b, h_out, c_out, B, H, W, (cudaStream_t)s);
}

size_t bin_backbone_packed_bytes(int nframes) { return valid_nframes(nframes) ? backbone_layout(nframes).bytes : 0; }

int bin_backbone_pack(int nframes, const float* const* w_host, const float* const* b_host, void* blob,
                      bin_stream_t s) {
  if (!valid_nframes(nframes)) return fail(BIN_ERR_ARG, "backbone_pack: nframes must be 2, 3 or 5");
  const BackboneLayout L = backbone_layout(nframes);
  for (int i = 0; i < BIN_BACKBONE_NCONV; ++i) {
    const ConvSpec& c = L.conv[i];
    BIN_TRY(launch_pack_weight(w_host[i], c.cout, c.cin, c.ks, c.cout_pad, c.cin_pad, BIN_CONV_DEFAULT,
                               (uint8_t*)blob + c.w_off,
                               (cudaStream_t)s));
    BIN_TRY(launch_pack_bias(b_host[i], c.cout, c.cout_pad, (float*)((uint8_t*)blob + c.b_off), (cudaStream_t)s));
  }
  return BIN_OK;
}

size_t bin_backbone_workspace_bytes(int nframes, int Btot, int H, int W) {
  return valid_nframes(nframes) ? backbone_ws(nframes, Btot, H, W, nullptr).bytes : 0;
}

int bin_backbone_fwd(int nframes, const void* blob, const bin_frames_t* fr, int H, int W, void* workspace,
                     size_t workspace_bytes, bin_stream_t s) {
  if (!fr || !blob) return fail(BIN_ERR_ARG, "backbone_fwd: null argument");
  return run_backbone(nframes, blob, *fr, H, W, workspace, workspace_bytes, (cudaStream_t)s);
}

int bin_rdb_fwd(const void* blob, int nframes, int index, const float* x, float* y, int B, int h, int w,
                void* workspace, size_t workspace_bytes, bin_stream_t s) {
  if (!valid_nframes(nframes) || index < 0 || index >= kD) return fail(BIN_ERR_ARG, "rdb_fwd: bad nframes/index");
  const BackboneLayout L = backbone_layout(nframes);
  size_t off = 0;
  auto carve = [&](int planes) {
    bin_act_t t;
    t.ptr = (uint8_t*)workspace + off; t.B = B; t.planes = planes; t.H = h; t.W = w;
    off = align_up(off + (size_t)B * planes * h * w * 16, 256);
    return t;
  };
  bin_act_t xin = carve(12), g = carve(16), out = carve(12);
  if (off > workspace_bytes) return fail(BIN_ERR_WORKSPACE, "rdb_fwd: workspace too small");
  BIN_TRY(launch_nchw_to_p8(x, kG0, xin, 0, (cudaStream_t)s));
  BIN_TRY(run_rdb(blob, L, index, xin, 0, g, out, 0, (cudaStream_t)s));
  return launch_p8_to_nchw(out, 0, kG0, y, (cudaStream_t)s);
}

size_t bin_window_workspace_bytes(int B, int H, int W) { return window_ws_bytes(B, H, W, BIN_MAX_CALLS, 9); }

int bin_window_fwd(const bin_net_t* net, const float* const* F, float* const* o, int B, int H, int W, void* workspace,
                   size_t workspace_bytes, bin_stream_t s_) {
  if (!net || !F || !o) return fail(BIN_ERR_ARG, "window_fwd: null argument");
  cudaStream_t s = (cudaStream_t)s_;
  const size_t need = window_ws_bytes(B, H, W, BIN_MAX_CALLS, 9);
  if (workspace_bytes < need) return fail(BIN_ERR_WORKSPACE, "window_fwd: workspace too small");
  const size_t fbytes = align_up((size_t)B * 3 * H * W * sizeof(float), 256);
  uint8_t* base = (uint8_t*)workspace;
  float* tmp[9];
  for (int i = 0; i < 9; ++i) tmp[i] = (float*)(base + i * fbytes);
  void* bws = base + 9 * fbytes;
  const size_t bws_bytes = workspace_bytes - 9 * fbytes;
  float *p4 = tmp[0], *p6 = tmp[1], *p8 = tmp[2], *p5 = tmp[3], *p7 = tmp[4], *p6b = tmp[5];
  float *t0 = tmp[6], *t1 = tmp[7], *t2 = tmp[8];
  auto lstm = [&](int k, const float* x, float* h) {
    return launch_convlstm(x, nullptr, nullptr, net->lstm_w[k], net->lstm_b[k], h, nullptr, B, H, W, s);
  };
  // Stage 1 (RDN.py:371-374): 4 calls of step 0 + the one stage-1 call of step 1 that is not a repeat.
  BIN_TRY(run_stage(net, 0, 2, {{{F[0], F[1]}, o[0]}, {{F[1], F[2]}, o[1]}, {{F[2], F[3]}, o[2]},
                               {{F[3], F[4]}, o[3]}, {{F[4], F[5]}, o[10]}}, B, H, W, bws, bws_bytes, s));
  // recurrent hand-off for the stage-1 outputs (RDN.py:451-453)
  BIN_TRY(lstm(0, o[1], p4)); BIN_TRY(lstm(1, o[2], p6)); BIN_TRY(lstm(2, o[3], p8));
  // Stage 2: step 0 (RDN.py:384-386, "prev" slot duplicated) + step 1 (RDN.py:377-379)
  BIN_TRY(run_stage(net, 1, 3, {{{o[0], o[0], o[1]}, o[4]}, {{o[1], o[1], o[2]}, o[5]}, {{o[2], o[2], o[3]}, o[6]}},
                    B, H, W, bws, bws_bytes, s));
  BIN_TRY(run_stage(net, 1, 3, {{{p4, o[1], o[2]}, t0}, {{p6, o[2], o[3]}, t1}, {{p8, o[3], o[10]}, o[11]}},
                    B, H, W, bws, bws_bytes, s));
  BIN_TRY(lstm(3, o[5], p5)); BIN_TRY(lstm(4, o[6], p7));                         // RDN.py:454-455
  // Stage 3: step 0 (RDN.py:387-388) + step 1 (RDN.py:380-381)
  BIN_TRY(run_stage(net, 2, 5, {{{o[4], F[1], o[4], o[5], F[2]}, o[7]}, {{o[5], F[2], o[5], o[6], F[3]}, o[8]},
                               {{p5, F[2], t0, t1, F[3]}, t2}, {{p7, F[3], t1, o[11], F[4]}, o[12]}},
                    B, H, W, bws, bws_bytes, s));
  BIN_TRY(lstm(5, o[8], p6b));                                                    // RDN.py:456
  // Stage 4: step 0 (RDN.py:389) + step 1 (RDN.py:382)
  BIN_TRY(run_stage(net, 3, 5, {{{o[1], o[1], o[7], o[8], o[2]}, o[9]}, {{p6b, o[2], t2, o[12], o[3]}, o[13]}},
                    B, H, W, bws, bws_bytes, s));
  return BIN_OK;
}

int bin_pyramid3_fwd(const bin_net_t* net, const float* const* F, float* const* o, int B, int H, int W,
                     void* workspace, size_t workspace_bytes, bin_stream_t s_) {
  if (!net || !F || !o) return fail(BIN_ERR_ARG, "pyramid3_fwd: null argument");
  cudaStream_t s = (cudaStream_t)s_;
  if (workspace_bytes < window_ws_bytes(B, H, W, BIN_MAX_CALLS, 9)) return fail(BIN_ERR_WORKSPACE, "pyramid3_fwd: workspace too small");
  BIN_TRY(run_stage(net, 0, 2, {{{F[0], F[1]}, o[0]}, {{F[1], F[2]}, o[1]}, {{F[2], F[3]}, o[2]}}, B, H, W, workspace,
                    workspace_bytes, s));
  BIN_TRY(run_stage(net, 1, 3, {{{o[0], o[0], o[1]}, o[3]}, {{o[1], o[1], o[2]}, o[4]}}, B, H, W, workspace,
                    workspace_bytes, s));
  return run_stage(net, 2, 5, {{{o[3], F[1], o[3], o[4], F[2]}, o[5]}}, B, H, W, workspace, workspace_bytes, s);
}

int bin_microbench_mma(int n, int iters, int mode, float* cycles_host) { return run_mma_bench(n, iters, mode, cycles_host); }

}  // extern "C"
