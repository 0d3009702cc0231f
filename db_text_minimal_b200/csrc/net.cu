// net.cu -- whole-network executor: DBTextModel forward / backward as one fixed graph of sm_100a kernels.
//
// Replaces src/models.py:34-48 (DBTextModel.forward) together with src/modules/resnet.py:231-242 (ResNet-18),
// src/modules/segmentation_body.py:64-87 (FPN) and src/modules/segmentation_head.py:35-45 (DBHead), and the
// autograd backward PyTorch would derive for them.  One C call enqueues the whole pass on the caller's stream:
// no per-op Python dispatch, no host synchronisation, CUDA-graph capturable.
//
// Data layout in HBM: module boundary tensors keep the reference's layout (x NCHW float32, out NCHW float32,
// parameters float32 in state_dict shapes); every internal activation is NHWC bf16 so that a pixel's channels are one
// contiguous TMA row.  All buffers live in a caller-owned workspace (bump-allocated at dbb_net_create).
// BatchNorm (training): raw conv output z (bf16) -> bn_stats -> finalize (scale/shift, running stats) -> bn_apply
// (+residual)(+ReLU) -> activation; backward mirrors it (reduce, finalize, apply).
#include "common.cuh"
#include "conv_ops.h"
#include "elementwise.h"
#include "head_tail.h"
#include <string>
#include <vector>
#include <map>
#include <type_traits>

namespace dbb {

struct PInfo { std::string name; int numel; };
static std::vector<PInfo> g_params, g_buffers;
static std::map<std::string, int> g_pidx, g_bidx;

static void add_p(const std::string& n, int numel) { g_pidx[n] = (int)g_params.size(); g_params.push_back({n, numel}); }
static void add_bn(const std::string& n, int c) {
  add_p(n + ".weight", c); add_p(n + ".bias", c);
  g_bidx[n + ".running_mean"] = (int)g_buffers.size(); g_buffers.push_back({n + ".running_mean", c});
  g_bidx[n + ".running_var"] = (int)g_buffers.size(); g_buffers.push_back({n + ".running_var", c});
}
static void build_tables() {
  if (!g_params.empty()) return;
  add_p("backbone.conv1.weight", 64 * 3 * 49);
  add_bn("backbone.bn1", 64);
  int inpl = 64;
  const int planes[4] = {64, 128, 256, 512};
  for (int li = 1; li <= 4; ++li)
    for (int bi = 0; bi < 2; ++bi) {
      const std::string pre = "backbone.layer" + std::to_string(li) + "." + std::to_string(bi);
      const int pl = planes[li - 1];
      const bool ds = (bi == 0 && li > 1);
      add_p(pre + ".conv1.weight", pl * inpl * 9); add_bn(pre + ".bn1", pl);
      add_p(pre + ".conv2.weight", pl * pl * 9); add_bn(pre + ".bn2", pl);
      if (ds) { add_p(pre + ".downsample.0.weight", pl * inpl); add_bn(pre + ".downsample.1", pl); }
      inpl = pl;
    }
  // present in the state dict, never used in forward: no gradient
  add_p("backbone.fc.weight", 1000 * 512); add_p("backbone.fc.bias", 1000);
  add_p("backbone.smooth.weight", 256 * 2048); add_p("backbone.smooth.bias", 256);
  const char* cn[4] = {"c2", "c3", "c4", "c5"};
  for (int i = 0; i < 4; ++i) {
    const std::string pre = std::string("segmentation_body.reduce_conv_") + cn[i];
    add_p(pre + ".conv.weight", 64 * planes[i]); add_p(pre + ".conv.bias", 64); add_bn(pre + ".bn", 64);
  }
  const char* pn[3] = {"p4", "p3", "p2"};
  for (int i = 0; i < 3; ++i) {
    const std::string pre = std::string("segmentation_body.smooth_") + pn[i];
    add_p(pre + ".conv.weight", 64 * 64 * 9); add_p(pre + ".conv.bias", 64); add_bn(pre + ".bn", 64);
  }
  add_p("segmentation_body.conv.0.weight", 256 * 256 * 9); add_p("segmentation_body.conv.0.bias", 256);
  add_bn("segmentation_body.conv.1", 256);
  for (int br = 0; br < 2; ++br) {
    const std::string pre = std::string("segmentation_head.") + (br ? "thresh" : "binarize");
    add_p(pre + ".0.weight", 64 * 256 * 9);
    if (br == 0) add_p(pre + ".0.bias", 64);      // thresh.0 has no bias (segmentation_head.py:64-68)
    add_bn(pre + ".1", 64);
    add_p(pre + ".3.weight", 64 * 64 * 4); add_p(pre + ".3.bias", 64);
    add_bn(pre + ".4", 64);
    add_p(pre + ".6.weight", 64 * 4); add_p(pre + ".6.bias", 1);
  }
}

// ------------------------------------------------------------------------------------------------
struct Buf { size_t off = 0; size_t bytes = 0; };

struct ConvBN {          // Conv2d (+bias) -> BatchNorm2d (-> ReLU handled by the caller's bn_apply flags)
  ConvGeom g{};
  int w = -1, b = -1, gamma = -1, beta = -1, rm = -1, rv = -1;   // param / buffer indices
  Buf wp, wpt;           // packed bf16 weights: fprop, dgrad
  Buf z, stats, coef, dz;
  int64_t P() const { return (int64_t)g.n * g.out_h() * g.out_w(); }
};

struct Block {
  ConvBN c1, c2, ds;
  bool has_ds = false;
  Buf a1, rd, out, d_a1, d_out;   // d_out: gradient w.r.t. the block output (allocated by the consumer side)
  int h_in, w_in, c_in;
};

}  // namespace dbb

using namespace dbb;

struct DbbNet {
  int n, h, w, training;
  int fp32 = 0;                          // DBB_PRECISION_FP32: float activations + CUDA-core convolutions (parity mode)
  size_t esz = sizeof(bf16);             // bytes per activation element
  int h1, w1, h2, w2, hh[4], ww[4];      // conv1 out, pool out (= c2), c2..c5 extents
  int ho, wo;                            // head output extent (4*h2, 4*w2)
  size_t ws_bytes = 0;
  uint64_t flops_fwd = 0;
  // stem
  Buf s2d, wp_conv1, z0, stats0, coef0, a0, x1, argmax, d_x1, d_a0, d_z0, dw_s2d;
  Block blocks[8];
  // FPN
  ConvBN lat[4];          // reduce_conv_c2..c5
  ConvBN smooth[3];       // smooth_p4, p3, p2   (index 0 -> p4)
  Buf l_act[4];           // lateral activations (l2, l3, l4, p5)
  Buf s_sum[3];           // s4, s3, s2 (inputs of the smooth convs)
  Buf p_act[2];           // p4, p3 (p2 lives in cat[..., 0:64])
  Buf cat, d_cat;
  ConvBN fconv;           // segmentation_body.conv
  Buf af, d_af;
  Buf d_p[3];             // d_p5, d_p4, d_p3
  Buf d_s[3];             // d_s4, d_s3, d_s2
  Buf d_c[4];             // gradients w.r.t. c2..c5
  // head
  ConvGeom hconv_g{}, tconv_g{};
  Buf wp_h, wpt_h, bias_h, zh, stats_h, coef_h, ah, d_ah, d_zh, dw_h, dbias_h;
  Buf wp_t[2], wpt_t[2], zt, stats_t, coef_t, d_zt, dbias_t;
  Buf head_out;           // fp32 (N, C, ho, wo) when a final bilinear resize is needed
  Buf d_head_out;
  Buf partials;           // shared reduction scratch
  Buf wg_scratch;         // split-K partials of the weight-gradient GEMMs
  Buf bn_acc;             // fp64 accumulators + ticket of the fused reduce+finalize kernels (kept zero between uses)
  int out_c;
  // side stream of the weight-gradient chain (backward): wgrad GEMMs are independent of the dgrad / BatchNorm chain of the
  // same layer, so they run concurrently with the (memory-bound) BatchNorm-backward kernels of the following layers
  cudaStream_t s2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // bump allocator
  size_t cur = 0;
  Buf alloc(size_t bytes) {
    Buf b; b.off = cur; b.bytes = bytes;
    cur += (bytes + 1023) / 1024 * 1024;   // 1 KB alignment (TMA needs 16 B; swizzled smem is unaffected)
    return b;
  }
};

static int P(const char* name) {
  auto it = g_pidx.find(name);
  return it == g_pidx.end() ? -1 : it->second;
}
static int P(const std::string& name) { return P(name.c_str()); }
static int B(const std::string& name) {
  auto it = g_bidx.find(name);
  return it == g_bidx.end() ? -1 : it->second;
}

static void setup_convbn(DbbNet* net, ConvBN& L, const std::string& conv, const std::string& bn, int n, int h, int w, int cin,
                         int cout, int ks, int stride, int pad, bool has_bias, bool need_dgrad) {
  L.g = ConvGeom{n, h, w, cin, cout, ks, stride, pad};
  L.w = P(conv + ".weight");
  L.b = has_bias ? P(conv + ".bias") : -1;
  L.gamma = P(bn + ".weight"); L.beta = P(bn + ".bias");
  L.rm = B(bn + ".running_mean"); L.rv = B(bn + ".running_var");
  const size_t wbytes = (size_t)cin * cout * ks * ks * sizeof(bf16);
  L.wp = net->alloc(wbytes);
  if (net->training && need_dgrad) L.wpt = net->alloc(wbytes);
  L.z = net->alloc((size_t)L.P() * cout * net->esz);
  L.stats = net->alloc(4 * cout * sizeof(float));
  if (net->training) {
    L.coef = net->alloc(3 * cout * sizeof(float));
    L.dz = net->alloc((size_t)L.P() * cout * net->esz);
  }
  net->flops_fwd += 2ull * (uint64_t)L.P() * cout * cin * ks * ks;
}

extern "C" DbbNet* dbb_net_create_ex(int64_t n, int64_t h, int64_t w, int training, int precision) {
  build_tables();
  if (n <= 0 || h < 32 || w < 32) { set_error(DBB_EINVAL, "net_create: need n >= 1 and h, w >= 32"); return nullptr; }
  if (precision != DBB_PRECISION_BF16 && precision != DBB_PRECISION_FP32) { set_error(DBB_EINVAL, "net_create: unknown precision"); return nullptr; }
  DbbNet* net = new DbbNet();
  net->n = (int)n; net->h = (int)h; net->w = (int)w; net->training = training ? 1 : 0;
  net->fp32 = precision == DBB_PRECISION_FP32;
  net->esz = net->fp32 ? sizeof(float) : sizeof(bf16);
  net->out_c = training ? 3 : 2;
  const int N = (int)n;
  const bool T = net->training;
  auto act_bytes = [&](int hh, int ww, int c) { return (size_t)N * hh * ww * c * net->esz; };
  // ---- stem
  net->h1 = ((int)h + 1) / 2; net->w1 = ((int)w + 1) / 2;                 // conv 7x7/2 pad 3
  net->h2 = (net->h1 + 2 - 3) / 2 + 1; net->w2 = (net->w1 + 2 - 3) / 2 + 1;   // maxpool 3/2 pad 1
  net->s2d = net->alloc((size_t)N * (net->h1 + 3) * (net->w1 + 3) * 16 * sizeof(bf16));
  net->wp_conv1 = net->alloc(64 * 256 * sizeof(bf16));
  net->z0 = net->alloc(act_bytes(net->h1, net->w1, 64));
  net->stats0 = net->alloc(4 * 64 * sizeof(float));
  net->a0 = net->alloc(act_bytes(net->h1, net->w1, 64));
  net->x1 = net->alloc(act_bytes(net->h2, net->w2, 64));
  net->flops_fwd += 2ull * N * net->h1 * net->w1 * 64 * 147;
  if (T) {
    net->coef0 = net->alloc(3 * 64 * sizeof(float));
    net->argmax = net->alloc((size_t)N * net->h2 * net->w2 * 64);
    net->d_x1 = net->alloc(act_bytes(net->h2, net->w2, 64));
    net->d_a0 = net->alloc(act_bytes(net->h1, net->w1, 64));
    net->d_z0 = net->alloc(act_bytes(net->h1, net->w1, 64));
    net->dw_s2d = net->alloc(64 * 64 * 4 * sizeof(float));
  }
  // ---- residual stages
  int hin = net->h2, win = net->w2, cin = 64;
  const int planes[4] = {64, 128, 256, 512};
  for (int li = 0; li < 4; ++li)
    for (int bi = 0; bi < 2; ++bi) {
      Block& bk = net->blocks[li * 2 + bi];
      const std::string pre = "backbone.layer" + std::to_string(li + 1) + "." + std::to_string(bi);
      const int pl = planes[li];
      const int stride = (li > 0 && bi == 0) ? 2 : 1;
      bk.h_in = hin; bk.w_in = win; bk.c_in = cin;
      bk.has_ds = (stride != 1 || cin != pl);
      // the very first block's input (pool output) needs a data gradient too (it flows on into conv1's BN)
      setup_convbn(net, bk.c1, pre + ".conv1", pre + ".bn1", N, hin, win, cin, pl, 3, stride, 1, false, true);
      const int ho = bk.c1.g.out_h(), wo = bk.c1.g.out_w();
      setup_convbn(net, bk.c2, pre + ".conv2", pre + ".bn2", N, ho, wo, pl, pl, 3, 1, 1, false, true);
      if (bk.has_ds) setup_convbn(net, bk.ds, pre + ".downsample.0", pre + ".downsample.1", N, hin, win, cin, pl, 1, stride, 0, false, true);
      bk.a1 = net->alloc(act_bytes(ho, wo, pl));
      if (bk.has_ds) bk.rd = net->alloc(act_bytes(ho, wo, pl));
      bk.out = net->alloc(act_bytes(ho, wo, pl));
      if (T) { bk.d_a1 = net->alloc(act_bytes(ho, wo, pl)); bk.d_out = net->alloc(act_bytes(ho, wo, pl)); }
      hin = ho; win = wo; cin = pl;
      if (bi == 1) { net->hh[li] = ho; net->ww[li] = wo; }
    }
  // ---- FPN
  const char* cn[4] = {"c2", "c3", "c4", "c5"};
  for (int i = 0; i < 4; ++i) {
    const std::string pre = std::string("segmentation_body.reduce_conv_") + cn[i];
    setup_convbn(net, net->lat[i], pre + ".conv", pre + ".bn", N, net->hh[i], net->ww[i], planes[i], 64, 1, 1, 0, true, true);
    net->l_act[i] = net->alloc(act_bytes(net->hh[i], net->ww[i], 64));
  }
  const char* pn[3] = {"p4", "p3", "p2"};
  for (int i = 0; i < 3; ++i) {
    const int lvl = 2 - i;   // p4 -> level index 2 (c4), p3 -> 1, p2 -> 0
    const std::string pre = std::string("segmentation_body.smooth_") + pn[i];
    setup_convbn(net, net->smooth[i], pre + ".conv", pre + ".bn", N, net->hh[lvl], net->ww[lvl], 64, 64, 3, 1, 1, true, true);
    net->s_sum[i] = net->alloc(act_bytes(net->hh[lvl], net->ww[lvl], 64));
    if (i < 2) net->p_act[i] = net->alloc(act_bytes(net->hh[lvl], net->ww[lvl], 64));
    if (T) net->d_s[i] = net->alloc(act_bytes(net->hh[lvl], net->ww[lvl], 64));
  }
  const int hf = net->hh[0], wf = net->ww[0];
  net->cat = net->alloc(act_bytes(hf, wf, 256));
  setup_convbn(net, net->fconv, "segmentation_body.conv.0", "segmentation_body.conv.1", N, hf, wf, 256, 256, 3, 1, 1, true, true);
  net->af = net->alloc(act_bytes(hf, wf, 256));
  if (T) {
    net->d_cat = net->alloc(act_bytes(hf, wf, 256));
    net->d_af = net->alloc(act_bytes(hf, wf, 256));
    net->d_p[0] = net->alloc(act_bytes(net->hh[3], net->ww[3], 64));   // d_p5
    net->d_p[1] = net->alloc(act_bytes(net->hh[2], net->ww[2], 64));   // d_p4
    net->d_p[2] = net->alloc(act_bytes(net->hh[1], net->ww[1], 64));   // d_p3
    for (int i = 0; i < 4; ++i) net->d_c[i] = net->blocks[i * 2 + 1].d_out;   // gradient of c2..c5 = d_out of the stage's last block
  }
  // ---- head: the two 3x3 convs share their input -> one 256->128 GEMM; the two ConvT(64,64,2,2) write one 128-wide tensor
  net->hconv_g = ConvGeom{N, hf, wf, 256, 128, 3, 1, 1};
  net->tconv_g = ConvGeom{N, hf, wf, 64, 64, 2, 2, 0};
  net->wp_h = net->alloc(128 * 256 * 9 * net->esz);      // fp32 mode: the two OIHW tensors side by side
  net->bias_h = net->alloc(128 * sizeof(float));
  net->zh = net->alloc(act_bytes(hf, wf, 128));
  net->stats_h = net->alloc(4 * 128 * sizeof(float));
  net->ah = net->alloc(act_bytes(hf, wf, 128));
  for (int br = 0; br < 2; ++br) net->wp_t[br] = net->alloc(4 * 64 * 64 * sizeof(bf16));
  net->zt = net->alloc(act_bytes(2 * hf, 2 * wf, 128));
  net->stats_t = net->alloc(4 * 128 * sizeof(float));
  net->flops_fwd += 2ull * N * hf * wf * 128 * 256 * 9 + 2ull * 2 * N * hf * wf * 64 * 256 + 2ull * 2 * N * 4 * hf * wf * 64 * 4;
  net->ho = 4 * hf; net->wo = 4 * wf;
  if (net->ho != net->h || net->wo != net->w) net->head_out = net->alloc((size_t)N * net->out_c * net->ho * net->wo * sizeof(float));
  if (T) {
    net->wpt_h = net->alloc(128 * 256 * 9 * sizeof(bf16));
    net->coef_h = net->alloc(3 * 128 * sizeof(float));
    net->d_ah = net->alloc(act_bytes(hf, wf, 128));
    net->d_zh = net->alloc(act_bytes(hf, wf, 128));
    net->dw_h = net->alloc(128 * 256 * 9 * sizeof(float));
    net->dbias_h = net->alloc(128 * sizeof(float));
    for (int br = 0; br < 2; ++br) net->wpt_t[br] = net->alloc(4 * 64 * 64 * sizeof(bf16));
    net->coef_t = net->alloc(3 * 128 * sizeof(float));
    net->d_zt = net->alloc(act_bytes(2 * hf, 2 * wf, 128));
    net->dbias_t = net->alloc(128 * sizeof(float));
    if (net->head_out.bytes) net->d_head_out = net->alloc((size_t)N * 3 * net->ho * net->wo * sizeof(float));
  }
  size_t pf = bn_partials_floats(512);
  if (head_tail_partials_floats() > pf) pf = head_tail_partials_floats();
  net->partials = net->alloc(pf * sizeof(float));
  net->bn_acc = net->alloc(BN_ACC_BYTES);
  if (T) net->wg_scratch = net->alloc(WGRAD_SCRATCH_BYTES);
  net->ws_bytes = net->cur;
  return net;
}

extern "C" DbbNet* dbb_net_create(int64_t n, int64_t h, int64_t w, int training) {
  return dbb_net_create_ex(n, h, w, training, DBB_PRECISION_BF16);
}
extern "C" int dbb_net_precision(const DbbNet* net) { return net ? (net->fp32 ? DBB_PRECISION_FP32 : DBB_PRECISION_BF16) : -1; }

extern "C" void dbb_net_destroy(DbbNet* net) {
  if (!net) return;
  if (net->ev_fork) cudaEventDestroy(net->ev_fork);
  if (net->ev_join) cudaEventDestroy(net->ev_join);
  if (net->s2) cudaStreamDestroy(net->s2);
  delete net;
}
extern "C" size_t dbb_net_workspace_bytes(const DbbNet* net) { return net ? net->ws_bytes : 0; }
extern "C" int64_t dbb_net_out_channels(const DbbNet* net) { return net ? net->out_c : 0; }
extern "C" uint64_t dbb_net_flops_fwd(const DbbNet* net) { return net ? net->flops_fwd : 0; }
extern "C" int dbb_net_num_params(void) { build_tables(); return (int)g_params.size(); }
extern "C" const char* dbb_net_param_name(int i) { build_tables(); return (i >= 0 && i < (int)g_params.size()) ? g_params[i].name.c_str() : nullptr; }
extern "C" int dbb_net_param_numel(int i) { build_tables(); return (i >= 0 && i < (int)g_params.size()) ? g_params[i].numel : -1; }
extern "C" int dbb_net_num_buffers(void) { build_tables(); return (int)g_buffers.size(); }
extern "C" const char* dbb_net_buffer_name(int i) { build_tables(); return (i >= 0 && i < (int)g_buffers.size()) ? g_buffers[i].name.c_str() : nullptr; }
extern "C" int dbb_net_buffer_numel(int i) { build_tables(); return (i >= 0 && i < (int)g_buffers.size()) ? g_buffers[i].numel : -1; }
extern "C" int dbb_net_num_segments(void) { return 3; }

namespace {

constexpr float BN_EPS = 1e-5f, BN_MOM = 0.1f, STEP_K = 50.f;

template <typename T>
struct Ctx {
  DbbNet* net;
  char* base;
  const float* const* params;
  float* const* buffers;
  float* const* grads;
  cudaStream_t s;
  cudaStream_t sw;        // weight-gradient stream (== s when the side stream is disabled)
  template <typename U = T> U* p(const Buf& b) const { return reinterpret_cast<U*>(base + b.off); }
  const float* par(int i) const { return i >= 0 ? params[i] : nullptr; }
  float* buf(int i) const { return (i >= 0 && buffers) ? buffers[i] : nullptr; }
  float* grad(int i) const { return (i >= 0 && grads) ? grads[i] : nullptr; }
  float* partials() const { return p<float>(net->partials); }
  float* wgs() const { return p<float>(net->wg_scratch); }
  double* acc() const { return this->template p<double>(net->bn_acc); }
  unsigned* ticket() const { return reinterpret_cast<unsigned*>(base + net->bn_acc.off + 2 * 2048 * sizeof(double)); }
};

#define RC(call) do { int rc__ = (call); if (rc__) return rc__; } while (0)

// The side stream of a plan (created on first use, never while the caller's stream is being captured); the caller's own
// stream when it is disabled (DBB_NO_WGRAD_STREAM) or while per-kernel profiling is on (CUDA-event times of overlapping
// kernels would be inflated).
cudaStream_t side_stream(DbbNet* net, cudaStream_t stream) {
  static const bool side = getenv("DBB_NO_WGRAD_STREAM") == nullptr;     // A/B switch
  if (!side) return stream;
  if (!net->s2) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone) {
      if (cudaStreamCreateWithFlags(&net->s2, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&net->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&net->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        if (net->s2) { cudaStreamDestroy(net->s2); net->s2 = nullptr; }
      }
    } else {
      cudaGetLastError();
    }
  }
  return (net->s2 && !prof_enabled()) ? net->s2 : stream;
}

// fork: work enqueued on c.sw from here on sees everything enqueued on c.s so far; join: the reverse
template <typename T>
int fork_w(const Ctx<T>& c) {
  if (c.sw == c.s) return 0;
  DBB_CUDA(cudaEventRecord(c.net->ev_fork, c.s));
  DBB_CUDA(cudaStreamWaitEvent(c.sw, c.net->ev_fork, 0));
  return 0;
}
template <typename T>
int join_w(const Ctx<T>& c) {
  if (c.sw == c.s) return 0;
  DBB_CUDA(cudaEventRecord(c.net->ev_join, c.sw));
  DBB_CUDA(cudaStreamWaitEvent(c.s, c.net->ev_join, 0));
  return 0;
}

// BatchNorm statistics of a raw conv output (training: one fused reduce+finalize launch) or running statistics (eval)
// -> stats4.  nseg = 2 for the 128-channel head tensors that carry two BatchNorm layers side by side.
struct BnIdx { int gamma, beta, rm, rv; };
template <typename T>
int bn_prepare(const Ctx<T>& c, const ND<T>* z, int64_t Pn, int ch, int nseg, const BnIdx* ix, float* stats4) {
  const int cn = ch / nseg;
  if (c.net->training) {
    BnFin fin;
    fin.nseg = nseg; fin.momentum = BN_MOM; fin.eps = BN_EPS; fin.stats4 = stats4;
    for (int i = 0; i < nseg; ++i)
      fin.seg[i] = BnFinSeg{c.par(ix[i].gamma), c.par(ix[i].beta), c.buf(ix[i].rm), c.buf(ix[i].rv), i * cn, cn};
    RC(bn_stats_finalize(z, Pn, ch, fin, c.acc(), c.ticket(), c.s));
  } else {
    for (int i = 0; i < nseg; ++i)
      RC(bn_finalize_eval(ch, i * cn, cn, c.par(ix[i].gamma), c.par(ix[i].beta), c.buf(ix[i].rm), c.buf(ix[i].rv), BN_EPS, stats4, c.s));
  }
  return 0;
}
// Training mode: the statistics ride in the producing convolution's epilogue (ConvStats) and no separate pass is
// launched.  Returns the descriptor to hand to conv_fprop, or nullptr (eval mode / DBB_NO_FUSED_STATS) in which case
// the caller runs bn_prepare after the convolution.  Segment i covers channels [coff0 + i*cn, +cn) of the y tensor.
template <typename T>
const ConvStats* bn_fused(const Ctx<T>& c, ConvStats& st, int cn, int nseg, const BnIdx* ix, float* stats4, int coff0 = 0) {
  static const bool off = getenv("DBB_NO_FUSED_STATS") != nullptr;     // A/B switch
  if (!c.net->training || off || c.net->fp32) return nullptr;
  memset(&st, 0, sizeof(st));
  st.enabled = 1; st.gacc = c.acc(); st.counter = c.ticket();
  st.fin.nseg = nseg; st.fin.momentum = BN_MOM; st.fin.eps = BN_EPS; st.fin.stats4 = stats4;
  for (int i = 0; i < nseg; ++i)
    st.fin.seg[i] = BnFinSeg{c.par(ix[i].gamma), c.par(ix[i].beta), c.buf(ix[i].rm), c.buf(ix[i].rv), coff0 + i * cn, cn};
  return &st;
}
template <typename T>
int bn_prepare(const Ctx<T>& c, const ND<T>* z, int64_t Pn, int ch, int gamma, int beta, int rm, int rv, float* stats4) {
  const BnIdx ix{gamma, beta, rm, rv};
  return bn_prepare(c, z, Pn, ch, 1, &ix, stats4);
}

// every weight tensor -> its T GEMM operand(s), in one or two launches at the start of the forward pass
// (fprop layout always; the transposed dgrad layout too in training mode, so backward launches no packing at all)
// stage 0: the stem weights, on the caller's stream; stage 1: everything else, on the side stream (it runs next to
// image_to_s2d / conv1 / max-pool and is joined before layer1)
// weight operand of a convolution: the packed bf16 GEMM matrix (product path) or the raw float32 parameter (fp32 mode)
inline const bf16* wsel(const Ctx<bf16>& c, const Buf& packed, int) { return c.p(packed); }
inline const float* wsel(const Ctx<float>& c, const Buf&, int param) { return c.par(param); }

template <typename T>
int pack_all(const Ctx<T>& c, int stage) {
  DbbNet* net = c.net;
  if constexpr (std::is_same<T, float>::value) {
    // fp32 mode: convolutions read the parameters directly; only the two head 3x3 weights are laid side by side
    // (one 256 -> 128 convolution for both branches, as on the product path)
    if (stage == 0) return 0;
    const size_t half = (size_t)64 * 256 * 9;
    for (int br = 0; br < 2; ++br) {
      const std::string pre = std::string("segmentation_head.") + (br ? "thresh" : "binarize");
      DBB_CUDA(cudaMemcpyAsync(c.p(net->wp_h) + br * half, c.par(P(pre + ".0.weight")), half * sizeof(float), cudaMemcpyDeviceToDevice, c.sw));
    }
    return 0;
  } else {
  PackBatch b;
  b.njobs = 0;
  cudaStream_t ps = stage == 0 ? c.s : c.sw;
  auto flush = [&]() -> int { int rc = pack_weights_batch(b, ps); b.njobs = 0; return rc; };
  auto add = [&](int mode, const float* w, T* out, int co, int ci, int ks, int co_total, int co_off, T* out2 = nullptr) -> int {
    if (b.njobs == PACK_BATCH) RC(flush());
    b.jobs[b.njobs++] = PackJob{w, out, mode, co, ci, ks, ks, co_total > 0 ? co_total : co, co_off, out2};
    return 0;
  };
  auto unit = [&](ConvBN& L) -> int {     // one tiled job writes the fprop layout and (training) the transposed dgrad layout
    return add(5, c.par(L.w), c.p(L.wp), L.g.cout, L.g.cin, L.g.ks, 0, 0, (net->training && L.wpt.bytes) ? c.p(L.wpt) : nullptr);
  };
  if (stage == 0) {
    RC(add(4, c.par(P("backbone.conv1.weight")), c.p(net->wp_conv1), 64, 3, 7, 0, 0));
    return flush();
  }
  for (int i = 0; i < 8; ++i) {
    RC(unit(net->blocks[i].c1)); RC(unit(net->blocks[i].c2));
    if (net->blocks[i].has_ds) RC(unit(net->blocks[i].ds));
  }
  for (int i = 0; i < 4; ++i) RC(unit(net->lat[i]));
  for (int i = 0; i < 3; ++i) RC(unit(net->smooth[i]));
  RC(unit(net->fconv));
  for (int br = 0; br < 2; ++br) {
    const std::string pre = std::string("segmentation_head.") + (br ? "thresh" : "binarize");
    RC(add(5, c.par(P(pre + ".0.weight")), c.p(net->wp_h) + (size_t)br * 64 * 256 * 9, 64, 256, 3, 128, br * 64,
           net->training ? c.p(net->wpt_h) : nullptr));
    RC(add(2, c.par(P(pre + ".3.weight")), c.p(net->wp_t[br]), 64, 64, 2, 0, 0));
    if (net->training) RC(add(3, c.par(P(pre + ".3.weight")), c.p(net->wpt_t[br]), 64, 64, 2, 0, 0));
  }
  return flush();
  }
}

template <typename T>
int convbn_fwd(const Ctx<T>& c, ConvBN& L, const ND<T>* x, int x_ctotal, int x_coff) {
  const BnIdx ix{L.gamma, L.beta, L.rm, L.rv};
  ConvStats st;
  const ConvStats* fused = bn_fused(c, st, L.g.cout, 1, &ix, c.template p<float>(L.stats));
  RC(conv_fprop(L.g, x, x_ctotal, x_coff, wsel(c, L.wp, L.w), c.par(L.b), c.p(L.z), L.g.cout, 0, c.s, fused));
  if (!fused) RC(bn_prepare(c, c.p(L.z), L.P(), L.g.cout, 1, &ix, c.template p<float>(L.stats)));
  return 0;
}

// conv -> BatchNorm -> [+ res] -> [ReLU] -> out.  Training: raw output z + fused statistics, then one apply pass.
// Inference: the fixed-statistics BatchNorm, the residual add and the ReLU run in the convolution's epilogue on the fp32
// accumulator (ConvEpi): no z tensor, no apply pass.
template <typename T>
int convbn_act_fwd(const Ctx<T>& c, ConvBN& L, const ND<T>* x, int x_ctotal, int x_coff, const ND<T>* res, int relu, ND<T>* out,
                   int out_ctotal, int out_coff) {
  static const bool no_epi = getenv("DBB_NO_EVAL_EPILOGUE") != nullptr;     // A/B switch
  const int ch = L.g.cout;
  if (c.net->training || no_epi || c.net->fp32) {
    RC(convbn_fwd(c, L, x, x_ctotal, x_coff));
    return bn_apply(c.p(L.z), L.P(), ch, c.template p<float>(L.stats), res, relu, out, out_ctotal, out_coff, c.s);
  }
  if constexpr (std::is_same<T, bf16>::value) {
    float* stats4 = c.template p<float>(L.stats);
    RC(bn_finalize_eval(ch, 0, ch, c.par(L.gamma), c.par(L.beta), c.buf(L.rm), c.buf(L.rv), BN_EPS, stats4, c.s));
    const ConvEpi epi{stats4, stats4 + ch, res, relu};
    return conv_fprop(L.g, x, x_ctotal, x_coff, wsel(c, L.wp, L.w), c.par(L.b), out, out_ctotal, out_coff, c.s, nullptr, &epi);
  }
  return DBB_OK;
}

// backward through BN (+ReLU mask) and the conv: dout -> dz -> (dW, dbias, dx)
static bool wgrad_fork_late() { static const bool v = getenv("DBB_WGRAD_FORK_LATE") != nullptr; return v; }   // A/B switch
static int self_mask() { static const int v = getenv("DBB_NO_SELF_MASK") ? 0 : 1; return v; }    // A/B switch

// mask_self: the ReLU mask is this layer's own output (no residual in between) -> re-derived from z instead of read
template <typename T>
int convbn_bwd(const Ctx<T>& c, ConvBN& L, const ND<T>* dout, int dout_ctotal, int dout_coff, const ND<T>* mask, int mask_ctotal,
               int mask_coff, const ND<T>* x, int x_ctotal, int x_coff, ND<T>* dx, int dx_accumulate, ND<T>* dsum, int mask_self = 0) {
  mask_self = mask_self && self_mask();
  const int64_t Pn = L.P();
  const int ch = L.g.cout;
  BnBwdFin fin;
  fin.nseg = 1; fin.coef3 = c.template p<float>(L.coef);
  fin.seg[0] = BnBwdFinSeg{c.par(L.gamma), c.grad(L.gamma), c.grad(L.beta), 0, ch};
  RC(bn_bwd_reduce_finalize(dout, dout_ctotal, dout_coff, mask, mask_ctotal, mask_coff, c.p(L.z), Pn, ch, c.template p<float>(L.stats), fin, c.acc(), c.ticket(), c.s, mask_self));
  RC(bn_bwd_apply(dout, dout_ctotal, dout_coff, mask, mask_ctotal, mask_coff, c.p(L.z), Pn, ch, c.template p<float>(L.stats), c.template p<float>(L.coef), c.p(L.dz), dsum, c.s, mask_self));
  // The weight gradient goes to the side stream as soon as dz is complete (measured: forking before the data gradient,
  // 7.85-7.92 ms/step, beats forking after it, 8.03; without the side stream 8.31).
  const bool fork_late = wgrad_fork_late();
  if (!fork_late) RC(fork_w(c));
  if (dx) {
    RC(conv_dgrad(L.g, c.p(L.dz), wsel(c, L.wpt, L.w), dx, c.s, dx_accumulate));
  }
  if (fork_late) RC(fork_w(c));
  RC(conv_wgrad(L.g, x, x_ctotal, x_coff, c.p(L.dz), ch, 0, c.grad(L.w), c.wgs(), WGRAD_SCRATCH_BYTES, c.sw));
  // the bias of a convolution that feeds a training-mode BatchNorm has an identically zero gradient
  // (sum_px dz = 0); the reference's value is float rounding noise.  Written as exact zeros.
  if (L.b >= 0) DBB_CUDA(cudaMemsetAsync(c.grad(L.b), 0, sizeof(float) * ch, c.sw));
  return 0;
}

template <typename T>
int block_fwd(const Ctx<T>& c, Block& bk, const ND<T>* x) {
  RC(convbn_act_fwd(c, bk.c1, x, bk.c_in, 0, nullptr, 1, c.p(bk.a1), bk.c1.g.cout, 0));
  const T* res = x;
  if (bk.has_ds) {
    RC(convbn_act_fwd(c, bk.ds, x, bk.c_in, 0, nullptr, 0, c.p(bk.rd), bk.ds.g.cout, 0));
    res = c.p(bk.rd);
  }
  RC(convbn_act_fwd(c, bk.c2, c.p(bk.a1), bk.c1.g.cout, 0, res, 1, c.p(bk.out), bk.c2.g.cout, 0));
  return 0;
}

// dx: gradient buffer of the block input; dx_has_content: it already holds another consumer's contribution
template <typename T>
int block_bwd(const Ctx<T>& c, Block& bk, const ND<T>* x, ND<T>* dx, int dx_has_content) {
  const int pl = bk.c2.g.cout;
  const T* dout = c.p(bk.d_out);
  const T* out = c.p(bk.out);
  // bn2 + conv2 ; identity skip: the ReLU-masked gradient goes straight to dx
  T* dsum = nullptr;
  if (!bk.has_ds) {
    if (dx_has_content) return set_error(DBB_EUNSUPPORTED, "block_bwd: identity-skip block with pre-filled dx");
    dsum = dx;
  }
  RC(convbn_bwd(c, bk.c2, dout, pl, 0, out, pl, 0, c.p(bk.a1), pl, 0, c.p(bk.d_a1), 0, dsum));
  int acc = (dsum != nullptr) || dx_has_content;
  if (bk.has_ds) {
    RC(convbn_bwd(c, bk.ds, dout, pl, 0, out, pl, 0, x, bk.c_in, 0, dx, acc, nullptr));
    acc = 1;
  }
  RC(convbn_bwd(c, bk.c1, c.p(bk.d_a1), pl, 0, c.p(bk.a1), pl, 0, x, bk.c_in, 0, dx, acc, nullptr, 1));
  return 0;
}

}  // namespace

template <typename T>
static int net_forward_t(DbbNet* net, const float* x, const float* const* params, float* const* buffers, float* out,
                         void* workspace, void* stream) {
  constexpr bool F32 = std::is_same<T, float>::value;
  Ctx<T> c{net, (char*)workspace, params, buffers, nullptr, (cudaStream_t)stream, side_stream(net, (cudaStream_t)stream)};
  const int N = net->n;
  DBB_CUDA(cudaMemsetAsync(c.template p<uint8_t>(net->bn_acc), 0, BN_ACC_BYTES, c.s));
  // ---- stem: conv 7x7/2 (space-to-depth form) -> BN -> ReLU -> maxpool
  RC(pack_all(c, 0));
  RC(fork_w(c));
  RC(pack_all(c, 1));
  if constexpr (!F32) RC(image_to_s2d(x, N, net->h, net->w, c.p(net->s2d), c.s));
  const int64_t P0 = (int64_t)N * net->h1 * net->w1;
  {
    const BnIdx ix{P("backbone.bn1.weight"), P("backbone.bn1.bias"), B("backbone.bn1.running_mean"), B("backbone.bn1.running_var")};
    ConvStats st;
    const ConvStats* fused = bn_fused(c, st, 64, 1, &ix, c.template p<float>(net->stats0));
    if constexpr (F32) RC(conv1_fprop_f32(N, net->h, net->w, x, c.par(P("backbone.conv1.weight")), c.p(net->z0), c.s));
    else RC(conv1_fprop(N, net->h, net->w, c.p(net->s2d), c.p(net->wp_conv1), c.p(net->z0), c.s, fused));
    if (!fused) RC(bn_prepare(c, c.p(net->z0), P0, 64, 1, &ix, c.template p<float>(net->stats0)));
  }
  // BN + ReLU are applied inside the pooling kernel: a0 is never written (debug reads materialise it on demand)
  RC(maxpool_fwd(c.p(net->z0), N, net->h1, net->w1, 64, c.p(net->x1), net->training ? c.template p<uint8_t>(net->argmax) : nullptr, c.s,
                 c.template p<float>(net->stats0)));
  RC(join_w(c));          // packed weights of every later layer are ready
  // ---- residual stages
  const T* cur = c.p(net->x1);
  for (int i = 0; i < 8; ++i) { RC(block_fwd(c, net->blocks[i], cur)); cur = c.p(net->blocks[i].out); }
  const T* feat[4];
  for (int i = 0; i < 4; ++i) feat[i] = c.p(net->blocks[i * 2 + 1].out);
  const int planes[4] = {64, 128, 256, 512};
  // ---- FPN top-down (segmentation_body.py:64-77)
  for (int i = 3; i >= 0; --i) {
    RC(convbn_act_fwd(c, net->lat[i], feat[i], planes[i], 0, nullptr, 1, c.p(net->l_act[i]), 64, 0));
  }
  const int hf = net->hh[0], wf = net->ww[0];
  const T* upper = c.p(net->l_act[3]);   // p5
  int uh = net->hh[3], uw = net->ww[3];
  for (int i = 0; i < 3; ++i) {             // p4, p3, p2
    const int lvl = 2 - i;
    RC(upsample_add_fwd(upper, uh, uw, c.p(net->l_act[lvl]), N, net->hh[lvl], net->ww[lvl], 64, c.p(net->s_sum[i]), c.s));
    T* dst = (i < 2) ? c.p(net->p_act[i]) : c.p(net->cat);
    RC(convbn_act_fwd(c, net->smooth[i], c.p(net->s_sum[i]), 64, 0, nullptr, 1, dst, i < 2 ? 64 : 256, 0));
    upper = dst; uh = net->hh[lvl]; uw = net->ww[lvl];
    if (i == 2) break;
  }
  // concat [p2, up(p3), up(p4), up(p5)] (segmentation_body.py:82-87); p2 is already in channels 0..63
  RC(upsample_into(c.p(net->p_act[1]), net->hh[1], net->ww[1], N, hf, wf, 64, c.p(net->cat), 256, 64, c.s));
  RC(upsample_into(c.p(net->p_act[0]), net->hh[2], net->ww[2], N, hf, wf, 64, c.p(net->cat), 256, 128, c.s));
  RC(upsample_into(c.p(net->l_act[3]), net->hh[3], net->ww[3], N, hf, wf, 64, c.p(net->cat), 256, 192, c.s));
  RC(convbn_act_fwd(c, net->fconv, c.p(net->cat), 256, 0, nullptr, 1, c.p(net->af), 256, 0));
  // ---- head (segmentation_head.py:35-45)
  DBB_CUDA(cudaMemsetAsync(c.template p<float>(net->bias_h), 0, 128 * sizeof(float), c.s));
  DBB_CUDA(cudaMemcpyAsync(c.template p<float>(net->bias_h), c.par(P("segmentation_head.binarize.0.bias")), 64 * sizeof(float), cudaMemcpyDeviceToDevice, c.s));
  const int64_t Ph = (int64_t)N * hf * wf;
  auto head_bn = [&](const char* idx) {
    std::vector<BnIdx> v;
    for (int br = 0; br < 2; ++br) {
      const std::string pre = std::string("segmentation_head.") + (br ? "thresh" : "binarize") + "." + idx;
      v.push_back(BnIdx{P(pre + ".weight"), P(pre + ".bias"), B(pre + ".running_mean"), B(pre + ".running_var")});
    }
    return v;
  };
  {
    const std::vector<BnIdx> ix = head_bn("1");
    static const bool no_epi = getenv("DBB_NO_EVAL_EPILOGUE") != nullptr;     // A/B switch
    if (!net->training && !no_epi && !F32) {     // inference: BatchNorm + ReLU in the convolution's epilogue (see convbn_act_fwd)
      float* stats4 = c.template p<float>(net->stats_h);
      RC(bn_prepare(c, (const T*)nullptr, Ph, 128, 2, ix.data(), stats4));
      const ConvEpi epi{stats4, stats4 + 128, nullptr, 1};
      RC(conv_fprop(net->hconv_g, c.p(net->af), 256, 0, c.p(net->wp_h), c.template p<float>(net->bias_h), c.p(net->ah), 128, 0, c.s, nullptr, &epi));
    } else {
      ConvStats st;
      const ConvStats* fused = bn_fused(c, st, 64, 2, ix.data(), c.template p<float>(net->stats_h));
      RC(conv_fprop(net->hconv_g, c.p(net->af), 256, 0, c.p(net->wp_h), c.template p<float>(net->bias_h), c.p(net->zh), 128, 0, c.s, fused));
      if (!fused) RC(bn_prepare(c, c.p(net->zh), Ph, 128, 2, ix.data(), c.template p<float>(net->stats_h)));
      RC(bn_apply(c.p(net->zh), Ph, 128, c.template p<float>(net->stats_h), nullptr, 1, c.p(net->ah), 128, 0, c.s));
    }
  }
  const int64_t Pt = Ph * 4;
  {
    const std::vector<BnIdx> ix = head_bn("4");
    bool all_fused = true;
    for (int br = 0; br < 2; ++br) {       // each branch launch finalizes its own 64-channel half of the 128-wide tensor
      const std::string pre = std::string("segmentation_head.") + (br ? "thresh" : "binarize");
      ConvStats st;
      const ConvStats* fused = bn_fused(c, st, 64, 1, &ix[br], c.template p<float>(net->stats_t), br * 64);
      all_fused = all_fused && fused;
      RC(convt_fprop(net->tconv_g, c.p(net->ah), 128, br * 64, wsel(c, net->wp_t[br], P(pre + ".3.weight")), c.par(P(pre + ".3.bias")), c.p(net->zt), 128, br * 64, c.s, fused));
    }
    if (!all_fused) RC(bn_prepare(c, c.p(net->zt), Pt, 128, 2, ix.data(), c.template p<float>(net->stats_t)));
  }
  float* hout = net->head_out.bytes ? c.template p<float>(net->head_out) : out;
  RC(head_tail_fwd(c.p(net->zt), N, 2 * hf, 2 * wf, c.template p<float>(net->stats_t), c.par(P("segmentation_head.binarize.6.weight")),
                   c.par(P("segmentation_head.thresh.6.weight")), c.par(P("segmentation_head.binarize.6.bias")),
                   c.par(P("segmentation_head.thresh.6.bias")), STEP_K, net->out_c, hout, c.s));
  if (net->head_out.bytes) RC(bilinear_fwd(hout, N * net->out_c, net->ho, net->wo, out, net->h, net->w, c.s));
  return DBB_OK;
}

template <typename T>
static int net_backward_t(DbbNet* net, const float* x_img, const float* out, const float* dout, const float* const* params,
                          float* const* grads, void* workspace, int segment, void* stream) {
  constexpr bool F32 = std::is_same<T, float>::value;
  Ctx<T> c{net, (char*)workspace, params, nullptr, grads, (cudaStream_t)stream, side_stream(net, (cudaStream_t)stream)};
  // (per-kernel profiling serialises the two chains: CUDA-event times of overlapping kernels would be inflated)
  const int N = net->n;
  const int hf = net->hh[0], wf = net->ww[0];
  const int planes[4] = {64, 128, 256, 512};
  const T* feat[4];
  for (int i = 0; i < 4; ++i) feat[i] = c.p(net->blocks[i * 2 + 1].out);
  const bool all = segment < 0;
  int nblk = 0;

  if (all || segment == 0) {
    // ---- fused head tail (ConvT2 + sigmoid + step, and the BatchNorm in front of them)
    const float* hout = out;
    const float* dhout = dout;
    if (net->head_out.bytes) {
      RC(bilinear_bwd(dout, N * 3, net->ho, net->wo, c.template p<float>(net->d_head_out), net->h, net->w, c.s));
      hout = c.template p<float>(net->head_out); dhout = c.template p<float>(net->d_head_out);
    }
    const std::string hb = "segmentation_head.binarize", ht = "segmentation_head.thresh";
    const int64_t Ph = (int64_t)N * hf * wf, Pt = Ph * 4;
    RC(head_tail_bwd_reduce(c.p(net->zt), N, 2 * hf, 2 * wf, c.template p<float>(net->stats_t), c.par(P(hb + ".6.weight")), c.par(P(ht + ".6.weight")),
                            hout, dhout, STEP_K, c.partials(), &nblk, c.s));
    RC(head_tail_bwd_finalize(c.partials(), nblk, Pt, c.par(P(hb + ".4.weight")), c.par(P(ht + ".4.weight")), c.template p<float>(net->stats_t),
                              c.grad(P(hb + ".4.weight")), c.grad(P(hb + ".4.bias")), c.grad(P(ht + ".4.weight")), c.grad(P(ht + ".4.bias")),
                              c.template p<float>(net->coef_t), c.grad(P(hb + ".6.weight")), c.grad(P(ht + ".6.weight")), c.grad(P(hb + ".6.bias")),
                              c.grad(P(ht + ".6.bias")), c.par(P(hb + ".6.weight")), c.par(P(ht + ".6.weight")), c.s));
    RC(head_tail_bwd_apply(c.p(net->zt), N, 2 * hf, 2 * wf, c.template p<float>(net->stats_t), c.template p<float>(net->coef_t), c.par(P(hb + ".6.weight")),
                           c.par(P(ht + ".6.weight")), hout, dhout, STEP_K, c.p(net->d_zt), c.s));
    // ---- ConvTranspose2d(64,64,2,2) x 2
    if (!wgrad_fork_late()) RC(fork_w(c));
    for (int br = 0; br < 2; ++br)
      RC(convt_dgrad(net->tconv_g, c.p(net->d_zt), 128, br * 64, wsel(c, net->wpt_t[br], P((br ? ht : hb) + ".3.weight")), c.p(net->d_ah), 128, br * 64, c.s));
    if (wgrad_fork_late()) RC(fork_w(c));
    for (int br = 0; br < 2; ++br) {
      const std::string pre = br ? ht : hb;
      // biases in front of a training-mode BatchNorm: identically zero gradient (see convbn_bwd)
      DBB_CUDA(cudaMemsetAsync(c.grad(P(pre + ".3.bias")), 0, 64 * sizeof(float), c.sw));
      RC(convt_wgrad(net->tconv_g, c.p(net->ah), 128, br * 64, c.p(net->d_zt), 128, br * 64, c.grad(P(pre + ".3.weight")), c.wgs(), WGRAD_SCRATCH_BYTES, c.sw));
    }
    // ---- BN(2 x 64) + ReLU + the fused 256->128 3x3 conv of the two branches
    {
      BnBwdFin fin;
      fin.nseg = 2; fin.coef3 = c.template p<float>(net->coef_h);
      for (int br = 0; br < 2; ++br) {
        const std::string pre = br ? ht : hb;
        fin.seg[br] = BnBwdFinSeg{c.par(P(pre + ".1.weight")), c.grad(P(pre + ".1.weight")), c.grad(P(pre + ".1.bias")), br * 64, 64};
      }
      RC(bn_bwd_reduce_finalize(c.p(net->d_ah), 128, 0, c.p(net->ah), 128, 0, c.p(net->zh), Ph, 128, c.template p<float>(net->stats_h), fin, c.acc(), c.ticket(), c.s, self_mask()));
    }
    RC(bn_bwd_apply(c.p(net->d_ah), 128, 0, c.p(net->ah), 128, 0, c.p(net->zh), Ph, 128, c.template p<float>(net->stats_h), c.template p<float>(net->coef_h),
                    c.p(net->d_zh), nullptr, c.s, self_mask()));
    if (!wgrad_fork_late()) RC(fork_w(c));
    RC(conv_dgrad(net->hconv_g, c.p(net->d_zh), F32 ? c.p(net->wp_h) : c.p(net->wpt_h), c.p(net->d_af), c.s, 0));
    if (wgrad_fork_late()) RC(fork_w(c));
    RC(conv_wgrad(net->hconv_g, c.p(net->af), 256, 0, c.p(net->d_zh), 128, 0, c.template p<float>(net->dw_h), c.wgs(), WGRAD_SCRATCH_BYTES, c.sw));
    const size_t half = (size_t)64 * 256 * 9;
    DBB_CUDA(cudaMemcpyAsync(c.grad(P(hb + ".0.weight")), c.template p<float>(net->dw_h), half * sizeof(float), cudaMemcpyDeviceToDevice, c.sw));
    DBB_CUDA(cudaMemcpyAsync(c.grad(P(ht + ".0.weight")), c.template p<float>(net->dw_h) + half, half * sizeof(float), cudaMemcpyDeviceToDevice, c.sw));
    DBB_CUDA(cudaMemsetAsync(c.grad(P(hb + ".0.bias")), 0, 64 * sizeof(float), c.sw));
    // ---- FPN output conv
    RC(convbn_bwd(c, net->fconv, c.p(net->d_af), 256, 0, c.p(net->af), 256, 0, c.p(net->cat), 256, 0, c.p(net->d_cat), 0, nullptr, 1));
    // ---- top-down path, bottom level first: smooth_p2, p3, p4 then the c5 lateral
    //      d_p[k]: gradient of p5 (k=0), p4 (1), p3 (2); d_s[i]: gradient of the sum feeding smooth[i]
    const T* dlevel = c.p(net->d_cat);   // gradient w.r.t. p2 = channels 0..63 of d_cat
    int dl_ct = 256;
    const T* mlevel = c.p(net->cat);     // p2 activation = channels 0..63 of cat
    int ml_ct = 256;
    for (int i = 2; i >= 0; --i) {
      const int lvl = 2 - i;               // smooth[2] -> level 0 (c2 resolution)
      const int up = lvl + 1;              // the level that was upsampled into this one
      RC(convbn_bwd(c, net->smooth[i], dlevel, dl_ct, 0, mlevel, ml_ct, 0, c.p(net->s_sum[i]), 64, 0, c.p(net->d_s[i]), 0, nullptr, 1));
      // gradient of the upsampled operand: concat slice (channels 64*up ..) + the upsample-add path
      T* dpu = c.p(net->d_p[3 - up]);   // up=1 -> d_p[2] (p3), up=2 -> d_p[1] (p4), up=3 -> d_p[0] (p5)
      RC(upsample_bwd(c.p(net->d_cat), 256, 64 * up, N, hf, wf, 64, dpu, net->hh[up], net->ww[up], 0, c.s));
      RC(upsample_bwd(c.p(net->d_s[i]), 64, 0, N, net->hh[lvl], net->ww[lvl], 64, dpu, net->hh[up], net->ww[up], 1, c.s));
      // lateral conv of this level: its output was the other operand of the sum
      RC(convbn_bwd(c, net->lat[lvl], c.p(net->d_s[i]), 64, 0, c.p(net->l_act[lvl]), 64, 0, feat[lvl], planes[lvl], 0, c.p(net->d_c[lvl]), 0, nullptr, 1));
      dlevel = dpu; dl_ct = 64;
      mlevel = (up < 3) ? c.p(net->p_act[2 - up]) : c.p(net->l_act[3]);   // p3 = p_act[1], p4 = p_act[0], p5 = l_act[3]
      ml_ct = 64;
    }
    RC(convbn_bwd(c, net->lat[3], dlevel, 64, 0, mlevel, 64, 0, feat[3], 512, 0, c.p(net->d_c[3]), 0, nullptr, 1));
  }
  auto run_block = [&](int i) -> int {
    const T* x = (i == 0) ? c.p(net->x1) : c.p(net->blocks[i - 1].out);
    T* dx = (i == 0) ? c.p(net->d_x1) : c.p(net->blocks[i - 1].d_out);
    const int has_content = (i >= 2 && (i % 2) == 0) ? 1 : 0;   // input is c2/c3/c4: the FPN lateral already wrote its share
    return block_bwd(c, net->blocks[i], x, dx, has_content);
  };
  if (all || segment == 1) for (int i = 7; i >= 4; --i) RC(run_block(i));
  if (all || segment == 2) {
    for (int i = 3; i >= 0; --i) RC(run_block(i));
    // ---- stem: maxpool -> ReLU/BN -> conv1 (no data gradient: the image needs none)
    RC(maxpool_bwd(c.p(net->d_x1), c.template p<uint8_t>(net->argmax), N, net->h1, net->w1, 64, c.p(net->d_a0), c.s));
    const int64_t P0 = (int64_t)N * net->h1 * net->w1;
    const int g1 = P("backbone.bn1.weight"), b1 = P("backbone.bn1.bias");
    {
      BnBwdFin fin;
      fin.nseg = 1; fin.coef3 = c.template p<float>(net->coef0);
      fin.seg[0] = BnBwdFinSeg{c.par(g1), c.grad(g1), c.grad(b1), 0, 64};
      // (the stem activation a0 is fused away in the forward pass, so its ReLU mask always comes from z0)
      RC(bn_bwd_reduce_finalize(c.p(net->d_a0), 64, 0, nullptr, 64, 0, c.p(net->z0), P0, 64, c.template p<float>(net->stats0), fin, c.acc(), c.ticket(), c.s, 1));
    }
    RC(bn_bwd_apply(c.p(net->d_a0), 64, 0, nullptr, 64, 0, c.p(net->z0), P0, 64, c.template p<float>(net->stats0), c.template p<float>(net->coef0),
                    c.p(net->d_z0), nullptr, c.s, 1));
    RC(fork_w(c));
    if constexpr (F32) {
      if (!x_img) return set_error(DBB_EINVAL, "net_backward: the fp32 mode needs the input image (dbb_net_backward_ex)");
      RC(conv1_wgrad_f32(N, net->h, net->w, x_img, c.p(net->d_z0), c.grad(P("backbone.conv1.weight")), c.wgs(), WGRAD_SCRATCH_BYTES, c.sw));
    } else {
      RC(conv1_wgrad(N, net->h, net->w, c.p(net->s2d), c.p(net->d_z0), c.template p<float>(net->dw_s2d), c.wgs(), WGRAD_SCRATCH_BYTES, c.sw));
      RC(conv1_wgrad_unpack(c.template p<float>(net->dw_s2d), c.grad(P("backbone.conv1.weight")), c.sw));
    }
  }
  RC(join_w(c));          // every gradient of this call is complete on the caller's stream
  return DBB_OK;
}

extern "C" int dbb_net_forward(DbbNet* net, const float* x, const float* const* params, float* const* buffers, float* out,
                               void* workspace, size_t workspace_bytes, void* stream) {
  if (!net || !x || !params || !out || !workspace) return set_error(DBB_EINVAL, "net_forward: null pointer");
  if (workspace_bytes < net->ws_bytes) return set_error(DBB_EWORKSPACE, "net_forward: workspace too small");
  if (!aligned16(x) || !aligned16(out) || (reinterpret_cast<uintptr_t>(workspace) & 1023)) return set_error(DBB_EALIGN, "net_forward: x/out need 16 B, workspace 1024 B alignment");
  return net->fp32 ? net_forward_t<float>(net, x, params, buffers, out, workspace, stream)
                   : net_forward_t<bf16>(net, x, params, buffers, out, workspace, stream);
}

// x: the input image of the matching forward call.  The bf16 path keeps its own staged copy (space-to-depth buffer) and
// ignores it; the fp32-parity mode reads it for conv1's weight gradient.
extern "C" int dbb_net_backward_ex(DbbNet* net, const float* x, const float* out, const float* dout, const float* const* params,
                                   float* const* grads, void* workspace, size_t workspace_bytes, int segment, void* stream) {
  if (!net || !out || !dout || !params || !grads || !workspace) return set_error(DBB_EINVAL, "net_backward: null pointer");
  if (!net->training) return set_error(DBB_EINVAL, "net_backward: network was planned in eval mode");
  if (workspace_bytes < net->ws_bytes) return set_error(DBB_EWORKSPACE, "net_backward: workspace too small");
  if (segment < -1 || segment > 2) return set_error(DBB_EINVAL, "net_backward: segment must be -1..2");
  return net->fp32 ? net_backward_t<float>(net, x, out, dout, params, grads, workspace, segment, stream)
                   : net_backward_t<bf16>(net, x, out, dout, params, grads, workspace, segment, stream);
}
extern "C" int dbb_net_backward(DbbNet* net, const float* out, const float* dout, const float* const* params,
                                float* const* grads, void* workspace, size_t workspace_bytes, int segment, void* stream) {
  return dbb_net_backward_ex(net, nullptr, out, dout, params, grads, workspace, workspace_bytes, segment, stream);
}

// ---- debugging / parity aid: copy a named internal NHWC bf16 tensor out as NCHW float32 (tests only)
static bool find_tensor(DbbNet* net, const std::string& name, Buf* b, int* h, int* w, int* ch) {
  auto blk = [&](int i, const char* what) -> bool {
    Block& k = net->blocks[i];
    const int ho = k.c1.g.out_h(), wo = k.c1.g.out_w(), pl = k.c2.g.cout;
    *h = ho; *w = wo; *ch = pl;
    std::string wname = what;
    if (wname == "out") *b = k.out; else if (wname == "a1") *b = k.a1; else if (wname == "z1") *b = k.c1.z;
    else if (wname == "z2") *b = k.c2.z; else if (wname == "d_out") *b = k.d_out; else if (wname == "d_a1") *b = k.d_a1;
    else if (wname == "dz1") *b = k.c1.dz; else if (wname == "dz2") *b = k.c2.dz;
    else if (wname == "zd" && k.has_ds) *b = k.ds.z; else if (wname == "dzd" && k.has_ds) *b = k.ds.dz; else return false;
    return b->bytes != 0;
  };
  if (name.rfind("block", 0) == 0) {
    const int i = name[5] - '0';
    if (i < 0 || i > 7 || name.size() < 8) return false;
    return blk(i, name.c_str() + 7);
  }
  {
    // "<unit>.z" / "<unit>.dz" for lat0..3, smooth0..2, fconv, blockK.ds
    auto unit = [&](const std::string& u) -> ConvBN* {
      if (u.rfind("lat", 0) == 0 && u.size() == 4) return &net->lat[u[3] - '0'];
      if (u.rfind("smooth", 0) == 0 && u.size() == 7) return &net->smooth[u[6] - '0'];
      if (u == "fconv") return &net->fconv;
      return nullptr;
    };
    const size_t dot = name.rfind('.');
    if (dot != std::string::npos) {
      ConvBN* L = unit(name.substr(0, dot));
      const std::string what = name.substr(dot + 1);
      if (L && (what == "z" || what == "dz")) {
        *b = (what == "z") ? L->z : L->dz;
        *h = L->g.out_h(); *w = L->g.out_w(); *ch = L->g.cout;
        return b->bytes != 0;
      }
    }
  }
  const int hf = net->hh[0], wf = net->ww[0];
  struct E { const char* n; Buf b; int h, w, c; };
  const E table[] = {
      {"z0", net->z0, net->h1, net->w1, 64}, {"a0", net->a0, net->h1, net->w1, 64}, {"x1", net->x1, net->h2, net->w2, 64},
      {"d_x1", net->d_x1, net->h2, net->w2, 64}, {"d_a0", net->d_a0, net->h1, net->w1, 64}, {"d_z0", net->d_z0, net->h1, net->w1, 64},
      {"l2", net->l_act[0], net->hh[0], net->ww[0], 64}, {"l3", net->l_act[1], net->hh[1], net->ww[1], 64},
      {"l4", net->l_act[2], net->hh[2], net->ww[2], 64}, {"p5", net->l_act[3], net->hh[3], net->ww[3], 64},
      {"p4", net->p_act[0], net->hh[2], net->ww[2], 64}, {"p3", net->p_act[1], net->hh[1], net->ww[1], 64},
      {"s4", net->s_sum[0], net->hh[2], net->ww[2], 64}, {"s3", net->s_sum[1], net->hh[1], net->ww[1], 64}, {"s2", net->s_sum[2], hf, wf, 64},
      {"cat", net->cat, hf, wf, 256}, {"zf", net->fconv.z, hf, wf, 256}, {"af", net->af, hf, wf, 256},
      {"zh", net->zh, hf, wf, 128}, {"ah", net->ah, hf, wf, 128}, {"zt", net->zt, 2 * hf, 2 * wf, 128},
      {"d_zt", net->d_zt, 2 * hf, 2 * wf, 128}, {"d_ah", net->d_ah, hf, wf, 128}, {"d_zh", net->d_zh, hf, wf, 128},
      {"d_af", net->d_af, hf, wf, 256}, {"d_cat", net->d_cat, hf, wf, 256},
      {"d_p5", net->d_p[0], net->hh[3], net->ww[3], 64}, {"d_p4", net->d_p[1], net->hh[2], net->ww[2], 64}, {"d_p3", net->d_p[2], net->hh[1], net->ww[1], 64},
      {"d_s4", net->d_s[0], net->hh[2], net->ww[2], 64}, {"d_s3", net->d_s[1], net->hh[1], net->ww[1], 64}, {"d_s2", net->d_s[2], hf, wf, 64},
  };
  for (const E& e : table)
    if (name == e.n) { *b = e.b; *h = e.h; *w = e.w; *ch = e.c; return e.b.bytes != 0; }
  return false;
}

extern "C" int dbb_net_debug_shape(DbbNet* net, const char* name, int64_t* shape4) {
  Buf b; int h, w, ch;
  if (!net || !name || !find_tensor(net, name, &b, &h, &w, &ch)) return set_error(DBB_EINVAL, "net_debug: unknown tensor");
  shape4[0] = net->n; shape4[1] = ch; shape4[2] = h; shape4[3] = w;
  return DBB_OK;
}
extern "C" int dbb_net_debug_read(DbbNet* net, const char* name, const void* workspace, float* out_nchw, void* stream) {
  Buf b; int h, w, ch;
  if (!net || !name || !find_tensor(net, name, &b, &h, &w, &ch)) return set_error(DBB_EINVAL, "net_debug: unknown tensor");
  if (std::string(name) == "a0") {   // the stem activation is fused away (maxpool_fwd applies BN + ReLU): materialise it for the reader
    char* base = (char*)const_cast<void*>(workspace);
    if (net->fp32)
      RC(bn_apply(reinterpret_cast<const float*>(base + net->z0.off), (int64_t)net->n * net->h1 * net->w1, 64,
                  reinterpret_cast<const float*>(base + net->stats0.off), nullptr, 1, reinterpret_cast<float*>(base + b.off), 64, 0,
                  (cudaStream_t)stream));
    else
    RC(bn_apply(reinterpret_cast<const bf16*>(base + net->z0.off), (int64_t)net->n * net->h1 * net->w1, 64,
                reinterpret_cast<const float*>(base + net->stats0.off), nullptr, 1, reinterpret_cast<bf16*>(base + b.off), 64, 0,
                (cudaStream_t)stream));
  }
  if (net->fp32) return nhwc_to_nchw_f32(reinterpret_cast<const float*>((const char*)workspace + b.off), out_nchw, net->n, ch, (int64_t)h * w, (cudaStream_t)stream);
  return nhwc_bf16_to_nchw_f32(reinterpret_cast<const bf16*>((const char*)workspace + b.off), out_nchw, net->n, ch, (int64_t)h * w, (cudaStream_t)stream);
}
