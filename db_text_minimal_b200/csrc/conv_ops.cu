// conv_ops.cu -- builds igemm / wgrad plans for the convolution flavours of the DB network and exposes the
// single-operator C ABI (dbb_conv2d, dbb_conv2d_wgrad) used by the parity tests and the module drop-ins.
#include "common.cuh"
#include "conv.h"
#include "conv_ops.h"
#include "elementwise.h"
#include <stdlib.h>

namespace dbb {

static int pick_block_n(int cout, int64_t m_tiles) {
  static const int forced = getenv("DBB_FORCE_BN") ? atoi(getenv("DBB_FORCE_BN")) : 0;   // tuning aid
  if (forced && cout >= forced) return forced;
  // wide N tiles amortise the A-tile fetch; fall back to narrower tiles when there are too few CTAs to fill the GPU
  if (cout >= 256 && m_tiles >= 2 * DBB_NUM_SMS) return 256;
  if (cout >= 128) return 128;
  return 64;
}

static bool use_halo(const ConvGeom& g) {
  static const bool off = getenv("DBB_NO_HALO") != nullptr;    // A/B switch for benchmarking
  return !off && g.ks == 3 && g.stride == 1 && g.pad == 1 && g.cin == 64 && g.cout == 64 && halo64_supported(g.h, g.w);
}

int conv_fprop(const ConvGeom& g, const bf16* x, int x_ctotal, int x_coff, const bf16* wp, const float* bias, bf16* y,
               int y_ctotal, int y_coff, cudaStream_t s, const ConvStats* st, const ConvEpi* epi) {
  const int ho = g.out_h(), wo = g.out_w();
  if (use_halo(g)) {
    HaloPlan hp;
    memset(&hp, 0, sizeof(hp));
    int rc = halo64_plan(&hp, x, g.n, g.h, g.w, x_ctotal, x_coff, wp, 0);
    if (rc) return rc;
    hp.y = y; hp.out_c = y_ctotal; hp.out_coff = y_coff; hp.bias = bias; hp.accumulate = 0;
    if (st) { hp.st = *st; hp.st.enabled = 1; hp.st.count = (double)g.n * ho * wo; }
    if (epi) hp.epi = *epi;
    return halo64_launch(hp, s);
  }
  IgemmPlan p;
  memset(&p, 0, sizeof(p));
  p.mn = g.n; p.mh = ho; p.mw = wo;
  p.in_sh = p.in_sw = g.stride;
  p.ntaps = g.ks * g.ks;
  for (int kh = 0; kh < g.ks; ++kh)
    for (int kw = 0; kw < g.ks; ++kw) {
      const int t = kh * g.ks + kw;
      p.dh[t] = (int8_t)(kh - g.pad); p.dw[t] = (int8_t)(kw - g.pad); p.wtap[t] = (uint8_t)t;
    }
  p.cout = g.cout;
  p.y = y; p.out_h = ho; p.out_w = wo; p.out_c = y_ctotal; p.out_coff = y_coff;
  p.out_sh = p.out_sw = 1; p.out_oh = p.out_ow = 0;
  p.bias = bias;
  if (st) { p.st = *st; p.st.enabled = 1; p.st.count = (double)g.n * ho * wo; }
  if (epi) p.epi = *epi;
  const int64_t m_tiles = ((int64_t)g.n * ho * wo + 127) / 128;
  int rc = igemm_plan_init(&p, x, g.n, g.h, g.w, x_ctotal, x_coff, g.cin, wp, p.ntaps * g.cin, g.cout, pick_block_n(g.cout, m_tiles));
  if (rc) return rc;
  return igemm_launch(p, s);
}

int conv_dgrad(const ConvGeom& g, const bf16* dy, const bf16* wp_t, bf16* dx, cudaStream_t s, int accumulate) {
  // dx[n,h,w,ci] = sum_{kh,kw,co} dy[n,(h+pad-kh)/s,(w+pad-kw)/s,co] * W[co,ci,kh,kw]   (where divisible)
  const int ho = g.out_h(), wo = g.out_w();
  const int st = g.stride;
  if (st != 1 && st != 2) return set_error(DBB_EUNSUPPORTED, "conv_dgrad: stride must be 1 or 2");
  if (use_halo(g)) {      // dx = conv(dy, W^T flipped): same kernel, mirrored tap offsets, weights packed by mode 1
    HaloPlan hp;
    memset(&hp, 0, sizeof(hp));
    int rc = halo64_plan(&hp, dy, g.n, g.h, g.w, g.cout, 0, wp_t, 1);
    if (rc) return rc;
    hp.y = dx; hp.out_c = g.cin; hp.out_coff = 0; hp.bias = nullptr; hp.accumulate = accumulate;
    return halo64_launch(hp, s);
  }
  bool need_zero = false;
  for (int a = 0; a < st; ++a)
    for (int b = 0; b < st; ++b) {
      int nt = 0;
      for (int kh = 0; kh < g.ks; ++kh) for (int kw = 0; kw < g.ks; ++kw)
        if ((a + g.pad - kh) % st == 0 && (b + g.pad - kw) % st == 0) ++nt;
      if (nt == 0) need_zero = true;
    }
  if (need_zero && !accumulate) DBB_CUDA(cudaMemsetAsync(dx, 0, sizeof(bf16) * (size_t)g.n * g.h * g.w * g.cin, s));
  for (int a = 0; a < st; ++a)
    for (int b = 0; b < st; ++b) {
      IgemmPlan p;
      memset(&p, 0, sizeof(p));
      p.mn = g.n; p.mh = (g.h - a + st - 1) / st; p.mw = (g.w - b + st - 1) / st;
      if (p.mh <= 0 || p.mw <= 0) continue;
      p.in_sh = p.in_sw = 1;
      int nt = 0;
      for (int kh = 0; kh < g.ks; ++kh)
        for (int kw = 0; kw < g.ks; ++kw) {
          const int eh = a + g.pad - kh, ew = b + g.pad - kw;
          if (eh % st != 0 || ew % st != 0) continue;
          p.dh[nt] = (int8_t)(eh / st); p.dw[nt] = (int8_t)(ew / st); p.wtap[nt] = (uint8_t)(kh * g.ks + kw);
          ++nt;
        }
      if (nt == 0) continue;
      p.ntaps = nt;
      p.cout = g.cin;
      p.y = dx; p.out_h = g.h; p.out_w = g.w; p.out_c = g.cin; p.out_coff = 0;
      p.out_sh = p.out_sw = st; p.out_oh = a; p.out_ow = b;
      p.bias = nullptr;
      p.accumulate = accumulate;
      const int64_t m_tiles = ((int64_t)p.mn * p.mh * p.mw + 127) / 128;
      int rc = igemm_plan_init(&p, dy, g.n, ho, wo, g.cout, 0, g.cout, wp_t, g.ks * g.ks * g.cout, g.cin, pick_block_n(g.cin, m_tiles));
      if (rc) return rc;
      rc = igemm_launch(p, s);
      if (rc) return rc;
    }
  return DBB_OK;
}

int convt_fprop(const ConvGeom& g, const bf16* x, int x_ctotal, int x_coff, const bf16* wp_cls, const float* bias, bf16* y,
                int y_ctotal, int y_coff, cudaStream_t s, const ConvStats* st) {
  // ConvTranspose2d(k=2, s=2): y[n,2i+a,2j+b,co] = sum_ci x[n,i,j,ci] * W[ci,co,a,b] + bias[co].
  // ONE GEMM [pixels, cin] x [cin, 4*cout] (packed weights are [class][co][ci] = 4*cout rows) with a pixel-shuffle epilogue.
  IgemmPlan p;
  memset(&p, 0, sizeof(p));
  p.mn = g.n; p.mh = g.h; p.mw = g.w;
  p.in_sh = p.in_sw = 1;
  p.ntaps = 1; p.dh[0] = 0; p.dw[0] = 0; p.wtap[0] = 0;
  p.cout = 4 * g.cout;
  p.cls_cols = g.cout;
  p.y = y; p.out_h = 2 * g.h; p.out_w = 2 * g.w; p.out_c = y_ctotal; p.out_coff = y_coff;
  p.out_sh = p.out_sw = 2; p.out_oh = p.out_ow = 0;
  p.bias = bias;
  if (st) { p.st = *st; p.st.enabled = 1; p.st.count = (double)g.n * g.h * g.w * 4.0; }
  // one k-iteration per tile and a store-bound pixel-shuffle epilogue: 64-wide tiles keep the fused BatchNorm statistics in
  // registers (measured 133 us vs 222 us per branch with 128-wide tiles at 16x160x160)
  static const int convt_bn = getenv("DBB_CONVT_BN") ? atoi(getenv("DBB_CONVT_BN")) : 64;     // tuning aid
  int rc = igemm_plan_init(&p, x, g.n, g.h, g.w, x_ctotal, x_coff, g.cin, wp_cls, g.cin, 4 * g.cout, convt_bn);
  if (rc) return rc;
  return igemm_launch(p, s);
}

int convt_dgrad(const ConvGeom& g, const bf16* dy, int dy_ctotal, int dy_coff, const bf16* wp_t, bf16* dx, int dx_ctotal,
                int dx_coff, cudaStream_t s) {
  // dx[n,i,j,ci] = sum_{a,b,co} dy[n,2i+a,2j+b,co] * W[ci,co,a,b]
  IgemmPlan p;
  memset(&p, 0, sizeof(p));
  p.mn = g.n; p.mh = g.h; p.mw = g.w;
  p.in_sh = p.in_sw = 2;
  p.ntaps = 4;
  for (int t = 0; t < 4; ++t) { p.dh[t] = (int8_t)(t >> 1); p.dw[t] = (int8_t)(t & 1); p.wtap[t] = (uint8_t)t; }
  p.cout = g.cin;
  p.y = dx; p.out_h = g.h; p.out_w = g.w; p.out_c = dx_ctotal; p.out_coff = dx_coff;
  p.out_sh = p.out_sw = 1; p.out_oh = p.out_ow = 0;
  const int64_t m_tiles = ((int64_t)g.n * g.h * g.w + 127) / 128;
  int rc = igemm_plan_init(&p, dy, g.n, 2 * g.h, 2 * g.w, dy_ctotal, dy_coff, g.cout, wp_t, 4 * g.cout, g.cin, pick_block_n(g.cin, m_tiles));
  if (rc) return rc;
  return igemm_launch(p, s);
}

// split-K policy: minimise (waves of CTAs) x (k-iterations per CTA).  A grid one CTA over a multiple of the resident slots
// costs a whole extra wave (306 CTAs on 148 slots ran at 69 % efficiency), so the split is chosen wave-aware; at least
// 8 k-iterations per CTA, partials must fit the scratch, ties go to the smaller split (less scratch traffic to reduce).
static int pick_split_k(int64_t out_tiles, int total_pixel_tiles, size_t tile_bytes_all, size_t scratch_bytes, int slots) {
  int64_t cap = total_pixel_tiles / 8 > 1 ? total_pixel_tiles / 8 : 1;
  const int64_t cap_ws = tile_bytes_all ? (int64_t)(scratch_bytes / tile_bytes_all) : 1;
  if (cap > cap_ws) cap = cap_ws;
  if (cap > total_pixel_tiles) cap = total_pixel_tiles;
  if (cap > 4096) cap = 4096;
  if (cap < 1) cap = 1;
  int64_t best = 1; double best_cost = 1e30;
  for (int64_t sp = 1; sp <= cap; ++sp) {
    const int64_t waves = (out_tiles * sp + slots - 1) / slots;
    const int64_t kiters = (total_pixel_tiles + sp - 1) / sp;
    const double cost = (double)waves * (double)(kiters + 6) + 0.05 * (double)sp;     // + pipeline fill/epilogue per CTA, + reduce
    if (cost < best_cost - 1e-9) { best_cost = cost; best = sp; }
  }
  return (int)best;
}

static int wgrad_finish_plan(WgradPlan& p, float* scratch, size_t scratch_bytes) {
  const int m_tiles = (p.m_total + 127) / 128, n_tiles = (p.n_total + p.n_tile - 1) / p.n_tile;
  p.m_pad = m_tiles * 128; p.n_pad = n_tiles * p.n_tile;
  const size_t one = (size_t)p.ntaps * p.m_pad * p.n_pad * sizeof(float);
  if (!scratch || scratch_bytes < one) return set_error(DBB_EWORKSPACE, "wgrad: split-K scratch too small");
  const int slots = DBB_NUM_SMS * (p.n_tile >= 256 ? 1 : 2);      // resident CTAs: 192 KB ring at N = 256, 96 KB below
  static const bool legacy = getenv("DBB_LEGACY_SPLITK") != nullptr;    // A/B switch: the old "2 CTAs per SM worth" rule
  if (legacy) {
    const int64_t out_tiles = (int64_t)p.ntaps * m_tiles * n_tiles;
    int64_t want = (2 * DBB_NUM_SMS + out_tiles - 1) / out_tiles;
    const int64_t T = p.tiles_n * p.tiles_h * p.tiles_w;
    if (want > (T / 8 > 1 ? T / 8 : 1)) want = (T / 8 > 1 ? T / 8 : 1);
    if (want > (int64_t)(scratch_bytes / one)) want = (int64_t)(scratch_bytes / one);
    p.split_k = (int)(want < 1 ? 1 : want);
  } else {
    p.split_k = pick_split_k((int64_t)p.ntaps * m_tiles * n_tiles, p.tiles_n * p.tiles_h * p.tiles_w, one, scratch_bytes, slots);
  }
  p.ws = scratch;
  return DBB_OK;
}

static int wgrad_common(WgradPlan& p, const bf16* a, int a_n, int a_h, int a_w, int a_ctotal, int a_coff, int a_c,
                        const bf16* b, int b_n, int b_h, int b_w, int b_ctotal, int b_coff, int b_c, float* scratch,
                        size_t scratch_bytes, cudaStream_t s) {
  choose_box(64, p.mn, p.mh, p.mw, &p.bn, &p.bh, &p.bw);
  p.tiles_n = (p.mn + p.bn - 1) / p.bn;
  p.tiles_h = (p.mh + p.bh - 1) / p.bh;
  p.tiles_w = (p.mw + p.bw - 1) / p.bw;
  p.m_tile = 128;
  p.n_tile = p.n_total >= 256 ? 256 : (p.n_total >= 128 ? 128 : 64);
  int rc = wgrad_finish_plan(p, scratch, scratch_bytes);
  if (rc) return rc;
  rc = encode_tmap_nhwc(&p.tmap_a, a, a_n, a_h, a_w, a_ctotal, a_coff, a_c, p.bn, p.bh, p.bw, p.a_sh, p.a_sw);
  if (rc) return rc;
  rc = encode_tmap_nhwc(&p.tmap_b, b, b_n, b_h, b_w, b_ctotal, b_coff, b_c, p.bn, p.bh, p.bw, p.b_sh, p.b_sw);
  if (rc) return rc;
  return wgrad_launch(p, s);
}

int conv_wgrad(const ConvGeom& g, const bf16* x, int x_ctotal, int x_coff, const bf16* dy, int dy_ctotal, int dy_coff, float* dw,
               float* scratch, size_t scratch_bytes, cudaStream_t s) {
  // dW[co,ci,kh,kw] = sum_{n,i,j} dy[n,i,j,co] * x[n, i*s+kh-pad, j*s+kw-pad, ci]
  const int ho = g.out_h(), wo = g.out_w();
  if (use_halo(g) && wgrad_row64_supported(g.w) && getenv("DBB_NO_ROW64") == nullptr)
    return wgrad_row64(x, x_ctotal, x_coff, dy, dy_ctotal, dy_coff, g.n, g.h, g.w, dw, scratch, scratch_bytes, s);
  WgradPlan p;
  memset(&p, 0, sizeof(p));
  p.mn = g.n; p.mh = ho; p.mw = wo;
  p.a_sh = p.a_sw = 1; p.b_sh = p.b_sw = g.stride;
  p.ntaps = g.ks * g.ks;
  for (int kh = 0; kh < g.ks; ++kh)
    for (int kw = 0; kw < g.ks; ++kw) {
      const int t = kh * g.ks + kw;
      p.a_dh[t] = 0; p.a_dw[t] = 0; p.b_dh[t] = (int8_t)(kh - g.pad); p.b_dw[t] = (int8_t)(kw - g.pad);
    }
  p.m_total = g.cout; p.n_total = g.cin;
  p.dw = dw; p.tap_stride = p.ntaps;
  return wgrad_common(p, dy, g.n, ho, wo, dy_ctotal, dy_coff, g.cout, x, g.n, g.h, g.w, x_ctotal, x_coff, g.cin, scratch, scratch_bytes, s);
}

int convt_wgrad(const ConvGeom& g, const bf16* x, int x_ctotal, int x_coff, const bf16* dy, int dy_ctotal, int dy_coff, float* dw,
                float* scratch, size_t scratch_bytes, cudaStream_t s) {
  // dW[ci,co,a,b] = sum_{n,i,j} x[n,i,j,ci] * dy[n,2i+a,2j+b,co]
  WgradPlan p;
  memset(&p, 0, sizeof(p));
  p.mn = g.n; p.mh = g.h; p.mw = g.w;
  p.a_sh = p.a_sw = 1; p.b_sh = p.b_sw = 2;
  p.ntaps = 4;
  for (int t = 0; t < 4; ++t) { p.a_dh[t] = 0; p.a_dw[t] = 0; p.b_dh[t] = (int8_t)(t >> 1); p.b_dw[t] = (int8_t)(t & 1); }
  p.m_total = g.cin; p.n_total = g.cout;
  p.dw = dw; p.tap_stride = 4;
  return wgrad_common(p, x, g.n, g.h, g.w, x_ctotal, x_coff, g.cin, dy, g.n, 2 * g.h, 2 * g.w, dy_ctotal, dy_coff, g.cout, scratch, scratch_bytes, s);
}

// ---- conv1: Conv2d(3, 64, 7, stride 2, pad 3) as a 4x4 stride-1 convolution over the space-to-depth image.
// The staging buffer [n][hs+3][ws+3][16] is viewed through an OVERLAPPING tensor map {64, ws, hs+3, n} with a 32-byte
// stride along W: one 128-byte "row" is 4 consecutive s2d pixels x 16 channels = the kw2 taps of one kh2 row.
static int conv1_tmap(CUtensorMap* m, const bf16* s2d, int n, int hs, int ws, int bn, int bh, int bw) {
  const cuuint64_t dims[4] = {64, (cuuint64_t)ws, (cuuint64_t)(hs + 3), (cuuint64_t)n};
  const cuuint64_t strides[3] = {32, (cuuint64_t)(ws + 3) * 32, (cuuint64_t)(hs + 3) * (ws + 3) * 32};
  const cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  return encode_tmap_raw4(m, s2d, dims, strides, box);
}

int conv1_fprop(int n, int h, int w, const bf16* s2d, const bf16* wp, bf16* y, cudaStream_t s, const ConvStats* st) {
  const int hs = (h + 1) / 2, ws = (w + 1) / 2;
  IgemmPlan p;
  memset(&p, 0, sizeof(p));
  p.mn = n; p.mh = hs; p.mw = ws;
  p.in_sh = p.in_sw = 1;
  p.ntaps = 4;
  for (int t = 0; t < 4; ++t) { p.dh[t] = (int8_t)t; p.dw[t] = 0; p.wtap[t] = (uint8_t)t; }
  p.cout = 64;
  p.y = y; p.out_h = hs; p.out_w = ws; p.out_c = 64; p.out_coff = 0;
  p.out_sh = p.out_sw = 1;
  p.cin = 64; p.block_n = 64;
  if (st) { p.st = *st; p.st.enabled = 1; p.st.count = (double)n * hs * ws; }
  choose_box(128, p.mn, p.mh, p.mw, &p.bn, &p.bh, &p.bw);
  p.tiles_n = (p.mn + p.bn - 1) / p.bn; p.tiles_h = (p.mh + p.bh - 1) / p.bh; p.tiles_w = (p.mw + p.bw - 1) / p.bw;
  int rc = conv1_tmap(&p.tmap_x, s2d, n, hs, ws, p.bn, p.bh, p.bw);
  if (rc) return rc;
  rc = encode_weights_public(&p.tmap_w, wp, 256, 64, 64);   // 2-D weight map: [64 rows][256 K]
  if (rc) return rc;
  return igemm_launch(p, s);
}

int conv1_wgrad(int n, int h, int w, const bf16* s2d, const bf16* dy, float* dw_s2d, float* scratch, size_t scratch_bytes, cudaStream_t s) {
  const int hs = (h + 1) / 2, ws = (w + 1) / 2;
  WgradPlan p;
  memset(&p, 0, sizeof(p));
  p.mn = n; p.mh = hs; p.mw = ws;
  p.a_sh = p.a_sw = 1; p.b_sh = p.b_sw = 1;
  p.ntaps = 4;
  for (int t = 0; t < 4; ++t) { p.b_dh[t] = (int8_t)t; p.b_dw[t] = 0; }
  p.m_total = 64; p.n_total = 64;
  p.dw = dw_s2d; p.tap_stride = 4;
  choose_box(64, p.mn, p.mh, p.mw, &p.bn, &p.bh, &p.bw);
  p.tiles_n = (p.mn + p.bn - 1) / p.bn; p.tiles_h = (p.mh + p.bh - 1) / p.bh; p.tiles_w = (p.mw + p.bw - 1) / p.bw;
  p.m_tile = 128; p.n_tile = 64;
  int rc = wgrad_finish_plan(p, scratch, scratch_bytes);
  if (rc) return rc;
  rc = encode_tmap_nhwc(&p.tmap_a, dy, n, hs, ws, 64, 0, 64, p.bn, p.bh, p.bw, 1, 1);
  if (rc) return rc;
  rc = conv1_tmap(&p.tmap_b, s2d, n, hs, ws, p.bn, p.bh, p.bw);
  if (rc) return rc;
  return wgrad_launch(p, s);
}

}  // namespace dbb

using namespace dbb;

// workspace of the single-operator entry: packed bf16 weights
extern "C" size_t dbb_conv2d_workspace(int kind, int64_t n, int64_t h, int64_t wdt, int cin, int cout, int ksize, int stride, int pad) {
  (void)kind; (void)n; (void)h; (void)wdt; (void)stride; (void)pad;
  return ((size_t)cin * cout * ksize * ksize * sizeof(bf16) + 255) / 256 * 256;
}

static int check_geom(const char* who, int64_t n, int64_t h, int64_t w, int cin, int cout, int ks, int stride) {
  if (n <= 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0) return set_error(DBB_EINVAL, who);
  if (cin % 64 || cout % 64) return set_error(DBB_EUNSUPPORTED, "conv: channel counts must be multiples of 64");
  if (ks != 1 && ks != 2 && ks != 3) return set_error(DBB_EUNSUPPORTED, "conv: kernel size must be 1, 2 (ConvT) or 3");
  if (stride != 1 && stride != 2) return set_error(DBB_EUNSUPPORTED, "conv: stride must be 1 or 2");
  return DBB_OK;
}

extern "C" int dbb_conv2d(int kind, const void* x, const float* w, const float* bias, void* y, int64_t n, int64_t h, int64_t wdt,
                          int cin, int cout, int ksize, int stride, int pad, void* workspace, size_t workspace_bytes, void* stream) {
  if (!x || !w || !y || !workspace) return set_error(DBB_EINVAL, "conv2d: null pointer");
  int rc = check_geom("conv2d: bad shape", n, h, wdt, cin, cout, ksize, stride);
  if (rc) return rc;
  if (workspace_bytes < dbb_conv2d_workspace(kind, n, h, wdt, cin, cout, ksize, stride, pad)) return set_error(DBB_EWORKSPACE, "conv2d: workspace too small");
  if (!aligned16(x) || !aligned16(y) || !aligned16(workspace)) return set_error(DBB_EALIGN, "conv2d: pointer not 16B aligned");
  cudaStream_t s = (cudaStream_t)stream;
  ConvGeom g{(int)n, (int)h, (int)wdt, cin, cout, ksize, stride, pad};
  bf16* wp = (bf16*)workspace;
  switch (kind) {
    case 0:
      if ((rc = pack_weights(0, w, wp, cout, cin, ksize, ksize, s))) return rc;
      return conv_fprop(g, (const bf16*)x, cin, 0, wp, bias, (bf16*)y, cout, 0, s);
    case 1:   // x := dy (n, ho, wo, cout), y := dx (n, h, w, cin)
      if ((rc = pack_weights(1, w, wp, cout, cin, ksize, ksize, s))) return rc;
      return conv_dgrad(g, (const bf16*)x, wp, (bf16*)y, s);
    case 2:
      if (ksize != 2 || stride != 2) return set_error(DBB_EUNSUPPORTED, "convT: only k=2, s=2");
      if ((rc = pack_weights(2, w, wp, cout, cin, 2, 2, s))) return rc;
      return convt_fprop(g, (const bf16*)x, cin, 0, wp, bias, (bf16*)y, cout, 0, s);
    case 3:
      if (ksize != 2 || stride != 2) return set_error(DBB_EUNSUPPORTED, "convT: only k=2, s=2");
      if ((rc = pack_weights(3, w, wp, cout, cin, 2, 2, s))) return rc;
      return convt_dgrad(g, (const bf16*)x, cout, 0, wp, (bf16*)y, cin, 0, s);
    default: return set_error(DBB_EINVAL, "conv2d: kind must be 0..3");
  }
}

extern "C" size_t dbb_conv2d_wgrad_workspace(void) { return WGRAD_SCRATCH_BYTES; }

extern "C" int dbb_conv2d_wgrad(int kind, const void* x, const void* dy, float* dw, int64_t n, int64_t h, int64_t wdt, int cin,
                                int cout, int ksize, int stride, int pad, void* workspace, size_t workspace_bytes, void* stream) {
  if (!workspace) return set_error(DBB_EINVAL, "conv2d_wgrad: null workspace (see dbb_conv2d_wgrad_workspace)");
  if (!x || !dy || !dw) return set_error(DBB_EINVAL, "conv2d_wgrad: null pointer");
  int rc = check_geom("conv2d_wgrad: bad shape", n, h, wdt, cin, cout, ksize, stride);
  if (rc) return rc;
  if (!aligned16(x) || !aligned16(dy) || !aligned16(dw)) return set_error(DBB_EALIGN, "conv2d_wgrad: pointer not 16B aligned");
  ConvGeom g{(int)n, (int)h, (int)wdt, cin, cout, ksize, stride, pad};
  if (kind == 0) return conv_wgrad(g, (const bf16*)x, cin, 0, (const bf16*)dy, cout, 0, dw, (float*)workspace, workspace_bytes, (cudaStream_t)stream);
  if (kind == 2) {
    if (ksize != 2 || stride != 2) return set_error(DBB_EUNSUPPORTED, "convT: only k=2, s=2");
    return convt_wgrad(g, (const bf16*)x, cin, 0, (const bf16*)dy, cout, 0, dw, (float*)workspace, workspace_bytes, (cudaStream_t)stream);
  }
  return set_error(DBB_EINVAL, "conv2d_wgrad: kind must be 0 (Conv2d) or 2 (ConvTranspose2d)");
}

// ---- stem: Conv2d(3, 64, 7, stride 2, pad 3, bias=False) of src/modules/resnet.py:171 on the NCHW float32 image.
// workspace = space-to-depth staging buffer + packed weights (+ weight-gradient scratch for the backward entry)
static size_t conv1_s2d_bytes(int64_t n, int64_t h, int64_t w) {
  const int64_t hs = (h + 1) / 2, ws = (w + 1) / 2;
  return ((size_t)n * (hs + 3) * (ws + 3) * 16 * sizeof(bf16) + 1023) / 1024 * 1024;
}
extern "C" size_t dbb_conv1_workspace(int64_t n, int64_t h, int64_t w, int backward) {
  return conv1_s2d_bytes(n, h, w) + (backward ? (size_t)(64 * 64 * 4 * sizeof(float) + 1024) + WGRAD_SCRATCH_BYTES : (size_t)(64 * 256 * sizeof(bf16) + 1024));
}
extern "C" int dbb_conv1_fwd(const float* img, const float* weight, void* y, int64_t n, int64_t h, int64_t w, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (!img || !weight || !y || !workspace) return set_error(DBB_EINVAL, "conv1_fwd: null pointer");
  if (n <= 0 || h < 8 || w < 8) return set_error(DBB_EINVAL, "conv1_fwd: bad shape");
  if (workspace_bytes < dbb_conv1_workspace(n, h, w, 0)) return set_error(DBB_EWORKSPACE, "conv1_fwd: workspace too small");
  if (!aligned16(img) || !aligned16(y) || (reinterpret_cast<uintptr_t>(workspace) & 1023)) return set_error(DBB_EALIGN, "conv1_fwd: alignment (workspace 1024 B)");
  cudaStream_t s = (cudaStream_t)stream;
  bf16* s2d = (bf16*)workspace;
  bf16* wp = (bf16*)((char*)workspace + conv1_s2d_bytes(n, h, w));
  int rc;
  if ((rc = image_to_s2d(img, (int)n, (int)h, (int)w, s2d, s))) return rc;
  if ((rc = pack_weights(4, weight, wp, 64, 3, 7, 7, s))) return rc;
  return conv1_fprop((int)n, (int)h, (int)w, s2d, wp, (bf16*)y, s);
}
extern "C" int dbb_conv1_wgrad(const float* img, const void* dy, float* dw, int64_t n, int64_t h, int64_t w, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (!img || !dy || !dw || !workspace) return set_error(DBB_EINVAL, "conv1_wgrad: null pointer");
  if (workspace_bytes < dbb_conv1_workspace(n, h, w, 1)) return set_error(DBB_EWORKSPACE, "conv1_wgrad: workspace too small");
  if (!aligned16(img) || !aligned16(dy) || (reinterpret_cast<uintptr_t>(workspace) & 1023)) return set_error(DBB_EALIGN, "conv1_wgrad: alignment (workspace 1024 B)");
  cudaStream_t s = (cudaStream_t)stream;
  bf16* s2d = (bf16*)workspace;
  float* dw_s2d = (float*)((char*)workspace + conv1_s2d_bytes(n, h, w));
  float* scratch = (float*)((char*)dw_s2d + 64 * 64 * 4 * sizeof(float) + 1024);
  int rc;
  if ((rc = image_to_s2d(img, (int)n, (int)h, (int)w, s2d, s))) return rc;
  if ((rc = conv1_wgrad((int)n, (int)h, (int)w, s2d, (const bf16*)dy, dw_s2d, scratch, WGRAD_SCRATCH_BYTES, s))) return rc;
  return conv1_wgrad_unpack(dw_s2d, dw, s);
}
