// db_loss.cu -- fused three-term DB loss (balanced BCE with OHEM top-k, Dice, masked L1), forward + backward.
//
// Replaces src/losses.py:18-40 (OHEMBalanceCrossEntropyLoss), :48-66 (DiceLoss), :75-82 (L1Loss) and
// :105-139 (DBLoss.forward) of the reference.  HBM-bound: one read of the 7 maps in the forward
// (28 B/px), one more read of P/gt/mask (12 B/px) for the exact radix select in 'none' mode, and
// 28 B/px read + 12 B/px write in the backward.  No host synchronisation anywhere: the three
// blocking syncs of the reference (losses.py:25,27,65) become device scalars in DbbLossState.
//
// Semantics restated from the reference:
//   pos = gt*mask, neg = (1-gt)*mask, n_pos = int(sum pos), n_neg = min(int(n_pos*ratio), int(sum neg))
//   bce_i = -(g*max(log p,-100) + (1-g)*max(log1p(-p),-100));  bce'_i = (p-g)/max(p(1-p),1e-12)
//   'mean': prob = mean(bce) * (sum pos + n_neg) / (n_pos + n_neg + eps)      (degenerate OHEM)
//   'none': prob = (sum pos*bce + sum of the n_neg largest neg*bce) / (n_pos + n_neg + eps)
//   thr  = sum |T-tg|*tm / (sum tm + eps);   bin = 1 - 2 sum(B g m) / (sum B m + sum g m + eps)
#include "common.cuh"

namespace dbb {

enum { S_POS = 0, S_NEG, S_BCE_ALL, S_BCE_POS, S_L1, S_TM, S_I, S_BM, S_NEGL, S_TOP_ABOVE, S_TOP_CAND, S_SPARE, NSUM };
static_assert(NSUM == 12, "DbbLossState.sums has 12 slots");

constexpr int LOSS_THREADS = 256;
constexpr int HIST1_BITS = 11;               // bits 30..20 of the (non-negative) float
constexpr int HIST1_BINS = 1 << HIST1_BITS;  // 2048
constexpr int HIST2_BINS = 2048;             // bits 19..9
constexpr int HIST3_BINS = 512;              // bits 8..0

// workspace layout (bytes): [partials double grid*NSUM][hist1 u32 2048][cand_count u32 (padded to 16)][taubin i32 ...][cands float px]
struct LossWs {
  double* partials;
  unsigned* hist1;
  unsigned* cand_count;   // [0] = count, [1] = taubin (as int), [2] = cnt_above_bin lo, [3] hi
  unsigned* hist2;
  unsigned* hist3;
  unsigned* sel_ctl;
  float* cands;
};
constexpr size_t LOSS_CTL_BYTES = sizeof(unsigned) * (HIST1_BINS + HIST2_BINS + HIST3_BINS) + 512;

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int loss_grid(int64_t px) {
  int64_t want = (px / 4 + LOSS_THREADS * 4 - 1) / (LOSS_THREADS * 4);   // ~4 float4 per thread minimum
  int64_t g = DBB_NUM_SMS * 4;
  if (want < g) g = want < 1 ? 1 : want;
  return (int)g;
}

static LossWs carve(void* ws, int grid, int64_t px) {
  LossWs w;
  char* p = (char*)ws;
  w.partials = (double*)p;            p += align_up(sizeof(double) * NSUM * (size_t)grid, 256);
  w.hist1 = (unsigned*)p;             p += sizeof(unsigned) * HIST1_BINS;
  w.cand_count = (unsigned*)p;        p += 256;
  w.hist2 = (unsigned*)p;             p += sizeof(unsigned) * HIST2_BINS;
  w.hist3 = (unsigned*)p;             p += sizeof(unsigned) * HIST3_BINS;
  w.sel_ctl = (unsigned*)p;           p += 256;
  w.cands = (float*)p;
  return w;
}

__device__ __forceinline__ float bce_fwd(float p, float g) {
  // ATen binary_cross_entropy: (t-1)*max(log1p(-x),-100) - t*max(log(x),-100)
  // explicit _rn ops: no FMA contraction, so every kernel (and the CPU oracle) rounds identically
  float l1 = fmaxf(log1pf(-p), -100.f);
  float l0 = fmaxf(logf(p), -100.f);
  return __fsub_rn(__fmul_rn(__fsub_rn(g, 1.f), l1), __fmul_rn(g, l0));
}
__device__ __forceinline__ float bce_bwd(float p, float g) {
  return (p - g) / fmaxf(p * (1.f - p), 1e-12f);
}
__device__ __forceinline__ float negl_of(float p, float g, float m) {
  float v = __fmul_rn(bce_fwd(p, g), __fmul_rn(__fsub_rn(1.f, g), m));   // loss * negative, fp32 like the reference
  return v > 0.f ? v : 0.f;                    // folds -0.0 into +0.0 so the bit pattern is monotone
}

template <int N>
__device__ __forceinline__ void block_reduce_store(float (&acc)[N], double* out) {
  __shared__ float red[LOSS_THREADS / 32][N];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) red[wid][i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < N) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < LOSS_THREADS / 32; ++w) s += (double)red[w][threadIdx.x];
    out[threadIdx.x] = s;
  }
}

struct Vec4 { float v[4]; };
template <int VEC> struct Loader;
template <> struct Loader<4> {
  static __device__ __forceinline__ Vec4 ld(const float* p) {
    float4 t = ldg_stream(reinterpret_cast<const float4*>(p));
    return Vec4{{t.x, t.y, t.z, t.w}};
  }
  static __device__ __forceinline__ void st(float* p, const Vec4& a) {
    stg_stream(reinterpret_cast<float4*>(p), make_float4(a.v[0], a.v[1], a.v[2], a.v[3]));
  }
};
template <> struct Loader<1> {
  static __device__ __forceinline__ Vec4 ld(const float* p) { return Vec4{{__ldg(p), 0.f, 0.f, 0.f}}; }
  static __device__ __forceinline__ void st(float* p, const Vec4& a) { *p = a.v[0]; }
};

// ---------------------------------------------------------------------------------------------
// pass 1: every reduction of the three losses in one read of the 7 maps (+ level-1 histogram)
// ---------------------------------------------------------------------------------------------
template <int VEC, bool HAS_B, bool SELECT>
__global__ void __launch_bounds__(LOSS_THREADS)
dbloss_reduce_kernel(const float* __restrict__ preds, const float* __restrict__ gts, int64_t n_img, int64_t hw, int cch,
                     double* __restrict__ partials, unsigned* __restrict__ hist1) {
  __shared__ unsigned sh_hist[SELECT ? HIST1_BINS : 1];
  if (SELECT) {
    for (int i = threadIdx.x; i < HIST1_BINS; i += LOSS_THREADS) sh_hist[i] = 0;
    __syncthreads();
  }
  const int64_t px = n_img * hw;
  const int64_t hwv = hw / VEC, nvec = n_img * hwv;
  float acc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = 0.f;

  for (int64_t v = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * LOSS_THREADS) {
    const unsigned n32 = (unsigned)v / (unsigned)hwv;          // nvec < 2^32 (host check): 32-bit divide
    const int64_t n = n32, r = (int64_t)((unsigned)v - n32 * (unsigned)hwv) * VEC;
    const float* pp = preds + (n * cch) * hw + r;
    const int64_t go = n * hw + r;
    Vec4 P = Loader<VEC>::ld(pp), T = Loader<VEC>::ld(pp + hw);
    Vec4 B; if (HAS_B) B = Loader<VEC>::ld(pp + 2 * hw);
    Vec4 G = Loader<VEC>::ld(gts + go), M = Loader<VEC>::ld(gts + px + go);
    Vec4 TG = Loader<VEC>::ld(gts + 2 * px + go), TM = Loader<VEC>::ld(gts + 3 * px + go);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float p = P.v[j], g = G.v[j], m = M.v[j];
      const float pos = __fmul_rn(g, m), neg = __fmul_rn(__fsub_rn(1.f, g), m);
      const float bce = bce_fwd(p, g);
      acc[S_POS] += pos;
      acc[S_NEG] += neg;
      acc[S_BCE_ALL] += bce;
      acc[S_BCE_POS] += bce * pos;
      acc[S_L1] += fabsf(T.v[j] - TG.v[j]) * TM.v[j];
      acc[S_TM] += TM.v[j];
      if (HAS_B) {
        acc[S_I] += B.v[j] * g * m;
        acc[S_BM] += B.v[j] * m;
      }
      if (SELECT) {
        float nl = __fmul_rn(bce, neg);
        if (nl > 0.f) {
          acc[S_NEGL] += nl;
          atomicAdd(&sh_hist[__float_as_uint(nl) >> 20], 1u);
        }
      }
    }
  }
  block_reduce_store<9>(acc, partials + (size_t)blockIdx.x * NSUM);
  if (SELECT) {
    __syncthreads();
    for (int i = threadIdx.x; i < HIST1_BINS; i += LOSS_THREADS) {
      unsigned c = sh_hist[i];
      if (c) atomicAdd(&hist1[i], c);
    }
  }
}

// deterministic sum of per-block partials: thread j (< NSUM*?) ...
__device__ void sum_partials(const double* partials, int nblk, double* sums /*smem NSUM*/, int first, int last) {
  // each warp handles one component at a time; fixed order -> bitwise reproducible
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = first + wid; c < last; c += nw) {
    double s = 0.0;
    for (int b = lane; b < nblk; b += 32) s += partials[(size_t)b * NSUM + c];
    s = warp_sum(s);
    if (lane == 0) sums[c] = s;
  }
}

struct LossParams {
  float alpha, beta, ratio, eps;
  int64_t px;
  int has_b;
};

__device__ void write_losses(const double* S, double prob, const LossParams& lp, float* losses5, DbbLossState* st,
                             double coef0) {
  const double thr = S[S_L1] / (S[S_TM] + (double)lp.eps);
  double bin = 0.0, U = 1.0, I = 0.0;
  if (lp.has_b) {
    I = S[S_I];
    U = S[S_BM] + S[S_POS] + (double)lp.eps;
    bin = 1.0 - 2.0 * I / U;
  }
  const double pt = prob + (double)lp.beta * thr;
  const double total = lp.has_b ? (double)lp.alpha * bin + pt : pt;
  losses5[0] = (float)prob; losses5[1] = (float)thr; losses5[2] = (float)bin; losses5[3] = (float)pt; losses5[4] = (float)total;
  st->coef[0] = (float)coef0;
  st->coef[1] = (float)(1.0 / (S[S_TM] + (double)lp.eps));
  st->coef[2] = lp.has_b ? (float)(2.0 / U) : 0.f;
  st->coef[3] = lp.has_b ? (float)(2.0 * I / (U * U)) : 0.f;
  for (int i = 0; i < NSUM; ++i) st->sums[i] = S[i];
  st->tie_ticket = 0;
}

// Walk a histogram (in shared memory) from the top until k entries are covered -- block-wide: every thread sums its slice
// of bins, a suffix scan over the threads locates the slice in which the running count crosses k, and that one thread
// finishes the walk inside its slice.  (The serial single-thread walk over 2048 bins cost 15-35 us per call.)
//   found : some bin b has  count(bins > b) < k <= count(bins >= b)  -> bin = b, above = count(bins > b)
//   k <= 0: bin = nbins - 1, above = 0 (found)          k > total: bin = 0, above = total (not found)
// Must be called by all threads of the block (blockDim.x <= 512).
struct PickRes { int bin; long long above; int found; };
__device__ PickRes pick_bin_block(const unsigned* sh, int nbins, long long k) {
  __shared__ long long part[512];
  __shared__ int s_bin, s_found;
  __shared__ long long s_above;
  const int nt = blockDim.x, t = threadIdx.x;
  const int per = (nbins + nt - 1) / nt;
  long long local = 0;
  for (int j = 0; j < per; ++j) { const int b = t * per + j; if (b < nbins) local += sh[b]; }
  part[t] = local;
  if (t == 0) { s_found = 0; s_bin = 0; s_above = 0; }
  __syncthreads();
  for (int off = 1; off < nt; off <<= 1) {            // part[t] <- sum of part[u], u >= t
    const long long v = (t + off < nt) ? part[t + off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  const long long incl = part[t], excl = incl - local;
  if (k <= 0) {
    if (t == 0) { s_found = 1; s_bin = nbins - 1; s_above = 0; }
  } else if (excl < k && incl >= k) {                 // exactly one thread
    long long acc = excl;
    for (int j = per - 1; j >= 0; --j) {
      const int b = t * per + j;
      if (b >= nbins) continue;
      const long long c = sh[b];
      if (acc + c >= k) { s_found = 1; s_bin = b; s_above = acc; break; }
      acc += c;
    }
  } else if (t == 0 && part[0] < k) {
    s_above = part[0];                                // k exceeds the total
  }
  __syncthreads();
  PickRes r{s_bin, s_above, s_found};
  __syncthreads();                                    // the shared result may be reused by a later call
  return r;
}

// finalize for 'mean', and level-1 bin search for 'none'
template <bool SELECT>
__global__ void __launch_bounds__(256)
dbloss_finalize1_kernel(const double* __restrict__ partials, int nblk, unsigned* __restrict__ hist1,
                        unsigned* __restrict__ cand_ctl, LossParams lp, float* losses5, DbbLossState* st) {
  __shared__ double S[NSUM];
  __shared__ unsigned sh[SELECT ? HIST1_BINS : 1];
  if (threadIdx.x < NSUM) S[threadIdx.x] = 0.0;
  __syncthreads();
  sum_partials(partials, nblk, S, 0, 9);
  if (SELECT) for (int i = threadIdx.x; i < HIST1_BINS; i += blockDim.x) sh[i] = hist1[i];
  __syncthreads();
  const long long n_pos = (long long)S[S_POS];                       // int(positive.sum())
  long long n_neg = (long long)((double)n_pos * (double)lp.ratio);   // int(no_positive * ratio)
  const long long n_neg_cur = (long long)S[S_NEG];
  if (n_neg_cur < n_neg) n_neg = n_neg_cur;
  const double D = (double)n_pos + (double)n_neg + (double)lp.eps;
  PickRes pick{0, 0, 0};
  if (SELECT) pick = pick_bin_block(sh, HIST1_BINS, n_neg);          // level-1: block-wide walk from the top bin down
  if (threadIdx.x != 0) return;
  st->n_pos = n_pos; st->n_neg = n_neg; st->reduction = SELECT ? 1 : 0;
  if (!SELECT) {
    const double mean_bce = S[S_BCE_ALL] / (double)lp.px;
    const double prob = mean_bce * (S[S_POS] + (double)n_neg) / D;
    st->tau = (float)mean_bce; st->tau_bits = __float_as_uint((float)mean_bce);
    st->n_above = 0; st->n_tie = n_neg;
    write_losses(S, prob, lp, losses5, st, (S[S_POS] + (double)n_neg) / D / (double)lp.px);
    return;
  }
  long long above = 0; int tb = -1;
  if (n_neg > 0) { above = pick.above; tb = pick.found ? pick.bin : -1; }
  // tb == -1: k == 0, or k exceeds the number of strictly positive entries -> tau = 0
  cand_ctl[0] = 0u;
  cand_ctl[1] = (unsigned)tb;
  cand_ctl[2] = (unsigned)(above & 0xffffffffll);
  cand_ctl[3] = (unsigned)(above >> 32);
  for (int i = 0; i < NSUM; ++i) st->sums[i] = S[i];
}

// pass 2 ('none'): sum of everything above the tau bin + compaction of the tau bin (one atomic per warp iteration)
template <int VEC>
__global__ void __launch_bounds__(LOSS_THREADS)
dbloss_select_pass2_kernel(const float* __restrict__ preds, const float* __restrict__ gts, int64_t n_img, int64_t hw, int cch,
                           double* __restrict__ partials, unsigned* __restrict__ cand_ctl, float* __restrict__ cands) {
  const int tb = (int)cand_ctl[1];
  const int64_t px = n_img * hw;
  const int64_t hwv = hw / VEC, nvec = n_img * hwv;
  __shared__ unsigned wtot[2][LOSS_THREADS / 32];      // double buffered by iteration parity
  __shared__ unsigned blk_base[2];
  float acc = 0.f;
  if (tb >= 0) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * LOSS_THREADS;
    const int64_t start = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x;
    const int64_t iters = (nvec + stride - 1) / stride;     // same trip count for every lane (shuffles below)
    for (int64_t it = 0; it < iters; ++it) {
      const int64_t v = start + it * stride;
      float hit[4];
      int cnt = 0;
      if (v < nvec) {
        const unsigned n32 = (unsigned)v / (unsigned)hwv;
        const int64_t n = n32, r = (int64_t)((unsigned)v - n32 * (unsigned)hwv) * VEC;
        Vec4 P = Loader<VEC>::ld(preds + (n * cch) * hw + r);
        Vec4 G = Loader<VEC>::ld(gts + n * hw + r), M = Loader<VEC>::ld(gts + px + n * hw + r);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float nl = negl_of(P.v[j], G.v[j], M.v[j]);
          const int bin = (nl > 0.f) ? (int)(__float_as_uint(nl) >> 20) : -1;
          if (bin > tb) acc += nl;
          if (bin == tb) hit[cnt++] = nl;
        }
      }
      // warp exclusive scan of the hit counts, warp totals combined per block -> ONE atomicAdd per block iteration
      // (the tau bin holds a large share of the negatives: one atomic per warp was ~50 k same-address atomics)
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const int par = (int)(it & 1);
      if (lane == 31) wtot[par][threadIdx.x >> 5] = (unsigned)incl;
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned sum = 0;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) { const unsigned t = wtot[par][w]; wtot[par][w] = sum; sum += t; }
        blk_base[par] = sum ? atomicAdd(&cand_ctl[0], sum) : 0u;
      }
      __syncthreads();
      const unsigned base = blk_base[par] + wtot[par][threadIdx.x >> 5] + (unsigned)(incl - cnt);
      for (int j = 0; j < cnt; ++j) cands[base + j] = hit[j];
    }
  }
  __shared__ float red[LOSS_THREADS / 32];
  const float s = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) t += (double)red[w];
    partials[(size_t)blockIdx.x * NSUM + S_TOP_ABOVE] = t;
  }
}

// exact select inside the tau bin, grid-wide over the compacted candidates:
//   level 2 histogram (bits 19..9) -> pick -> level 3 histogram (bits 8..0 of the matching prefix) + sum above -> pick.
// hist2 / hist3 live behind hist1 in the workspace; sel_ctl: [0] = bin2, [1] = k_rem after level 2 (lo), [2] hi, [3] above lo, [4] above hi
__global__ void __launch_bounds__(LOSS_THREADS)
dbloss_l2hist_kernel(const unsigned* __restrict__ cand_ctl, const float* __restrict__ cands, unsigned* __restrict__ hist2) {
  __shared__ unsigned sh[HIST2_BINS];
  for (int i = threadIdx.x; i < HIST2_BINS; i += LOSS_THREADS) sh[i] = 0;
  __syncthreads();
  const unsigned n = ((int)cand_ctl[1] >= 0) ? cand_ctl[0] : 0u;
  for (unsigned i = blockIdx.x * LOSS_THREADS + threadIdx.x; i < n; i += gridDim.x * LOSS_THREADS)
    atomicAdd(&sh[(__float_as_uint(cands[i]) >> 9) & 0x7ffu], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < HIST2_BINS; i += LOSS_THREADS) { const unsigned c = sh[i]; if (c) atomicAdd(&hist2[i], c); }
}

__global__ void __launch_bounds__(256)
dbloss_pick2_kernel(const unsigned* __restrict__ cand_ctl, const unsigned* __restrict__ hist2, const DbbLossState* __restrict__ st,
                    unsigned* __restrict__ sel_ctl) {
  __shared__ unsigned sh[HIST2_BINS];
  for (int i = threadIdx.x; i < HIST2_BINS; i += blockDim.x) sh[i] = hist2[i];
  __syncthreads();
  const int tb = (int)cand_ctl[1];
  if (tb < 0) {
    if (threadIdx.x == 0) { sel_ctl[0] = 0; sel_ctl[1] = sel_ctl[2] = sel_ctl[3] = sel_ctl[4] = 0; }
    return;
  }
  const long long above1 = ((long long)cand_ctl[3] << 32) | cand_ctl[2];
  const long long krem = st->n_neg - above1;
  const PickRes pk = pick_bin_block(sh, HIST2_BINS, krem);
  if (threadIdx.x != 0) return;
  const int b2 = pk.bin; const long long a2 = pk.above;
  const long long k3 = krem - a2, ab = above1 + a2;
  sel_ctl[0] = (unsigned)b2;
  sel_ctl[1] = (unsigned)(k3 & 0xffffffffll); sel_ctl[2] = (unsigned)(k3 >> 32);
  sel_ctl[3] = (unsigned)(ab & 0xffffffffll); sel_ctl[4] = (unsigned)(ab >> 32);
}

__global__ void __launch_bounds__(LOSS_THREADS)
dbloss_l3hist_kernel(const unsigned* __restrict__ cand_ctl, const float* __restrict__ cands, const unsigned* __restrict__ sel_ctl,
                     unsigned* __restrict__ hist3, double* __restrict__ partials) {
  __shared__ unsigned sh[HIST3_BINS];
  __shared__ double red[LOSS_THREADS / 32];
  for (int i = threadIdx.x; i < HIST3_BINS; i += LOSS_THREADS) sh[i] = 0;
  __syncthreads();
  const unsigned n = ((int)cand_ctl[1] >= 0) ? cand_ctl[0] : 0u;
  const unsigned b2 = sel_ctl[0];
  double acc = 0.0;   // candidates whose level-2 prefix is above the picked bin (all strictly above tau)
  for (unsigned i = blockIdx.x * LOSS_THREADS + threadIdx.x; i < n; i += gridDim.x * LOSS_THREADS) {
    const float c = cands[i];
    const unsigned u = __float_as_uint(c), p2 = (u >> 9) & 0x7ffu;
    if (p2 > b2) acc += (double)c;
    else if (p2 == b2) atomicAdd(&sh[u & 0x1ffu], 1u);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  for (int i = threadIdx.x; i < HIST3_BINS; i += LOSS_THREADS) { const unsigned c = sh[i]; if (c) atomicAdd(&hist3[i], c); }
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) t += red[w];
    partials[(size_t)blockIdx.x * NSUM + S_TOP_CAND] = t;
  }
}

// final pick (bits 8..0) + the five losses for 'none'
__global__ void __launch_bounds__(512)
dbloss_select_final_kernel(double* __restrict__ partials, int nblk, const unsigned* __restrict__ cand_ctl,
                           const unsigned* __restrict__ sel_ctl, const unsigned* __restrict__ hist3, LossParams lp,
                           float* losses5, DbbLossState* st) {
  __shared__ unsigned sh[HIST3_BINS];
  __shared__ double S[NSUM];
  const int tid = threadIdx.x;
  const int tb = (int)cand_ctl[1];
  if (tid < NSUM) S[tid] = st->sums[tid];
  for (int i = tid; i < HIST3_BINS; i += blockDim.x) sh[i] = hist3[i];
  __syncthreads();
  sum_partials(partials, nblk, S, S_TOP_ABOVE, S_TOP_CAND + 1);
  __syncthreads();
  const long long n_pos = st->n_pos, n_neg = st->n_neg;
  const double D = (double)n_pos + (double)n_neg + (double)lp.eps;
  if (tb < 0) {
    if (tid != 0) return;
    // tau = 0: every strictly positive entry is taken, the rest of the picks are zeros
    const long long nz = ((long long)cand_ctl[3] << 32) | cand_ctl[2];
    st->tau = 0.f; st->tau_bits = 0u;
    st->n_above = n_neg > 0 ? nz : 0; st->n_tie = n_neg > 0 ? n_neg - nz : 0;
    const double top = n_neg > 0 ? S[S_NEGL] : 0.0;
    S[S_TOP_CAND] = 0.0; S[S_TOP_ABOVE] = top;
    write_losses(S, (S[S_BCE_POS] + top) / D, lp, losses5, st, 1.0 / D);
    return;
  }
  const unsigned b2 = sel_ctl[0];
  const long long k3 = ((long long)sel_ctl[2] << 32) | sel_ctl[1];
  long long above = ((long long)sel_ctl[4] << 32) | sel_ctl[3];
  const PickRes pk = pick_bin_block(sh, HIST3_BINS, k3);
  const int b3 = pk.bin;
  above += pk.above;
  const unsigned prefix = ((unsigned)tb << 20) | (b2 << 9);
  const unsigned tau_bits = prefix | (unsigned)b3;
  const float tau = __uint_as_float(tau_bits);
  // candidates sharing the 23-bit prefix and a larger low field are single values: count x value is exact
  // (one bin per thread, warp sums combined in warp order: a fixed summation order)
  __shared__ double wsum[16];
  double term = 0.0;
  for (int b = tid; b < HIST3_BINS; b += blockDim.x)
    if (b > b3) term += (double)sh[b] * (double)__uint_as_float(prefix | (unsigned)b);
  term = warp_sum(term);
  if ((tid & 31) == 0) wsum[tid >> 5] = term;
  __syncthreads();
  if (tid != 0) return;
  double same_prefix = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) same_prefix += wsum[w];
  const long long n_tie = n_neg - above;
  st->tau = tau; st->tau_bits = tau_bits; st->n_above = above; st->n_tie = n_tie;
  S[S_TOP_CAND] += same_prefix;
  const double top = S[S_TOP_ABOVE] + S[S_TOP_CAND] + (double)tau * (double)n_tie;
  write_losses(S, (S[S_BCE_POS] + top) / D, lp, losses5, st, 1.0 / D);
}

// ---------------------------------------------------------------------------------------------
// backward: per-pixel gradient of sum_j grad_out[j] * loss[j]
// ---------------------------------------------------------------------------------------------
template <int VEC, bool HAS_B, bool SELECT>
__global__ void __launch_bounds__(LOSS_THREADS)
dbloss_bwd_kernel(const float* __restrict__ preds, const float* __restrict__ gts, int64_t n_img, int64_t hw, int cch,
                  float alpha, float beta, const float* __restrict__ grad_out5, DbbLossState* __restrict__ st,
                  float* __restrict__ dpreds) {
  const int64_t px = n_img * hw;
  const int64_t hwv = hw / VEC, nvec = n_img * hwv;
  const float go0 = grad_out5[0], go1 = grad_out5[1], go2 = grad_out5[2], go3 = grad_out5[3], go4 = grad_out5[4];
  const float c_prob = (go0 + go3 + go4) * st->coef[0];
  const float c_thr = (go1 + beta * (go3 + go4)) * st->coef[1];
  const float c_bin = HAS_B ? (go2 + alpha * go4) : 0.f;
  const float dice_a = st->coef[2], dice_b = st->coef[3];
  const float tau = st->tau;
  const int n_tie = (int)(st->n_tie > 0x7fffffffll ? 0x7fffffffll : st->n_tie);
  const bool any_neg = st->n_neg > 0;

  for (int64_t v = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * LOSS_THREADS) {
    const unsigned n32 = (unsigned)v / (unsigned)hwv;
    const int64_t n = n32, r = (int64_t)((unsigned)v - n32 * (unsigned)hwv) * VEC;
    const int64_t po = (n * cch) * hw + r, go = n * hw + r;
    Vec4 P = Loader<VEC>::ld(preds + po), T = Loader<VEC>::ld(preds + po + hw);
    Vec4 G = Loader<VEC>::ld(gts + go), M = Loader<VEC>::ld(gts + px + go);
    Vec4 TG = Loader<VEC>::ld(gts + 2 * px + go), TM = Loader<VEC>::ld(gts + 3 * px + go);
    Vec4 dP, dT, dB;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float p = P.v[j], g = G.v[j], m = M.v[j];
      float w = 1.f;
      if (SELECT) {
        const float pos = g * m, neg = (1.f - g) * m;
        float sel = 0.f;
        if (any_neg) {
          const float nl = negl_of(p, g, m);
          if (nl > tau) sel = 1.f;
          else if (nl == tau && neg != 0.f && n_tie > 0) {
            // torch.topk leaves the order among equal values unspecified; take n_tie of them by ticket
            if (atomicAdd(&st->tie_ticket, 1) < n_tie) sel = 1.f;
          }
        }
        w = pos + sel * neg;
      }
      dP.v[j] = c_prob * w * bce_bwd(p, g);
      const float d = T.v[j] - TG.v[j];
      dT.v[j] = c_thr * TM.v[j] * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
      if (HAS_B) dB.v[j] = c_bin * m * (dice_b - g * dice_a);
    }
    Loader<VEC>::st(dpreds + po, dP);
    Loader<VEC>::st(dpreds + po + hw, dT);
    if (HAS_B) Loader<VEC>::st(dpreds + po + 2 * hw, dB);
  }
}

// ---------------------------------------------------------------------------------------------
// step function, standalone (DBHead.step_function drop-in; the fused head tail has its own copy)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) step_fwd_kernel(const float* __restrict__ p, const float* __restrict__ t,
                                                       float* __restrict__ b, int64_t n, float k) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    b[i] = 1.f / (1.f + expf(-k * (p[i] - t[i])));
}
__global__ void __launch_bounds__(256) step_bwd_kernel(const float* __restrict__ p, const float* __restrict__ t,
                                                       const float* __restrict__ db, float* __restrict__ dp,
                                                       float* __restrict__ dt, int64_t n, float k) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float e = expf(-k * (p[i] - t[i]));
    const float bb = 1.f / (1.f + e);
    const float d = db[i] * k * bb * bb * e;     // k B^2 e, not k B (1-B): SURVEY.md section 9
    dp[i] = d; dt[i] = -d;
  }
}

// ---------------------------------------------------------------------------------------------
// per-step pixel metric (SURVEY f-3): 2x2 confusion matrix of (P*mask > thresh) vs int(gt*mask)
// replaces src/text_metrics.py:63-82 (cal_text_score) + RunningScore._fast_hist (:14-23), which copy the full P map to
// the host every training step (src/train.py:176-181).  12 B/px read, 4 integers out.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LOSS_THREADS)
text_score_hist_kernel(const float* __restrict__ p, int64_t p_img_stride, const float* __restrict__ gt, const float* __restrict__ mask,
                       int64_t n_img, int64_t hw, float thresh, unsigned long long* __restrict__ hist4) {
  unsigned c[4] = {0u, 0u, 0u, 0u};
  const int64_t total = n_img * hw;
  for (int64_t i = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * LOSS_THREADS) {
    const unsigned n32 = (unsigned)i / (unsigned)hw;
    const int64_t r = (int64_t)((unsigned)i - n32 * (unsigned)hw);
    const float m = mask[i];
    const float pv = p[(int64_t)n32 * p_img_stride + r] * m;
    const int g = (int)(gt[i] * m);                 // .astype(np.int32): truncation
    if (g >= 0 && g < 2) c[g * 2 + (pv > thresh ? 1 : 0)] += 1u;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int s = warp_sum((int)c[k]);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(&hist4[k], (unsigned long long)s);
  }
}

// ---------------------------------------------------------------------------------------------
// optimizer step (SURVEY f-2): Adam over the flat parameter / gradient buffers of the executor
// replaces torch.optim.Adam(dbnet.parameters(), lr=0.005, amsgrad=False).step() (src/train.py:114-117,172): one launch over
// 12.27 M elements (28 B/element) instead of one multi-tensor pass per chunk list.  The step counter lives on the device
// so that the whole training step stays CUDA-graph capturable.
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)   (+ L2 weight decay into g)
// ---------------------------------------------------------------------------------------------
__global__ void adam_tick_kernel(long long* step) { *step += 1; }

__global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n4,
                 float lr, float b1, float b2, float eps, float wd, float grad_scale, const long long* __restrict__ step) {
  const double t = (double)*step;
  const float bc1 = (float)(1.0 - pow((double)b1, t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
  const float step_size = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = ldg_stream(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* P = &pp.x; const float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr = G[k] * grad_scale;
      if (wd != 0.f) gr = fmaf(wd, P[k], gr);
      M[k] = fmaf(1.f - b1, gr - M[k], M[k]);                 // lerp, as torch's fused kernel
      V[k] = fmaf(b2, V[k], (1.f - b2) * gr * gr);
      const float denom = sqrtf(V[k]) / bc2_sqrt + eps;
      P[k] -= step_size * (M[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
}

}  // namespace dbb

using namespace dbb;

// p, g, m, v: flat float32 device buffers of n elements (n % 4 == 0, 16-byte aligned); step: device int64, incremented first
extern "C" int dbb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                             float weight_decay, float grad_scale, long long* step, void* stream) {
  if (!p || !g || !m || !v || !step || n <= 0 || (n & 3)) return set_error(DBB_EINVAL, "adam_step: bad argument (n must be a multiple of 4)");
  if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) return set_error(DBB_EALIGN, "adam_step: buffers need 16-byte alignment");
  cudaStream_t s = (cudaStream_t)stream;
  DBB_LAUNCH("adam_tick", s, adam_tick_kernel<<<1, 1, 0, s>>>(step));
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > DBB_NUM_SMS * 8) blocks = DBB_NUM_SMS * 8;
  DBB_LAUNCH("adam_flat", s, adam_flat_kernel<<<(int)blocks, 256, 0, s>>>(p, g, m, v, n / 4, lr, beta1, beta2, eps, weight_decay, grad_scale, step));
  return DBB_OK;
}

// hist4 (device, 4 x uint64, row-major [gt][pred]) is ACCUMULATED into (zero it to start a new RunningScore)
extern "C" int dbb_text_score_hist(const float* p, int64_t p_img_stride, const float* gt, const float* mask, int64_t n, int64_t h,
                                   int64_t w, float thresh, unsigned long long* hist4, void* stream) {
  if (!p || !gt || !mask || !hist4 || n <= 0 || h <= 0 || w <= 0) return set_error(DBB_EINVAL, "text_score_hist: bad argument");
  if (n * h * w >= ((int64_t)1 << 32)) return set_error(DBB_EUNSUPPORTED, "text_score_hist: more than 2^32 pixels");
  const int grid = loss_grid(n * h * w);
  DBB_LAUNCH("text_score_hist", (cudaStream_t)stream, text_score_hist_kernel<<<grid, LOSS_THREADS, 0, (cudaStream_t)stream>>>(p, p_img_stride, gt, mask, n, h * w, thresh, hist4));
  return DBB_OK;
}

extern "C" size_t dbb_dbloss_workspace(int64_t n, int c, int64_t h, int64_t w, int reduction) {
  (void)c;
  const int64_t px = n * h * w;
  const int grid = loss_grid(px);
  size_t b = align_up(sizeof(double) * NSUM * (size_t)grid, 256) + LOSS_CTL_BYTES;
  if (reduction == 1) b += sizeof(float) * (size_t)px;
  return align_up(b, 256);
}

template <int VEC, bool HAS_B, bool SELECT>
static int loss_fwd_impl(const float* preds, const float* gts, int64_t n, int c, int64_t hw, LossParams lp, float* losses5,
                         DbbLossState* state, void* workspace, cudaStream_t s) {
  const int64_t px = n * hw;
  const int grid = loss_grid(px);
  LossWs ws = carve(workspace, grid, px);
  if (SELECT) DBB_CUDA(cudaMemsetAsync(ws.hist1, 0, LOSS_CTL_BYTES, s));
  DBB_LAUNCH("dbloss_reduce", s, dbloss_reduce_kernel<VEC, HAS_B, SELECT><<<grid, LOSS_THREADS, 0, s>>>(preds, gts, n, hw, c, ws.partials, ws.hist1));
  DBB_LAUNCH("dbloss_finalize1", s, dbloss_finalize1_kernel<SELECT><<<1, 256, 0, s>>>(ws.partials, grid, ws.hist1, ws.cand_count, lp, losses5, state));
  if (SELECT) {
    DBB_LAUNCH("dbloss_select_pass2", s, dbloss_select_pass2_kernel<VEC><<<grid, LOSS_THREADS, 0, s>>>(preds, gts, n, hw, c, ws.partials, ws.cand_count, ws.cands));
    DBB_LAUNCH("dbloss_l2hist", s, dbloss_l2hist_kernel<<<grid, LOSS_THREADS, 0, s>>>(ws.cand_count, ws.cands, ws.hist2));
    DBB_LAUNCH("dbloss_pick2", s, dbloss_pick2_kernel<<<1, 256, 0, s>>>(ws.cand_count, ws.hist2, state, ws.sel_ctl));
    DBB_LAUNCH("dbloss_l3hist", s, dbloss_l3hist_kernel<<<grid, LOSS_THREADS, 0, s>>>(ws.cand_count, ws.cands, ws.sel_ctl, ws.hist3, ws.partials));
    DBB_LAUNCH("dbloss_select_final", s, dbloss_select_final_kernel<<<1, 512, 0, s>>>(ws.partials, grid, ws.cand_count, ws.sel_ctl, ws.hist3, lp, losses5, state));
  }
  return DBB_OK;
}

extern "C" int dbb_dbloss_fwd(const float* preds, const float* gts, int64_t n, int c, int64_t h, int64_t w, float alpha,
                              float beta, int reduction, float negative_ratio, float eps, float* losses5,
                              DbbLossState* state, void* workspace, size_t workspace_bytes, void* stream) {
  if (!preds || !gts || !losses5 || !state || !workspace) return set_error(DBB_EINVAL, "dbloss_fwd: null pointer");
  if (n <= 0 || h <= 0 || w <= 0 || (c != 2 && c != 3) || (reduction != 0 && reduction != 1))
    return set_error(DBB_EINVAL, "dbloss_fwd: bad shape or reduction");
  if (!aligned16(preds) || !aligned16(gts) || !aligned16(workspace)) return set_error(DBB_EALIGN, "dbloss_fwd: pointer not 16B aligned");
  if (n * h * w >= ((int64_t)1 << 32)) return set_error(DBB_EUNSUPPORTED, "dbloss_fwd: more than 2^32 pixels");
  if (workspace_bytes < dbb_dbloss_workspace(n, c, h, w, reduction)) return set_error(DBB_EWORKSPACE, "dbloss_fwd: workspace too small");
  const int64_t hw = h * w;
  LossParams lp{alpha, beta, negative_ratio, eps, n * hw, c == 3};
  cudaStream_t s = (cudaStream_t)stream;
  const bool v4 = (hw % 4) == 0;
#define DISPATCH(V, B, S) return loss_fwd_impl<V, B, S>(preds, gts, n, c, hw, lp, losses5, state, workspace, s)
  if (v4) {
    if (c == 3) { if (reduction) DISPATCH(4, true, true); else DISPATCH(4, true, false); }
    else        { if (reduction) DISPATCH(4, false, true); else DISPATCH(4, false, false); }
  } else {
    if (c == 3) { if (reduction) DISPATCH(1, true, true); else DISPATCH(1, true, false); }
    else        { if (reduction) DISPATCH(1, false, true); else DISPATCH(1, false, false); }
  }
#undef DISPATCH
}

extern "C" int dbb_dbloss_bwd(const float* preds, const float* gts, int64_t n, int c, int64_t h, int64_t w, float alpha,
                              float beta, int reduction, float eps, const float* grad_out5, DbbLossState* state,
                              float* dpreds, void* stream) {
  (void)eps;
  if (!preds || !gts || !grad_out5 || !state || !dpreds) return set_error(DBB_EINVAL, "dbloss_bwd: null pointer");
  if (n <= 0 || h <= 0 || w <= 0 || (c != 2 && c != 3) || (reduction != 0 && reduction != 1))
    return set_error(DBB_EINVAL, "dbloss_bwd: bad shape or reduction");
  if (!aligned16(preds) || !aligned16(gts) || !aligned16(dpreds)) return set_error(DBB_EALIGN, "dbloss_bwd: pointer not 16B aligned");
  const int64_t hw = h * w;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = loss_grid(n * hw);
  const bool v4 = (hw % 4) == 0;
  if (reduction) DBB_CUDA(cudaMemsetAsync(&state->tie_ticket, 0, sizeof(int), s));
#define LAUNCH(V, B, S) DBB_LAUNCH("dbloss_bwd", s, dbloss_bwd_kernel<V, B, S><<<grid, LOSS_THREADS, 0, s>>>(preds, gts, n, hw, c, alpha, beta, grad_out5, state, dpreds))
  if (v4) {
    if (c == 3) { if (reduction) LAUNCH(4, true, true); else LAUNCH(4, true, false); }
    else        { if (reduction) LAUNCH(4, false, true); else LAUNCH(4, false, false); }
  } else {
    if (c == 3) { if (reduction) LAUNCH(1, true, true); else LAUNCH(1, true, false); }
    else        { if (reduction) LAUNCH(1, false, true); else LAUNCH(1, false, false); }
  }
#undef LAUNCH
  return DBB_OK;
}

extern "C" int dbb_step_fwd(const float* p, const float* t, float* b, int64_t numel, float k, void* stream) {
  if (!p || !t || !b || numel < 0) return set_error(DBB_EINVAL, "step_fwd: bad argument");
  if (numel == 0) return DBB_OK;
  int64_t g = (numel + 255) / 256; if (g > DBB_NUM_SMS * 8) g = DBB_NUM_SMS * 8;
  DBB_LAUNCH("step_fwd", (cudaStream_t)stream, step_fwd_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(p, t, b, numel, k));
  return DBB_OK;
}
extern "C" int dbb_step_bwd(const float* p, const float* t, const float* db, float* dp, float* dt, int64_t numel, float k,
                            void* stream) {
  if (!p || !t || !db || !dp || !dt || numel < 0) return set_error(DBB_EINVAL, "step_bwd: bad argument");
  if (numel == 0) return DBB_OK;
  int64_t g = (numel + 255) / 256; if (g > DBB_NUM_SMS * 8) g = DBB_NUM_SMS * 8;
  DBB_LAUNCH("step_bwd", (cudaStream_t)stream, step_bwd_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(p, t, db, dp, dt, numel, k));
  return DBB_OK;
}
