// ccl.cu -- GPU front of SegDetectorRepresenter: binarize + connected components + box scores, contour-free.
//
// Replaces, for every image of the batch at once and without leaving the device:
//   src/postprocess.py:51-52   binarize:            bitmap = P > thresh   (strict, float32)
//   src/postprocess.py:67,116  cv2.findContours(RETR_LIST): one OUTER border per 8-connected foreground component and
//                              one HOLE border per 4-connected background region enclosed by foreground
//   src/postprocess.py:186-198 box_score_fast:      float64 mean of P over fillPoly(contour)
//   src/postprocess.py:80,129  the score filter     keep = not (box_thresh > score)
// using the set identities verified against OpenCV in tests/test_oracle_golden.py:
//   fill(outer border of F) = F + everything in the containment tree below F
//   fill(hole border of G)  = G + everything below G + the pixels of G's parent component that are 4-adjacent to G
//
// Pipeline (all images in one grid; labels are pixel indices, root = smallest index of the component, so the root IS the
// raster-first pixel = cv2's discovery point).  A warp owns one 32-pixel row segment of a bit-packed bitmap:
//   A pack_init  binarize (strict >), pack 32 px / word, byte bitmap, label = first pixel of the segment run   (4 B read, 5 B written / px)
//   B link       one union per pair of touching runs (fg: vertical + the two diagonals, bg: vertical), segment seams,
//                frame runs -> virtual outside node; two-level (32-row strips, then strip seams) to keep chains short
//   C flatten    run starts -> root; roots zero their statistics slot
//   D stats      final labels + per-run float64 sums (warp prefix sum) -> one atomic set per run; hole boundary rings;
//                the outside region is skipped (it is no candidate and would serialise the atomics)
//   E tree       every node adds its statistics to all its ancestors (parent = region north of the root pixel)
//   F rank/emit  suffix count of roots in raster order = position in cv2's reverse-discovery order -> candidates are
//                written already sorted, the first max_cands of them (src/postprocess.py:70,119)
#include "common.cuh"

namespace dbb {

constexpr int CCL_THREADS = 256;

struct CompStat {          // one slot per pixel index, touched only at roots
  double sum;              // own pixels
  double acc_sum;          // descendants (+ boundary ring for holes)
  int count, acc_count;
  int x0, y0, x1, y1;
};

struct CclWs {
  unsigned* bits;          // [n][h][wq]    packed bitmap, 32 pixels per word
  int* label;              // [n][hw + 1]   (+1: virtual outside node)
  CompStat* stat;          // [n][hw]
  int* blk_count;          // [n][nblk]
  int* blk_off;            // [n][nblk]
  unsigned* rootbits;      // [n][h][wq]    1 where the pixel is the root of its component (written by the final flatten)
};

__host__ __device__ inline size_t ccl_align(size_t v) { return (v + 255) / 256 * 256; }

static int ccl_nblk(int64_t hw) { return (int)((hw + CCL_THREADS - 1) / CCL_THREADS); }

static CclWs ccl_carve(void* ws, int64_t n, int64_t hw, int64_t h, int64_t wq) {
  CclWs w;
  char* p = (char*)ws;
  w.bits = (unsigned*)p;    p += ccl_align(sizeof(unsigned) * (size_t)n * h * wq);
  w.label = (int*)p;        p += ccl_align(sizeof(int) * (size_t)n * (hw + 1));
  w.stat = (CompStat*)p;    p += ccl_align(sizeof(CompStat) * (size_t)n * hw);
  w.blk_count = (int*)p;    p += ccl_align(sizeof(int) * (size_t)n * ccl_nblk(hw));
  w.blk_off = (int*)p;      p += ccl_align(sizeof(int) * (size_t)n * ccl_nblk(hw));
  w.rootbits = (unsigned*)p;
  return w;
}

__device__ __forceinline__ int uf_find(const int* L, int i) {
  int p = L[i];
  while (p != i) { i = p; p = L[i]; }
  return i;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  while (true) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }     // a > b: hang the larger root under the smaller
    const int old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;                                          // somebody else re-parented a meanwhile: retry from there
  }
}

// ---------------------------------------------------------------------------------------------
// Run-based labelling.  A warp owns one 32-pixel segment of a row; the segment's bits come from a packed bitmap
// (1 bit / pixel), so every neighbourhood test is a few word operations and unions are issued once per pair of
// touching runs instead of once per pixel.
// ---------------------------------------------------------------------------------------------
struct Seg { int img, y, s, x0, nvalid; };
// (32-bit arithmetic on purpose: the entry points check n*h*wq < 2^32, and a 64-bit divide is a ~100-instruction loop --
//  ncu showed ccl_pack_init at 150 warp-instructions per word with 64-bit decodes)
__device__ __forceinline__ bool seg_of(int64_t widx, int h, int wq, int w, Seg& g) {
  const unsigned u = (unsigned)widx;
  g.s = (int)(u % (unsigned)wq); const unsigned t = u / (unsigned)wq;
  g.y = (int)(t % (unsigned)h); g.img = (int)(t / (unsigned)h);
  g.x0 = g.s * 32;
  g.nvalid = w - g.x0 < 32 ? w - g.x0 : 32;
  return true;
}
__device__ __forceinline__ unsigned valid_mask(int nvalid) { return nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u); }
// lanes where a run (maximal stretch of equal class inside the segment) starts
__device__ __forceinline__ unsigned run_starts(unsigned cur, int nvalid) {
  unsigned st = (cur ^ (cur << 1)) | 1u;
  if (nvalid < 32) st |= 1u << nvalid;     // sentinel: terminates the last valid run
  return st;
}

// A: binarize (strict >), pack, write the byte bitmap, label every pixel with the first pixel of its segment run
__global__ void __launch_bounds__(CCL_THREADS)
ccl_pack_init_kernel(const float* __restrict__ pred, int c, int n, int h, int w, int wq, float thresh, uint8_t* __restrict__ bitmap,
                     unsigned* __restrict__ bits, int* __restrict__ label) {
  const int64_t hw = (int64_t)h * w;
  const int lane = threadIdx.x & 31;
  const int64_t nseg = (int64_t)n * h * wq;
  const int64_t wstride = (int64_t)gridDim.x * (CCL_THREADS / 32);
  for (int64_t widx = (int64_t)blockIdx.x * (CCL_THREADS / 32) + (threadIdx.x >> 5); widx < nseg; widx += wstride) {
    Seg g; seg_of(widx, h, wq, w, g);
    const bool in = lane < g.nvalid;
    const int64_t pix = (int64_t)g.y * w + g.x0 + lane;
    const float p = in ? pred[(int64_t)g.img * c * hw + pix] : 0.f;
    const unsigned cur = __ballot_sync(0xffffffffu, in && p > thresh);
    if (lane == 0) bits[widx] = cur;
    if (in) {
      bitmap[g.img * hw + pix] = (cur >> lane) & 1u;
      const unsigned st = run_starts(cur, g.nvalid) & ((2u << lane) - 1u);
      label[g.img * (hw + 1) + pix] = (int)((int64_t)g.y * w + g.x0 + (31 - __clz(st)));
    }
    if (g.y == 0 && g.s == 0 && lane == 0) label[g.img * (hw + 1) + hw] = (int)hw;   // virtual outside node
  }
}

// B: unions between touching runs: across segment boundaries, with the row above (4-connectivity for background,
// 8-connectivity for foreground), and background runs on the image frame with the virtual outside node.
// ONE THREAD PER 32-PIXEL WORD: the neighbourhood masks are a dozen word operations computed once, and the thread then walks
// the set bits (typically 0-2 unions per word).  The first version gave every pixel a lane that recomputed the same masks and
// diverged on its own union: 30 warp-instructions per pixel at 5.7 active lanes (ncu, profiles/prof_ccl_r02.md) -- issue-bound.
__device__ __forceinline__ void link_word(const unsigned* __restrict__ bits, int64_t widx, int h, int w, int wq, int* __restrict__ label, int phase) {
  const int64_t hw = (int64_t)h * w;
  Seg g; seg_of(widx, h, wq, w, g);
  // two-level merge keeps union-find chains short: phase 0 links everything except across the boundaries of
  // 32-row strips (chains <= 32), the strips are flattened, phase 1 links the strip boundaries (chains <= H/32)
  const bool strip_edge = (g.y & 31) == 0;
  int* L = label + g.img * (hw + 1);
  const unsigned cur = bits[widx];
  const unsigned prv = g.s > 0 ? bits[widx - 1] : 0u, nxt = g.s < wq - 1 ? bits[widx + 1] : 0u;
  const bool has_up = g.y > 0 && (phase == 1 || !strip_edge);
  const unsigned vm = valid_mask(g.nvalid);
  const int i0 = g.y * w + g.x0;
  if (has_up) {
    const unsigned up = bits[widx - wq];
    const unsigned upp = g.s > 0 ? bits[widx - wq - 1] : 0u, upn = g.s < wq - 1 ? bits[widx - wq + 1] : 0u;
    // neighbours shifted into lane position: L = pixel x-1, R = pixel x+1
    const unsigned curL = (cur << 1) | (prv >> 31), curR = (cur >> 1) | (nxt << 31);
    const unsigned upL = (up << 1) | (upp >> 31), upR = (up >> 1) | (upn << 31);
    const unsigned hasL = g.s > 0 ? 0xffffffffu : 0xfffffffeu;                         // pixel x-1 exists
    unsigned V = cur & up & ~(curL & upL & hasL);            // first column of a vertical overlap
    unsigned DL = cur & ~up & upL & ~curL & hasL;            // diagonal up-left, not implied by a neighbour
    unsigned DR = cur & ~up & upR & ~curR;                   // diagonal up-right (bits beyond the row end are 0)
    if (g.s == wq - 1 && g.nvalid >= 1) DR &= ~(1u << (g.nvalid - 1));     // x + 1 < w
    while (V) { const int b = __ffs(V) - 1; V &= V - 1; uf_union(L, i0 + b, i0 + b - w); }
    while (DL) { const int b = __ffs(DL) - 1; DL &= DL - 1; uf_union(L, i0 + b, i0 + b - w - 1); }
    while (DR) { const int b = __ffs(DR) - 1; DR &= DR - 1; uf_union(L, i0 + b, i0 + b - w + 1); }
    const unsigned ncur = ~cur & vm;
    unsigned VB = ncur & ~up & ~(~curL & ~upL & hasL);       // background: vertical links only (4-connectivity)
    while (VB) { const int b = __ffs(VB) - 1; VB &= VB - 1; uf_union(L, i0 + b, i0 + b - w); }
  }
  if (phase != 0) return;
  // horizontal link across the segment boundary (inside a segment the run label already encodes it)
  if (g.s > 0 && (((cur & 1u) != 0) == ((prv >> 31) != 0))) uf_union(L, i0, i0 - 1);
  // frame pixels of the background belong to the outside region: one union per background run start on the first / last
  // row, the first / last pixel of every other row
  const unsigned ncur = ~cur & vm;
  if (g.y == 0 || g.y == h - 1) {
    unsigned st = run_starts(cur, g.nvalid) & ncur;
    while (st) { const int b = __ffs(st) - 1; st &= st - 1; uf_union(L, (int)hw, i0 + b); }
  } else {
    if (g.s == 0 && (ncur & 1u)) uf_union(L, (int)hw, i0);
  }
  if (g.s == wq - 1 && ((ncur >> (g.nvalid - 1)) & 1u)) uf_union(L, (int)hw, i0 + g.nvalid - 1);
}
__global__ void __launch_bounds__(CCL_THREADS)
ccl_link_kernel(const unsigned* __restrict__ bits, int n, int h, int w, int wq, int* __restrict__ label, int phase) {
  const int64_t stride = (int64_t)gridDim.x * CCL_THREADS;
  if (phase == 0) {
    const int64_t nseg = (int64_t)n * h * wq;
    for (int64_t widx = (int64_t)blockIdx.x * CCL_THREADS + threadIdx.x; widx < nseg; widx += stride) link_word(bits, widx, h, w, wq, label, 0);
  } else {
    // only the first row of every 32-row strip (except row 0) takes part
    const int nedge = (h - 1) / 32;                          // rows 32, 64, ...
    const int64_t total = (int64_t)n * nedge * wq;
    for (int64_t t = (int64_t)blockIdx.x * CCL_THREADS + threadIdx.x; t < total; t += stride) {
      const int s = (int)(t % wq); const int64_t q = t / wq;
      const int e = (int)(q % nedge), img = (int)(q / nedge);
      link_word(bits, ((int64_t)img * h + (int64_t)(e + 1) * 32) * wq + s, h, w, wq, label, 1);
    }
  }
}

// C: run-start pixels jump straight to their root; roots zero their statistics slot.  One thread per word (see B).
__global__ void __launch_bounds__(CCL_THREADS)
ccl_flatten_kernel(const unsigned* __restrict__ bits, int n, int h, int w, int wq, int* __restrict__ label, CompStat* __restrict__ stat,
                   int init_stats, unsigned* __restrict__ rootbits) {
  const int64_t hw = (int64_t)h * w;
  const int64_t nseg = (int64_t)n * h * wq;
  for (int64_t widx = (int64_t)blockIdx.x * CCL_THREADS + threadIdx.x; widx < nseg; widx += (int64_t)gridDim.x * CCL_THREADS) {
    Seg g; seg_of(widx, h, wq, w, g);
    int* L = label + g.img * (hw + 1);
    if (g.y == 0 && g.s == 0) L[hw] = uf_find(L, (int)hw);
    unsigned st = run_starts(bits[widx], g.nvalid) & valid_mask(g.nvalid);
    unsigned roots = 0;
    while (st) {
      const int b = __ffs(st) - 1; st &= st - 1;
      const int i = g.y * w + g.x0 + b;
      const int r = uf_find(L, i);
      L[i] = r;
      if (r == i && init_stats) {
        CompStat z;
        z.sum = 0.0; z.acc_sum = 0.0; z.count = 0; z.acc_count = 0;
        z.x0 = w; z.y0 = h; z.x1 = -1; z.y1 = -1;
        stat[g.img * hw + i] = z;
        roots |= 1u << b;
      }
    }
    if (init_stats) rootbits[widx] = roots;
  }
}

// parent region of a root pixel r: the region containing the pixel north of it (outside for the first row).
// two-hop lookup: a label is either already the root or a run start whose label is the root.
__device__ __forceinline__ int root2(const int* L, int i) { return L[L[i]]; }
__device__ __forceinline__ int parent_of(const int* L, int r, int w, int r_out) { return r < w ? r_out : root2(L, r - w); }

// D: final labels + per-component statistics.  One atomic set per (segment run) instead of per pixel: run sums come from
// a warp prefix sum in float64; the outside region is skipped; foreground pixels that 4-touch an enclosed background
// region add themselves to that hole's boundary ring.
__global__ void __launch_bounds__(CCL_THREADS)
ccl_stats_kernel(const float* __restrict__ pred, int c, const unsigned* __restrict__ bits, int n, int h, int w, int wq,
                 int* __restrict__ label, CompStat* __restrict__ stat) {
  const int64_t hw = (int64_t)h * w;
  const int lane = threadIdx.x & 31;
  const int64_t nseg = (int64_t)n * h * wq;
  const int64_t wstride = (int64_t)gridDim.x * (CCL_THREADS / 32);
  for (int64_t widx = (int64_t)blockIdx.x * (CCL_THREADS / 32) + (threadIdx.x >> 5); widx < nseg; widx += wstride) {
    Seg g; seg_of(widx, h, wq, w, g);
    int* L = label + g.img * (hw + 1);
    CompStat* S = stat + g.img * hw;
    const int r_out = L[hw];
    const bool in = lane < g.nvalid;
    const int x = g.x0 + lane;
    const int i = g.y * w + x;
    const unsigned cur = bits[widx];
    const unsigned st = run_starts(cur, g.nvalid);
    // this lane's run: [a, b]
    const int a = 31 - __clz(st & ((2u << lane) - 1u));
    const unsigned later = (lane >= 31) ? 0u : (st & ~((2u << lane) - 1u));
    const int b = later ? (__ffs(later) - 2) : 31;
    int r = -1;
    if (in) r = root2(L, g.y * w + g.x0 + a);            // every lane of a run reads the same word
    const float p = in ? pred[(int64_t)g.img * c * hw + i] : 0.f;
    // inclusive prefix sum over the warp in float64
    double pre = (double)p;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    const double pre_b = __shfl_sync(0xffffffffu, pre, b < g.nvalid ? b : g.nvalid - 1 < 0 ? 0 : (b > 31 ? 31 : b));
    if (!in) continue;
    if (r == r_out) continue;
    if (lane == a) {
      const int be = b < g.nvalid ? b : g.nvalid - 1;
      const double run_sum = pre_b - pre + (double)p;
      CompStat* t = S + r;
      atomicAdd(&t->sum, run_sum);
      atomicAdd(&t->count, be - a + 1);
      atomicMin(&t->x0, g.x0 + a); atomicMax(&t->x1, g.x0 + be); atomicMin(&t->y0, g.y); atomicMax(&t->y1, g.y);
    }
    // boundary ring of holes: a foreground pixel contributes once to every DISTINCT enclosed region it 4-touches
    if ((cur >> lane) & 1u) {
      const unsigned prv = g.s > 0 ? bits[widx - 1] : 0xffffffffu, nxt = g.s < wq - 1 ? bits[widx + 1] : 0xffffffffu;
      const bool bgL = x > 0 && !(lane > 0 ? (cur >> (lane - 1)) & 1u : (prv >> 31) & 1u);
      const bool bgR = x < w - 1 && !(lane < 31 ? (cur >> (lane + 1)) & 1u : nxt & 1u);
      const bool bgU = g.y > 0 && !((bits[widx - wq] >> lane) & 1u);
      const bool bgD = g.y < h - 1 && !((bits[widx + wq] >> lane) & 1u);
      if (bgL | bgR | bgU | bgD) {
        const int pr = parent_of(L, r, w, r_out);
        int gq[4] = {-1, -1, -1, -1};
        if (bgL) gq[0] = root2(L, i - 1);
        if (bgR) gq[1] = root2(L, i + 1);
        if (bgU) gq[2] = root2(L, i - w);
        if (bgD) gq[3] = root2(L, i + w);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int gk = gq[k];
          if (gk < 0 || gk == pr || gk == r_out) continue;
          bool dup = false;
#pragma unroll
          for (int q = 0; q < 4; ++q) if (q < k && gq[q] == gk) dup = true;
          if (dup) continue;
          atomicAdd(&S[gk].acc_sum, (double)p);
          atomicAdd(&S[gk].acc_count, 1);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(CCL_THREADS)
ccl_tree_kernel(const uint8_t* __restrict__ bitmap, int h, int w, const int* __restrict__ label, CompStat* __restrict__ stat) {
  const int img = blockIdx.y;
  const int64_t hw = (int64_t)h * w;
  const int* L = label + img * (hw + 1);
  CompStat* S = stat + img * hw;
  const int r_out = L[hw];
  for (int64_t i = (int64_t)blockIdx.x * CCL_THREADS + threadIdx.x; i < hw; i += (int64_t)gridDim.x * CCL_THREADS) {
    if (L[i] != (int)i || (int)i == r_out) continue;
    const double s = S[i].sum;
    const int cnt = S[i].count;
    int a = parent_of(L, (int)i, w, r_out);
    while (a != r_out) {
      atomicAdd(&S[a].acc_sum, s);
      atomicAdd(&S[a].acc_count, cnt);
      a = parent_of(L, a, w, r_out);
    }
  }
}

// ---- ranking: candidates in cv2 order = roots by DESCENDING pixel index.  Roots are known as one bit per pixel
// (ccl_flatten_kernel), so counting and ranking touch 1/32 of a word per pixel instead of the 4-byte label.
// root bits of word wi of an image, the outside region's root removed
__device__ __forceinline__ unsigned cand_bits(const unsigned* __restrict__ rootbits, int64_t base, int wi, int nw, int w, int wq, int r_out) {
  if (wi >= nw) return 0u;
  unsigned rb = rootbits[base + wi];
  const int y = wi / wq, x0 = (wi - y * wq) * 32;
  const int o = r_out - (y * w + x0);
  if (o >= 0 && o < 32 && r_out < (y + 1) * w) rb &= ~(1u << o);
  return rb;
}
__global__ void __launch_bounds__(CCL_THREADS)
ccl_count_kernel(int h, int w, int wq, const int* __restrict__ label, const unsigned* __restrict__ rootbits, int* __restrict__ blk_count, int nblk) {
  const int img = blockIdx.y, b = blockIdx.x;
  const int nw = h * wq;
  const int r_out = label[img * ((int64_t)h * w + 1) + (int64_t)h * w];
  const int c = __popc(cand_bits(rootbits, (int64_t)img * nw, b * CCL_THREADS + threadIdx.x, nw, w, wq, r_out));
  __shared__ int wsum[CCL_THREADS / 32];
  const int ws = warp_sum(c);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = ws;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < CCL_THREADS / 32; ++k) t += wsum[k];
    blk_count[img * nblk + b] = t;
  }
}
__global__ void __launch_bounds__(1024)
ccl_scan_kernel(const int* __restrict__ blk_count, int* __restrict__ blk_off, int nblk, int* __restrict__ n_cands) {
  // one CTA per image: exclusive SUFFIX sum over blocks
  const int img = blockIdx.x;
  __shared__ int carry;
  __shared__ int sh[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = ((nblk - 1) / 1024) * 1024; base >= 0; base -= 1024) {
    const int b = base + threadIdx.x;
    const int v = b < nblk ? blk_count[img * nblk + b] : 0;
    // inclusive suffix scan inside the chunk
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = (threadIdx.x + o < 1024) ? sh[threadIdx.x + o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (b < nblk) blk_off[img * nblk + b] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += sh[0];
    __syncthreads();
  }
  if (threadIdx.x == 0) n_cands[img] = carry;
}
__global__ void __launch_bounds__(CCL_THREADS)
ccl_emit_kernel(const uint8_t* __restrict__ bitmap, int h, int w, int wq, const int* __restrict__ label, CompStat* __restrict__ stat,
                const unsigned* __restrict__ rootbits, const int* __restrict__ blk_off, int nblk, double box_thresh,
                DbbCandidate* __restrict__ cands, int max_cands) {
  const int img = blockIdx.y, b = blockIdx.x;
  const int64_t hw = (int64_t)h * w;
  const int nw = h * wq;
  const uint8_t* bm = bitmap + img * hw;
  const int* L = label + img * (hw + 1);
  CompStat* S = stat + img * hw;
  const int r_out = L[hw];
  const int wi = b * CCL_THREADS + threadIdx.x;
  unsigned rb = cand_bits(rootbits, (int64_t)img * nw, wi, nw, w, wq, r_out);
  const int c = __popc(rb);
  // roots in LATER words of this block (suffix count): warp suffix by shuffles + totals of the later warps
  __shared__ int wcount[CCL_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int suf = c;                                   // inclusive suffix over the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_down_sync(0xffffffffu, suf, o);
    if (lane + o < 32) suf += t;
  }
  if (lane == 0) wcount[wid] = suf;
  __syncthreads();
  if (!rb) return;
  int rank = suf - c;
  for (int k = wid + 1; k < CCL_THREADS / 32; ++k) rank += wcount[k];
  rank += blk_off[img * nblk + b];
  const int y = wi / wq, x0 = (wi - y * wq) * 32;
  while (rb) {                                   // highest pixel index first
    const int bit = 31 - __clz(rb);
    rb &= ~(1u << bit);
    const int i = y * w + x0 + bit;
    const CompStat s = S[i];
    // (last reader of the statistics slot: acc_count now carries the candidate's output slot for ccl_points_kernel,
    //  -1 = not an emitted, kept candidate)
    const double sc = (s.sum + s.acc_sum) / (double)(s.count + s.acc_count);
    const int keep = (box_thresh > sc) ? 0 : 1;
    S[i].acc_count = (rank < max_cands && keep) ? rank : -1;
    if (rank < max_cands) {
      DbbCandidate cd;
      cd.kind = bm[i] ? 0 : 1;
      cd.first_y = y; cd.first_x = x0 + bit;
      if (cd.kind == 0) { cd.x0 = s.x0; cd.y0 = s.y0; cd.x1 = s.x1; cd.y1 = s.y1; }
      else { cd.x0 = s.x0 - 1; cd.y0 = s.y0 - 1; cd.x1 = s.x1 + 1; cd.y1 = s.y1 + 1; }   // + the ring of parent pixels
      cd.count = s.count + s.acc_count;
      cd.sum = s.sum + s.acc_sum;
      cd.keep = keep;
      cd.pad_ = 0;
      cands[(int64_t)img * max_cands + rank] = cd;
    }
    ++rank;
  }
}
// optional per-pixel label map (tests / debugging): fg 1 + root index, bg -(1 + root index)
__global__ void __launch_bounds__(CCL_THREADS)
ccl_labels_kernel(const uint8_t* __restrict__ bitmap, int64_t hw, const int* __restrict__ label, int32_t* __restrict__ labels_out) {
  const int img = blockIdx.y;
  const int* L = label + img * (hw + 1);
  for (int64_t i = (int64_t)blockIdx.x * CCL_THREADS + threadIdx.x; i < hw; i += (int64_t)gridDim.x * CCL_THREADS) {
    const int rr = L[L[i]];                      // run starts carry their root; every other pixel points at its run start
    labels_out[img * hw + i] = bitmap[img * hw + i] ? (rr + 1) : -(rr + 1);
  }
}

// G: border points of the KEPT candidates (src/postprocess.py:119-121: the contour handed to get_mini_boxes).  cv2.minAreaRect
// only sees the convex hull of the contour, and hull(outer border of F) = hull(pixels of F) = hull(end pixels of F's runs),
// hull(hole border of G) = hull(foreground pixels 4-adjacent to G) -- so the device emits, per kept candidate, the two end
// pixels of each of its segment runs (holes: the foreground pixels left / right of each run and the first / last foreground
// pixel above / below it).  A few thousand (slot, x, y) triples per image cross to the host instead of the bitmap.
// points of one word; WRITE = false only counts them
template <bool WRITE>
__device__ __forceinline__ int word_points(const unsigned* __restrict__ bits, int64_t widx, const Seg& g, int h, int w, int wq, const int* __restrict__ L,
                                           const CompStat* __restrict__ S, int r_out, int32_t* __restrict__ out, int base, int cap) {
  const unsigned cur = bits[widx];
  const unsigned vm = valid_mask(g.nvalid);
  const unsigned st = run_starts(cur, g.nvalid);
  unsigned todo = st & vm;
  int total = 0;
  while (todo) {
    const int a = __ffs(todo) - 1; todo &= todo - 1;
    const unsigned later = (a >= 31) ? 0u : (st & ~((2u << a) - 1u));
    int b = later ? (__ffs(later) - 2) : 31;
    if (b > g.nvalid - 1) b = g.nvalid - 1;
    const int r = L[g.y * w + g.x0 + a];                      // run starts carry their root (ccl_flatten_kernel)
    if (r == r_out) continue;
    const int slot = S[r].acc_count;
    if (slot < 0) continue;
    int px[6], py[6], k = 0;
    const int xs = g.x0 + a, xe = g.x0 + b;
    // does the run really start / end here, or does it continue in the neighbouring 32-pixel word?
    const bool same_class_left = a == 0 && g.s > 0 && (((bits[widx - 1] >> 31) & 1u) == ((cur >> a) & 1u));
    const bool same_class_right = b == g.nvalid - 1 && g.s < wq - 1 && ((bits[widx + 1] & 1u) == ((cur >> b) & 1u));
    if ((cur >> a) & 1u) {
      if (!same_class_left) { px[k] = xs; py[k++] = g.y; }
      if (!same_class_right && (xe != xs || same_class_left)) { px[k] = xe; py[k++] = g.y; }
    } else {
      const unsigned runmask = ((b >= 31) ? 0xffffffffu : ((2u << b) - 1u)) & ~((1u << a) - 1u);
      // left / right foreground neighbours (a hole never touches the frame, so they exist whenever the run really ends here)
      if (!same_class_left && xs > 0) { px[k] = xs - 1; py[k++] = g.y; }
      if (!same_class_right && xe < w - 1) { px[k] = xe + 1; py[k++] = g.y; }
      if (g.y > 0) {
        const unsigned m = bits[widx - wq] & runmask;
        if (m) { px[k] = g.x0 + __ffs(m) - 1; py[k++] = g.y - 1; if (m & (m - 1)) { px[k] = g.x0 + 31 - __clz(m); py[k++] = g.y - 1; } }
      }
      if (g.y < h - 1) {
        const unsigned m = bits[widx + wq] & runmask;
        if (m) { px[k] = g.x0 + __ffs(m) - 1; py[k++] = g.y + 1; if (m & (m - 1)) { px[k] = g.x0 + 31 - __clz(m); py[k++] = g.y + 1; } }
      }
    }
    if (WRITE) {
      for (int j = 0; j < k; ++j) {
        if (base + total + j < cap) { out[2 * (base + total + j)] = slot; out[2 * (base + total + j) + 1] = (py[j] << 16) | px[j]; }
      }
    }
    total += k;
  }
  return total;
}
__global__ void __launch_bounds__(CCL_THREADS)
ccl_points_kernel(const unsigned* __restrict__ bits, int n, int h, int w, int wq, const int* __restrict__ label,
                  const CompStat* __restrict__ stat, int32_t* __restrict__ points, int32_t* __restrict__ n_points, int cap) {
  const int64_t hw = (int64_t)h * w;
  const int64_t nseg = (int64_t)n * h * wq;
  const int lane = threadIdx.x & 31;
  // one thread per 32-pixel word (see ccl_link_kernel); count, reserve with ONE atomic per warp and image, then write
  const int64_t nround = (nseg + CCL_THREADS - 1) / CCL_THREADS * CCL_THREADS;
  for (int64_t widx = (int64_t)blockIdx.x * CCL_THREADS + threadIdx.x; widx < nround; widx += (int64_t)gridDim.x * CCL_THREADS) {
    const bool in = widx < nseg;
    Seg g; g.img = -1;
    int cnt = 0;
    const int* L = nullptr; const CompStat* S = nullptr; int r_out = 0;
    if (in) {
      seg_of(widx, h, wq, w, g);
      L = label + g.img * (hw + 1); S = stat + g.img * hw; r_out = L[hw];
      cnt = word_points<false>(bits, widx, g, h, w, wq, L, S, r_out, nullptr, 0, 0);
    }
    // exclusive prefix of cnt among the lanes of the same image
    const unsigned peers = __match_any_sync(0xffffffffu, g.img);
    int pre = 0, tot = 0;
    for (unsigned m = peers; m; m &= m - 1) {
      const int src = __ffs(m) - 1;
      const int v = __shfl_sync(peers, cnt, src);
      if (src < lane) pre += v;
      tot += v;
    }
    int base = 0;
    const int leader = __ffs(peers) - 1;
    if (lane == leader && tot > 0 && in) base = atomicAdd(&n_points[g.img], tot);
    base = __shfl_sync(peers, base, leader);
    if (in && cnt > 0) word_points<true>(bits, widx, g, h, w, wq, L, S, r_out, points + (int64_t)g.img * cap * 2, base + pre, cap);
  }
}

}  // namespace dbb

using namespace dbb;

// Must follow dbb_binarize_ccl_score on the SAME workspace and stream (it reads the packed bitmap, the final labels and the
// per-root candidate slots left there).  points: (N, cap, 2) int32 {candidate slot, (y << 16) | x}; n_points: (N) int32, the
// number of points the image produced (> cap means the buffer was too small: call again with a larger one).
extern "C" int dbb_ccl_border_points(const void* workspace, size_t workspace_bytes, int64_t n, int64_t h, int64_t w, int32_t* points,
                                     int32_t* n_points, int cap, void* stream) {
  if (!workspace || !points || !n_points || cap <= 0) return set_error(DBB_EINVAL, "ccl_border_points: bad argument");
  if (h >= 32768 || w >= 65536) return set_error(DBB_EUNSUPPORTED, "ccl_border_points: points are packed as (y << 16) | x");
  if (workspace_bytes < dbb_postprocess_workspace(n, h, w)) return set_error(DBB_EWORKSPACE, "ccl_border_points: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t hw = h * w;
  const int wq = (int)((w + 31) / 32);
  CclWs ws = ccl_carve(const_cast<void*>(workspace), n, hw, h, wq);
  const int64_t nseg = n * h * wq;
  int gseg = (int)((nseg + CCL_THREADS - 1) / CCL_THREADS);
  if (gseg > DBB_NUM_SMS * 8) gseg = DBB_NUM_SMS * 8;
  DBB_CUDA(cudaMemsetAsync(n_points, 0, sizeof(int32_t) * (size_t)n, s));
  DBB_LAUNCH("ccl_points", s, ccl_points_kernel<<<gseg, CCL_THREADS, 0, s>>>(ws.bits, (int)n, (int)h, (int)w, wq, ws.label, ws.stat, points, n_points, cap));
  return DBB_OK;
}

extern "C" size_t dbb_postprocess_workspace(int64_t n, int64_t h, int64_t w) {
  const int64_t hw = h * w, wq = (w + 31) / 32;
  return ccl_align(sizeof(unsigned) * (size_t)n * h * wq) + ccl_align(sizeof(int) * (size_t)n * (hw + 1)) +
         ccl_align(sizeof(CompStat) * (size_t)n * hw) + 2 * ccl_align(sizeof(int) * (size_t)n * ccl_nblk(hw)) +
         ccl_align(sizeof(unsigned) * (size_t)n * h * wq) + 256;
}

extern "C" int dbb_binarize_ccl_score(const float* pred, int64_t n, int c, int64_t h, int64_t w, float thresh, double box_thresh,
                                      uint8_t* bitmap, int32_t* labels, DbbCandidate* cands, int32_t* n_cands, int max_cands,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  if (!pred || !bitmap || !cands || !n_cands || !workspace) return set_error(DBB_EINVAL, "binarize_ccl_score: null pointer");
  if (n <= 0 || c <= 0 || h <= 0 || w <= 0 || max_cands <= 0 || n > 65535) return set_error(DBB_EINVAL, "binarize_ccl_score: bad shape");
  if (h * w >= (int64_t)1 << 31) return set_error(DBB_EUNSUPPORTED, "binarize_ccl_score: image too large for 32-bit labels");
  if (n * h * ((w + 31) / 32) >= (int64_t)1 << 32) return set_error(DBB_EUNSUPPORTED, "binarize_ccl_score: batch too large for 32-bit word indices");
  if (workspace_bytes < dbb_postprocess_workspace(n, h, w)) return set_error(DBB_EWORKSPACE, "binarize_ccl_score: workspace too small");
  if (!aligned16(workspace)) return set_error(DBB_EALIGN, "binarize_ccl_score: workspace not 16B aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t hw = h * w;
  const int wq = (int)((w + 31) / 32);
  CclWs ws = ccl_carve(workspace, n, hw, h, wq);
  const int64_t nseg = n * h * wq;
  int gseg = (int)((nseg + CCL_THREADS / 32 - 1) / (CCL_THREADS / 32));
  if (gseg > DBB_NUM_SMS * 16) gseg = DBB_NUM_SMS * 16;
  const int nblk = ccl_nblk(hw);
  int gx = nblk < DBB_NUM_SMS * 8 ? nblk : DBB_NUM_SMS * 8;
  const dim3 grid((unsigned)gx, (unsigned)n), gridb((unsigned)nblk, (unsigned)n);
  DBB_LAUNCH("ccl_pack_init", s, ccl_pack_init_kernel<<<gseg, CCL_THREADS, 0, s>>>(pred, c, (int)n, (int)h, (int)w, wq, thresh, bitmap, ws.bits, ws.label));
  // thread-per-word kernels
  int gword = (int)((nseg + CCL_THREADS - 1) / CCL_THREADS);
  if (gword > DBB_NUM_SMS * 8) gword = DBB_NUM_SMS * 8;
  DBB_LAUNCH("ccl_link", s, ccl_link_kernel<<<gword, CCL_THREADS, 0, s>>>(ws.bits, (int)n, (int)h, (int)w, wq, ws.label, 0));
  if (h > 32) {
    const int64_t nedge_words = n * ((h - 1) / 32) * wq;
    int gedge = (int)((nedge_words + CCL_THREADS - 1) / CCL_THREADS);
    if (gedge > DBB_NUM_SMS * 8) gedge = DBB_NUM_SMS * 8;
    if (gedge < 1) gedge = 1;
    DBB_LAUNCH("ccl_flatten", s, ccl_flatten_kernel<<<gword, CCL_THREADS, 0, s>>>(ws.bits, (int)n, (int)h, (int)w, wq, ws.label, ws.stat, 0, ws.rootbits));
    DBB_LAUNCH("ccl_link", s, ccl_link_kernel<<<gedge, CCL_THREADS, 0, s>>>(ws.bits, (int)n, (int)h, (int)w, wq, ws.label, 1));
  }
  DBB_LAUNCH("ccl_flatten", s, ccl_flatten_kernel<<<gword, CCL_THREADS, 0, s>>>(ws.bits, (int)n, (int)h, (int)w, wq, ws.label, ws.stat, 1, ws.rootbits));
  DBB_LAUNCH("ccl_stats", s, ccl_stats_kernel<<<gseg, CCL_THREADS, 0, s>>>(pred, c, ws.bits, (int)n, (int)h, (int)w, wq, ws.label, ws.stat));
  DBB_LAUNCH("ccl_tree", s, ccl_tree_kernel<<<grid, CCL_THREADS, 0, s>>>(bitmap, (int)h, (int)w, ws.label, ws.stat));
  const int nwblk = (int)((h * wq + CCL_THREADS - 1) / CCL_THREADS);        // blocks of 256 words in raster order
  const dim3 gridw((unsigned)nwblk, (unsigned)n);
  DBB_LAUNCH("ccl_count", s, ccl_count_kernel<<<gridw, CCL_THREADS, 0, s>>>((int)h, (int)w, wq, ws.label, ws.rootbits, ws.blk_count, nwblk));
  DBB_LAUNCH("ccl_scan", s, ccl_scan_kernel<<<(unsigned)n, 1024, 0, s>>>(ws.blk_count, ws.blk_off, nwblk, n_cands));
  DBB_LAUNCH("ccl_emit", s, ccl_emit_kernel<<<gridw, CCL_THREADS, 0, s>>>(bitmap, (int)h, (int)w, wq, ws.label, ws.stat, ws.rootbits, ws.blk_off, nwblk, box_thresh, cands, max_cands));
  if (labels) DBB_LAUNCH("ccl_labels", s, ccl_labels_kernel<<<grid, CCL_THREADS, 0, s>>>(bitmap, hw, ws.label, labels));
  return DBB_OK;
}

// src/postprocess.py:51-52 on its own: bitmap = P[:, 0] > thresh   (5 B per pixel)
__global__ void dbb_binarize_kernel(const float* __restrict__ pred, int c, int64_t hw, float thresh, uint8_t* __restrict__ bitmap) {
  const int img = blockIdx.y;
  const float* P = pred + (int64_t)img * c * hw;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < hw; i += (int64_t)gridDim.x * 256) bitmap[img * hw + i] = P[i] > thresh ? 1 : 0;
}
extern "C" int dbb_binarize(const float* pred, int64_t n, int c, int64_t h, int64_t w, float thresh, uint8_t* bitmap, void* stream) {
  if (!pred || !bitmap || n <= 0 || c <= 0 || h <= 0 || w <= 0 || n > 65535) return set_error(DBB_EINVAL, "binarize: bad argument");
  const int64_t hw = h * w;
  int gx = (int)((hw + 255) / 256); if (gx > DBB_NUM_SMS * 8) gx = DBB_NUM_SMS * 8;
  DBB_LAUNCH("binarize", (cudaStream_t)stream, dbb_binarize_kernel<<<dim3((unsigned)gx, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(pred, c, hw, thresh, bitmap));
  return DBB_OK;
}
