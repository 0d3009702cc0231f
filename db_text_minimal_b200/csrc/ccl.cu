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
// raster-first pixel = cv2's discovery point):
//   1 init      bitmap, label[i] = i                                     read P 4 B, write 1 + 4 B per pixel
//   2 merge     union-find over backward neighbours (fg: W,NW,N,NE; bg: W,N; border bg pixels join a virtual outside node)
//   3 compress  label[i] = root(i); roots zero their statistics slot
//   4 stats     per-component count / float64 sum / bbox with warp-aggregated atomics (match.any), plus the boundary
//               ring of every hole; the outside region is skipped (it is no candidate and would serialise the atomics)
//   5 tree      every node adds its own statistics to all its ancestors (parent = region north of the root pixel)
//   6 rank/emit suffix count of roots in raster order = position in cv2's reverse-discovery order -> candidates are
//               written already sorted, the first max_cands of them (src/postprocess.py:70,119)
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
  int* label;              // [n][hw + 1]   (+1: virtual outside node)
  CompStat* stat;          // [n][hw]
  int* blk_count;          // [n][nblk]
  int* blk_off;            // [n][nblk]
};

__host__ __device__ inline size_t ccl_align(size_t v) { return (v + 255) / 256 * 256; }

static int ccl_nblk(int64_t hw) { return (int)((hw + CCL_THREADS - 1) / CCL_THREADS); }

static CclWs ccl_carve(void* ws, int64_t n, int64_t hw) {
  CclWs w;
  char* p = (char*)ws;
  w.label = (int*)p;        p += ccl_align(sizeof(int) * (size_t)n * (hw + 1));
  w.stat = (CompStat*)p;    p += ccl_align(sizeof(CompStat) * (size_t)n * hw);
  w.blk_count = (int*)p;    p += ccl_align(sizeof(int) * (size_t)n * ccl_nblk(hw));
  w.blk_off = (int*)p;
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

__global__ void __launch_bounds__(CCL_THREADS)
ccl_init_kernel(const float* __restrict__ pred, int c, int h, int w, float thresh, uint8_t* __restrict__ bitmap, int* __restrict__ label) {
  const int img = blockIdx.y;
  const int64_t hw = (int64_t)h * w;
  const float* P = pred + (int64_t)img * c * hw;
  uint8_t* bm = bitmap + img * hw;
  int* L = label + img * (hw + 1);
  for (int64_t i = (int64_t)blockIdx.x * CCL_THREADS + threadIdx.x; i <= hw; i += (int64_t)gridDim.x * CCL_THREADS) {
    L[i] = (int)i;
    if (i < hw) bm[i] = P[i] > thresh ? 1 : 0;
  }
}

__global__ void __launch_bounds__(CCL_THREADS)
ccl_merge_kernel(const uint8_t* __restrict__ bitmap, int h, int w, int* __restrict__ label) {
  const int img = blockIdx.y;
  const int64_t hw = (int64_t)h * w;
  const uint8_t* bm = bitmap + img * hw;
  int* L = label + img * (hw + 1);
  for (int64_t i = (int64_t)blockIdx.x * CCL_THREADS + threadIdx.x; i < hw; i += (int64_t)gridDim.x * CCL_THREADS) {
    const int y = (int)(i / w), x = (int)(i - (int64_t)y * w);
    const uint8_t v = bm[i];
    if (v) {   // foreground: 8-connectivity
      if (x > 0 && bm[i - 1]) uf_union(L, (int)i, (int)i - 1);
      if (y > 0) {
        if (bm[i - w]) uf_union(L, (int)i, (int)(i - w));
        if (x > 0 && bm[i - w - 1]) uf_union(L, (int)i, (int)(i - w - 1));
        if (x < w - 1 && bm[i - w + 1]) uf_union(L, (int)i, (int)(i - w + 1));
      }
    } else {   // background: 4-connectivity; the image frame belongs to the outside region
      if (x > 0 && !bm[i - 1]) uf_union(L, (int)i, (int)i - 1);
      if (y > 0 && !bm[i - w]) uf_union(L, (int)i, (int)(i - w));
      if (x == 0 || y == 0 || x == w - 1 || y == h - 1) uf_union(L, (int)hw, (int)i);
    }
  }
}

__global__ void __launch_bounds__(CCL_THREADS)
ccl_compress_kernel(int h, int w, int* __restrict__ label, CompStat* __restrict__ stat) {
  const int img = blockIdx.y;
  const int64_t hw = (int64_t)h * w;
  int* L = label + img * (hw + 1);
  CompStat* S = stat + img * hw;
  for (int64_t i = (int64_t)blockIdx.x * CCL_THREADS + threadIdx.x; i <= hw; i += (int64_t)gridDim.x * CCL_THREADS) {
    const int r = uf_find(L, (int)i);
    L[i] = r;       // benign race: concurrent finds still terminate at the same root
    if (r == (int)i && i < hw) {
      CompStat z;
      z.sum = 0.0; z.acc_sum = 0.0; z.count = 0; z.acc_count = 0;
      z.x0 = w; z.y0 = h; z.x1 = -1; z.y1 = -1;
      S[i] = z;
    }
  }
}

// parent region of a root pixel r: the region containing the pixel north of it (outside for the first row)
__device__ __forceinline__ int parent_of(const int* L, int r, int w, int r_out) { return r < w ? r_out : L[r - w]; }

__global__ void __launch_bounds__(CCL_THREADS)
ccl_stats_kernel(const float* __restrict__ pred, int c, const uint8_t* __restrict__ bitmap, int h, int w, const int* __restrict__ label,
                 CompStat* __restrict__ stat) {
  const int img = blockIdx.y;
  const int64_t hw = (int64_t)h * w;
  const float* P = pred + (int64_t)img * c * hw;
  const uint8_t* bm = bitmap + img * hw;
  const int* L = label + img * (hw + 1);
  CompStat* S = stat + img * hw;
  const int r_out = L[hw];
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * CCL_THREADS;
  const int64_t iters = (hw + stride - 1) / stride;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t i = (int64_t)blockIdx.x * CCL_THREADS + threadIdx.x + it * stride;
    const bool valid = i < hw;
    int r = -1, x = 0, y = 0;
    float p = 0.f;
    if (valid) {
      r = L[i];
      if (r == r_out) r = -1;      // the outside region is not a candidate
      else { y = (int)(i / w); x = (int)(i - (int64_t)y * w); p = P[i]; }
    }
    // warp aggregation: lanes with the same root elect a leader that issues one set of atomics
    const unsigned act = __ballot_sync(0xffffffffu, r >= 0);
    if (r >= 0) {
      const unsigned grp = __match_any_sync(act, r);
      const int leader = __ffs(grp) - 1;
      double gs; int gc, gx0, gx1, gy0, gy1;
      if (grp == 0xffffffffu) {          // the common case inside a region: the whole warp is one run -> butterfly
        gs = (double)p; gc = 32; gx0 = x; gx1 = x; gy0 = y; gy1 = y;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          gs += __shfl_xor_sync(0xffffffffu, gs, o);
          gx0 = min(gx0, __shfl_xor_sync(0xffffffffu, gx0, o)); gx1 = max(gx1, __shfl_xor_sync(0xffffffffu, gx1, o));
          gy0 = min(gy0, __shfl_xor_sync(0xffffffffu, gy0, o)); gy1 = max(gy1, __shfl_xor_sync(0xffffffffu, gy1, o));
        }
      } else {                           // arbitrary lane subset: every member walks the member list
        gs = 0.0; gc = 0; gx0 = w; gx1 = -1; gy0 = h; gy1 = -1;
        for (unsigned m = grp; m; m &= m - 1) {
          const int src = __ffs(m) - 1;
          const double so = __shfl_sync(grp, (double)p, src);
          const int ax = __shfl_sync(grp, x, src), ay = __shfl_sync(grp, y, src);
          gs += so; gc += 1;
          gx0 = min(gx0, ax); gx1 = max(gx1, ax); gy0 = min(gy0, ay); gy1 = max(gy1, ay);
        }
      }
      if (lane == leader) {
        CompStat* t = S + r;
        atomicAdd(&t->sum, gs);
        atomicAdd(&t->count, gc);
        atomicMin(&t->x0, gx0); atomicMax(&t->x1, gx1); atomicMin(&t->y0, gy0); atomicMax(&t->y1, gy1);
      }
      // boundary ring of holes: a foreground pixel contributes once to every DISTINCT enclosed region it 4-touches
      if (bm[i]) {
        const int pr = parent_of(L, r, w, r_out);
        int g[4] = {-1, -1, -1, -1};
        if (x > 0 && !bm[i - 1]) g[0] = L[i - 1];
        if (x < w - 1 && !bm[i + 1]) g[1] = L[i + 1];
        if (y > 0 && !bm[i - w]) g[2] = L[i - w];
        if (y < h - 1 && !bm[i + w]) g[3] = L[i + w];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int gk = g[k];
          if (gk < 0 || gk == pr || gk == r_out) continue;
          bool dup = false;
#pragma unroll
          for (int q = 0; q < 4; ++q) if (q < k && g[q] == gk) dup = true;
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

// ---- ranking: candidates in cv2 order = roots by DESCENDING pixel index
__global__ void __launch_bounds__(CCL_THREADS)
ccl_count_kernel(int h, int w, const int* __restrict__ label, int* __restrict__ blk_count, int nblk) {
  const int img = blockIdx.y, b = blockIdx.x;
  const int64_t hw = (int64_t)h * w;
  const int* L = label + img * (hw + 1);
  const int r_out = L[hw];
  const int64_t i = (int64_t)b * CCL_THREADS + threadIdx.x;
  const int flag = (i < hw && L[i] == (int)i && (int)i != r_out) ? 1 : 0;
  const int c = __syncthreads_count(flag);
  if (threadIdx.x == 0) blk_count[img * nblk + b] = c;
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
ccl_emit_kernel(const uint8_t* __restrict__ bitmap, int h, int w, const int* __restrict__ label, const CompStat* __restrict__ stat,
                const int* __restrict__ blk_off, int nblk, double box_thresh, DbbCandidate* __restrict__ cands, int max_cands,
                int32_t* __restrict__ labels_out) {
  const int img = blockIdx.y, b = blockIdx.x;
  const int64_t hw = (int64_t)h * w;
  const uint8_t* bm = bitmap + img * hw;
  const int* L = label + img * (hw + 1);
  const CompStat* S = stat + img * hw;
  const int r_out = L[hw];
  const int64_t i = (int64_t)b * CCL_THREADS + threadIdx.x;
  const bool in = i < hw;
  const int r = in ? L[i] : -1;
  if (in && labels_out) labels_out[img * hw + i] = bm[i] ? (r + 1) : -(r + 1);
  const int flag = (in && r == (int)i && (int)i != r_out) ? 1 : 0;
  // rank inside the block among HIGHER thread indices (suffix), via warp ballots
  __shared__ int wcount[CCL_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) wcount[wid] = __popc(bal);
  __syncthreads();
  if (!flag) return;
  int rank = __popc(bal & ~((2u << lane) - 1u));      // lanes above me in my warp
  for (int k = wid + 1; k < CCL_THREADS / 32; ++k) rank += wcount[k];
  rank += blk_off[img * nblk + b];
  if (rank >= max_cands) return;
  const CompStat s = S[i];
  DbbCandidate cd;
  const int y = (int)(i / w), x = (int)(i - (int64_t)y * w);
  cd.kind = bm[i] ? 0 : 1;
  cd.first_y = y; cd.first_x = x;
  if (cd.kind == 0) { cd.x0 = s.x0; cd.y0 = s.y0; cd.x1 = s.x1; cd.y1 = s.y1; }
  else { cd.x0 = s.x0 - 1; cd.y0 = s.y0 - 1; cd.x1 = s.x1 + 1; cd.y1 = s.y1 + 1; }   // + the ring of parent pixels
  cd.count = s.count + s.acc_count;
  cd.sum = s.sum + s.acc_sum;
  const double score = cd.sum / (double)cd.count;
  cd.keep = (box_thresh > score) ? 0 : 1;
  cd.pad_ = 0;
  cands[(int64_t)img * max_cands + rank] = cd;
}

}  // namespace dbb

using namespace dbb;

extern "C" size_t dbb_postprocess_workspace(int64_t n, int64_t h, int64_t w) {
  const int64_t hw = h * w;
  return ccl_align(sizeof(int) * (size_t)n * (hw + 1)) + ccl_align(sizeof(CompStat) * (size_t)n * hw) +
         2 * ccl_align(sizeof(int) * (size_t)n * ccl_nblk(hw)) + 256;
}

extern "C" int dbb_binarize_ccl_score(const float* pred, int64_t n, int c, int64_t h, int64_t w, float thresh, double box_thresh,
                                      uint8_t* bitmap, int32_t* labels, DbbCandidate* cands, int32_t* n_cands, int max_cands,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  if (!pred || !bitmap || !cands || !n_cands || !workspace) return set_error(DBB_EINVAL, "binarize_ccl_score: null pointer");
  if (n <= 0 || c <= 0 || h <= 0 || w <= 0 || max_cands <= 0 || n > 65535) return set_error(DBB_EINVAL, "binarize_ccl_score: bad shape");
  if (h * w >= (int64_t)1 << 31) return set_error(DBB_EUNSUPPORTED, "binarize_ccl_score: image too large for 32-bit labels");
  if (workspace_bytes < dbb_postprocess_workspace(n, h, w)) return set_error(DBB_EWORKSPACE, "binarize_ccl_score: workspace too small");
  if (!aligned16(workspace)) return set_error(DBB_EALIGN, "binarize_ccl_score: workspace not 16B aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t hw = h * w;
  CclWs ws = ccl_carve(workspace, n, hw);
  const int nblk = ccl_nblk(hw);
  int gx = nblk < DBB_NUM_SMS * 8 ? nblk : DBB_NUM_SMS * 8;
  const dim3 grid((unsigned)gx, (unsigned)n), gridb((unsigned)nblk, (unsigned)n);
  DBB_LAUNCH("ccl_init", s, ccl_init_kernel<<<grid, CCL_THREADS, 0, s>>>(pred, c, (int)h, (int)w, thresh, bitmap, ws.label));
  DBB_LAUNCH("ccl_merge", s, ccl_merge_kernel<<<grid, CCL_THREADS, 0, s>>>(bitmap, (int)h, (int)w, ws.label));
  DBB_LAUNCH("ccl_compress", s, ccl_compress_kernel<<<grid, CCL_THREADS, 0, s>>>((int)h, (int)w, ws.label, ws.stat));
  DBB_LAUNCH("ccl_stats", s, ccl_stats_kernel<<<grid, CCL_THREADS, 0, s>>>(pred, c, bitmap, (int)h, (int)w, ws.label, ws.stat));
  DBB_LAUNCH("ccl_tree", s, ccl_tree_kernel<<<grid, CCL_THREADS, 0, s>>>(bitmap, (int)h, (int)w, ws.label, ws.stat));
  DBB_LAUNCH("ccl_count", s, ccl_count_kernel<<<gridb, CCL_THREADS, 0, s>>>((int)h, (int)w, ws.label, ws.blk_count, nblk));
  DBB_LAUNCH("ccl_scan", s, ccl_scan_kernel<<<(unsigned)n, 1024, 0, s>>>(ws.blk_count, ws.blk_off, nblk, n_cands));
  DBB_LAUNCH("ccl_emit", s, ccl_emit_kernel<<<gridb, CCL_THREADS, 0, s>>>(bitmap, (int)h, (int)w, ws.label, ws.stat, ws.blk_off, nblk, box_thresh, cands, max_cands, labels));
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
