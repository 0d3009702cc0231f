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
// raster-first pixel = cv2's discovery point).  Everything works on a bit-packed bitmap, 32 pixels per word:
//   strip    16-row strips in SHARED memory: binarize (strict >), pack, byte bitmap, row-run labels, unions between touching
//            runs (fg: vertical + the two diagonals, bg: vertical; frame runs -> virtual outside node), flatten; every pixel
//            leaves the SM once, carrying its strip root                                     (4 B read, ~9 B written / px)
//            (rows too wide for shared memory: pack_init + link + flatten as three kernels on global memory)
//   seams    unions across the strip boundaries + each strip's outside root with the image's outside node
//   flatten  strip roots -> root; roots zero their statistics slot, store their parent region, publish the root bitmap
//   stats    one thread per word: per-run float64 sums -> shared-memory table per block -> one atomic set per region and
//            block; hole boundary rings; words of the outside region are skipped before their probabilities are read
//   tree     every region adds its statistics to all its ancestors (parent = region north of the root pixel)
//   rank/emit  suffix count of roots in raster order = position in cv2's reverse-discovery order -> candidates are
//            written already sorted, the first max_cands of them (src/postprocess.py:70,119)
#include "common.cuh"

namespace dbb {

constexpr int CCL_THREADS = 256;

struct CompStat {          // one slot per pixel index, touched only at roots
  double sum;              // own pixels
  double acc_sum;          // descendants (+ boundary ring for holes)
  int count, acc_count;
  int x0, parent, x1, y1;  // bounding box without y0 (the root is the raster-first pixel: its row IS y0); parent = root of
                           // the region north of the root pixel (written by the final flatten)
};

struct CclWs {
  unsigned* bits;          // [n][h][wq]    packed bitmap, 32 pixels per word
  int* label;              // [n][lab_stride(hw)]   (hw pixels + the virtual outside node at index hw, padded to 16 bytes)
  CompStat* stat;          // [n][hw]
  int* blk_count;          // [n][nblk]
  int* blk_off;            // [n][nblk]
  unsigned* rootbits;      // [n][h][wq]    1 where the pixel is the root of its component (written by the final flatten)
};

__host__ __device__ inline size_t ccl_align(size_t v) { return (v + 255) / 256 * 256; }
__host__ __device__ inline int64_t lab_stride(int64_t hw) { return (hw + 4) & ~(int64_t)3; }

static int ccl_nblk(int64_t hw) { return (int)((hw + CCL_THREADS - 1) / CCL_THREADS); }

static CclWs ccl_carve(void* ws, int64_t n, int64_t hw, int64_t h, int64_t wq) {
  CclWs w;
  char* p = (char*)ws;
  w.bits = (unsigned*)p;    p += ccl_align(sizeof(unsigned) * (size_t)n * h * wq);
  w.label = (int*)p;        p += ccl_align(sizeof(int) * (size_t)n * lab_stride(hw));
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

// A: binarize (strict >), pack, write the byte bitmap, label every pixel with the first pixel of its ROW run.
// One warp per image row, 32 words (1,024 pixels) per pass: the 32 coalesced 128-byte loads of a pass are all issued before
// the first ballot (the first version had one load in flight per warp: 2.2 TB/s); then lane k owns word k and a max-scan over
// the lanes carries run starts across word boundaries, so horizontal neighbours are linked by construction and ccl_link has
// no seam unions left to do.  (A label that points into an earlier word is an ordinary union-find parent pointer: it is
// smaller than the pixel index and is itself the start of a segment run.)
// One image row by one warp.  Lrow / Wrow2 may point to shared memory (the strip kernel keeps strip-local labels there);
// label_base is the index of the row's first pixel in the index space of Lrow's array.
__device__ __forceinline__ void pack_row(const float* __restrict__ P, int w, int wq, float thresh, int* Lrow, int label_base,
                                         uint8_t* __restrict__ Brow, unsigned* __restrict__ Wrow, unsigned* Wrow2, int word_stores, int lane) {
  int carry_start = -1;            // start (x) of the row run that reaches the end of the previous pass
  unsigned carry_bit = 0;          // class of the last pixel of the previous pass
  for (int s0 = 0; s0 < wq; s0 += 32) {
    const int kk = wq - s0 < 32 ? wq - s0 : 32;
    const bool full = (s0 + 32) * 32 <= w;         // all 32 words of the pass are whole: constant offsets, no bounds tests
    float v[32];
    if (full) {
      const float* Pl = P + s0 * 32 + lane;
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = __ldg(Pl + k * 32);
    } else {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int x = (s0 + k) * 32 + lane;
        v[k] = (k < kk && x < w) ? __ldg(P + x) : 0.f;
      }
    }
    unsigned cw[32];
    unsigned mine = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      cw[k] = __ballot_sync(0xffffffffu, v[k] > thresh);               // lanes beyond the row hold 0 -> bit 0
      if (lane == k) mine = cw[k];
    }
    // ---- lane k owns word s0 + k
    const int x0 = (s0 + lane) * 32;
    const int nvalid = lane < kk ? (w - x0 < 32 ? w - x0 : 32) : 0;
    const unsigned vm = valid_mask(nvalid);
    if (lane < kk) { Wrow[s0 + lane] = mine; if (Wrow2) Wrow2[s0 + lane] = mine; }
    unsigned prev_word = __shfl_up_sync(0xffffffffu, mine, 1);
    const unsigned prev_bit = lane == 0 ? carry_bit : (prev_word >> 31);
    const bool cont = (s0 + lane > 0) && nvalid > 0 && ((mine & 1u) == prev_bit);
    const unsigned stv = run_starts(mine, nvalid) & vm;
    const bool single = (stv & (stv - 1)) == 0;
    int val = (nvalid == 0 || (single && cont)) ? -1 : x0 + (31 - __clz(stv));
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, val, o);
      if (lane >= o && t > val) val = t;
    }
    if (val < carry_start) val = carry_start;                          // nothing defined up to here: the run comes from the previous pass
    int before = __shfl_up_sync(0xffffffffu, val, 1);
    if (lane == 0) before = carry_start;
    const int first_start = cont ? before : x0;                        // start of the row run the word's FIRST segment run belongs to
    // ---- labels (lane = pixel again) and the byte bitmap
    if (full) {
      int* Ll = Lrow + s0 * 32 + lane;
      const unsigned upto = (2u << lane) - 1u;
      const int xl = label_base + s0 * 32;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int fs = __shfl_sync(0xffffffffu, first_start, k);
        const unsigned st = (((cw[k] ^ (cw[k] << 1)) | 1u)) & upto;
        const int a = 31 - __clz(st);
        Ll[k * 32] = a == 0 ? label_base + fs : xl + k * 32 + a;
        if (!word_stores) Brow[(s0 + k) * 32 + lane] = (cw[k] >> lane) & 1u;
      }
    } else
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int fs = __shfl_sync(0xffffffffu, first_start, k);
      if (k < kk) {
        const int xk = (s0 + k) * 32;
        const int nv = w - xk < 32 ? w - xk : 32;
        if (lane < nv) {
          const unsigned st = run_starts(cw[k], nv) & ((2u << lane) - 1u);
          const int a = 31 - __clz(st);
          Lrow[xk + lane] = label_base + (a == 0 ? fs : xk + a);
          if (!word_stores) Brow[xk + lane] = (cw[k] >> lane) & 1u;
        }
      }
    }
    if (word_stores) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {                                    // 4 words = 128 pixels = one 128-byte store per warp
        const unsigned wv = __shfl_sync(0xffffffffu, mine, 4 * q + (lane >> 3));
        const int x = (s0 + 4 * q) * 32 + 4 * lane;
        if (4 * q < kk && x < w) {
          const unsigned nib = (wv >> ((lane & 7) * 4)) & 0xfu;
          *reinterpret_cast<unsigned*>(Brow + x) = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
        }
      }
    }
    carry_start = __shfl_sync(0xffffffffu, val, kk - 1);
    carry_bit = __shfl_sync(0xffffffffu, mine, kk - 1) >> 31;
  }
}
__global__ void __launch_bounds__(CCL_THREADS)
ccl_pack_init_kernel(const float* __restrict__ pred, int c, int n, int h, int w, int wq, float thresh, uint8_t* __restrict__ bitmap,
                     unsigned* __restrict__ bits, int* __restrict__ label, int word_stores) {
  const int64_t hw = (int64_t)h * w;
  const int lane = threadIdx.x & 31;
  const int nrows = n * h;
  const int wstride = gridDim.x * (CCL_THREADS / 32);
  for (int row = blockIdx.x * (CCL_THREADS / 32) + (threadIdx.x >> 5); row < nrows; row += wstride) {
    const int img = row / h, y = row - img * h;
    if (y == 0 && lane == 0) label[img * lab_stride(hw) + hw] = (int)hw;       // virtual outside node
    pack_row(pred + (int64_t)img * c * hw + (int64_t)y * w, w, wq, thresh, label + img * lab_stride(hw) + (int64_t)y * w, y * w,
             bitmap + img * hw + (int64_t)y * w, bits + (int64_t)row * wq, nullptr, word_stores, lane);
  }
}

// B: unions between touching runs: across segment boundaries, with the row above (4-connectivity for background,
// 8-connectivity for foreground), and background runs on the image frame with the virtual outside node.
// ONE THREAD PER 32-PIXEL WORD: the neighbourhood masks are a dozen word operations computed once, and the thread then walks
// the set bits (typically 0-2 unions per word).  The first version gave every pixel a lane that recomputed the same masks and
// diverged on its own union: 30 warp-instructions per pixel at 5.7 active lanes (ncu, profiles/prof_ccl_r02.md) -- issue-bound.
// B points at the word itself: B[-1] / B[1] are its row neighbours, B[-wq] the word above (global bits and the strip kernel's
// shared copy have the same layout).  i0 = index of the word's first pixel in L's index space (row pitch w), out_node = index
// of the virtual outside node there.  vertical: link with the row above; frame: horizontal-frame / outside unions.
__device__ __forceinline__ void link_core(const unsigned* B, int s, int wq, int w, int nvalid, bool vertical, bool frame, bool frame_row,
                                          int* L, int i0, int out_node) {
  const unsigned cur = B[0];
  const unsigned prv = s > 0 ? B[-1] : 0u, nxt = s < wq - 1 ? B[1] : 0u;
  const unsigned vm = valid_mask(nvalid);
  if (vertical) {
    const unsigned up = B[-wq];
    const unsigned upp = s > 0 ? B[-wq - 1] : 0u, upn = s < wq - 1 ? B[-wq + 1] : 0u;
    // neighbours shifted into lane position: L = pixel x-1, R = pixel x+1
    const unsigned curL = (cur << 1) | (prv >> 31), curR = (cur >> 1) | (nxt << 31);
    const unsigned upL = (up << 1) | (upp >> 31), upR = (up >> 1) | (upn << 31);
    const unsigned hasL = s > 0 ? 0xffffffffu : 0xfffffffeu;                         // pixel x-1 exists
    unsigned V = cur & up & ~(curL & upL & hasL);            // first column of a vertical overlap
    unsigned DL = cur & ~up & upL & ~curL & hasL;            // diagonal up-left, not implied by a neighbour
    unsigned DR = cur & ~up & upR & ~curR;                   // diagonal up-right (bits beyond the row end are 0)
    if (s == wq - 1 && nvalid >= 1) DR &= ~(1u << (nvalid - 1));     // x + 1 < w
    while (V) { const int b = __ffs(V) - 1; V &= V - 1; uf_union(L, i0 + b, i0 + b - w); }
    while (DL) { const int b = __ffs(DL) - 1; DL &= DL - 1; uf_union(L, i0 + b, i0 + b - w - 1); }
    while (DR) { const int b = __ffs(DR) - 1; DR &= DR - 1; uf_union(L, i0 + b, i0 + b - w + 1); }
    const unsigned ncur = ~cur & vm;
    unsigned VB = ncur & ~up & ~(~curL & ~upL & hasL);       // background: vertical links only (4-connectivity)
    while (VB) { const int b = __ffs(VB) - 1; VB &= VB - 1; uf_union(L, i0 + b, i0 + b - w); }
  }
  if (!frame) return;
  // (no horizontal unions: pack_row labels every pixel with the start of its ROW run)
  // frame pixels of the background belong to the outside region: one union per background run start on the first / last
  // row, the first / last pixel of every other row
  const unsigned ncur = ~cur & vm;
  if (frame_row) {
    unsigned st = run_starts(cur, nvalid) & ncur;
    while (st) { const int b = __ffs(st) - 1; st &= st - 1; uf_union(L, out_node, i0 + b); }
  } else {
    if (s == 0 && (ncur & 1u)) uf_union(L, out_node, i0);
  }
  if (s == wq - 1 && ((ncur >> (nvalid - 1)) & 1u)) uf_union(L, out_node, i0 + nvalid - 1);
}
// two-level merge keeps union-find chains short: phase 0 links everything except across the boundaries of strips of `sh`
// rows (chains <= sh), the strips are flattened, phase 1 links the strip boundaries (chains <= H / sh)
__device__ __forceinline__ void link_word(const unsigned* __restrict__ bits, int64_t widx, int h, int w, int wq, int* __restrict__ label, int phase, int sh) {
  const int64_t hw = (int64_t)h * w;
  Seg g; seg_of(widx, h, wq, w, g);
  const bool strip_edge = (g.y % sh) == 0;
  link_core(bits + widx, g.s, wq, w, g.nvalid, g.y > 0 && (phase == 1 || !strip_edge), phase == 0, g.y == 0 || g.y == h - 1,
            label + g.img * lab_stride(hw), g.y * w + g.x0, (int)hw);
}
// phase 0: every word.  phase 1: only the first row of every strip (except row 0); strip_out (strip kernel path): the root each
// strip found for its frame background, to be joined with the image's outside node.
__global__ void __launch_bounds__(CCL_THREADS)
ccl_link_kernel(const unsigned* __restrict__ bits, int n, int h, int w, int wq, int* __restrict__ label, int phase, int sh,
                const int* __restrict__ strip_out) {
  const int64_t stride = (int64_t)gridDim.x * CCL_THREADS;
  if (phase == 0) {
    const int64_t nseg = (int64_t)n * h * wq;
    for (int64_t widx = (int64_t)blockIdx.x * CCL_THREADS + threadIdx.x; widx < nseg; widx += stride) link_word(bits, widx, h, w, wq, label, 0, sh);
  } else {
    const int nedge = (h - 1) / sh;                          // rows sh, 2 sh, ...
    const int64_t total = (int64_t)n * nedge * wq;
    for (int64_t t = (int64_t)blockIdx.x * CCL_THREADS + threadIdx.x; t < total; t += stride) {
      const int s = (int)(t % wq); const int64_t q = t / wq;
      const int e = (int)(q % nedge), img = (int)(q / nedge);
      link_word(bits, ((int64_t)img * h + (int64_t)(e + 1) * sh) * wq + s, h, w, wq, label, 1, sh);
    }
    if (strip_out) {
      const int nstrips = (h + sh - 1) / sh;
      const int64_t hw = (int64_t)h * w;
      for (int64_t t = (int64_t)blockIdx.x * CCL_THREADS + threadIdx.x; t < (int64_t)n * nstrips; t += stride) {
        const int r = strip_out[t];
        if (r >= 0) uf_union(label + (t / nstrips) * lab_stride(hw), (int)hw, r);
      }
    }
  }
}

// A+B+C for one strip of STRIP_ROWS rows in ONE kernel: the strip's labels live in shared memory while its runs are linked
// (a union-find hop costs ~30 cycles there instead of an L2 / DRAM round trip), and leave the SM already flattened:
// every pixel is written once, carrying the strip-local root.  Replaces pack_init + link (phase 0) + flatten (strips):
// 0.55 ms -> see profiles/prof_ccl_r02.md for 64 x 1024^2.
constexpr int STRIP_ROWS = 16;
constexpr int STRIP_THREADS = 512;
__global__ void __launch_bounds__(STRIP_THREADS)
ccl_strip_kernel(const float* __restrict__ pred, int c, int n, int h, int w, int wq, float thresh, uint8_t* __restrict__ bitmap,
                 unsigned* __restrict__ bits, int* __restrict__ label, int* __restrict__ strip_out, unsigned* __restrict__ rootbits, int word_stores) {
  extern __shared__ int strip_smem[];
  int* sl = strip_smem;                                             // [STRIP_ROWS * w + 1] strip-local labels
  unsigned* sb = (unsigned*)(strip_smem + STRIP_ROWS * w + 1);      // [STRIP_ROWS * wq]    the strip's words
  const int64_t hw = (int64_t)h * w;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int nstrips = (h + STRIP_ROWS - 1) / STRIP_ROWS;
  for (int sid = blockIdx.x; sid < n * nstrips; sid += gridDim.x) {
    const int img = sid / nstrips, y0 = (sid - img * nstrips) * STRIP_ROWS;
    const int rows = h - y0 < STRIP_ROWS ? h - y0 : STRIP_ROWS;
    const int out_node = rows * w;
    int* Lg = label + img * lab_stride(hw);
    // ---- A: binarize + pack + row-run labels
    for (int r = wrp; r < rows; r += STRIP_THREADS / 32)
      pack_row(pred + (int64_t)img * c * hw + (int64_t)(y0 + r) * w, w, wq, thresh, sl + r * w, r * w,
               bitmap + img * hw + (int64_t)(y0 + r) * w, bits + ((int64_t)img * h + y0 + r) * wq, sb + r * wq, word_stores, lane);
    if (threadIdx.x == 0) { sl[out_node] = out_node; if (y0 == 0) Lg[hw] = (int)hw; }
    __syncthreads();
    // ---- B: unions inside the strip
    for (int t = threadIdx.x; t < rows * wq; t += STRIP_THREADS) {
      const int r = t / wq, sw = t - r * wq;
      const int x0 = sw * 32;
      const int y = y0 + r;
      link_core(sb + t, sw, wq, w, w - x0 < 32 ? w - x0 : 32, r > 0, true, y == 0 || y == h - 1, sl, r * w + x0, out_node);
    }
    __syncthreads();
    // ---- C: segment run starts -> strip root, then every pixel -> root of its run start, written out once
    for (int t = threadIdx.x; t < rows * wq; t += STRIP_THREADS) {
      const int r = t / wq, sw = t - r * wq;
      const int x0 = sw * 32;
      const int nvalid = w - x0 < 32 ? w - x0 : 32;
      unsigned st = run_starts(sb[t], nvalid) & valid_mask(nvalid);
      unsigned roots = 0;
      while (st) {
        const int b = __ffs(st) - 1; st &= st - 1;
        const int i = r * w + x0 + b;
        const int rt = uf_find(sl, i);
        sl[i] = rt;
        if (rt == i) roots |= 1u << b;
      }
      rootbits[((int64_t)img * h + y0 + r) * wq + sw] = roots;      // strip roots: the only pixels the final flatten has to visit
    }
    __syncthreads();
    const int base = y0 * w;
    if ((w & 3) == 0) {                                             // label rows are 16-byte aligned (lab_stride)
      for (int i = threadIdx.x * 4; i < rows * w; i += STRIP_THREADS * 4) {
        const int4 l = *reinterpret_cast<const int4*>(sl + i);
        int4 o;
        o.x = base + sl[l.x]; o.y = base + sl[l.y]; o.z = base + sl[l.z]; o.w = base + sl[l.w];
        *reinterpret_cast<int4*>(Lg + base + i) = o;
      }
    } else {
      for (int i = threadIdx.x; i < rows * w; i += STRIP_THREADS) Lg[base + i] = base + sl[sl[i]];
    }
    if (threadIdx.x == 0) {
      const int ro = uf_find(sl, out_node);
      strip_out[sid] = ro != out_node ? base + ro : -1;
    }
    __syncthreads();                                                // the next strip reuses the arrays
  }
}

// C: run-start pixels jump straight to their root; roots zero their statistics slot.  One thread per word (see B).
__global__ void __launch_bounds__(CCL_THREADS)
ccl_flatten_kernel(const unsigned* __restrict__ bits, int n, int h, int w, int wq, int* __restrict__ label, CompStat* __restrict__ stat,
                   int init_stats, unsigned* __restrict__ rootbits, int from_roots) {
  const int64_t hw = (int64_t)h * w;
  const int64_t nseg = (int64_t)n * h * wq;
  for (int64_t widx = (int64_t)blockIdx.x * CCL_THREADS + threadIdx.x; widx < nseg; widx += (int64_t)gridDim.x * CCL_THREADS) {
    Seg g; seg_of(widx, h, wq, w, g);
    int* L = label + g.img * lab_stride(hw);
    if (g.y == 0 && g.s == 0) L[hw] = uf_find(L, (int)hw);
    // strip-kernel path (from_roots): every pixel already carries its strip root, so only the strip roots move
    unsigned st = from_roots ? rootbits[widx] : (run_starts(bits[widx], g.nvalid) & valid_mask(g.nvalid));
    unsigned roots = 0;
    while (st) {
      const int b = __ffs(st) - 1; st &= st - 1;
      const int i = g.y * w + g.x0 + b;
      const int p0 = L[i];
      const int r = uf_find(L, p0);
      if (p0 != r) L[i] = r;                        // (after the strip kernel most run starts already hold their final root)
      if (r == i && init_stats) {
        CompStat z;
        z.sum = 0.0; z.acc_sum = 0.0; z.count = 0; z.acc_count = 0;
        z.x0 = w; z.x1 = -1; z.y1 = -1;
        z.parent = uf_find(L, i < w ? (int)hw : i - w);   // all unions are done: uf_find is final whatever the flatten has reached
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

// D: per-component statistics.  ONE THREAD PER 32-PIXEL WORD (the warp-per-word version spent 313 warp-instructions per word
// on a float64 warp scan and per-lane ring tests: issue-bound at 0.85 ms for 64 x 1024^2, profiles/prof_ccl_r02.md):
//   * a word whose runs all belong to the outside region (most of a text map) is dropped after one label lookup -- its
//     probabilities are never read;
//   * the probabilities of the other words of a warp are staged through shared memory (coalesced 128-byte rows in, one
//     row of 32 floats per thread out, pitch 33 -> conflict-free), summed per run in float64, one atomic set per run;
//   * hole rings: foreground pixels that 4-touch an enclosed background region add themselves to that hole's boundary ring.
//     Neighbouring background RUNS are classified first (one label lookup per run); only pixels that touch a run which is
//     not the outside region take the per-pixel path (distinct regions, parent test).
__device__ __forceinline__ unsigned run_mask(int a, int b) {      // bits a..b inclusive
  return (b >= 31 ? 0xffffffffu : ((2u << b) - 1u)) & ~((1u << a) - 1u);
}
// bits of `word` (background = 0 bits, restricted to `touch`) that lie in a background run whose region is not r_out
__device__ __forceinline__ unsigned inner_bg(unsigned word, unsigned touch, int nvalid, const int* L, int ibase, int r_out) {
  unsigned out = 0;
  unsigned cand = ~word & touch;
  const unsigned st = run_starts(word, nvalid);
  while (cand) {
    const int b = __ffs(cand) - 1;
    const int a = 31 - __clz(st & ((2u << b) - 1u));
    const unsigned later = b >= 31 ? 0u : (st & ~((2u << b) - 1u));
    const int e = later ? (__ffs(later) - 2) : 31;
    const unsigned m = run_mask(a, e);
    if (root2(L, ibase + a) != r_out) out |= m;
    cand &= ~m;
  }
  return out;
}
// Per-block aggregation of the run statistics.  A block owns a CONTIGUOUS range of words (STATS_WORDS_PER_BLOCK: 64 rows of a
// 1,024-pixel image), so the runs of one region that fall into the range are summed in a small shared-memory hash table
// and reach the region's global slot as ONE atomic set per block instead of one per run.  (Direct atomics: 19 M RED requests
// for 64 x 1024^2, the busiest L2 slice 32 % occupied by its atomic unit, 0.62 of the kernel's 0.97 ms.)
constexpr int AGG_N = 128;
constexpr int STATS_WORDS_PER_BLOCK = 256;
struct AggEntry { unsigned long long key; double sum, acc_sum; int count, x0, x1, y1, acc_count, pad; };
constexpr unsigned long long AGG_EMPTY = ~0ull;
__device__ __forceinline__ void stat_add_global(CompStat* t, double sum, int count, int xa, int xb, int y) {
  atomicAdd(&t->sum, sum);
  atomicAdd(&t->count, count);
  atomicMin(&t->x0, xa); atomicMax(&t->x1, xb); atomicMax(&t->y1, y);
}
__device__ __forceinline__ void agg_add(AggEntry* agg, CompStat* stat, unsigned long long key, double sum, int count, int xa, int xb, int y) {
  const unsigned hsh = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 40);
#pragma unroll 1
  for (int probe = 0; probe < 4; ++probe) {
    AggEntry* e = agg + ((hsh + probe) & (AGG_N - 1));
    const unsigned long long old = atomicCAS(&e->key, AGG_EMPTY, key);
    if (old == AGG_EMPTY || old == key) {
      atomicAdd(&e->sum, sum);
      atomicAdd(&e->count, count);
      atomicMin(&e->x0, xa); atomicMax(&e->x1, xb); atomicMax(&e->y1, y);
      return;
    }
  }
  stat_add_global(stat + key, sum, count, xa, xb, y);       // table full around this hash: straight to the slot
}
// hole-ring contributions (acc_sum / acc_count of the hole's slot) through the same table
__device__ __forceinline__ void agg_add_ring(AggEntry* agg, CompStat* stat, unsigned long long key, double sum, int count) {
  const unsigned hsh = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 40);
#pragma unroll 1
  for (int probe = 0; probe < 4; ++probe) {
    AggEntry* e = agg + ((hsh + probe) & (AGG_N - 1));
    const unsigned long long old = atomicCAS(&e->key, AGG_EMPTY, key);
    if (old == AGG_EMPTY || old == key) {
      atomicAdd(&e->acc_sum, sum);
      atomicAdd(&e->acc_count, count);
      return;
    }
  }
  atomicAdd(&stat[key].acc_sum, sum);
  atomicAdd(&stat[key].acc_count, count);
}
__global__ void __launch_bounds__(CCL_THREADS, 4)
ccl_stats_kernel(const float* __restrict__ pred, int c, const unsigned* __restrict__ bits, int n, int h, int w, int wq,
                 const int* __restrict__ label, CompStat* __restrict__ stat) {
  __shared__ float tile[CCL_THREADS / 32][32 * 33];
  __shared__ AggEntry agg[AGG_N];
  const int64_t hw = (int64_t)h * w;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int64_t nseg = (int64_t)n * h * wq;
  float* T = tile[wrp];
  for (int t = threadIdx.x; t < AGG_N; t += CCL_THREADS) { AggEntry z; z.key = AGG_EMPTY; z.sum = 0.0; z.acc_sum = 0.0; z.count = 0; z.x0 = 0x7fffffff; z.x1 = -1; z.y1 = -1; z.acc_count = 0; z.pad = 0; agg[t] = z; }
  __syncthreads();
  const int64_t range0 = (int64_t)blockIdx.x * STATS_WORDS_PER_BLOCK;
  const int64_t range1 = range0 + STATS_WORDS_PER_BLOCK < nseg ? range0 + STATS_WORDS_PER_BLOCK : nseg;
  for (int64_t base = range0; base < range1; base += CCL_THREADS) {
    const int64_t widx = base + threadIdx.x;
    const bool live = widx < nseg;
    Seg g; g.img = 0; g.y = 0; g.s = 0; g.x0 = 0; g.nvalid = 0;
    if (live) seg_of(widx, h, wq, w, g);
    const int* L = label + g.img * lab_stride(hw);
    CompStat* S = stat + g.img * hw;
    const int r_out = live ? L[hw] : 0;
    const int i0 = g.y * w + g.x0;
    const unsigned vm = valid_mask(g.nvalid);
    const unsigned cur = live ? bits[widx] : 0u;
    const unsigned st = run_starts(cur, g.nvalid) & vm;
    // the four neighbour words (ring test below), fetched with the word itself
    const bool has_fg = (cur & vm) != 0;
    const unsigned prv = has_fg && g.s > 0 ? bits[widx - 1] : 0xffffffffu, nxt = has_fg && g.s < wq - 1 ? bits[widx + 1] : 0xffffffffu;
    const unsigned up = has_fg && g.y > 0 ? (bits[widx - wq] | ~vm) : 0xffffffffu, dn = has_fg && g.y < h - 1 ? (bits[widx + wq] | ~vm) : 0xffffffffu;
    // roots of the first four runs, looked up together
    int ra[4], rr[4];
    {
      unsigned m = st;
#pragma unroll
      for (int k = 0; k < 4; ++k) { ra[k] = m ? __ffs(m) - 1 : -1; m &= m - 1; }
#pragma unroll
      int l1[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) l1[k] = (live && ra[k] >= 0) ? L[i0 + ra[k]] : -1;          // run start -> (strip) root -> root
#pragma unroll
      for (int k = 0; k < 4; ++k) rr[k] = l1[k] >= 0 ? L[l1[k]] : r_out;
    }
    const int nruns = __popc(st);
    const bool need = live && (nruns > 4 || rr[0] != r_out || rr[1] != r_out || rr[2] != r_out || rr[3] != r_out);
    // ---- stage the probabilities of the needed words of this warp (uniform per k)
    const unsigned needm = __ballot_sync(0xffffffffu, need);
    const int64_t pbase = (int64_t)g.img * c * hw + i0;
    __syncwarp();                                   // the previous chunk's reads of T are done
    if (needm) {
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int64_t pb = __shfl_sync(0xffffffffu, pbase, k);
        const int nv = __shfl_sync(0xffffffffu, g.nvalid, k);
        v[k] = ((needm >> k) & 1u) && lane < nv ? __ldg(pred + pb + lane) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) T[k * 33 + lane] = v[k];
    }
    __syncwarp();
    if (!need) continue;
    const float* row = T + lane * 33;
    // ---- per-run sums -> one atomic set per run that is not the outside region
    unsigned inner_cur = 0;                          // background pixels of this word in enclosed regions (for the ring test)
    {
      double acc = 0.0, sums[4] = {0.0, 0.0, 0.0, 0.0};
      int ends[4] = {0, 0, 0, 0};
      int a = 0, k = 0;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        acc += (double)row[j];
        const bool last = j + 1 >= g.nvalid;
        const bool ends_here = j < g.nvalid && (last || ((st >> (j + 1)) & 1u));
        if (ends_here) {
          if (k < 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) if (q == k) { sums[q] = acc; ends[q] = j; }
          } else {                                   // fifth and later runs of a word (noise): looked up on the spot
            const int r = root2(L, i0 + a);
            if (r != r_out) {
              agg_add(agg, stat, (unsigned long long)(g.img * hw + r), acc, j - a + 1, g.x0 + a, g.x0 + j, g.y);
              if (!((cur >> a) & 1u)) inner_cur |= run_mask(a, j);
            }
          }
          acc = 0.0; a = j + 1; ++k;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q < nruns && rr[q] != r_out) {
          agg_add(agg, stat, (unsigned long long)(g.img * hw + rr[q]), sums[q], ends[q] - ra[q] + 1, g.x0 + ra[q], g.x0 + ends[q], g.y);
          if (!((cur >> ra[q]) & 1u)) inner_cur |= run_mask(ra[q], ends[q]);
        }
      }
    }
    // ---- boundary rings of holes
    const unsigned fg = cur & vm;
    if (!fg) continue;
    const unsigned ones = 0xffffffffu;
    const unsigned curx = cur | ~vm;                                          // beyond the row end: not background
    const int nv_up = g.nvalid;                                               // same column range in the rows above / below
    const unsigned curL = (curx << 1) | (prv >> 31), curR = (curx >> 1) | (nxt << 31);
    if (!(fg & (~curL | ~curR | ~up | ~dn))) continue;                        // interior word
    // background runs around the word that are NOT the outside region
    unsigned innL = inner_cur << 1, innR = inner_cur >> 1;
    const int prv_start = 31 - __clz(run_starts(prv, 32));                    // start of the run that ends the previous word
    if ((fg & 1u) && !(prv >> 31) && root2(L, i0 - 32 + prv_start) != r_out) innL |= 1u;
    if ((fg >> 31) && !(nxt & 1u) && root2(L, i0 + 32) != r_out) innR |= 0x80000000u;
    const unsigned innU = (fg & ~up) ? inner_bg(up, fg, nv_up, L, i0 - w, r_out) : 0u;
    const unsigned innD = (fg & ~dn) ? inner_bg(dn, fg, nv_up, L, i0 + w, r_out) : 0u;
    unsigned slow = fg & (innL | innR | innU | innD);
    if (!slow) continue;
    const unsigned stU = run_starts(up, nv_up), stD = run_starts(dn, nv_up);
    // consecutive ring pixels mostly feed the same hole: keep one pending (region, sum, count) per thread
    long long pend_key = -1; double pend_sum = 0.0; int pend_cnt = 0;
    int last_r = -1, last_pr = -1;
    while (slow) {
      const int b = __ffs(slow) - 1; slow &= slow - 1;
      const unsigned upto = (2u << b) - 1u;
      const int r = root2(L, i0 + 31 - __clz(st & upto));
      if (r != last_r) { last_r = r; last_pr = S[r].parent; }
      const int pr = last_pr;
      int gq[4] = {-1, -1, -1, -1};
      if ((innL >> b) & 1u) gq[0] = b > 0 ? root2(L, i0 + 31 - __clz(st & (upto >> 1))) : root2(L, i0 - 32 + prv_start);
      if ((innR >> b) & 1u) gq[1] = root2(L, i0 + b + 1);
      if ((innU >> b) & 1u) gq[2] = root2(L, i0 - w + 31 - __clz(stU & upto));
      if ((innD >> b) & 1u) gq[3] = root2(L, i0 + w + 31 - __clz(stD & upto));
      const double p = (double)row[b];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int gk = gq[k];
        if (gk < 0 || gk == pr || gk == r_out) continue;
        bool dup = false;
#pragma unroll
        for (int q = 0; q < 4; ++q) if (q < k && gq[q] == gk) dup = true;
        if (dup) continue;
        const long long key = (long long)(g.img * hw + gk);
        if (key != pend_key) {
          if (pend_cnt) agg_add_ring(agg, stat, (unsigned long long)pend_key, pend_sum, pend_cnt);
          pend_key = key; pend_sum = 0.0; pend_cnt = 0;
        }
        pend_sum += p; ++pend_cnt;
      }
    }
    if (pend_cnt) agg_add_ring(agg, stat, (unsigned long long)pend_key, pend_sum, pend_cnt);
  }
  // ---- one atomic set per region and block
  __syncthreads();
  for (int t = threadIdx.x; t < AGG_N; t += CCL_THREADS) {
    const AggEntry e = agg[t];
    if (e.key == AGG_EMPTY) continue;
    if (e.count) stat_add_global(stat + e.key, e.sum, e.count, e.x0, e.x1, e.y1);
    if (e.acc_count) { atomicAdd(&stat[e.key].acc_sum, e.acc_sum); atomicAdd(&stat[e.key].acc_count, e.acc_count); }
  }
}

// E: every region adds its own statistics to all its ancestors.  Roots come from the root bitmap (1/32 word per pixel instead
// of the 4-byte label: 8 MB instead of 268 MB for 64 x 1024^2).
__global__ void __launch_bounds__(CCL_THREADS)
ccl_tree_kernel(int h, int w, int wq, const int* __restrict__ label, const unsigned* __restrict__ rootbits, CompStat* __restrict__ stat) {
  const int img = blockIdx.y;
  const int64_t hw = (int64_t)h * w;
  const int* L = label + img * lab_stride(hw);
  CompStat* S = stat + img * hw;
  const int r_out = L[hw];
  const int nw = h * wq;
  for (int wi = blockIdx.x * CCL_THREADS + threadIdx.x; wi < nw; wi += gridDim.x * CCL_THREADS) {
    unsigned rb = rootbits[(int64_t)img * nw + wi];
    const int y = wi / wq, i0 = y * w + (wi - y * wq) * 32;
    while (rb) {
      const int i = i0 + __ffs(rb) - 1; rb &= rb - 1;
      if (i == r_out) continue;
      const double s = S[i].sum;
      const int cnt = S[i].count;
      int a = S[i].parent;
      while (a != r_out) {
        atomicAdd(&S[a].acc_sum, s);
        atomicAdd(&S[a].acc_count, cnt);
        a = S[a].parent;
      }
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
  const int r_out = label[img * lab_stride((int64_t)h * w) + (int64_t)h * w];
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
  const int* L = label + img * lab_stride(hw);
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
      if (cd.kind == 0) { cd.x0 = s.x0; cd.y0 = y; cd.x1 = s.x1; cd.y1 = s.y1; }
      else { cd.x0 = s.x0 - 1; cd.y0 = y - 1; cd.x1 = s.x1 + 1; cd.y1 = s.y1 + 1; }   // + the ring of parent pixels
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
  const int* L = label + img * lab_stride(hw);
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
    const int r = root2(L, g.y * w + g.x0 + a);               // pixel -> run start or strip root -> root
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
      L = label + g.img * lab_stride(hw); S = stat + g.img * hw; r_out = L[hw];
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
  return ccl_align(sizeof(unsigned) * (size_t)n * h * wq) + ccl_align(sizeof(int) * (size_t)n * lab_stride(hw)) +
         ccl_align(sizeof(CompStat) * (size_t)n * hw) + 2 * ccl_align(sizeof(int) * (size_t)n * ccl_nblk(hw)) +
         ccl_align(sizeof(unsigned) * (size_t)n * h * wq) + 256;
}

extern "C" int dbb_binarize_ccl_score(const float* pred, int64_t n, int c, int64_t h, int64_t w, float thresh, double box_thresh,
                                      uint8_t* bitmap, int32_t* labels, DbbCandidate* cands, int32_t* n_cands, int max_cands,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  if (!pred || !bitmap || !cands || !n_cands || !workspace) return set_error(DBB_EINVAL, "binarize_ccl_score: null pointer");
  if (n <= 0 || c <= 0 || h <= 0 || w <= 0 || max_cands <= 0 || n > 65535) return set_error(DBB_EINVAL, "binarize_ccl_score: bad shape");
  if (h * w >= (int64_t)1 << 31) return set_error(DBB_EUNSUPPORTED, "binarize_ccl_score: image too large for 32-bit labels");
  if (n * h >= (int64_t)1 << 31) return set_error(DBB_EUNSUPPORTED, "binarize_ccl_score: too many rows for 32-bit row indices");
  if (n * h * ((w + 31) / 32) >= (int64_t)1 << 32) return set_error(DBB_EUNSUPPORTED, "binarize_ccl_score: batch too large for 32-bit word indices");
  if (workspace_bytes < dbb_postprocess_workspace(n, h, w)) return set_error(DBB_EWORKSPACE, "binarize_ccl_score: workspace too small");
  if (!aligned16(workspace)) return set_error(DBB_EALIGN, "binarize_ccl_score: workspace not 16B aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t hw = h * w;
  const int wq = (int)((w + 31) / 32);
  CclWs ws = ccl_carve(workspace, n, hw, h, wq);
  const int64_t nseg = n * h * wq;
  int grow = (int)((n * h + CCL_THREADS / 32 - 1) / (CCL_THREADS / 32));
  if (grow > DBB_NUM_SMS * 8) grow = DBB_NUM_SMS * 8;
  const int nblk = ccl_nblk(hw);
  int gx = nblk < DBB_NUM_SMS * 8 ? nblk : DBB_NUM_SMS * 8;
  const dim3 grid((unsigned)gx, (unsigned)n), gridb((unsigned)nblk, (unsigned)n);
  // thread-per-word kernels
  int gword = (int)((nseg + CCL_THREADS - 1) / CCL_THREADS);
  if (gword > DBB_NUM_SMS * 8) gword = DBB_NUM_SMS * 8;
  const int word_stores = (w % 4 == 0 && ((uintptr_t)bitmap & 3) == 0) ? 1 : 0;
  const size_t strip_smem = sizeof(int) * ((size_t)STRIP_ROWS * w + 1 + (size_t)STRIP_ROWS * wq);
  const bool fused = w >= 16 && strip_smem <= 200 * 1024 && !getenv("DBB_CCL_NO_STRIP");
  const int sh = fused ? STRIP_ROWS : 32;                    // strip height of the two-level merge
  if (fused) {
    // strips labelled in shared memory (pack + link + flatten in one pass over the probabilities)
    static size_t attr_set = 0;
    if (strip_smem > attr_set) {
      DBB_CUDA(cudaFuncSetAttribute(ccl_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)strip_smem));
      attr_set = strip_smem;
    }
    const int64_t nstrips = n * ((h + STRIP_ROWS - 1) / STRIP_ROWS);
    const int gstrip = (int)(nstrips < (int64_t)DBB_NUM_SMS * 64 ? nstrips : (int64_t)DBB_NUM_SMS * 64);
    DBB_LAUNCH("ccl_strip", s, ccl_strip_kernel<<<gstrip, STRIP_THREADS, strip_smem, s>>>(pred, c, (int)n, (int)h, (int)w, wq, thresh, bitmap, ws.bits, ws.label,
                                                                                      ws.blk_count, ws.rootbits, word_stores));
  } else {
    DBB_LAUNCH("ccl_pack_init", s, ccl_pack_init_kernel<<<grow, CCL_THREADS, 0, s>>>(pred, c, (int)n, (int)h, (int)w, wq, thresh, bitmap, ws.bits, ws.label, word_stores));
    DBB_LAUNCH("ccl_link", s, ccl_link_kernel<<<gword, CCL_THREADS, 0, s>>>(ws.bits, (int)n, (int)h, (int)w, wq, ws.label, 0, sh, nullptr));
    if (h > sh) DBB_LAUNCH("ccl_flatten_strips", s, ccl_flatten_kernel<<<gword, CCL_THREADS, 0, s>>>(ws.bits, (int)n, (int)h, (int)w, wq, ws.label, ws.stat, 0, ws.rootbits, 0));
  }
  if (h > sh || fused) {
    const int64_t nedge_words = n * ((h - 1) / sh) * wq;
    int gedge = (int)((nedge_words + CCL_THREADS - 1) / CCL_THREADS);
    if (gedge > DBB_NUM_SMS * 8) gedge = DBB_NUM_SMS * 8;
    if (gedge < 1) gedge = 1;
    DBB_LAUNCH("ccl_link_seams", s, ccl_link_kernel<<<gedge, CCL_THREADS, 0, s>>>(ws.bits, (int)n, (int)h, (int)w, wq, ws.label, 1, sh, fused ? ws.blk_count : nullptr));
  }
  DBB_LAUNCH("ccl_flatten", s, ccl_flatten_kernel<<<gword, CCL_THREADS, 0, s>>>(ws.bits, (int)n, (int)h, (int)w, wq, ws.label, ws.stat, 1, ws.rootbits, fused ? 1 : 0));
  DBB_LAUNCH("ccl_stats", s, ccl_stats_kernel<<<(unsigned)((nseg + STATS_WORDS_PER_BLOCK - 1) / STATS_WORDS_PER_BLOCK), CCL_THREADS, 0, s>>>(pred, c, ws.bits, (int)n, (int)h, (int)w, wq, ws.label, ws.stat));
  
  const int nwblk = (int)((h * wq + CCL_THREADS - 1) / CCL_THREADS);        // blocks of 256 words in raster order
  const dim3 gridw((unsigned)nwblk, (unsigned)n);
  DBB_LAUNCH("ccl_tree", s, ccl_tree_kernel<<<gridw, CCL_THREADS, 0, s>>>((int)h, (int)w, wq, ws.label, ws.rootbits, ws.stat));
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
