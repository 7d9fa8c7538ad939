// Fused CenterNet decode for sm_100a: sigmoid + kxk max-pool-equals-keep pseudo-NMS + class arg-max
// (one streaming pass over the head output) followed by a per-image top-k select + box / embedding gather.
//
// Replaces (reference paths): centernet_lightning/models/centernet.py:229-304
//   get_topk_from_heatmap   :243-261  (max_pool2d == heatmap, heatmap*mask, max over classes, topk, gather)
//   gather_and_decode_boxes :263-304
// and centernet_lightning/models/fairmot.py:63-73 (embedding gather).
//
// Kernel 1 (peaks):   reads the (N,C,H,W) fp32 map exactly once from HBM (halo rows are L2 hits) and writes one
//                     (score, label) candidate per pixel: 6 B/pixel instead of 4*C B/pixel.  HBM-bound.
// Kernel 2 (select):  one CTA per image: exact radix select of the k-th largest candidate, ordered tie handling
//                     (score desc, flat index asc), bitonic sort of the k winners, box decode with the reference's
//                     op-by-op fp32 rounding (no FMA contraction), optional reid gather.
#include <cfloat>
#include <cmath>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include "cnl_common.h"

#ifndef CNL_NEARTIE_MODE
#define CNL_NEARTIE_MODE 1
#endif
#ifndef CNL_PEAKS_PIPELINE
#define CNL_PEAKS_PIPELINE 0      // measured on B200: prefetching the next class inside the walk costs registers (spills at 64-72) and is slower
#endif
#ifndef CNL_PEAKS_MINB
#define CNL_PEAKS_MINB 8
#endif

namespace cnl {

static thread_local char g_err[512] = "";
char* error_buffer() { return g_err; }
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ------------------------------------------------------------------------------------------------------------
// Scalar semantics shared by every peaks kernel
// ------------------------------------------------------------------------------------------------------------

// (sigmoid32, the logistic the from_logits path specifies, lives in cnl_common.h)

// from_logits semantics.  The reference compares PROBABILITIES (centernet.py:252-254): a pixel is kept when
// sigmoid(x) == max over the window of sigmoid(.), and the label is the FIRST class whose kept probability equals the
// per-pixel maximum.  The logistic is monotone, so sigmoid(window max of the logits) is that window maximum and the
// streaming pass works on logits: x is a peak iff x == m or sigmoid32(x) == sigmoid32(m), m = window max of the logits.
// fp32 rounding maps many logits to one probability - 8 consecutive floats at x = 3, 211 at x = 8, every x >= 16.64 to
// exactly 1.0f - so the second clause is real (two neighbours 8.0 / 8.00001 are BOTH peaks in the reference).  It is
// evaluated only where m - x is below collapse_thr(m), a cheap conservative bound on the width of such a collapse;
// the per-pixel maximum is converted to a probability once, at the end of the pass, and everything downstream (histogram,
// select, ordering) runs on probabilities exactly like the from_logits = 0 entry point.
//
// collapse_thr: sigmoid32(x) == sigmoid32(m) with x < m  implies  m - x < collapse_thr(m).  With up to 3 ulp of error per
// evaluated logistic, equal results mean the true values differ by at most 6 ulp(p): for m >= 0 (p in [0.5,1], ulp 2^-24,
// slope >= 0.25 e^-m) that is m - x <= 1.43e-6 e^m, for m < 0 (ulp <= 2^-23 p, slope >= p/2) m - x <= 1.43e-6.
// Returned: 2^-18 * 2^ceil(max(m,0) * log2 e) (>= 2.6x the bound).  Saturated (>= 16.5: p is 1 - 2^-24 or 1) and
// underflowing (< -80: p is denormal or 0, absolute spacing) logits have no useful bound: +inf, always take the exact test.
// collapse_exp: the exponent e of that power of two (thr = 2^e), or kNoBound.
constexpr int kNoBound = 1 << 20;
__device__ __forceinline__ int collapse_exp(float m) {
  if (!(m > -80.0f && m < 16.5f)) return kNoBound;            // also catches NaN
  return (int)ceilf(fmaxf(m, 0.0f) * 1.4426950f) - 18;       // -18 .. 6
}
__device__ __forceinline__ float pow2f(int e) { return __int_as_float((127 + e) << 23); }
__device__ __forceinline__ float collapse_thr(float m) {
  const int e = collapse_exp(m);
  return e == kNoBound ? INFINITY : pow2f(e);
}
// exact peak test of the from_logits path for a centre x below its window max m
__device__ __forceinline__ bool same_probability(float x, float m) { return sigmoid32(x) == sigmoid32(m); }

// x: centre value, m: kxk window max (m >= x).  Reference: mask = (maxpool(h) == h); h*mask; max over classes.
// The streaming pass keeps only the per-pixel MAXIMUM of the masked values (3 ALU ops per element); which class
// attained it is recovered afterwards for the k winners only (select kernel), by the reference's rule "first class
// whose masked value equals the maximum".
// `best` starts at -inf for logits ("no peak yet" = probability 0) and at 0 for probabilities (a non-peak contributes
// h*mask = 0; exact for non-negative inputs, i.e. probabilities).  One compare + one predicated max per element.
template <bool LOGITS>
__device__ __forceinline__ void masked_max(float& best, float x, float m) {
  if (x == m) best = fmaxf(best, x);
}
template <bool LOGITS>
__device__ __forceinline__ float best_init() { return LOGITS ? -INFINITY : 0.0f; }
// value written per pixel: always a probability (0 = no peak, as h*mask leaves it in the reference)
template <bool LOGITS>
__device__ __forceinline__ float best_to_prob(float best) {
  if (!LOGITS) return best;
  return (best == -INFINITY) ? 0.0f : sigmoid32(best);
}

__device__ __forceinline__ uint32_t sortable_key(float f) {       // larger float <=> larger key (NaN-free input)
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}
constexpr int kHistBins = 4096;             // per-image histogram of the candidates, built by the select kernel in shared memory
// Candidates are probabilities.  Bin = [exponent | top 8 mantissa bits] of the float, offset so that p = 1.0f is the last
// bin and everything below 2^-16 the first: 256 bins per octave between 2^-16 and 1 (0.4 % wide - the top-k of a map almost
// always ends inside one or two of them, where float's own exponent/3-mantissa-bit split put all of [0.5, 1) into 8 bins).
// Monotone in the key; bins 1 .. kHistBins-2 hold one exponent each, so the next 11 mantissa bits refine them.
constexpr uint32_t kBinBase = ((127u - 16u) << 8) + 1u;
constexpr int kSubShift = 4;                // second-level digit of an un-clamped bin: key bits [14:4]
__device__ __forceinline__ uint32_t cand_bin(uint32_t key) {
  if (!(key & 0x80000000u)) return 0u;                        // negative inputs (not probabilities) sort below everything
  const int v = (int)((key & 0x7fffffffu) >> 15) - (int)kBinBase;
  return (uint32_t)min(max(v, 0), kHistBins - 1);
}

constexpr int kNoPeakBin = 0;                // cand_bin(key(0.f)): pixels without any peak

// ------------------------------------------------------------------------------------------------------------
// Kernel 1a: fast streaming peaks kernel (W % 4 == 0).  One warp = RxTW pixel strip x one class group.
// ------------------------------------------------------------------------------------------------------------
constexpr int kTW = 128;      // columns per warp tile: 32 lanes x float4

__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// Class loop of one warp: rows [r0-P, r0+R+P) x columns [x0-P, x0+4*VEC+P) of classes [c_begin, c_end).
// A lane owns 4*VEC consecutive columns (VEC float4 loads per row), the warp a tile of 128*VEC columns.
// EDGE=false is the interior fast path (every row and column of the strip is inside the map: no predicates).
//
// CNL_PEAKS_PIPELINE=1 (build-time experiment, off): output row i reads strip rows i .. i+2P, so strip row i is dead once
// output row i is done and can be refilled with the SAME row of the NEXT class right there (loads in flight while the
// warp computes, without a second set of registers).  Measured on B200 at 32x80x128x128: 32.3 us against 28.3 us for the
// plain load-then-walk loop - the extra address registers spill at the 64-register budget that keeps 8 CTAs per SM.
//
// LOGITS: the walk applies the x == m test and tracks the smallest non-zero gap m - x; only when some lane of the warp saw
// a near tie (see the semantics note above) the class is re-read and walked a second time with the exact probability
// test for the centres below their window max.
// WC: compile-time row length (128, the 512x512-input map) or 0 = runtime W; a constant turns every row offset into an
// immediate and frees the registers the pipelined loads need.
template <int P, bool LOGITS, int R, bool MT, bool EDGE, int VEC, int WC>
__device__ __forceinline__ void peaks_class_loop(const float* __restrict__ base, size_t plane, int c_begin, int c_end,
                                                 int H, int W_rt, int r0, int x0, int lane,
                                                 float (&best)[R][4 * VEC]) {
  const int W = WC ? WC : W_rt;
  constexpr int ROWS = R + 2 * P;
  constexpr int PH = (P > 0) ? P : 1;
  constexpr int NC = 4 * VEC;                        // columns per lane
  const float NEG = -INFINITY;
  bool row_ok[ROWS];
#pragma unroll
  for (int j = 0; j < ROWS; ++j) { int r = r0 - P + j; row_ok[j] = !EDGE || (r >= 0 && r < H); }
  bool col_ok[VEC];
#pragma unroll
  for (int u = 0; u < VEC; ++u) col_ok[u] = !EDGE || (x0 + 4 * u < W);
  const float edge_l = (lane == 0) ? NEG : 0.0f, edge_r = (lane == 31) ? NEG : 0.0f;
  // Halo columns of neighbouring column tiles (rows wider than one warp tile).  The neighbour shuffles of the walk are
  // rotations, so lane 31's slot in the "from the left" shuffle and lane 0's slot in the "from the right" shuffle are
  // free: lane 31 carries the tile's LEFT halo columns, lane 0 its RIGHT halo columns.
  const bool halo_lane = (lane == 0 || lane == 31);

  float4 v[ROWS][VEC];
  float hx[ROWS][PH];
  auto load_row = [&](const float* pl, int j) {
#pragma unroll
    for (int u = 0; u < VEC; ++u) {
      if (EDGE) v[j][u] = (row_ok[j] && col_ok[u]) ? ld_stream4(pl + (long long)j * W + 4 * u) : make_float4(NEG, NEG, NEG, NEG);
      else      v[j][u] = ld_stream4(pl + (long long)j * W + 4 * u);
    }
    if constexpr (P > 0 && MT) {
#pragma unroll
      for (int q = 0; q < P; ++q) {
        hx[j][q] = NEG;
        const int xt = (lane == 31) ? (x0 - 31 * NC - 1 - q) : (x0 + 32 * NC + q);     // tile's first column - 1 - q / last + 1 + q
        if (halo_lane && row_ok[j] && xt >= 0 && xt < W) hx[j][q] = __ldg(pl + (long long)j * W + (xt - x0));
      }
    }
  };
  auto load_strip = [&](const float* pl) {
#pragma unroll
    for (int j = 0; j < ROWS; ++j) load_row(pl, j);
  };

  if (c_begin >= c_end) return;
#if CNL_PEAKS_PIPELINE
  load_strip(base + (size_t)c_begin * plane);
#endif
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    const float* pl = base + (size_t)c * plane;
    const float* pl_next = pl + plane;
#if CNL_PEAKS_PIPELINE
    const bool has_next = (c + 1 < c_end);
#else
    const bool has_next = false;
    load_strip(pl);
#endif
    // Near-tie detector of the fast walk: the smallest non-zero gap m - x it saw, compared after the walk with the collapse
    // bound of the largest window max this lane can have met (the vertical maxima of its own columns, its neighbour lanes'
    // and the halo columns).  CNL_NEARTIE_MODE 0 disables it (timing experiments only: not reference-exact).
    float gap = INFINITY;
    float tmax = -INFINITY;
    // one output row; EXACT = false: x == m test (+ detector), EXACT = true: probability test for the centres that are
    // below their window max by less than thr
    auto walk_row = [&](auto exact_tag, int i, float thr) {
      constexpr bool EXACT = decltype(exact_tag)::value;
      // vertical max over the window rows
      float4 vm[VEC];
#pragma unroll
      for (int u = 0; u < VEC; ++u) vm[u] = v[i][u];
      float vh[PH];
      if constexpr (P > 0 && MT) {
#pragma unroll
        for (int q = 0; q < P; ++q) vh[q] = hx[i][q];
      }
#pragma unroll
      for (int j = 1; j <= 2 * P; ++j) {
#pragma unroll
        for (int u = 0; u < VEC; ++u) vm[u] = max4(vm[u], v[i + j][u]);
        if constexpr (P > 0 && MT) {
#pragma unroll
          for (int q = 0; q < P; ++q) vh[q] = fmaxf(vh[q], hx[i + j][q]);
        }
      }
      if constexpr (LOGITS && P > 0 && !EXACT && CNL_NEARTIE_MODE != 0) {
#pragma unroll
        for (int u = 0; u < VEC; ++u) tmax = fmaxf(fmaxf(tmax, fmaxf(vm[u].x, vm[u].y)), fmaxf(vm[u].z, vm[u].w));
        if constexpr (MT) {
#pragma unroll
          for (int q = 0; q < P; ++q) tmax = fmaxf(tmax, vh[q]);
        }
      }
      // horizontal: e[0..P-1] left neighbours (nearest last), e[P..P+NC-1] own, e[P+NC..] right neighbours
      float e[NC + 2 * P];
#pragma unroll
      for (int u = 0; u < VEC; ++u) {
        e[P + 4 * u + 0] = vm[u].x; e[P + 4 * u + 1] = vm[u].y; e[P + 4 * u + 2] = vm[u].z; e[P + 4 * u + 3] = vm[u].w;
      }
      if constexpr (P > 0) {
#pragma unroll
        for (int q = 0; q < P; ++q) {
          // q-th column to the left of x0 is the (NC-1-q)-th column of lane-1; to the right of x0+NC-1 it is column q of lane+1
          float src_l = e[P + NC - 1 - q], src_r = e[P + q];
          if constexpr (MT) { src_l = (lane == 31) ? vh[q] : src_l; src_r = (lane == 0) ? vh[q] : src_r; }
          float fl = __shfl_sync(0xffffffffu, src_l, (lane + 31) & 31);
          float fr = __shfl_sync(0xffffffffu, src_r, (lane + 1) & 31);
          if constexpr (!MT) { fl += edge_l; fr += edge_r; }      // -inf beyond the row ends (FADD: keeps the ALU pipe free)
          e[P - 1 - q] = fl;
          e[P + NC + q] = fr;
        }
      }
#pragma unroll
      for (int u = 0; u < VEC; ++u) {
        const float ctr[4] = {v[i + P][u].x, v[i + P][u].y, v[i + P][u].z, v[i + P][u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float m = e[4 * u + j];
#pragma unroll
          for (int q = 1; q <= 2 * P; ++q) m = fmaxf(m, e[4 * u + j + q]);
          if constexpr (!EXACT) {
            masked_max<LOGITS>(best[i][4 * u + j], ctr[j], m);
            if constexpr (LOGITS && P > 0 && CNL_NEARTIE_MODE != 0) gap = fminf(gap, (ctr[j] == m) ? INFINITY : m - ctr[j]);
          } else {
            // (an out-of-map centre is -inf: m - ctr = inf or NaN, never below thr)
            if (ctr[j] != m && m - ctr[j] < thr && same_probability(ctr[j], m)) best[i][4 * u + j] = fmaxf(best[i][4 * u + j], ctr[j]);
          }
        }
      }
    };
#pragma unroll
    for (int i = 0; i < R; ++i) {
      walk_row(std::false_type{}, i, 0.0f);
      if (has_next) load_row(pl_next, i);            // strip row i is dead: refill it with the next class (prefetch)
    }
    if (has_next) {
#pragma unroll
      for (int j = R; j < ROWS; ++j) load_row(pl_next, j);
    }
    if constexpr (LOGITS && P > 0 && CNL_NEARTIE_MODE != 0) {
      const float tl = __shfl_sync(0xffffffffu, tmax, (lane + 31) & 31), tr = __shfl_sync(0xffffffffu, tmax, (lane + 1) & 31);
      const float thr = collapse_thr(fmaxf(tmax, fmaxf(tl, tr)));
      if (__any_sync(0xffffffffu, gap < thr)) {       // near tie somewhere in this warp's strip: rare
        load_strip(pl);                               // re-read this class from L2 (its registers hold the next one by now)
#pragma unroll
        for (int i = 0; i < R; ++i) walk_row(std::true_type{}, i, thr);
        if (has_next) load_strip(pl_next);
      }
    }
  }
}

template <int P, bool LOGITS, int R, int G, bool MT, int VEC, int WC = 0>
__global__ void __launch_bounds__(G * 32, (R * VEC <= 4) ? CNL_PEAKS_MINB : 4)
peaks_fast_kernel(const float* __restrict__ heat, float* __restrict__ cbest, uint8_t* __restrict__ cgroup,
                  int C, int H, int W_rt) {
  const int W = WC ? WC : W_rt;
  constexpr int NC = 4 * VEC;
  constexpr int TW = kTW * VEC;                  // columns per warp tile
  const int lane = threadIdx.x & 31;
  const int g = threadIdx.x >> 5;
  const int n = blockIdx.z;
  const int r0 = blockIdx.y * R;
  const int x0 = blockIdx.x * TW + lane * NC;
  const size_t plane = (size_t)H * W;
  const int cg = (C + G - 1) / G;
  const int c_begin = g * cg;
  const int c_end = min(C, c_begin + cg);

  float best[R][NC];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < NC; ++j) best[i][j] = best_init<LOGITS>();

  const float* base = heat + (size_t)n * C * plane + (long long)(r0 - P) * W + x0;
  const bool interior = (r0 - P >= 0) && (r0 + R + P <= H) && ((int)(blockIdx.x + 1) * TW <= W);   // block-uniform
  if (interior) peaks_class_loop<P, LOGITS, R, MT, false, VEC, WC>(base, plane, c_begin, c_end, H, W, r0, x0, lane, best);
  else          peaks_class_loop<P, LOGITS, R, MT, true, VEC, WC>(base, plane, c_begin, c_end, H, W, r0, x0, lane, best);

  // every CTA is past its streaming loop: the select kernel's CTAs may be scheduled while this grid drains (they wait
  // in griddepcontrol.wait until this grid has completed and flushed)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // merge the G class groups through shared memory (a plain max), then emit one candidate per pixel
  __shared__ __align__(16) float s_v[G][R][TW];
  if (G > 1) {
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int u = 0; u < VEC; ++u)
        *reinterpret_cast<float4*>(&s_v[g][i][lane * NC + 4 * u]) =
            make_float4(best[i][4 * u], best[i][4 * u + 1], best[i][4 * u + 2], best[i][4 * u + 3]);
    __syncthreads();
  }
  for (int i = g; i < R; i += G) {          // warp g finishes rows g, g+G, ...  (warp-uniform trip count)
    const int r = r0 + i;
#pragma unroll
    for (int u = 0; u < VEC; ++u) {
      const int xu = x0 + 4 * u;
      const bool ok = (r < H) && (xu < W);
      float bv[4];
      uint32_t grp = 0;                      // which class group attained the maximum (first group on ties), 8 bits per pixel
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (G > 1) {
          bv[j] = s_v[0][i][lane * NC + 4 * u + j];
          uint32_t gj = 0;
#pragma unroll
          for (int gg = 1; gg < G; ++gg) {
            const float ov = s_v[gg][i][lane * NC + 4 * u + j];
            if (ov > bv[j]) { bv[j] = ov; gj = (uint32_t)gg; }
          }
          if (LOGITS) {
            // an EARLIER group whose best logit is smaller but may round to the same probability owns the first maximal
            // class (reference: first class whose kept probability equals the maximum): group unknown, scan every class
            const float thr = collapse_thr(bv[j]);
#pragma unroll
            for (int gg = 0; gg < G - 1; ++gg)
              if ((uint32_t)gg < gj && bv[j] - s_v[gg][i][lane * NC + 4 * u + j] < thr) gj = 0xffu;
          }
          grp |= gj << (8 * j);
        } else {
          bv[j] = best[i][4 * u + j];
        }
        bv[j] = best_to_prob<LOGITS>(bv[j]);
      }
      if (ok) {
        *reinterpret_cast<uint32_t*>(cgroup + (size_t)n * plane + (size_t)r * W + xu) = grp;
        *reinterpret_cast<float4*>(cbest + (size_t)n * plane + (size_t)r * W + xu) = make_float4(bv[0], bv[1], bv[2], bv[3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Kernel 1b: generic peaks kernel (any W, any odd window up to 7): one thread per pixel.
// ------------------------------------------------------------------------------------------------------------
template <bool LOGITS>
__global__ void __launch_bounds__(256)
peaks_generic_kernel(const float* __restrict__ heat, float* __restrict__ cbest, uint8_t* __restrict__ cgroup,
                     int C, int H, int W, int P) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int n = blockIdx.z;
  if (x >= W || y >= H) return;
  const size_t plane = (size_t)H * W;
  const float* img = heat + (size_t)n * C * plane;
  float best = best_init<LOGITS>();
  const int y_lo = max(0, y - P), y_hi = min(H - 1, y + P);
  const int x_lo = max(0, x - P), x_hi = min(W - 1, x + P);
  for (int c = 0; c < C; ++c) {
    const float* pl = img + (size_t)c * plane;
    const float ctr = __ldg(pl + (size_t)y * W + x);
    float m = ctr;
    for (int yy = y_lo; yy <= y_hi; ++yy)
      for (int xx = x_lo; xx <= x_hi; ++xx) m = fmaxf(m, __ldg(pl + (size_t)yy * W + xx));
    if (LOGITS) {
      if (ctr == m || (m - ctr < collapse_thr(m) && same_probability(ctr, m))) best = fmaxf(best, ctr);
    } else {
      masked_max<LOGITS>(best, ctr, m);
    }
  }
  best = best_to_prob<LOGITS>(best);
  cbest[(size_t)n * plane + (size_t)y * W + x] = best;
  cgroup[(size_t)n * plane + (size_t)y * W + x] = 0xff;
}

// ------------------------------------------------------------------------------------------------------------
// Kernel 2: per-image exact top-k and gather/decode.  One CTA per image.
//   fast path:  kernel 1 left a 4096-bin histogram of the candidate keys; a suffix scan finds the bin holding the
//               k-th largest, every candidate in that bin or above (usually k .. a few hundred) is collected and
//               bitonic-sorted on (key desc, index asc).
//   fallback:   when that set does not fit (large plateaus of equal scores) an exact 3-pass radix select with
//               ordered tie handling runs instead.
// ------------------------------------------------------------------------------------------------------------
constexpr int kSelThreads = 1024;
constexpr int kMaxK = 1024;
constexpr int kListCap = 2048;
constexpr int kBins = 2048;
constexpr int kRefineAbove = 256;
constexpr int kLabelBatch = 3;               // classes per lane and batch in the label recovery (8 lanes x 3 = 24 classes at once)            // sort directly when bin_k-and-above holds at most this many

// inclusive block scan over kSelThreads ints (warp shuffles + one smem hop)
__device__ __forceinline__ int block_inclusive_scan(int v, int* s_warp /*[32]*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) s_warp[wid] = v;
  __syncthreads();
  if (wid == 0) {
    int w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  if (wid > 0) v += s_warp[wid - 1];
  __syncthreads();          // s_warp may be reused by the caller right away
  return v;
}

// reference centernet.py:278-303 for one detection: every fp32 op rounded separately (no FMA contraction)
__device__ __forceinline__ float4 decode_box_from(const float (&g)[4], int idx, int H, int W,
                                                  int normalize, int box_log, float mult, float stride_f) {
  float cx = __fadd_rn((float)(idx % W), 0.5f);
  float cy = __fadd_rn((float)(idx / W), 0.5f);
  float s[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float gv = g[c];
    if (box_log) gv = expf(gv);
    gv = __fmul_rn(gv, mult);
    s[c] = (gv < 0.0f) ? 0.0f : gv;            // torch.clamp_min(x, 0): a NaN stays NaN (fmaxf would return 0)
  }
  float x1 = __fsub_rn(cx, s[0]), y1 = __fsub_rn(cy, s[1]);
  float x2 = __fadd_rn(cx, s[2]), y2 = __fadd_rn(cy, s[3]);
  if (normalize) {
    x1 = __fdiv_rn(x1, (float)W); x2 = __fdiv_rn(x2, (float)W);
    y1 = __fdiv_rn(y1, (float)H); y2 = __fdiv_rn(y2, (float)H);
  } else {
    x1 = __fmul_rn(x1, stride_f); y1 = __fmul_rn(y1, stride_f);
    x2 = __fmul_rn(x2, stride_f); y2 = __fmul_rn(y2, stride_f);
  }
  return make_float4(x1, y1, x2, y2);
}
__device__ __forceinline__ float4 decode_box(const float* box_img, size_t plane, int idx, int H, int W,
                                             int normalize, int box_log, float mult, float stride_f) {
  const float g[4] = {__ldg(box_img + idx), __ldg(box_img + plane + idx), __ldg(box_img + 2 * plane + idx),
                      __ldg(box_img + 3 * plane + idx)};
  return decode_box_from(g, idx, H, W, normalize, box_log, mult, stride_f);
}

// Stand-alone gather for caller-supplied indices (the reference's staticmethod is also used outside decode).
__global__ void gather_boxes_kernel(const float* __restrict__ box, const long long* __restrict__ indices,
                                    int N, int H, int W, int k, int normalize, int box_log, float mult, float stride_f,
                                    float* __restrict__ boxes) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * k) return;
  int n = t / k;
  size_t plane = (size_t)H * W;
  long long idx = indices[t];
  if (idx < 0 || idx >= (long long)plane) {      // torch.gather would raise; mark the row instead of faulting
    *reinterpret_cast<float4*>(boxes + (size_t)t * 4) = make_float4(NAN, NAN, NAN, NAN);
    return;
  }
  *reinterpret_cast<float4*>(boxes + (size_t)t * 4) =
      decode_box(box + (size_t)n * 4 * plane, plane, (int)idx, H, W, normalize, box_log, mult, stride_f);
}

// torchvision.ops.box_convert(boxes, 'xyxy', 'xywh') as the reference's validation_step applies it (centernet.py:207)
__global__ void xyxy_to_xywh_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 b = in[i];
  out[i] = make_float4(b.x, b.y, __fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

// Elementwise logistic for the G1 forward() alias (heads return probabilities there): same sigmoid32 as the fused decode.
__global__ void sigmoid_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = sigmoid32(in[i]);
}

struct DecodeParams {
  const float* cscore;                               // per-pixel best kept probability (0 = no peak)
  const float* heat; int C, P, from_logits;          // the head map itself: labels are recovered for the k winners only
  const uint8_t* cgroup; int group_classes;          // class group (of group_classes classes) that holds the winner; 255 = unknown
  const float* box; const float* reid;
  int H, W, E, k;
  int normalize, box_log; float mult; float stride_f;
  float* boxes; float* scores; long long* labels; long long* indices; float* emb;
  float* packed; int packed_w;                       // optional packed rows [x1,y1,x2,y2,score,index bits,label lo,label hi,emb...]
  unsigned long long* sortbuf; int sort_cap;         // k > kMaxK: per-image global sort buffer of sort_cap = pow2(H*W) entries
};

__device__ __forceinline__ unsigned long long pack_entry(uint32_t key, int idx) {
  return ((unsigned long long)key << 32) | (uint32_t)(0xffffffffu - (uint32_t)idx);   // descending = score desc, index asc
}

// exact radix select + ordered ties: fills s_list[0..k) (unsorted).  Used only when the histogram path overflows.
__device__ void radix_select_fallback(const float* sc, int HW, int k, unsigned long long* s_list, int* s_hist,
                                      int* s_warp, int* s_scalars /*[3]: bin, above, cnt*/) {
  const int tid = threadIdx.x;
  uint32_t prefix = 0, mask = 0;
  int need = k;
  int count_at_T = 0;
  const int shifts[3] = {21, 10, 0};
  const int nbits[3] = {11, 11, 10};
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = shifts[pass];
    const uint32_t dm = (1u << nbits[pass]) - 1u;
    for (int i = tid; i < kBins; i += kSelThreads) s_hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < HW; i += kSelThreads) {
      uint32_t u = sortable_key(sc[i]);
      if ((u & mask) == prefix) atomicAdd(&s_hist[(u >> shift) & dm], 1);
    }
    __syncthreads();
    int b_hi = s_hist[kBins - 1 - 2 * tid];
    int b_lo = s_hist[kBins - 2 - 2 * tid];
    int incl = block_inclusive_scan(b_hi + b_lo, s_warp);
    int excl = incl - (b_hi + b_lo);
    if (excl < need && need <= incl) {            // exactly one thread
      if (need <= excl + b_hi) { s_scalars[0] = kBins - 1 - 2 * tid; s_scalars[1] = excl; s_scalars[2] = b_hi; }
      else                     { s_scalars[0] = kBins - 2 - 2 * tid; s_scalars[1] = excl + b_hi; s_scalars[2] = b_lo; }
    }
    __syncthreads();
    need -= s_scalars[1];
    prefix |= ((uint32_t)s_scalars[0]) << shift;
    mask |= dm << shift;
    count_at_T = s_scalars[2];
    __syncthreads();
  }
  const uint32_t T = prefix;              // key of the k-th largest candidate
  const int n_greater = k - need;         // all keys > T are selected; `need` (>=1) of the keys == T
  if (tid == 0) s_scalars[2] = 0;
  __syncthreads();
  const bool all_ties_taken = (count_at_T == need);
  for (int i = tid; i < HW; i += kSelThreads) {
    uint32_t u = sortable_key(sc[i]);
    if (u > T || (all_ties_taken && u == T)) s_list[atomicAdd(&s_scalars[2], 1)] = pack_entry(u, i);
  }
  if (!all_ties_taken) {
    // more candidates equal to T than slots left: take the lowest flat indices (deterministic tie rule)
    const int L = (HW + kSelThreads - 1) / kSelThreads;
    const int b = tid * L, e = min(HW, b + L);
    int mine = 0;
    for (int i = b; i < e; ++i) mine += (sortable_key(sc[i]) == T);
    __syncthreads();
    int incl = block_inclusive_scan(mine, s_warp);
    int rank = incl - mine;
    for (int i = b; i < e && rank < need; ++i) {
      if (sortable_key(sc[i]) == T) { s_list[n_greater + rank] = pack_entry(T, i); ++rank; }
    }
  }
  __syncthreads();
}

// descending bitonic sort of n (power of two) 64-bit keys in shared memory; `pay` (optional) is permuted alongside
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* keys, uint32_t* pay, int n, int tid) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (n >> 1); t += kSelThreads) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool desc = ((lo & size) == 0);
        unsigned long long a = keys[lo], b = keys[hi];
        if ((a < b) == desc) {
          keys[lo] = b; keys[hi] = a;
          if (pay != nullptr) { uint32_t t2 = pay[lo]; pay[lo] = pay[hi]; pay[hi] = t2; }
        }
      }
      __syncthreads();
    }
  }
}

// Out-of-place descending rank sort of n <= kSelThreads DISTINCT 64-bit keys (the flat index is part of the key):
// rank(i) = #{j : keys[j] > keys[i]}.  1024/pow2(n) threads (at most 32) share one element; one barrier instead of the
// ~log2(n)^2/2 barriers of the bitonic network, which dominated the select kernel for the usual n of 100..256.
template <typename PAY>
__device__ __forceinline__ void rank_sort_desc(const unsigned long long* keys, const PAY* pay, int n,
                                               unsigned long long* out_keys, PAY* out_pay, int tid) {
  int np = 1;
  while (np < n) np <<= 1;
  int tpe = kSelThreads / np;                     // threads per element (power of two)
  if (tpe > 32) tpe = 32;
  const int i = tid / tpe, part = tid & (tpe - 1);
  const bool act = i < n;
  const unsigned long long mine = act ? keys[i] : 0ull;
  int cnt = 0;
  for (int j = part; j < n; j += tpe) cnt += (keys[j] > mine) ? 1 : 0;
  for (int o = tpe >> 1; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (act && part == 0) {
    out_keys[cnt] = mine;
    if (pay != nullptr) out_pay[cnt] = pay[i];
  }
  __syncthreads();
}

// CACHE: H*W <= 16384 and a multiple of 4 - every thread keeps its 16 candidates in registers (loaded once, before
// the histogram scan, so the L2 latency hides behind it); otherwise the passes re-read the candidate map from L2.
template <bool CACHE>
__global__ void __launch_bounds__(kSelThreads)
select_gather_kernel(DecodeParams p) {
  __shared__ unsigned long long s_list[kListCap];
  __shared__ __align__(16) int s_hist[kHistBins];      // first level: 4096 bins of cand_bin; reused (2048 entries) by the refinement / fallback
  __shared__ int s_warp[32];
  __shared__ int s_scalars[3];
  __shared__ int s_n;
  // class group of every collected candidate (CACHE path): carried through the sort so that the label recovery does not
  // have to fetch it from global memory first (one dependent round trip less)
  __shared__ uint8_t s_grp_in[kListCap];
  __shared__ uint8_t s_grp_sorted[kSelThreads];

  const int n = blockIdx.x;
  const int tid = threadIdx.x;
  const int HW = p.H * p.W;
  *reinterpret_cast<int4*>(&s_hist[4 * tid]) = make_int4(0, 0, 0, 0);
  // launched with programmatic stream serialization: this grid may start while the peaks kernel drains; nothing the
  // peaks kernel wrote may be read before it has completed and flushed
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const float* sc = p.cscore + (size_t)n * HW;
  const int k = p.k;
  const bool vec4 = (HW & 3) == 0;
  const float4* sc4 = reinterpret_cast<const float4*>(sc);
  const int n_vec = vec4 ? (HW >> 2) : 0;

  float4 cache[4];
  uint32_t cgrp[4];
  if (CACHE) {
    const uint32_t* cg32 = reinterpret_cast<const uint32_t*>(p.cgroup + (size_t)n * HW);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = tid + q * kSelThreads;
      cache[q] = (i < n_vec) ? sc4[i] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);   // -inf sorts below every candidate
      cgrp[q] = (i < n_vec) ? cg32[i] : 0u;
    }
  }
  // visit(f): f(key, flat_index, valid) for every candidate of this thread, called the SAME number of times by every
  // thread of the CTA (valid = false pads the tail) so that f may use warp collectives
  auto visit = [&](auto&& f) {
    if (CACHE) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = tid + q * kSelThreads;
        const float fv[4] = {cache[q].x, cache[q].y, cache[q].z, cache[q].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) f(sortable_key(fv[j]), 4 * i + j, i < n_vec, (cgrp[q] >> (8 * j)) & 0xffu);
      }
    } else {
      // larger maps: 4 independent 16-byte loads per thread and round trip (the candidate map is L2-resident; one load
      // per trip made every pass cost H*W/4096 L2 latencies)
      for (int i0 = 0; i0 < n_vec; i0 += 4 * kSelThreads) {
        float4 v4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = i0 + q * kSelThreads + tid;
          v4[q] = (i < n_vec) ? sc4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = i0 + q * kSelThreads + tid;
          const float fv[4] = {v4[q].x, v4[q].y, v4[q].z, v4[q].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) f(sortable_key(fv[j]), 4 * i + j, i < n_vec, 0xffu);
        }
      }
      for (int i0 = 4 * n_vec; i0 < HW; i0 += kSelThreads) {
        const int i = i0 + tid;
        const bool ok = i < HW;
        f(sortable_key(ok ? sc[i] : 0.f), i, ok, 0xffu);
      }
    }
  };
  // warp-aggregated append to s_list: one shared-memory atomic per warp and step instead of one per element
  auto append = [&](bool pred, unsigned long long entry, uint32_t grp) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return;
    const int lane = tid & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&s_n, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) {
      const int pos = base + __popc(m & ((1u << lane) - 1u));
      s_list[pos] = entry;
      if (CACHE) s_grp_in[pos] = (uint8_t)grp;
    }
  };

  // collect(pred): append every candidate whose key satisfies pred to s_list.  Winners are ~1 % of the candidates, so the
  // 16 candidates a thread holds are first only counted (one compare each); a warp scan of the counts gives every lane its
  // write position behind ONE shared-memory atomic per warp and chunk.
  auto collect_chunk = [&](const uint32_t (&u)[16], const int (&idx)[16], const bool (&ok)[16], const uint32_t (&grp)[16], auto&& pred) {
    int c = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) c += (ok[q] && pred(u[q])) ? 1 : 0;
    if (!__any_sync(0xffffffffu, c != 0)) return;
    const int lane = tid & 31;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    int base = 0;
    if (lane == 31) base = atomicAdd(&s_n, incl);
    base = __shfl_sync(0xffffffffu, base, 31);
    int pos = base + incl - c;
    if (c == 0) return;
#pragma unroll
    for (int q = 0; q < 16; ++q)
      if (ok[q] && pred(u[q])) {
        s_list[pos] = pack_entry(u[q], idx[q]);
        if (CACHE) s_grp_in[pos] = (uint8_t)grp[q];
        ++pos;
      }
  };
  auto collect = [&](auto&& pred) {
    uint32_t u[16], grp[16];
    int idx[16];
    bool ok[16];
    if (CACHE) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = tid + q * kSelThreads;
        const float fv[4] = {cache[q].x, cache[q].y, cache[q].z, cache[q].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          u[4 * q + j] = sortable_key(fv[j]); idx[4 * q + j] = 4 * i + j; ok[4 * q + j] = i < n_vec; grp[4 * q + j] = (cgrp[q] >> (8 * j)) & 0xffu;
        }
      }
      collect_chunk(u, idx, ok, grp, pred);
    } else {
      for (int i0 = 0; i0 < n_vec; i0 += 4 * kSelThreads) {        // uniform trip count
        float4 v4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = i0 + q * kSelThreads + tid;
          v4[q] = (i < n_vec) ? sc4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = i0 + q * kSelThreads + tid;
          const float fv[4] = {v4[q].x, v4[q].y, v4[q].z, v4[q].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) { u[4 * q + j] = sortable_key(fv[j]); idx[4 * q + j] = 4 * i + j; ok[4 * q + j] = i < n_vec; grp[4 * q + j] = 0xffu; }
        }
        collect_chunk(u, idx, ok, grp, pred);
      }
      for (int i0 = 4 * n_vec; i0 < HW; i0 += kSelThreads) {         // maps whose size is not a multiple of 4 (small, generic kernel)
        const int i = i0 + tid;
        const bool in = i < HW;
        const uint32_t uu = sortable_key(in ? sc[i] : 0.f);
        append(in && pred(uu), pack_entry(uu, i), 0xffu);
      }
    }
  };
  // smallest sortable key of a histogram bin (bin 0 holds everything below 2^-16, negative inputs included)
  auto bin_floor_key = [](uint32_t bin) -> uint32_t { return bin == 0 ? 0u : (0x80000000u | ((bin + kBinBase) << 15)); };

  // ---- histogram of the image's candidates (shared memory), then the bin of the k-th largest (thread t owns the 4 bins
  //      4092-4t .. 4095-4t).  Pixels without a peak (probability 0, often the majority) are counted per warp, not per pixel.
  __syncthreads();                                   // s_hist was zeroed before the dependency wait
  {
    int zeros = 0;
    visit([&](uint32_t u, int i, bool ok, uint32_t g) {
      if (!ok) return;
      if (u == 0x80000000u) ++zeros; else atomicAdd(&s_hist[cand_bin(u)], 1);
    });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    if ((tid & 31) == 0 && zeros) atomicAdd(&s_hist[kNoPeakBin], zeros);
  }
  __syncthreads();
  const int4 h4 = *reinterpret_cast<const int4*>(&s_hist[kHistBins - 4 - 4 * tid]);
  int cnt[4] = {h4.w, h4.z, h4.y, h4.x};            // descending bin order
  int sum4 = cnt[0] + cnt[1] + cnt[2] + cnt[3];
  const unsigned long long* s_sorted = s_list;
  bool have_grp = false;                             // s_grp_sorted[j] = class group of winner j
  if (k > kMaxK) {
    // More winners than the shared-memory lists hold (the reference accepts any k <= H*W, centernet.py:259): sort ALL
    // candidates of the image in a global buffer (L2-resident; one CTA, so __syncthreads orders the passes).  Rare path.
    unsigned long long* buf = p.sortbuf + (size_t)n * p.sort_cap;
    for (int i = tid; i < p.sort_cap; i += kSelThreads) buf[i] = (i < HW) ? pack_entry(sortable_key(sc[i]), i) : 0ull;
    __syncthreads();
    bitonic_sort_desc(buf, nullptr, p.sort_cap, tid);
    s_sorted = buf;
  } else {
  int incl = block_inclusive_scan(sum4, s_warp);
  int run = incl - sum4;
  if (run < k && k <= incl) {                    // exactly one thread: the k-th largest lies in one of its bins
    int b = 0;
    while (run + cnt[b] < k) { run += cnt[b]; ++b; }
    s_scalars[0] = kHistBins - 1 - 4 * tid - b;  // bin
    s_scalars[1] = run + cnt[b];                 // candidates in this bin or above
    s_n = 0;
  }
  __syncthreads();
  const uint32_t bin_k = (uint32_t)s_scalars[0];
  const int n_in_or_above = s_scalars[1];
  __syncthreads();                                  // everyone has read s_scalars before it is reused
  int n_sort;
  bool done = false;
  if (n_in_or_above <= kRefineAbove) {
    // few enough: collect every candidate in bin_k or above and sort them all
    const uint32_t floor_k = bin_floor_key(bin_k);
    collect([&](uint32_t u) { return u >= floor_k; });
    n_sort = n_in_or_above;
    done = true;
  } else if (bin_k > 0 && bin_k < (uint32_t)(kHistBins - 1)) {
    // Scores crowd into bin_k: refine with the next 11 key bits (the clamped end bins span several exponents and go to the
    // exact radix select below instead).  One pass
    // collects everything above bin_k and histograms the members of bin_k; a second pass collects the members
    // of bin_k at or above the sub-bin that holds the k-th largest.
    for (int i = tid; i < kBins; i += kSelThreads) s_hist[i] = 0;
    __syncthreads();
    visit([&](uint32_t u, int i, bool ok, uint32_t g) {
      const uint32_t b = cand_bin(u);
      append(ok && b > bin_k, pack_entry(u, i), g);
      if (ok && b == bin_k) atomicAdd(&s_hist[(u >> kSubShift) & (kBins - 1)], 1);
    });
    __syncthreads();
    const int n_above = s_n;                         // < k by construction
    const int need = k - n_above;
    const int b_hi = s_hist[kBins - 1 - 2 * tid];
    const int b_lo = s_hist[kBins - 2 - 2 * tid];
    const int incl2 = block_inclusive_scan(b_hi + b_lo, s_warp);
    const int excl2 = incl2 - (b_hi + b_lo);
    if (excl2 < need && need <= incl2) {
      if (need <= excl2 + b_hi) { s_scalars[0] = kBins - 1 - 2 * tid; s_scalars[1] = excl2 + b_hi; }
      else                      { s_scalars[0] = kBins - 2 - 2 * tid; s_scalars[1] = incl2; }
    }
    __syncthreads();
    const uint32_t sub_k = (uint32_t)s_scalars[0];
    const int n_total = n_above + s_scalars[1];
    if (n_total <= kListCap) {
      // members of bin_k at or above sub-bin sub_k: one key interval (an un-clamped bin is one exponent, so its keys order
      // like their mantissa bits)
      const uint32_t lo_k = bin_floor_key(bin_k) | (sub_k << kSubShift), hi_k = bin_floor_key(bin_k + 1);
      collect([&](uint32_t u) { return u >= lo_k && u < hi_k; });
      n_sort = n_total;
      done = true;
    }
    __syncthreads();
  }
  if (!done) {                                       // large plateaus of equal scores: exact radix select
    radix_select_fallback(sc, HW, k, s_list, s_hist, s_warp, s_scalars);
    n_sort = k;
  }
  // ---- sort, descending -------------------------------------------------------------------------------------
  __syncthreads();
  if (n_sort <= kSelThreads) {                       // the usual case: k .. a few hundred collected candidates
    have_grp = CACHE && done;
    rank_sort_desc<uint8_t>(s_list, have_grp ? s_grp_in : nullptr, n_sort, s_list + kSelThreads, s_grp_sorted, tid);
    s_sorted = s_list + kSelThreads;
  } else {
    int kp = 1;
    while (kp < n_sort) kp <<= 1;
    for (int i = n_sort + tid; i < kp; i += kSelThreads) s_list[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc(s_list, nullptr, kp, tid);
  }
  }   // k <= kMaxK

  // ---- label recovery, box decode and output of the k winners ---------------------------------------------------------------
  // Reference: labels = argmax over classes of the MASKED map (first maximal class).  For a winner with best value v
  // that is the first class c with value(c) == v that is a kxk peak.  8 lanes share one winner: each loads C/8
  // classes (independent loads, one latency round trip); a second look at the 3x3 window is needed only when
  // several classes hold exactly the same value.
  const size_t plane = (size_t)HW;
  const float* img = p.heat + (size_t)n * p.C * plane;
  const int sub = tid & 7;
  for (int j0 = 0; j0 < k; j0 += kSelThreads / 8) {
    const int j = j0 + (tid >> 3);
    const bool act = j < k;
    int idx = 0;
    float v = 0.f;
    if (act) {
      const unsigned long long w = s_sorted[j];
      idx = (int)(0xffffffffu - (uint32_t)(w & 0xffffffffull));
      v = key_to_float((uint32_t)(w >> 32));
    }
    int first = 0x7fffffff;
    const bool trivial = (v == 0.0f);                                        // all-zero column: argmax is class 0
    // kernel 1 recorded which class group attained the maximum: only those classes are re-read (the C planes of one
    // pixel are 4*H*W bytes apart - same DRAM bank - so every class read costs a row activation)
    int c_lo = 0, c_hi = p.C;
    if (act && p.group_classes > 0) {
      const int gid = have_grp ? (int)s_grp_sorted[j] : (int)p.cgroup[(size_t)n * HW + idx];
      if (gid != 0xff) { c_lo = gid * p.group_classes; c_hi = min(p.C, c_lo + p.group_classes); }    // 0xff: group unknown, scan all
    }
    // classes to re-read: warp-uniform trip count (the 8 lanes of a winner exchange their results by shuffles)
    const int scan_len = __reduce_max_sync(0xffffffffu, act ? (c_hi - c_lo) : 0);
    // the winners' box-map values travel with their class values (independent loads, same round trip)
    float braw = 0.f;
    if (p.box != nullptr && act && sub < 4) braw = __ldg(p.box + ((size_t)n * 4 + sub) * plane + idx);
    const bool multi_batch = scan_len > 8 * kLabelBatch;
    for (int off = 0; off < scan_len; off += 8 * kLabelBatch) {              // uniform trip count for the whole CTA
      float xv[kLabelBatch];
      int nmatch = 0;
#pragma unroll
      for (int q = 0; q < kLabelBatch; ++q) {                                 // batch the loads: one latency round trip
        const int c = c_lo + off + sub + 8 * q;
        xv[q] = (act && !trivial && c < c_hi) ? __ldg(img + (size_t)c * plane + idx) : NAN;
      }
      // from_logits: the candidate is a probability, the map holds logits - a class matches when its logit rounds to
      // that probability (distinct logits can: see the semantics note at the top)
      bool hit[kLabelBatch];
#pragma unroll
      for (int q = 0; q < kLabelBatch; ++q) {
        const int c = c_lo + off + sub + 8 * q;
        hit[q] = act && !trivial && c < c_hi && ((p.from_logits ? sigmoid32(xv[q]) : xv[q]) == v);   // (padding lanes hold NaN)
        nmatch += hit[q] ? 1 : 0;
      }
      int total = nmatch;                                                     // matches among the 8 lanes of this winner
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
#pragma unroll
      for (int q = 0; q < kLabelBatch; ++q) {
        const int c = c_lo + off + sub + 8 * q;
        if (!hit[q] || c >= first) continue;
        bool peak = true;                                                     // a unique match IS the peak that produced v
        if (total > 1 || multi_batch) {
          const int y = idx / p.W, xx0 = idx - y * p.W;
          float m = xv[q];
          for (int yy = max(0, y - p.P); yy <= min(p.H - 1, y + p.P); ++yy)
            for (int xx = max(0, xx0 - p.P); xx <= min(p.W - 1, xx0 + p.P); ++xx) {
              m = fmaxf(m, __ldg(img + (size_t)c * plane + (size_t)yy * p.W + xx));
            }
          peak = (m == xv[q]) || (p.from_logits && sigmoid32(m) == v);      // kept iff its probability is the window's maximum
        }
        if (peak) first = c;
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    // lanes 0..3 of the winner hold its four box-map values: hand them to lane 0, which decodes and writes the row
    const int seg = (tid & 31) & ~7;
    const float g0 = __shfl_sync(0xffffffffu, braw, seg), g1 = __shfl_sync(0xffffffffu, braw, seg + 1);
    const float g2 = __shfl_sync(0xffffffffu, braw, seg + 2), g3 = __shfl_sync(0xffffffffu, braw, seg + 3);
    if (act && sub == 0) {
      long long label = (first == 0x7fffffff) ? 0 : (long long)first;
      if (v == 0.0f) label = 0;                            // underflowed / zero candidates arg-max to class 0
      const size_t o = (size_t)n * k + j;                  // the winners are in canonical order (probability desc, index asc)
      p.scores[o] = v;
      p.indices[o] = idx;
      p.labels[o] = label;
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.box != nullptr) {
        const float g[4] = {g0, g1, g2, g3};
        b4 = decode_box_from(g, idx, p.H, p.W, p.normalize, p.box_log, p.mult, p.stride_f);
        *reinterpret_cast<float4*>(p.boxes + o * 4) = b4;
      }
      if (p.packed != nullptr) {                           // row for the cross-rank all_gather (bit-exact integer lanes)
        float* row = p.packed + o * (size_t)p.packed_w;
        *reinterpret_cast<float4*>(row) = b4;
        *reinterpret_cast<float4*>(row + 4) = make_float4(v, __int_as_float(idx), __int_as_float((int)(label & 0xffffffffll)),
                                                          __int_as_float((int)(label >> 32)));
      }
    }
  }
  if (p.reid != nullptr) {                    // fairmot.py:63-73: emb[n, j, e] = reid[n, e, idx_j]
    const int E = p.E;
    for (int t = tid; t < k * E; t += kSelThreads) {
      int j = t / E, e = t - j * E;
      int idx = (int)(0xffffffffu - (uint32_t)(s_sorted[j] & 0xffffffffull));
      const float val = __ldg(p.reid + ((size_t)n * E + e) * plane + idx);
      p.emb[((size_t)n * k + j) * E + e] = val;
      if (p.packed != nullptr) p.packed[((size_t)n * k + j) * p.packed_w + 8 + e] = val;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------------
// Strip geometry.  R = rows per strip (the 2P halo rows are re-read by the neighbouring strips through L2: 50 % extra
// L2 traffic at R = 4, 25 % at R = 8), VEC = float4 loads per lane and row (warp tile = 128*VEC columns; rows of up to
// 256 columns then need no halo columns from a neighbouring tile).  CNL_PEAKS_R / CNL_PEAKS_VEC override the choice.
template <int P, bool LOGITS, int R, int VEC>
static int launch_fast_rv(const float* heat, float* cscore, uint8_t* cgroup, int N, int C, int H, int W,
                          cudaStream_t st) {
  dim3 grid((W + kTW * VEC - 1) / (kTW * VEC), (H + R - 1) / R, N);
  const bool mt = grid.x > 1;        // rows wider than one warp tile need halo columns from neighbours
#define CNL_LAUNCH_PEAKS(G_)                                                                                      \
  do {                                                                                                            \
    if (mt) peaks_fast_kernel<P, LOGITS, R, G_, true, VEC><<<grid, G_ * 32, 0, st>>>(heat, cscore, cgroup, C, H, W);  \
    else if (P == 1 && VEC == 1 && R == 4 && W == kTW)                                                             \
      peaks_fast_kernel<P, LOGITS, R, G_, false, VEC, (P == 1 && VEC == 1 && R == 4) ? kTW : 0><<<grid, G_ * 32, 0, st>>>(heat, cscore, cgroup, C, H, W); \
    else    peaks_fast_kernel<P, LOGITS, R, G_, false, VEC><<<grid, G_ * 32, 0, st>>>(heat, cscore, cgroup, C, H, W); \
  } while (0)
  int groups = 1;
  if (C >= 32) { CNL_LAUNCH_PEAKS(4); groups = 4; }
  else if (C >= 2) { CNL_LAUNCH_PEAKS(2); groups = 2; }
  else CNL_LAUNCH_PEAKS(1);
#undef CNL_LAUNCH_PEAKS
  return (C + groups - 1) / groups;          // classes per group, for the label recovery
}

template <int P, bool LOGITS>
static int launch_fast(const float* heat, float* cscore, uint8_t* cgroup, int N, int C, int H, int W,
                       cudaStream_t st) {
  static const int env_r = getenv("CNL_PEAKS_R") ? atoi(getenv("CNL_PEAKS_R")) : 0;
  static const int env_v = getenv("CNL_PEAKS_VEC") ? atoi(getenv("CNL_PEAKS_VEC")) : 0;
  int vec = (W > kTW && W <= 2 * kTW) ? 2 : 1;
  if (env_v == 1 || env_v == 2) vec = env_v;
  int rows = 4;
  if (env_r == 4 || env_r == 8) rows = env_r;
  if constexpr (P == 1) {
    if (vec == 2) return launch_fast_rv<P, LOGITS, 4, 2>(heat, cscore, cgroup, N, C, H, W, st);
    if (rows == 8) return launch_fast_rv<P, LOGITS, 8, 1>(heat, cscore, cgroup, N, C, H, W, st);
  }
  return launch_fast_rv<P, LOGITS, 4, 1>(heat, cscore, cgroup, N, C, H, W, st);
}

// returns the number of classes per recorded class group (0 = no group information, scan every class)
template <bool LOGITS>
static int launch_peaks(const float* heat, float* cscore, uint8_t* cgroup, int N, int C, int H, int W,
                        int P, bool force_generic, cudaStream_t st) {
  bool fast = !force_generic && (W % 4 == 0) && P <= 2 && ((reinterpret_cast<uintptr_t>(heat) & 15) == 0);
  if (fast) {
    switch (P) {
      case 0: return launch_fast<0, LOGITS>(heat, cscore, cgroup, N, C, H, W, st);
      case 1: return launch_fast<1, LOGITS>(heat, cscore, cgroup, N, C, H, W, st);
      case 2: return launch_fast<2, LOGITS>(heat, cscore, cgroup, N, C, H, W, st);
    }
  }
  dim3 grid((W + 31) / 32, (H + 7) / 8, N);
  peaks_generic_kernel<LOGITS><<<grid, 256, 0, st>>>(heat, cscore, cgroup, C, H, W, P);
  return 0;
}

static size_t sort_capacity(int h, int w) { size_t c = 1; while (c < (size_t)h * w) c <<= 1; return c; }
static size_t hist_bytes(int n) { return align_up((size_t)n * kHistBins * sizeof(unsigned int), 256); }
static size_t score_bytes(int n, int h, int w) { return align_up((size_t)n * h * w * sizeof(float), 256); }

}  // namespace cnl

using namespace cnl;

extern "C" {

const char* cnl_last_error(void) { return error_buffer(); }
int cnl_version(void) { return 1002; }   // 1.2: cnl_conv_desc kinds 2-4 (depthwise, fuse, 3x3 stem) + trailing fields, relu = 2 (ReLU6)
int cnl_compiled_sm(void) { return 100; }

size_t cnl_decode_workspace_bytes(int n, int h, int w) {
  if (n <= 0 || h <= 0 || w <= 0) return 0;
  return hist_bytes(n) + score_bytes(n, h, w) + align_up((size_t)n * h * w, 256);
}

size_t cnl_decode_workspace_bytes_k(int n, int h, int w, int num_detections) {
  const size_t base = cnl_decode_workspace_bytes(n, h, w);
  if (base == 0 || num_detections <= kMaxK) return base;
  return base + (size_t)n * sort_capacity(h, w) * sizeof(unsigned long long);     // global sort buffer of the large-k path
}

int cnl_decode_detections_packed(const float* heatmap, const float* box_offsets, const float* reid,
                                 int n, int c, int h, int w, int reid_dim,
                                 int from_logits, int nms_kernel, int num_detections,
                                 int normalize_boxes, int box_log, float box_multiplier, int stride,
                                 float* boxes, float* scores, int64_t* labels, int64_t* indices, float* embeddings,
                                 float* packed, int packed_width,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!heatmap || !scores || !labels || !indices || !workspace)
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: null pointer argument");
  if ((box_offsets != nullptr) != (boxes != nullptr))
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: box_offsets and boxes must be given together");
  if (n <= 0 || c <= 0 || h <= 0 || w <= 0)
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: bad shape (%d,%d,%d,%d)", n, c, h, w);
  if (c > 65535) return fail(CNL_ERR_UNSUPPORTED, "cnl_decode_detections: more than 65535 classes");
  if ((long long)h * w > (1ll << 30)) return fail(CNL_ERR_UNSUPPORTED, "cnl_decode_detections: map too large");
  if (from_logits & ~(1 | CNL_DECODE_WORKSPACE_CLEAN | CNL_DECODE_PEAKS_ONLY)) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: bad from_logits flags %d", from_logits);
  const bool workspace_clean = (from_logits & CNL_DECODE_WORKSPACE_CLEAN) != 0;
  const bool peaks_only = (from_logits & CNL_DECODE_PEAKS_ONLY) != 0;
  from_logits &= 1;
  // a negative or even kernel changes the pooled map's size in the reference (F.max_pool2d would not broadcast)
  int force_generic = 0;
  if (nms_kernel < 0) { force_generic = 1; nms_kernel = -nms_kernel; }   // test hook: negative selects the generic kernel
  if (nms_kernel < 1 || nms_kernel > 7 || (nms_kernel % 2) == 0)
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: nms_kernel must be odd and in 1..7 (got %d)", nms_kernel);
  if (num_detections < 1 || (long long)num_detections > (long long)h * w)
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: num_detections=%d must be in 1..H*W=%d (torch.topk raises)",
                num_detections, h * w);
  if (packed != nullptr && (packed_width != 8 + (reid ? reid_dim : 0) || (packed_width & 3) || (reinterpret_cast<uintptr_t>(packed) & 15)))
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: packed rows are 8 + reid_dim floats (a multiple of 4), 16-byte aligned");
  if ((reid != nullptr) != (embeddings != nullptr) || (reid && reid_dim <= 0))
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: reid, embeddings and reid_dim must be given together");
  if (workspace_bytes < cnl_decode_workspace_bytes_k(n, h, w, num_detections))
    return fail(CNL_ERR_WORKSPACE, "cnl_decode_detections: workspace %zu < %zu bytes (cnl_decode_workspace_bytes_k)", workspace_bytes,
                cnl_decode_workspace_bytes_k(n, h, w, num_detections));
  if (reinterpret_cast<uintptr_t>(workspace) & 255)
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: workspace must be 256-byte aligned");
  if (reinterpret_cast<uintptr_t>(boxes) & 15)
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: boxes must be 16-byte aligned");

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* cscore = reinterpret_cast<float*>(static_cast<char*>(workspace) + hist_bytes(n));
  const int P = (nms_kernel - 1) / 2;
  uint8_t* cgroup = reinterpret_cast<uint8_t*>(static_cast<char*>(workspace) + hist_bytes(n) + score_bytes(n, h, w));
  const int group_classes = from_logits ? launch_peaks<true>(heatmap, cscore, cgroup, n, c, h, w, P, force_generic, st)
                                        : launch_peaks<false>(heatmap, cscore, cgroup, n, c, h, w, P, force_generic, st);
  CNL_CUDA_CHECK(cudaGetLastError());
  if (peaks_only) return CNL_OK;

  DecodeParams p;
  p.cscore = cscore; p.box = box_offsets; p.reid = reid;
  p.heat = heatmap; p.C = c; p.P = P; p.from_logits = from_logits ? 1 : 0;
  p.cgroup = cgroup; p.group_classes = group_classes;
  p.H = h; p.W = w; p.E = reid_dim; p.k = num_detections;
  p.normalize = normalize_boxes; p.box_log = box_log; p.mult = box_multiplier; p.stride_f = (float)stride;
  p.boxes = boxes; p.scores = scores; p.labels = reinterpret_cast<long long*>(labels);
  p.indices = reinterpret_cast<long long*>(indices); p.emb = embeddings;
  p.packed = packed; p.packed_w = packed_width;
  p.sortbuf = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + cnl_decode_workspace_bytes(n, h, w));
  p.sort_cap = (int)sort_capacity(h, w);
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n);
    cfg.blockDim = dim3(kSelThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // overlap this launch with the peaks kernel's tail
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if ((h * w) % 4 == 0 && h * w <= 4 * 4 * kSelThreads) CNL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, select_gather_kernel<true>, p));
    else                                                  CNL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, select_gather_kernel<false>, p));
  }
  CNL_CUDA_CHECK(cudaGetLastError());
  return CNL_OK;
}

int cnl_decode_detections(const float* heatmap, const float* box_offsets, const float* reid,
                          int n, int c, int h, int w, int reid_dim,
                          int from_logits, int nms_kernel, int num_detections,
                          int normalize_boxes, int box_log, float box_multiplier, int stride,
                          float* boxes, float* scores, int64_t* labels, int64_t* indices, float* embeddings,
                          void* workspace, size_t workspace_bytes, void* stream) {
  return cnl_decode_detections_packed(heatmap, box_offsets, reid, n, c, h, w, reid_dim, from_logits, nms_kernel, num_detections,
                                      normalize_boxes, box_log, box_multiplier, stride, boxes, scores, labels, indices, embeddings,
                                      nullptr, 0, workspace, workspace_bytes, stream);
}

int cnl_boxes_xyxy_to_xywh(const float* boxes_xyxy, float* boxes_xywh, size_t n_boxes, void* stream) {
  if (!boxes_xyxy || !boxes_xywh) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_boxes_xyxy_to_xywh: null pointer argument");
  if ((reinterpret_cast<uintptr_t>(boxes_xyxy) | reinterpret_cast<uintptr_t>(boxes_xywh)) & 15)
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_boxes_xyxy_to_xywh: boxes must be 16-byte aligned");
  if (n_boxes == 0) return CNL_OK;
  xyxy_to_xywh_kernel<<<(unsigned)((n_boxes + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(boxes_xyxy), reinterpret_cast<float4*>(boxes_xywh), n_boxes);
  CNL_CUDA_CHECK(cudaGetLastError());
  return CNL_OK;
}

int cnl_sigmoid(const float* in, float* out, size_t n, void* stream) {
  if (!in || !out) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_sigmoid: null pointer argument");
  if (n == 0) return CNL_OK;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
  sigmoid_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, n);
  CNL_CUDA_CHECK(cudaGetLastError());
  return CNL_OK;
}

int cnl_gather_boxes(const float* box_offsets, const int64_t* indices, int n, int h, int w, int k,
                     int normalize_boxes, int box_log, float box_multiplier, int stride, float* boxes, void* stream) {
  if (!box_offsets || !indices || !boxes) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_gather_boxes: null pointer argument");
  if (n <= 0 || h <= 0 || w <= 0 || k <= 0) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_gather_boxes: bad shape");
  if (reinterpret_cast<uintptr_t>(boxes) & 15) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_gather_boxes: boxes must be 16-byte aligned");
  int total = n * k;
  gather_boxes_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      box_offsets, reinterpret_cast<const long long*>(indices), n, h, w, k, normalize_boxes, box_log, box_multiplier,
      (float)stride, boxes);
  CNL_CUDA_CHECK(cudaGetLastError());
  return CNL_OK;
}

}  // extern "C"
