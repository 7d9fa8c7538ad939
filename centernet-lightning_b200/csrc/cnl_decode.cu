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
#include <cstring>
#include "cnl_common.h"

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

// The logistic the from_logits path specifies: three separately rounded fp32 operations.
__device__ __forceinline__ float sigmoid32(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// Logits closer than this (or above kSatLogit) may round to the same fp32 probability; only then is the
// probability-space comparison of the reference evaluated explicitly.  d(sigmoid)/dx >= 2^-24*(1+e^x) bounds it.
constexpr float kNearTie = 5e-5f;
constexpr float kSatLogit = 4.0f;

struct Cand {           // running per-pixel winner across classes
  float v;              // LOGITS: best peak logit (-inf = none yet).  PROBS: best heatmap*mask value so far.
  int label;
};

// x: centre value, m: kxk window max (m >= x).  Reference: mask = (maxpool(p) == p); p*mask; max over classes (first wins).
template <bool LOGITS>
__device__ __forceinline__ void update_cand(Cand& s, float x, float m, int c) {
  if (LOGITS) {
    bool peak = (x == m);
    if (!peak) {
      float d = m - x;
      if (d < kNearTie || x > kSatLogit) peak = (sigmoid32(x) == sigmoid32(m));   // fp32 probability plateau
    }
    if (peak) {
      bool take = x > s.v;
      if (take && s.v != -INFINITY) {
        float e = x - s.v;
        if (e < kNearTie || s.v > kSatLogit) take = sigmoid32(x) > sigmoid32(s.v);
      }
      if (take) { s.v = x; s.label = c; }
    }
  } else {
    float cand = (x == m) ? x : 0.0f;
    if (cand > s.v) { s.v = cand; s.label = c; }
  }
}

// Merge a later class group into an earlier one (groups are visited in class order, so "first max wins" holds).
template <bool LOGITS>
__device__ __forceinline__ void merge_cand(Cand& s, const Cand& o) {
  if (LOGITS) {
    if (o.v == -INFINITY) return;
    bool take = o.v > s.v;
    if (take && s.v != -INFINITY) {
      float e = o.v - s.v;
      if (e < kNearTie || s.v > kSatLogit) take = sigmoid32(o.v) > sigmoid32(s.v);
    }
    if (take) s = o;
  } else {
    if (o.v > s.v) s = o;
  }
}

template <bool LOGITS>
__device__ __forceinline__ void finish_cand(const Cand& s, float& score, int& label) {
  if (LOGITS) {
    score = sigmoid32(s.v);                 // -inf (no peak at this pixel) -> exactly 0
    label = (score == 0.0f) ? 0 : s.label;  // an all-zero column arg-maxes to class 0 in the reference
  } else {
    score = s.v;
    label = s.label;
  }
}

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

template <int P, bool LOGITS, int R, int G>
__global__ void __launch_bounds__(G * 32)
peaks_fast_kernel(const float* __restrict__ heat, float* __restrict__ cscore, uint16_t* __restrict__ clabel,
                  int C, int H, int W) {
  constexpr int ROWS = R + 2 * P;
  const int lane = threadIdx.x & 31;
  const int g = threadIdx.x >> 5;
  const int n = blockIdx.z;
  const int r0 = blockIdx.y * R;
  const int x0 = blockIdx.x * kTW + lane * 4;
  const bool col_ok = x0 < W;
  const bool multi_tile = gridDim.x > 1;
  const size_t plane = (size_t)H * W;
  const int cg = (C + G - 1) / G;
  const int c_begin = g * cg;
  const int c_end = min(C, c_begin + cg);
  const float NEG = -INFINITY;

  Cand st[R][4];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { st[i][j].v = NEG; st[i][j].label = 0; }

  const float* img = heat + (size_t)n * C * plane;
  for (int c = c_begin; c < c_end; ++c) {
    const float* pl = img + (size_t)c * plane;
    float4 v[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      int r = r0 - P + j;
      v[j] = (col_ok && r >= 0 && r < H) ? ld_stream4(pl + (size_t)r * W + x0) : make_float4(NEG, NEG, NEG, NEG);
    }
    // halo columns owned by neighbouring column tiles (only when a row is wider than one warp tile)
    constexpr int PH = (P > 0) ? P : 1;
    float hl[ROWS][PH], hr[ROWS][PH];
#pragma unroll
    for (int j = 0; j < ROWS; ++j)
#pragma unroll
      for (int q = 0; q < PH; ++q) { hl[j][q] = NEG; hr[j][q] = NEG; }
    if constexpr (P > 0) {
      if (multi_tile && (lane == 0 || lane == 31)) {
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
          int r = r0 - P + j;
          if (r >= 0 && r < H) {
#pragma unroll
            for (int q = 0; q < P; ++q) {
              int xl = x0 - 1 - q, xr = x0 + 4 + q;
              if (lane == 0 && xl >= 0) hl[j][q] = __ldg(pl + (size_t)r * W + xl);
              if (lane == 31 && xr < W) hr[j][q] = __ldg(pl + (size_t)r * W + xr);
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
      // vertical max over the window rows
      float4 vm = v[i];
      float vl[PH], vr[PH];
#pragma unroll
      for (int q = 0; q < PH; ++q) { vl[q] = hl[i][q]; vr[q] = hr[i][q]; }
#pragma unroll
      for (int j = 1; j <= 2 * P; ++j) {
        vm = max4(vm, v[i + j]);
#pragma unroll
        for (int q = 0; q < PH; ++q) { vl[q] = fmaxf(vl[q], hl[i + j][q]); vr[q] = fmaxf(vr[q], hr[i + j][q]); }
      }
      // horizontal: e[0..P-1] left neighbours (nearest last), e[P..P+3] own, e[P+4..] right neighbours
      float e[4 + 2 * P];
      e[P + 0] = vm.x; e[P + 1] = vm.y; e[P + 2] = vm.z; e[P + 3] = vm.w;
      if constexpr (P > 0) {
        const float own[4] = {vm.x, vm.y, vm.z, vm.w};
#pragma unroll
        for (int q = 0; q < P; ++q) {
          // q-th column to the left of x0 is component (3-q) of lane-1; to the right of x0+3 it is component q of lane+1
          float fl = __shfl_up_sync(0xffffffffu, own[3 - q], 1);
          float fr = __shfl_down_sync(0xffffffffu, own[q], 1);
          e[P - 1 - q] = (lane == 0) ? vl[q] : fl;
          e[P + 4 + q] = (lane == 31) ? vr[q] : fr;
        }
      }
      const float ctr[4] = {v[i + P].x, v[i + P].y, v[i + P].z, v[i + P].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float m = e[j];
#pragma unroll
        for (int q = 1; q <= 2 * P; ++q) m = fmaxf(m, e[j + q]);
        update_cand<LOGITS>(st[i][j], ctr[j], m, c);
      }
    }
  }

  // merge the G class groups (in class order) through shared memory, then emit one candidate per pixel
  __shared__ float s_v[G][R][kTW];
  __shared__ uint16_t s_l[G][R][kTW];
  if (G > 1) {
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s_v[g][i][lane * 4 + j] = st[i][j].v;
        s_l[g][i][lane * 4 + j] = (uint16_t)st[i][j].label;
      }
    __syncthreads();
  }
  for (int i = g; i < R; i += G) {          // warp g finishes rows g, g+G, ...
    int r = r0 + i;
    if (r >= H || !col_ok) continue;
    float sc[4];
    int lb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      Cand a;
      if (G > 1) {
        a.v = s_v[0][i][lane * 4 + j]; a.label = s_l[0][i][lane * 4 + j];
#pragma unroll
        for (int gg = 1; gg < G; ++gg) {
          Cand o; o.v = s_v[gg][i][lane * 4 + j]; o.label = s_l[gg][i][lane * 4 + j];
          merge_cand<LOGITS>(a, o);
        }
      } else {
        a = st[i][j];
      }
      finish_cand<LOGITS>(a, sc[j], lb[j]);
    }
    size_t o = (size_t)n * plane + (size_t)r * W + x0;
    *reinterpret_cast<float4*>(cscore + o) = make_float4(sc[0], sc[1], sc[2], sc[3]);
    ushort4 l4 = make_ushort4((uint16_t)lb[0], (uint16_t)lb[1], (uint16_t)lb[2], (uint16_t)lb[3]);
    *reinterpret_cast<ushort4*>(clabel + o) = l4;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Kernel 1b: generic peaks kernel (any W, any odd window up to 7): one thread per pixel.
// ------------------------------------------------------------------------------------------------------------
template <bool LOGITS>
__global__ void __launch_bounds__(256)
peaks_generic_kernel(const float* __restrict__ heat, float* __restrict__ cscore, uint16_t* __restrict__ clabel,
                     int C, int H, int W, int P) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int n = blockIdx.z;
  if (x >= W || y >= H) return;
  const size_t plane = (size_t)H * W;
  const float* img = heat + (size_t)n * C * plane;
  Cand s; s.v = -INFINITY; s.label = 0;
  const int y_lo = max(0, y - P), y_hi = min(H - 1, y + P);
  const int x_lo = max(0, x - P), x_hi = min(W - 1, x + P);
  for (int c = 0; c < C; ++c) {
    const float* pl = img + (size_t)c * plane;
    float ctr = __ldg(pl + (size_t)y * W + x);
    float m = ctr;
    for (int yy = y_lo; yy <= y_hi; ++yy)
      for (int xx = x_lo; xx <= x_hi; ++xx) m = fmaxf(m, __ldg(pl + (size_t)yy * W + xx));
    update_cand<LOGITS>(s, ctr, m, c);
  }
  float score; int label;
  finish_cand<LOGITS>(s, score, label);
  size_t o = (size_t)n * plane + (size_t)y * W + x;
  cscore[o] = score;
  clabel[o] = (uint16_t)label;
}

// ------------------------------------------------------------------------------------------------------------
// Kernel 2: per-image exact top-k (radix select + ordered ties + bitonic sort) and gather/decode.
// ------------------------------------------------------------------------------------------------------------
constexpr int kSelThreads = 1024;
constexpr int kMaxK = 1024;
constexpr int kBins = 2048;

__device__ __forceinline__ uint32_t sortable_key(float f) {       // larger float <=> larger key (NaN-free input)
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}

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
__device__ __forceinline__ float4 decode_box(const float* box_img, size_t plane, int idx, int H, int W,
                                             int normalize, int box_log, float mult, float stride_f) {
  const float* bx = box_img + idx;
  float cx = __fadd_rn((float)(idx % W), 0.5f);
  float cy = __fadd_rn((float)(idx / W), 0.5f);
  float s[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float gv = __ldg(bx + c * plane);
    if (box_log) gv = expf(gv);
    gv = __fmul_rn(gv, mult);
    s[c] = fmaxf(gv, 0.0f);
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

struct DecodeParams {
  const float* cscore; const uint16_t* clabel;
  const float* box; const float* reid;
  int H, W, E, k;
  int normalize, box_log; float mult; float stride_f;
  float* boxes; float* scores; long long* labels; long long* indices; float* emb;
};

__global__ void __launch_bounds__(kSelThreads)
select_gather_kernel(DecodeParams p) {
  __shared__ int s_hist[kBins];
  __shared__ int s_warp[32];
  __shared__ unsigned long long s_list[kMaxK];
  __shared__ int s_bin, s_above, s_cnt;

  const int n = blockIdx.x;
  const int tid = threadIdx.x;
  const int HW = p.H * p.W;
  const float* sc = p.cscore + (size_t)n * HW;
  const int k = p.k;

  // ---- 3-pass MSB radix select of the k-th largest key --------------------------------------------------
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
    // suffix counts from the top bin: thread t owns reversed bins 2t, 2t+1  (bin = kBins-1-rev)
    int b_hi = s_hist[kBins - 1 - 2 * tid];
    int b_lo = s_hist[kBins - 2 - 2 * tid];
    int incl = block_inclusive_scan(b_hi + b_lo, s_warp);
    int excl = incl - (b_hi + b_lo);
    if (excl < need && need <= incl) {            // exactly one thread
      if (need <= excl + b_hi) { s_bin = kBins - 1 - 2 * tid; s_above = excl; s_cnt = b_hi; }
      else                     { s_bin = kBins - 2 - 2 * tid; s_above = excl + b_hi; s_cnt = b_lo; }
    }
    __syncthreads();
    need -= s_above;
    prefix |= ((uint32_t)s_bin) << shift;
    mask |= dm << shift;
    count_at_T = s_cnt;
    __syncthreads();
  }
  const uint32_t T = prefix;              // key of the k-th largest candidate
  const int n_greater = k - need;         // all keys > T are selected; `need` (>=1) of the keys == T

  // ---- collect winners as (key << 32 | ~index): descending order of this word = score desc, index asc ----
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  const bool all_ties_taken = (count_at_T == need);
  for (int i = tid; i < HW; i += kSelThreads) {
    uint32_t u = sortable_key(sc[i]);
    if (u > T || (all_ties_taken && u == T)) {
      int pos = atomicAdd(&s_cnt, 1);
      s_list[pos] = ((unsigned long long)u << 32) | (uint32_t)(0xffffffffu - (uint32_t)i);
    }
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
      if (sortable_key(sc[i]) == T) {
        s_list[n_greater + rank] = ((unsigned long long)T << 32) | (uint32_t)(0xffffffffu - (uint32_t)i);
        ++rank;
      }
    }
  }
  int kp = 1;
  while (kp < k) kp <<= 1;
  for (int i = k + tid; i < kp; i += kSelThreads) s_list[i] = 0ull;
  __syncthreads();

  // ---- bitonic sort, descending ---------------------------------------------------------------------------
  for (int size = 2; size <= kp; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (kp >> 1); t += kSelThreads) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool desc = ((lo & size) == 0);
        unsigned long long a = s_list[lo], b = s_list[hi];
        if ((a < b) == desc) { s_list[lo] = b; s_list[hi] = a; }
      }
      __syncthreads();
    }
  }

  // ---- gather + decode (reference centernet.py:278-303: every op rounded separately) ---------------------
  const size_t plane = (size_t)HW;
  for (int j = tid; j < k; j += kSelThreads) {
    unsigned long long w = s_list[j];
    int idx = (int)(0xffffffffu - (uint32_t)(w & 0xffffffffull));
    float score = key_to_float((uint32_t)(w >> 32));
    size_t o = (size_t)n * k + j;
    p.scores[o] = score;
    p.indices[o] = idx;
    p.labels[o] = p.clabel[(size_t)n * HW + idx];
    if (p.box == nullptr) continue;
    float4 b4 = decode_box(p.box + (size_t)n * 4 * plane, plane, idx, p.H, p.W, p.normalize, p.box_log, p.mult, p.stride_f);
    *reinterpret_cast<float4*>(p.boxes + o * 4) = b4;
  }
  if (p.reid != nullptr) {                    // fairmot.py:63-73: emb[n, j, e] = reid[n, e, idx_j]
    const int E = p.E;
    for (int t = tid; t < k * E; t += kSelThreads) {
      int j = t / E, e = t - j * E;
      int idx = (int)(0xffffffffu - (uint32_t)(s_list[j] & 0xffffffffull));
      p.emb[((size_t)n * k + j) * E + e] = __ldg(p.reid + ((size_t)n * E + e) * plane + idx);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------------
template <int P, bool LOGITS>
static void launch_fast(const float* heat, float* cscore, uint16_t* clabel, int N, int C, int H, int W, cudaStream_t st) {
  constexpr int R = 4;
  dim3 grid((W + kTW - 1) / kTW, (H + R - 1) / R, N);
  if (C >= 32) {
    peaks_fast_kernel<P, LOGITS, R, 4><<<grid, 4 * 32, 0, st>>>(heat, cscore, clabel, C, H, W);
  } else if (C >= 2) {
    peaks_fast_kernel<P, LOGITS, R, 2><<<grid, 2 * 32, 0, st>>>(heat, cscore, clabel, C, H, W);
  } else {
    peaks_fast_kernel<P, LOGITS, R, 1><<<grid, 32, 0, st>>>(heat, cscore, clabel, C, H, W);
  }
}

template <bool LOGITS>
static void launch_peaks(const float* heat, float* cscore, uint16_t* clabel, int N, int C, int H, int W, int P,
                         bool force_generic, cudaStream_t st) {
  bool fast = !force_generic && (W % 4 == 0) && P <= 2 && ((reinterpret_cast<uintptr_t>(heat) & 15) == 0);
  if (fast) {
    switch (P) {
      case 0: launch_fast<0, LOGITS>(heat, cscore, clabel, N, C, H, W, st); return;
      case 1: launch_fast<1, LOGITS>(heat, cscore, clabel, N, C, H, W, st); return;
      case 2: launch_fast<2, LOGITS>(heat, cscore, clabel, N, C, H, W, st); return;
    }
  }
  dim3 grid((W + 31) / 32, (H + 7) / 8, N);
  peaks_generic_kernel<LOGITS><<<grid, 256, 0, st>>>(heat, cscore, clabel, C, H, W, P);
}

static size_t score_bytes(int n, int h, int w) { return align_up((size_t)n * h * w * sizeof(float), 256); }

}  // namespace cnl

using namespace cnl;

extern "C" {

const char* cnl_last_error(void) { return error_buffer(); }
int cnl_version(void) { return 1000; }
int cnl_compiled_sm(void) { return 100; }

size_t cnl_decode_workspace_bytes(int n, int h, int w) {
  if (n <= 0 || h <= 0 || w <= 0) return 0;
  return score_bytes(n, h, w) + align_up((size_t)n * h * w * sizeof(uint16_t), 256);
}

int cnl_decode_detections(const float* heatmap, const float* box_offsets, const float* reid,
                          int n, int c, int h, int w, int reid_dim,
                          int from_logits, int nms_kernel, int num_detections,
                          int normalize_boxes, int box_log, float box_multiplier, int stride,
                          float* boxes, float* scores, int64_t* labels, int64_t* indices, float* embeddings,
                          void* workspace, size_t workspace_bytes, void* stream) {
  if (!heatmap || !scores || !labels || !indices || !workspace)
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: null pointer argument");
  if ((box_offsets != nullptr) != (boxes != nullptr))
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: box_offsets and boxes must be given together");
  if (n <= 0 || c <= 0 || h <= 0 || w <= 0)
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: bad shape (%d,%d,%d,%d)", n, c, h, w);
  if (c > 65535) return fail(CNL_ERR_UNSUPPORTED, "cnl_decode_detections: more than 65535 classes");
  if ((long long)h * w > (1ll << 30)) return fail(CNL_ERR_UNSUPPORTED, "cnl_decode_detections: map too large");
  // a negative or even kernel changes the pooled map's size in the reference (F.max_pool2d would not broadcast)
  int force_generic = 0;
  if (nms_kernel < 0) { force_generic = 1; nms_kernel = -nms_kernel; }   // test hook: negative selects the generic kernel
  if (nms_kernel < 1 || nms_kernel > 7 || (nms_kernel % 2) == 0)
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: nms_kernel must be odd and in 1..7 (got %d)", nms_kernel);
  if (num_detections < 1 || (long long)num_detections > (long long)h * w)
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: num_detections=%d must be in 1..H*W=%d (torch.topk raises)",
                num_detections, h * w);
  if (num_detections > kMaxK)
    return fail(CNL_ERR_UNSUPPORTED, "cnl_decode_detections: num_detections=%d exceeds %d", num_detections, kMaxK);
  if ((reid != nullptr) != (embeddings != nullptr) || (reid && reid_dim <= 0))
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: reid, embeddings and reid_dim must be given together");
  if (workspace_bytes < cnl_decode_workspace_bytes(n, h, w))
    return fail(CNL_ERR_WORKSPACE, "cnl_decode_detections: workspace %zu < %zu bytes", workspace_bytes,
                cnl_decode_workspace_bytes(n, h, w));
  if (reinterpret_cast<uintptr_t>(workspace) & 255)
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: workspace must be 256-byte aligned");
  if (reinterpret_cast<uintptr_t>(boxes) & 15)
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_decode_detections: boxes must be 16-byte aligned");

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* cscore = static_cast<float*>(workspace);
  uint16_t* clabel = reinterpret_cast<uint16_t*>(static_cast<char*>(workspace) + score_bytes(n, h, w));
  const int P = (nms_kernel - 1) / 2;
  if (from_logits) launch_peaks<true>(heatmap, cscore, clabel, n, c, h, w, P, force_generic, st);
  else             launch_peaks<false>(heatmap, cscore, clabel, n, c, h, w, P, force_generic, st);
  CNL_CUDA_CHECK(cudaGetLastError());

  DecodeParams p;
  p.cscore = cscore; p.clabel = clabel; p.box = box_offsets; p.reid = reid;
  p.H = h; p.W = w; p.E = reid_dim; p.k = num_detections;
  p.normalize = normalize_boxes; p.box_log = box_log; p.mult = box_multiplier; p.stride_f = (float)stride;
  p.boxes = boxes; p.scores = scores; p.labels = reinterpret_cast<long long*>(labels);
  p.indices = reinterpret_cast<long long*>(indices); p.emb = embeddings;
  select_gather_kernel<<<n, kSelThreads, 0, st>>>(p);
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
