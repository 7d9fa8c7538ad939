// sm_100a forward engine: backbone -> neck -> heads as a list of fused implicit-GEMM convolutions.
//
// Replaces GenericModel.forward (reference centernet_lightning/models/meta.py:41-47) and the vision_toolbox
// backbone / FPN / ConvBnAct modules it calls (meta.py:9-10, 21-30, 87-96).
//
// Data layout in HBM
//   activations : NHWC, fp16.  CNL_PRECISION_SPLIT keeps two planes per tensor, [plane][N][H][W][C]:
//                 plane 0 = hi = fp16(x), plane 1 = lo = fp16(x - hi); x ~ hi + lo carries ~22 mantissa bits.
//   weights     : [plane][tap][Cout_pad][Cin] fp16 (K-major rows for the UMMA B operand), BatchNorm folded,
//                 scaled by a per-op power of two so that the lo parts stay in fp16's normal range.
//   head outputs: (N,C,H,W) fp32, the layout the decode kernel and the reference API use.
//
// conv_tc_kernel (one launch per ConvOp, persistent, 1 CTA/SM, 192 threads)
//   GEMM view: M = 128 output pixels (a th x tw patch of one image), N = Cout tile (<= 256), K = taps x Cin.
//   warp 0    : TMA producer.  For every (tap, 64-channel block) one 4-D tiled TMA load brings the shifted input
//               patch [th][tw][64ch] straight into the 128B-swizzled K-major A tile (out-of-bounds rows/columns are
//               zero-filled by the TMA unit = conv padding; stride-2 convs use the tensor map's element strides),
//               and one 3-D load brings the [Cout_tile][64] weight tile.  mbarrier ring of `num_stages` stages.
//   warp 1    : allocates 512 TMEM columns (2 x 256-column fp32 accumulators) and issues tcgen05.mma
//               (cta_group::1, kind::f16, M=128, N=Cout_tile, K=16).  SPLIT precision issues hi*hi, hi*lo, lo*hi
//               into the same accumulator.  tcgen05.commit frees smem stages and publishes finished accumulators.
//   warps 2-5 : epilogue.  tcgen05.ld the accumulator (one pixel row per thread), scale + bias (+ residual,
//               optionally nearest-upsampled for the FPN top-down add) + ReLU, split into hi/lo fp16, stage in
//               128B-swizzled smem and TMA-store NHWC; head outputs are written directly as fp32 NCHW.
//               The double-buffered accumulator lets the epilogue of tile i overlap the MMAs of tile i+1.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "cnl_common.h"
#include "cnl_tcgen05.cuh"

namespace cnl {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                  // 64 fp16 = one 128-byte swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2;
constexpr int kMaxStages = 8;
constexpr int kConvThreads = 320;           // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int kEpiWarps = 8;
constexpr int kSmemLimit = 232448 - 1024;    // 227 KB opt-in shared memory per CTA minus the static part (barriers)
constexpr int kStageWarpBytes = 32 * 64;     // epilogue staging: 32 pixels x 32 channels fp16 per warp and plane
constexpr int kParamBias = 1024;             // bias values carried in the kernel parameters (constant bank)

struct ConvParams {
  int n_img, out_h, out_w;
  int tw, th, tiles_w, tiles_h, m_tiles;
  int n_tiles, n_tile;
  int kh, kw, stride, pad_h, pad_w, kblocks;
  int src_c_off, dst_c_off;
  int relu;
  int sigmoid;                  // out_mode 1 only: store 1/(1+exp(-x)) (the `.sigmoid()` the reference applies to the heat map, fused)
  int out_mode;                 // 0 = NHWC fp16 planes via TMA store, 1 = NCHW fp32 direct stores
  int cout_real;
  float wscale_inv;
  const float* bias;            // [n_tiles * n_tile]
  float* out_nchw;
  const __half* res;            // residual NHWC planes (or nullptr)
  int res_up, res_c, res_h, res_w;
  long long res_plane_elems;
  int num_stages;
  int acc_stages;               // TMEM accumulator stages (2 = the epilogue of tile i overlaps the MMAs of tile i+1)
  int corr_off;                 // column offset of the correction accumulator inside an accumulator stage (CORR only)
  int store_w, store_h;         // per-warp TMA store box in pixels (store_w * store_h == 32)
  int cluster;                  // CTAs per cluster sharing one multicast weight tile (1 = no cluster)
  int cat;                      // CORR with Cout tile <= 128: hi*[hi|lo] issued as ONE MMA of N = 2*n_tile (see the MMA warp)
  int dst_up;                   // 2: the destination has twice the conv's output resolution (5-D store map, see below)
  int dst_phase;                // dst_up == 2: -1 = write every pixel to its whole 2x2 block (conv + nearest x2 upsample),
                                //              0..3 = write only sub-pixel (py, px) = (phase >> 1, phase & 1)
  int dst_c;                    // channels of the destination buffer (dst_up == 2: sub-pixel px is folded into dim 0)
  // ROWS mode (see conv_tc_kernel): input rows live in a ring of shared-memory slots and are reused by every tap
  int a_slots;                  // row slots in the ring (>= kh + 1)
  int a_slot_bytes;             // bytes of one slot and plane: (tw + kw - 1) pixels x 128 B, rounded up to 1024
  int box_w;                    // pixels per loaded row: tw + kw - 1
  int stage_depth;              // epilogue staging buffers per warp (ring): TMA stores of the last depth-1 chunks may still be in flight
  int w_resident;               // ROWS: every tap's weight tile stays in its own stage for the whole launch (num_stages == taps)
  // The per-channel bias travels in the kernel parameters: the epilogue reads it through the constant cache (one broadcast
  // LDC per value).  As __ldg loads from global memory the same values missed L1 on every tile - 226 KB of the SM's 256 KB
  // are shared memory, and the epilogue's own stores / residual loads evict the three bias lines - and 30-35 % of all warp
  // stall samples of the bandwidth-bound ops (head out-convs, FPN laterals) sat on those loads (ncu source view, r2).
  // Direct epilogue stores (CNL_DIRECT_STORE=1, off by default): NHWC outputs go from registers to global memory as 256-bit
  // stores instead of smem staging + TMA stores.  Background (r2, CNL_DEBUG_EPI experiments): NOT issuing the epilogue's
  // stores makes the whole forward 19 % faster (13.8 -> 11.15 ms) - their waits and the proxy fence cost nothing, the
  // stores themselves do.  Direct stores measured slower still (14.65 ms).
  int direct_store;
  int early_release;            // single accumulator stage, Cout tile 256: drain TMEM into registers, release it, THEN convert / store
  __half* dst;                  // destination buffer (plane 0), its plane stride in elements and its geometry
  long long dst_plane_elems;
  int dst_h, dst_w;
  int debug;                    // CNL_DEBUG_EPI bit mask, timing experiments only (WRONG results): 1 = epilogue skips its stores
  int bias_in_params;           // 0: more than kParamBias channels, read p.bias from global memory instead
  float bias_c[kParamBias];
};

// ------------------------------------------------------------------------------------------------------------
// The implicit-GEMM convolution kernel
// ------------------------------------------------------------------------------------------------------------
// CORR = true (CNL_PRECISION_SPLIT): the hi*lo and lo*hi correction products accumulate in their own TMEM accumulator
// (p.corr_off columns after the main one) and are added to the hi*hi sum in the epilogue.  The tensor core truncates
// (RZ) after every accumulate step, which biases long sums; keeping the 2^-11-times-smaller correction stream out of
// the main accumulator cuts the number of roundings at full magnitude by 3x.  With Cout tiles of 256 it costs the
// accumulator double buffering (512 TMEM columns); narrower tiles keep two stages.
//
// Epilogue: 8 warps.  Warps w and w+4 own the same TMEM lane quarter (w & 3) and split the tile's columns in halves,
// so two tcgen05.ld -> convert -> store chains per scheduler overlap their latencies.
//
// ROWS = true ("row-rolling", stride-1 convs with Cin = 64 on maps at least 65 pixels wide, i.e. one output row of 128
// pixels per tile): the im2col view re-reads every input pixel once per tap - 9x for a 3x3 conv - and for Cin = Cout = 64
// that L2 -> shared-memory traffic (432 KB per 128-pixel tile), not the tensor pipe or HBM, bounds the kernel.  Here a
// CTA owns a run of consecutive output rows of one image column strip and keeps the last kh input rows (each tw + kw - 1
// pixels, both planes) in a ring of shared-memory slots; every output row needs ONE new input row.  The A operand of
// tap (r, s) is a shifted view of a slot: start address = slot of input row (h + r - pad_h) + s pixels x 128 B - the
// 128-byte swizzle is a function of the shared-memory address bits, so a view that starts s rows into a swizzle atom reads
// exactly what TMA wrote there.  Only the weight tiles still stream through the stage ring.
//
// PAIR = true (Cout tile 256, split precision): the two CTAs of a cluster form a tcgen05 CTA pair (cta_group::2).  Each
// CTA loads its own 128-pixel A tile and HALF of the weight tile (128 of the 256 Cout rows); the leader CTA issues
// M = 256 instructions that read A and B from both CTAs' shared memory and write each CTA's 128 accumulator rows into
// its own TMEM.  Per k-step a CTA then takes in 64 KB instead of 96 KB - the im2col kernel with full weight tiles sits
// at ~52 B/clk of shared-memory fill per SM, which is what capped its tensor pipe at ~83 % - and three stages fit
// instead of two.  Barrier protocol: both CTAs' TMA loads credit the LEADER's full barrier; the leader's
// tcgen05.commit multicasts to both CTAs' empty / accumulator-full barriers; the peer's epilogue warps arrive remotely on
// the leader's accumulator-empty barrier.
template <int NPLANE, bool CORR, bool ROWS, bool PAIR>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ CUtensorMap w_map,
               const __grid_constant__ CUtensorMap dst_map, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t a_full_bar[kMaxStages];       // ROWS: input-row slots
  __shared__ __align__(8) uint64_t a_empty_bar[kMaxStages];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

  const int b_tile_bytes = p.n_tile * kBlockK * 2;
  const int b_half_bytes = b_tile_bytes / 2;                            // PAIR: this CTA's half of the weight tile
  const int stage_bytes = ROWS ? NPLANE * b_tile_bytes                  // ROWS: stages hold weights only
                        : PAIR ? NPLANE * (kATileBytes + b_half_bytes)
                               : NPLANE * (kATileBytes + b_tile_bytes);
  const int a_ring_bytes = ROWS ? p.a_slots * NPLANE * p.a_slot_bytes : 0;
  uint8_t* a_ring = smem;                                               // ROWS: [a_slots][NPLANE][a_slot_bytes]
  uint8_t* stages = smem + a_ring_bytes;
  uint8_t* staging = stages + (size_t)p.num_stages * stage_bytes;      // [8 warps][NPLANE][2048], 1024-aligned
  const int taps = p.kh * p.kw;
  const int k_iters = taps * p.kblocks;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&src_map);
    ptx::prefetch_tmap(&w_map);
    if (p.out_mode == 0) ptx::prefetch_tmap(&dst_map);
    for (int i = 0; i < p.num_stages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], PAIR ? 1 : p.cluster); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tmem_full_bar[i], 1); ptx::mbar_init(&tmem_empty_bar[i], PAIR ? 2 * kEpiWarps : kEpiWarps); }
    if (ROWS) for (int i = 0; i < p.a_slots; ++i) { ptx::mbar_init(&a_full_bar[i], 1); ptx::mbar_init(&a_empty_bar[i], 1); }
    ptx::fence_barrier_init();
  }
  if (PAIR) ptx::cluster_sync_all();                // both CTAs are resident before the pair allocates tensor memory
  if (warp == 1) {
    if (PAIR) { ptx::tmem_alloc_pair(&tmem_base_smem, 512); ptx::tmem_relinquish_pair(); }
    else      { ptx::tmem_alloc(&tmem_base_smem, 512); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) ptx::cluster_sync_all();       // peers' barriers are initialised before any multicast / remote arrive
  ptx::tc_fence_after();
  // Programmatic dependent launch: this CTA became resident as soon as an SM of the previous launch drained, and its
  // set-up above (tensor-map prefetch, barrier init, TMEM allocation) overlapped that launch's tail.  Everything the
  // previous launch wrote (this op's input, residual) is visible only after the wait.  No activation buffer is ever
  // rewritten inside one forward, so there is no write-after-read hazard to order.
  ptx::grid_dependents_launch();
  ptx::grid_dependency_wait();
  const uint32_t tmem_base = tmem_base_smem;
  // Work distribution.  A cluster of `csize` CTAs takes `csize` consecutive pixel tiles of the SAME Cout tile, so the
  // weight tile is fetched from L2 once per cluster: CTA r loads rows [r, r+1) * n_tile/csize and multicasts them.
  // A pixel tile past the end (m_tiles not a multiple of csize) is a dummy: its image coordinate is out of bounds, so
  // TMA loads return zeros and stores are clipped, while the CTA keeps its part in the shared barrier protocol.
  const int csize = p.cluster;
  const int crank = (csize > 1) ? (int)ptx::cluster_ctarank() : 0;
  const int m_groups = (p.m_tiles + csize - 1) / csize;
  const int total_groups = m_groups * p.n_tiles;
  const int group0 = blockIdx.x / csize, group_step = gridDim.x / csize;
  const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const bool two_acc = p.acc_stages == 2;
  // ROWS: tiles are numbered (image, column strip, row) and every CTA takes one contiguous run of them
  const int rows_t0 = ROWS ? (int)((long long)p.m_tiles * blockIdx.x / gridDim.x) : 0;
  const int rows_t1 = ROWS ? (int)((long long)p.m_tiles * (blockIdx.x + 1) / gridDim.x) : 0;

  if (warp == 0) {
    // ===================================== TMA producer ==============================================
    if (ROWS && ptx::elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      int ld_slot = 0;                                     // ring position of the next input-row load and the parity of its
      uint32_t ld_par = 0;                                 // slot's barriers (counters instead of n_loads % / a_slots)
      const int row_bytes = p.box_w * kBlockK * 2;
      if (p.w_resident && rows_t0 < rows_t1) {
        // the whole weight tensor (taps x [hi|lo] tiles) fits beside the row ring: load it ONCE; the issue loop then never
        // waits for a weight tile again (per output row the stem re-fetched 64 KB of weights for 32 KB of input)
        for (int tap = 0; tap < taps; ++tap) {
          ptx::mbar_arrive_expect_tx(&full_bar[tap], stage_bytes);
          uint8_t* st = stages + (size_t)tap * stage_bytes;
#pragma unroll
          for (int pl = 0; pl < NPLANE; ++pl)
            ptx::tma_load_3d(st + pl * b_tile_bytes, &w_map, &full_bar[tap], 0, 0, tap + pl * taps);
        }
      }
      // (image, column strip, row) of the next tile, advanced by adds
      int strip = rows_t0 / p.tiles_h, h = rows_t0 - strip * p.tiles_h;
      int img = strip / p.tiles_w, w0 = (strip - img * p.tiles_w) * p.tw;
      for (int t = rows_t0; t < rows_t1; ++t) {
        const bool fresh = (t == rows_t0) || (h == 0);     // first output row of this CTA's run or of a new strip
        const int n_new = fresh ? p.kh : 1;
        for (int q = 0; q < n_new; ++q) {
          const int in_row = h - p.pad_h + (fresh ? q : p.kh - 1);
          const int slot = ld_slot;
          ptx::mbar_wait(&a_empty_bar[slot], ld_par ^ 1);
          ptx::mbar_arrive_expect_tx(&a_full_bar[slot], NPLANE * row_bytes);
#pragma unroll
          for (int pl = 0; pl < NPLANE; ++pl)
            ptx::tma_load_4d(a_ring + ((size_t)slot * NPLANE + pl) * p.a_slot_bytes, &src_map, &a_full_bar[slot], p.src_c_off,
                             w0 - p.pad_w, in_row, img + pl * p.n_img);
          if (++ld_slot == p.a_slots) { ld_slot = 0; ld_par ^= 1; }
        }
        for (int tap = 0; tap < taps && !p.w_resident; ++tap) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
          uint8_t* st = stages + (size_t)stage * stage_bytes;
#pragma unroll
          for (int pl = 0; pl < NPLANE; ++pl)
            ptx::tma_load_3d(st + pl * b_tile_bytes, &w_map, &full_bar[stage], 0, 0, tap + pl * taps);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
        if (++h == p.tiles_h) { h = 0; w0 += p.tw; if (w0 >= p.tiles_w * p.tw) { w0 = 0; ++img; } }
      }
    }
    if (PAIR && ptx::elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      const int b_rows = p.n_tile / 2;
      for (int grp = group0; grp < total_groups; grp += group_step) {
        const int n_idx = grp % p.n_tiles;
        int m_idx = (grp / p.n_tiles) * 2 + crank;
        int img = m_idx / tiles_per_img;
        m_idx -= img * tiles_per_img;
        if (img >= p.n_img) img = 1 << 20;                // dummy tile: every coordinate out of bounds
        const int h0 = (m_idx / p.tiles_w) * p.th;
        const int w0 = (m_idx % p.tiles_w) * p.tw;
        for (int tap = 0; tap < taps; ++tap) {
          const int r = tap / p.kw, sx = tap - r * p.kw;
          const int cw = w0 * p.stride + sx - p.pad_w;
          const int ch = h0 * p.stride + r - p.pad_h;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            // the leader's barrier collects the bytes of BOTH CTAs' loads for this stage
            const bool skip_a = (p.debug & 32) && sx != 0;          // timing experiment: what would sharing A across the kw taps buy?
            if (crank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * (stage_bytes - (skip_a ? NPLANE * kATileBytes : 0)));
            const uint32_t bar = ptx::mapa_u32(&full_bar[stage], 0);
            uint8_t* st = stages + (size_t)stage * stage_bytes;
#pragma unroll
            for (int pl = 0; pl < NPLANE; ++pl)
              if (!skip_a) ptx::tma_load_4d_pair(st + pl * kATileBytes, &src_map, bar, p.src_c_off + kb * kBlockK, cw, ch,
                                                 (((p.debug & 256) && img < p.n_img) ? (img & 1) : img) + pl * p.n_img);   // debug 256: L2-resident input (timing experiment)
#pragma unroll
            for (int pl = 0; pl < NPLANE; ++pl)
              ptx::tma_load_3d_pair(st + NPLANE * kATileBytes + pl * b_half_bytes, &w_map, bar, kb * kBlockK,
                                    n_idx * p.n_tile + crank * b_rows, tap + pl * taps);
            if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    if (!ROWS && !PAIR && ptx::elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      const int b_rows = p.n_tile / csize;
      for (int grp = group0; grp < total_groups; grp += group_step) {
        const int n_idx = grp % p.n_tiles;
        int m_idx = (grp / p.n_tiles) * csize + crank;
        int img = m_idx / tiles_per_img;
        m_idx -= img * tiles_per_img;
        if (img >= p.n_img) img = 1 << 20;                // dummy tile: every coordinate out of bounds
        const int h0 = (m_idx / p.tiles_w) * p.th;
        const int w0 = (m_idx % p.tiles_w) * p.tw;
        for (int tap = 0; tap < taps; ++tap) {
          const int r = tap / p.kw, s = tap - r * p.kw;
          const int cw = w0 * p.stride + s - p.pad_w;
          const int ch = h0 * p.stride + r - p.pad_h;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            ptx::mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
            uint8_t* st = stages + (size_t)stage * stage_bytes;
#pragma unroll
            for (int pl = 0; pl < NPLANE; ++pl)
              ptx::tma_load_4d(st + pl * kATileBytes, &src_map, &full_bar[stage], p.src_c_off + kb * kBlockK, cw, ch,
                               img + pl * p.n_img);
            if (csize == 1) {
#pragma unroll
              for (int pl = 0; pl < NPLANE; ++pl)
                ptx::tma_load_3d(st + NPLANE * kATileBytes + pl * b_tile_bytes, &w_map, &full_bar[stage], kb * kBlockK,
                                 n_idx * p.n_tile, tap + pl * taps);
            } else {
#pragma unroll
              for (int pl = 0; pl < NPLANE; ++pl)
                ptx::tma_load_3d_mc(st + NPLANE * kATileBytes + pl * b_tile_bytes + crank * b_rows * (kBlockK * 2), &w_map,
                                    &full_bar[stage], kb * kBlockK, n_idx * p.n_tile + crank * b_rows, tap + pl * taps, cmask);
            }
            if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =================================================
    if (ROWS && ptx::elect_one_sync()) {
      // row-rolling issue loop: always the cat form (N = 2*n_tile for hi x [hi|lo], N = n_tile for lo x hi).
      // ONE thread issues 8 small MMAs (N = 128 / 64: 64 / 32 tensor-pipe cycles each) per tap, so the scalar work between
      // them is what bounds this kernel: measured (ncu source view, r2) 181 SASS instructions per tap - two integer
      // divisions (tap / kw, load index % slots) and an ELECT / BRA.U.ANY loop around every uniform-datapath instruction of
      // the divergent `lane == 0` branch - kept the tensor pipe at 37 %.  Hence elect.sync, slot / parity counters instead
      // of div / mod, and descriptors advanced by adds.
      const uint32_t idesc = ptx::make_idesc_f16_m128((uint32_t)p.n_tile);
      const uint32_t idesc_cat = ptx::make_idesc_f16_m128((uint32_t)p.n_tile * 2u);
      const uint64_t desc_hi_bits = ptx::make_sw128_kmajor_desc(0);          // everything but the 14 address bits
      const uint32_t ring_addr = ptx::smem_u32(a_ring), stages_addr = ptx::smem_u32(stages);
      const uint32_t slot_stride = (uint32_t)(NPLANE * p.a_slot_bytes);
      int stage = 0;
      uint32_t phase = 0;
      int ld_slot = 0, first_slot = 0;                      // ring position of the next load / of input row h - pad_h
      uint32_t ld_par = 0;
      int h = rows_t0 % p.tiles_h;
      uint32_t acc_it = 0;
      for (int t = rows_t0; t < rows_t1; ++t, ++acc_it) {
        const bool fresh = (t == rows_t0) || (h == 0);
        const bool last = (t + 1 == rows_t1) || (h == p.tiles_h - 1);   // the rows in the ring die with this output row
        const int n_new = fresh ? p.kh : 1;
        const uint32_t as = two_acc ? (acc_it & 1u) : 0u;
        const uint32_t aphase = two_acc ? ((acc_it >> 1) & 1u) : (acc_it & 1u);
        ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        if (fresh) first_slot = ld_slot;
        else if (++first_slot == p.a_slots) first_slot = 0;
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        const uint32_t d_corr = d_tmem + p.corr_off;
        uint32_t first_mma = 0;                               // 0 only for the first MMA of the tile (overwrite the accumulator)
        int slot = first_slot;
        for (int r = 0; r < p.kh; ++r) {
          // The new input rows are the LAST n_new rows of the window.  Each is waited for only when its taps come up: in
          // steady state (one new row per output row) that is after the (kh-1)*kw taps that still read the old rows, which
          // doubles the time the row's TMA load has to land (the slot it refills is released after the r = 0 taps of the
          // PREVIOUS output row).  Waiting for it at the top of the tile stalled every output row of the stem and layer1.
          if (r >= p.kh - n_new) {
            ptx::mbar_wait(&a_full_bar[ld_slot], ld_par);
            if (++ld_slot == p.a_slots) { ld_slot = 0; ld_par ^= 1; }
            ptx::tc_fence_after();
          }
          const uint32_t a_row = ring_addr + (uint32_t)slot * slot_stride;
          for (int sx = 0; sx < p.kw; ++sx) {
            ptx::mbar_wait(&full_bar[stage], phase);           // (resident weights: phase stays 0, the wait falls through)
            ptx::tc_fence_after();
            // (descriptor base-offset field stays 0: measured on B200, the swizzle phase comes from the address bits)
            const uint32_t a_addr = a_row + (uint32_t)sx * (kBlockK * 2);
            const uint32_t b_addr = stages_addr + (uint32_t)stage * (uint32_t)stage_bytes;
            const uint64_t a_hi = desc_hi_bits | (uint64_t)((a_addr & 0x3ffffu) >> 4);
            const uint64_t a_lo = desc_hi_bits | (uint64_t)(((a_addr + (uint32_t)p.a_slot_bytes) & 0x3ffffu) >> 4);
            const uint64_t b_hi = desc_hi_bits | (uint64_t)((b_addr & 0x3ffffu) >> 4);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              const uint64_t koff = (uint64_t)((k * 32) >> 4);
              ptx::umma_f16(d_tmem, a_hi + koff, b_hi + koff, idesc_cat, first_mma);
              first_mma = 1;
              if (NPLANE == 2) ptx::umma_f16(d_corr, a_lo + koff, b_hi + koff, idesc, 1);
            }
            if (!p.w_resident) ptx::umma_commit(&empty_bar[stage]);
            if (++stage == p.num_stages) { stage = 0; if (!p.w_resident) phase ^= 1; }
          }
          // The oldest input row (h - pad_h) is read by the r = 0 taps only: hand its slot back as soon as they are issued, so
          // that the producer can refill it with the row the NEXT output row needs while the remaining taps still run.  A
          // ring of kh slots then suffices and the shared memory goes to weight stages instead.  At the end of a strip / run
          // the other rows die as well.
          if (r == 0 || last) ptx::umma_commit(&a_empty_bar[slot]);
          if (++slot == p.a_slots) slot = 0;
        }
        ptx::umma_commit(&tmem_full_bar[as]);
        if (++h == p.tiles_h) h = 0;
      }
    }
    if (PAIR && crank == 0 && ptx::elect_one_sync()) {
      // leader of the CTA pair: M = 256 (128 accumulator rows in each CTA's tensor memory), N = n_tile
      const uint32_t idesc = ptx::make_idesc_f16_m256((uint32_t)p.n_tile);
      int stage = 0;
      uint32_t phase = 0;
      int acc_it = 0;
      for (int grp = group0; grp < total_groups; grp += group_step, ++acc_it) {
        const int as = two_acc ? (acc_it & 1) : 0;
        const uint32_t aphase = two_acc ? ((acc_it >> 1) & 1) : (acc_it & 1);
        ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1);       // the epilogue warps of BOTH CTAs have drained their halves
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        const uint32_t d_corr = CORR ? (d_tmem + p.corr_off) : d_tmem;
        for (int it = 0; it < k_iters; ++it) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(stages + (size_t)stage * stage_bytes);
          const uint32_t b_addr = a_addr + NPLANE * kATileBytes;
          const uint64_t a_hi = ptx::make_sw128_kmajor_desc(a_addr);
          const uint64_t a_lo = ptx::make_sw128_kmajor_desc(a_addr + kATileBytes);
          const uint64_t b_hi = ptx::make_sw128_kmajor_desc(b_addr);
          const uint64_t b_lo = ptx::make_sw128_kmajor_desc(b_addr + b_half_bytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t koff = (uint64_t)((k * 32) >> 4);
            ptx::umma_f16_pair(d_tmem, a_hi + koff, b_hi + koff, idesc, (it | k) != 0);
            if (NPLANE == 2) {
              ptx::umma_f16_pair(d_corr, a_hi + koff, b_lo + koff, idesc, CORR ? (uint32_t)((it | k) != 0) : 1u);
              ptx::umma_f16_pair(d_corr, a_lo + koff, b_hi + koff, idesc, 1);
            }
          }
          ptx::umma_commit_pair(&empty_bar[stage], 0b11);       // the stage is reusable in both CTAs
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_pair(&tmem_full_bar[as], 0b11);        // both CTAs' epilogues may read their accumulator rows
      }
    }
    if (!ROWS && !PAIR && ptx::elect_one_sync()) {
      const uint32_t idesc = ptx::make_idesc_f16_m128((uint32_t)p.n_tile);
      // cat: the hi and lo weight tiles sit back to back in the stage (n_tile rows of 128 B each, whole swizzle atoms),
      // so A_hi x [B_hi | B_lo] is ONE instruction of N = 2*n_tile whose columns [n_tile, 2*n_tile) are the correction
      // accumulator; only A_lo x B_hi is left as a second instruction.  The A tile is streamed from shared memory twice
      // instead of three times per K step - narrow tiles (N = 64) are bound by exactly that operand bandwidth.
      const uint32_t idesc_cat = ptx::make_idesc_f16_m128((uint32_t)p.n_tile * 2u);
      const bool cat = CORR && p.cat;
      int stage = 0;
      uint32_t phase = 0;
      int acc_it = 0;
      for (int grp = group0; grp < total_groups; grp += group_step, ++acc_it) {
        const int as = two_acc ? (acc_it & 1) : 0;
        const uint32_t aphase = two_acc ? ((acc_it >> 1) & 1) : (acc_it & 1);
        ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        const uint32_t d_corr = CORR ? (d_tmem + p.corr_off) : d_tmem;
        for (int it = 0; it < k_iters; ++it) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(stages + (size_t)stage * stage_bytes);
          const uint32_t b_addr = a_addr + NPLANE * kATileBytes;
          const uint64_t a_hi = ptx::make_sw128_kmajor_desc(a_addr);
          const uint64_t b_hi = ptx::make_sw128_kmajor_desc(b_addr);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t koff = (uint64_t)((k * 32) >> 4);        // +32 bytes along K inside the swizzled row
            if (NPLANE == 2 && cat) {
              const uint64_t a_lo = ptx::make_sw128_kmajor_desc(a_addr + kATileBytes);
              ptx::umma_f16(d_tmem, a_hi + koff, b_hi + koff, idesc_cat, (it | k) != 0);
              ptx::umma_f16(d_corr, a_lo + koff, b_hi + koff, idesc, 1);
              continue;
            }
            ptx::umma_f16(d_tmem, a_hi + koff, b_hi + koff, idesc, (it | k) != 0);
            if (NPLANE == 2) {
              const uint64_t a_lo = ptx::make_sw128_kmajor_desc(a_addr + kATileBytes);
              const uint64_t b_lo = ptx::make_sw128_kmajor_desc(b_addr + b_tile_bytes);
              ptx::umma_f16(d_corr, a_hi + koff, b_lo + koff, idesc, CORR ? (uint32_t)((it | k) != 0) : 1u);
              ptx::umma_f16(d_corr, a_lo + koff, b_hi + koff, idesc, 1);
            }
          }
          if (csize == 1) ptx::umma_commit(&empty_bar[stage]);          // smem stage reusable once these MMAs retire
          else            ptx::umma_commit_mc(&empty_bar[stage], cmask); // ... in every CTA of the cluster (shared weight tile)
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tmem_full_bar[as]);           // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================================== epilogue (warps 2..9) =====================================
    const int q = warp & 3;                              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                    // which half of the tile's columns this warp drains
    const int lane_base = q * 32;
    uint8_t* const my_stage_base = staging + (size_t)(warp - 2) * p.stage_depth * NPLANE * kStageWarpBytes;
    int sbuf = 0;                                       // ring position inside this warp's staging buffers
    int acc_it = 0;
    const int it_begin = ROWS ? rows_t0 : group0, it_end = ROWS ? rows_t1 : total_groups, it_step = ROWS ? 1 : group_step;
    // this thread's pixel and this warp's first pixel inside a tile: the same for every tile (the short-K tiles of the stem /
    // layer1 run one tile per ~1500-3500 tensor-pipe cycles, where five integer divisions per tile in every epilogue thread showed)
    const int pix = lane_base + lane;
    const int pix_dh = pix / p.tw, pix_dw = pix - pix_dh * p.tw;
    const int lb_dh = lane_base / p.tw, lb_dw = lane_base - lb_dh * p.tw;
    int r_strip = ROWS ? it_begin / p.tiles_h : 0;
    int r_h = ROWS ? it_begin - r_strip * p.tiles_h : 0;
    int r_img = ROWS ? r_strip / p.tiles_w : 0;
    int r_w0 = ROWS ? (r_strip - r_img * p.tiles_w) * p.tw : 0;
    for (int grp = it_begin; grp < it_end; grp += it_step, ++acc_it) {
      const int as = two_acc ? (acc_it & 1) : 0;
      const uint32_t aphase = two_acc ? ((acc_it >> 1) & 1) : (acc_it & 1);
      int n_idx, img, h0, w0;
      bool dummy = false;
      if (ROWS) {
        n_idx = 0;
        h0 = r_h; img = r_img; w0 = r_w0;                  // th == 1; advanced by adds at the end of the iteration
        if (++r_h == p.tiles_h) { r_h = 0; r_w0 += p.tw; if (r_w0 >= p.tiles_w * p.tw) { r_w0 = 0; ++r_img; } }
      } else {
        n_idx = grp % p.n_tiles;
        int m_idx = (grp / p.n_tiles) * csize + crank;
        img = m_idx / tiles_per_img;
        m_idx -= img * tiles_per_img;
        dummy = img >= p.n_img;
        if (dummy) img = 1 << 20;
        h0 = (m_idx / p.tiles_w) * p.th;
        w0 = (m_idx % p.tiles_w) * p.tw;
      }
      const int h = h0 + pix_dh, w = w0 + pix_dw;
      const bool valid = !dummy && (h < p.out_h) && (w < p.out_w);
      const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + as * 256;

      if (p.out_mode == 0) {
        const int n32 = p.n_tile / 32;                   // NHWC outputs have Cout tiles that are multiples of 64
        const int c_lo = half * (n32 / 2), c_hi = c_lo + n32 / 2;
        // residual (ResNet identity / FPN top-down map): its global loads are issued one 32-channel chunk ahead - the
        // first chunk even before the accumulator is ready - so their latency hides behind the MMAs and the TMEM reads
        const __half* res_px = nullptr;
        if (!PAIR && p.res != nullptr && valid)          // (residual ops never run as CTA pairs: see prepare_conv)
          res_px = p.res + (((long long)img * p.res_h + (h / p.res_up)) * p.res_w + (w / p.res_up)) * p.res_c;
        uint4 res_a[NPLANE * 4], res_b[NPLANE * 4];
        auto prefetch_res = [&](int c32, uint4 (&dst)[NPLANE * 4]) {
          if (res_px != nullptr) {
            const int cb = n_idx * p.n_tile + c32 * 32;
#pragma unroll
            for (int pl = 0; pl < NPLANE; ++pl) {
              const uint4* rp = reinterpret_cast<const uint4*>(res_px + pl * p.res_plane_elems + cb);
#pragma unroll
              for (int j = 0; j < 4; ++j) dst[pl * 4 + j] = __ldg(rp + j);
            }
          }
        };
        prefetch_res(c_lo, res_a);
        // the later chunks' residual lines are pulled into L2 now (no registers): their loads, issued one chunk ahead, then see
        // L2 latency instead of DRAM latency (34 % of the stall samples of neck.lateral.0 sat on these loads)
        if (res_px != nullptr) {
          for (int c32 = c_lo + 1; c32 < c_hi; ++c32)
#pragma unroll
            for (int pl = 0; pl < NPLANE; ++pl)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(res_px + pl * p.res_plane_elems + n_idx * p.n_tile + c32 * 32));
        }

        ptx::mbar_wait(&tmem_full_bar[as], aphase);
        ptx::tc_fence_after();

        const int hq = h0 + lb_dh, wq = w0 + lb_dw;           // first pixel of this warp's 32 (store box origin)
        // the accumulator values of one 32-channel chunk of this thread's pixel (main + correction accumulator)
        auto drain = [&](int c32, float (&v)[32]) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(taddr + c32 * 32, r);
          if (CORR) {
            uint32_t rc[32];
            ptx::tmem_ld_32x32b_x32(taddr + p.corr_off + c32 * 32, rc);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + __uint_as_float(rc[j]);
          } else {
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          }
        };
        auto chunk = [&](int c32, const uint4 (&res)[NPLANE * 4], float (&v)[32]) {
          const int cb = n_idx * p.n_tile + c32 * 32;
          if (p.bias_in_params) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], p.wscale_inv, p.bias_c[cb + j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cb + j));
              v[j + 0] = fmaf(v[j + 0], p.wscale_inv, b4.x);
              v[j + 1] = fmaf(v[j + 1], p.wscale_inv, b4.y);
              v[j + 2] = fmaf(v[j + 2], p.wscale_inv, b4.z);
              v[j + 3] = fmaf(v[j + 3], p.wscale_inv, b4.w);
            }
          }
          if (res_px != nullptr) {
#pragma unroll
            for (int pl = 0; pl < NPLANE; ++pl) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const __half2* h2 = reinterpret_cast<const __half2*>(&res[pl * 4 + j]);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float2 f = __half22float2(h2[t]);
                  v[j * 8 + 2 * t] += f.x;
                  v[j * 8 + 2 * t + 1] += f.y;
                }
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
            if (p.relu == 2) {                             // ReLU6 (separable convs, MobileNetV2)
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fminf(v[j], 6.0f);
            }
          }
          // The TMA unit serves this warp's small stores behind the producer's 64-96 KB loads, so a store takes ~1 us to
          // finish reading its staging buffer; with ONE buffer per warp every chunk waited for the previous chunk's store
          // (80 % of the epilogue's busy samples, and the epilogue is exposed whenever the accumulator is not double
          // buffered).  With a ring of `stage_depth` buffers only the store issued depth chunks ago has to be done.
          if (p.debug & 1) return;
          if (p.direct_store) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) {
              const float a = v[2 * t], b = v[2 * t + 1];
              const __half2 h2 = __floats2half2_rn(a, b);
              hi[t] = *reinterpret_cast<const uint32_t*>(&h2);
              if (NPLANE == 2) {
                const float2 back = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn(a - back.x, b - back.y);
                lo[t] = *reinterpret_cast<const uint32_t*>(&l2);
              }
            }
            if (valid) {
              // dst_up == 2: the destination has twice the resolution; phase -1 fills the 2x2 block (conv + nearest x2), a
              // phase 0..3 writes one sub-pixel (one phase of a stride-2 transposed conv)
              const int up = p.dst_up == 2 ? 2 : 1;
              for (int ph = 0; ph < up * up; ++ph) {
                if (up == 2 && p.dst_phase >= 0 && ph != p.dst_phase) continue;
                const int oh = h * up + (ph >> 1), ow = w * up + (ph & 1);
                __half* dp = p.dst + (((long long)img * p.dst_h + oh) * p.dst_w + ow) * p.dst_c + p.dst_c_off + cb;
                ptx::st_global_256(dp, reinterpret_cast<const uint32_t(&)[8]>(hi[0]));
                ptx::st_global_256(dp + 16, reinterpret_cast<const uint32_t(&)[8]>(hi[8]));
                if (NPLANE == 2) {
                  ptx::st_global_256(dp + p.dst_plane_elems, reinterpret_cast<const uint32_t(&)[8]>(lo[0]));
                  ptx::st_global_256(dp + p.dst_plane_elems + 16, reinterpret_cast<const uint32_t(&)[8]>(lo[8]));
                }
              }
            }
            return;
          }
          uint8_t* my_stage = my_stage_base + (size_t)sbuf * NPLANE * kStageWarpBytes;
          if (++sbuf == p.stage_depth) sbuf = 0;
          if (lane == 0) {
            if (p.debug & 4) {}
            else if (p.stage_depth >= 3) ptx::tma_store_wait_read<2>();
            else if (p.stage_depth == 2) ptx::tma_store_wait_read<1>();
            else                         ptx::tma_store_wait_read<0>();
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {                  // 8 channels = one 16-byte chunk; rows are 64 B, SWIZZLE_64B
            uint4 hi, lo;
            __half2* hh = reinterpret_cast<__half2*>(&hi);
            __half2* ll = reinterpret_cast<__half2*>(&lo);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float a = v[j * 8 + 2 * t], b = v[j * 8 + 2 * t + 1];
              const __half2 h2 = __floats2half2_rn(a, b);
              hh[t] = h2;
              if (NPLANE == 2) {
                const float2 back = __half22float2(h2);
                ll[t] = __floats2half2_rn(a - back.x, b - back.y);
              }
            }
            const int off = lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
            *reinterpret_cast<uint4*>(my_stage + off) = hi;
            if (NPLANE == 2) *reinterpret_cast<uint4*>(my_stage + kStageWarpBytes + off) = lo;
          }
          if (!(p.debug & 2)) ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && !(p.debug & 8)) {
            if (p.dst_up == 2) {
              // destination at twice the resolution, viewed as [n][y][py][x][px*C + c]: one store per sub-pixel.
              // All four = conv followed by a nearest x2 upsample; a single one = one phase of a stride-2 transposed conv.
              for (int ph = 0; ph < 4; ++ph) {
                if (p.dst_phase >= 0 && ph != p.dst_phase) continue;
#pragma unroll
                for (int pl = 0; pl < NPLANE; ++pl)
                  ptx::tma_store_5d(&dst_map, my_stage + pl * kStageWarpBytes, (ph & 1) * p.dst_c + p.dst_c_off + cb, wq, ph >> 1,
                                    hq, img + pl * p.n_img);
              }
            } else {
              ptx::tma_store_5d(&dst_map, my_stage, p.dst_c_off + cb, wq, hq, (p.debug & 64) ? (1 << 20) : ((p.debug & 128) ? (img & 1) : img), 0);        // both planes in one request (debug 64: fully out of bounds = clipped, timing experiment)
            }
            ptx::tma_store_commit();
          }
        };
        if constexpr (PAIR) {
          if (p.early_release) {
            // Single accumulator stage (main + correction accumulators of a 256-wide tile fill the 512 TMEM columns): the
            // MMA warp cannot start the next tile before this epilogue has released the accumulator, and the store path
            // (convert, stage, wait for the previous chunk's TMA store, store) is most of the epilogue.  So the warp first
            // drains its 4 x 32 columns into registers - the main accumulator straight into its final registers, the
            // correction accumulator in 16-column pieces, 128 + 16 live values under the 168-register cap of a 10-warp CTA -
            // hands the accumulator back, and only then converts and stores, overlapped with the next tile's MMAs.
            float acc[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c) ptx::tmem_ld_32x32b_x32_f(taddr + (c_lo + c) * 32, acc[c]);
            ptx::tmem_ld_wait();
            if (CORR) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                  uint32_t rc[16];
                  ptx::tmem_ld_32x32b_x16(taddr + p.corr_off + (c_lo + c) * 32 + hf * 16, rc);
                  ptx::tmem_ld_wait();
#pragma unroll
                  for (int j = 0; j < 16; ++j) acc[c][hf * 16 + j] += __uint_as_float(rc[j]);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (crank != 0) ptx::mbar_arrive_cluster(ptx::mapa_u32(&tmem_empty_bar[as], 0));
              else            ptx::mbar_arrive(&tmem_empty_bar[as]);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) chunk(c_lo + c, res_a, acc[c]);
            continue;                                     // (the accumulator was released above)
          }
        }
        for (int c32 = c_lo; c32 < c_hi; c32 += 2) {
          float v[32];
          if (c32 + 1 < c_hi) prefetch_res(c32 + 1, res_b);
          drain(c32, v);
          chunk(c32, res_a, v);
          if (c32 + 1 < c_hi) {
            if (c32 + 2 < c_hi) prefetch_res(c32 + 2, res_a);
            drain(c32 + 1, v);
            chunk(c32 + 1, res_b, v);
          }
        }
      } else {
        ptx::mbar_wait(&tmem_full_bar[as], aphase);
        ptx::tc_fence_after();
        const int n16 = p.n_tile / 16;
        const int c_lo = half * ((n16 + 1) / 2), c_hi = half ? n16 : (n16 + 1) / 2;
        float* out_px = p.out_nchw + ((long long)img * p.cout_real * p.out_h + h) * p.out_w + w;
        const long long cstride = (long long)p.out_h * p.out_w;
        for (int c16 = c_lo; c16 < c_hi; ++c16) {
          uint32_t r[16];
          float v[16];
          ptx::tmem_ld_32x32b_x16(taddr + c16 * 16, r);
          if (CORR) {
            uint32_t rc[16];
            ptx::tmem_ld_32x32b_x16(taddr + p.corr_off + c16 * 16, rc);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + __uint_as_float(rc[j]);
          } else {
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          }
          const int cb = n_idx * p.n_tile + c16 * 16;
          // Branch-free arithmetic over the whole 16-channel chunk (the bias array is padded to the Cout tile), stores
          // predicated afterwards: with the channel test around each element the 16 bias loads, logistics (MUFU.EX2 + a
          // correctly rounded division) and stores ran one after the other, and this epilogue - not HBM - set the tile period
          // of the head out-convs (6.8 us per 128-pixel tile).
          float bv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) bv[j] = p.bias_in_params ? p.bias_c[cb + j] : __ldg(p.bias + cb + j);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float val = fmaf(v[j], p.wscale_inv, bv[j]);
            if (p.relu) val = fmaxf(val, 0.0f);
            if (p.relu == 2) val = fminf(val, 6.0f);
            v[j] = val;
          }
          if (p.sigmoid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = sigmoid32(v[j]);
          }
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (cb + j < p.cout_real) out_px[(cb + j) * cstride] = v[j];
          }
        }
      }
      // accumulator drained: hand the TMEM stage back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && crank != 0) ptx::mbar_arrive_cluster(ptx::mapa_u32(&tmem_empty_bar[as], 0));   // the leader issues the MMAs
        else                    ptx::mbar_arrive(&tmem_empty_bar[as]);
      }
    }
    if (lane == 0) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) ptx::cluster_sync_all();       // no peer may still multicast into / arrive on this CTA's shared memory
  if (warp == 1) {
    ptx::tc_fence_after();
    if (PAIR) ptx::tmem_dealloc_pair(tmem_base, 512);
    else      ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Stem: conv7x7/2 (3->64) + BN + ReLU + max-pool3x3/2 on the tensor cores.
//   Cin = 3 has no UMMA shape, so the image is first rewritten as an "im2row" tensor T at half resolution:
//   space-to-depth by 2 turns the 7x7/2 conv into a 4x4/1 conv over 12 channels (3 colours x 2x2 phases); the four
//   horizontal taps are unrolled into the channel dimension (4 x 12 = 48, zero-padded to 64 = one 128-byte swizzle
//   row).  What is left is a conv with 4 VERTICAL taps, Cin = 64, Cout = 64, which the generic tcgen05 kernel above
//   runs as kh=4, kw=1 (vertical taps are plain TMA row offsets; rows outside the image are zero-filled = padding).
//   A plane-aware 3x3/2 max-pool finishes the stem.
// ------------------------------------------------------------------------------------------------------------
// T[n][sy][sx][dxi*12 + c*4 + py*2 + px] = image[n][c][2*sy + py][2*(sx + dxi - 2) + px]   (0 outside the image)
// One thread builds the 64-channel row of one pixel; the rows of a warp's 32 consecutive pixels are contiguous in T
// (4 KB per plane), so they are staged in shared memory and written back as eight fully coalesced 512-byte stores
// (a thread storing its own 128-byte row touches 32 different lines per store instruction).
template <int NPLANE>
__global__ void __launch_bounds__(256)
stem_im2row_kernel(const float* __restrict__ image, __half* __restrict__ T, int N, int H, int W, long long plane_elems) {
  __shared__ uint4 s_rows[8][32][9];                 // [warp][pixel][8 chunks of 8 channels + 1 pad: conflict-free both ways]
  const int SH = H / 2, SW = W / 2;
  const long long total = (long long)N * SH * SW;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pix0 = t - lane;                   // first pixel of this warp
  float v[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) v[j] = 0.0f;
  if (t < total) {
    const int sx = (int)(t % SW);
    const int sy = (int)((t / SW) % SH);
    const int n = (int)(t / ((long long)SW * SH));
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int py = 0; py < 2; ++py) {
        const float* row = image + (((size_t)n * 3 + c) * H + (2 * sy + py)) * W;
#pragma unroll
        for (int dxi = 0; dxi < 4; ++dxi) {
          const int xs = sx + dxi - 2;
          float2 f = make_float2(0.0f, 0.0f);
          if (xs >= 0 && xs < SW) f = __ldg(reinterpret_cast<const float2*>(row + 2 * xs));
          v[dxi * 12 + c * 4 + py * 2 + 0] = f.x;
          v[dxi * 12 + c * 4 + py * 2 + 1] = f.y;
        }
      }
  }
#pragma unroll
  for (int pl = 0; pl < NPLANE; ++pl) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint4 u;
      __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float a = v[j * 8 + 2 * q], b = v[j * 8 + 2 * q + 1];
        const __half2 h2 = __floats2half2_rn(a, b);
        if (pl == 0) {
          hh[q] = h2;
        } else {
          const float2 back = __half22float2(h2);
          hh[q] = __floats2half2_rn(a - back.x, b - back.y);
        }
      }
      s_rows[warp][lane][j] = u;
    }
    __syncwarp();
    uint4* dst = reinterpret_cast<uint4*>(T + (size_t)pl * plane_elems + (size_t)pix0 * 64);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int q = it * 32 + lane;                  // 16-byte chunk index inside the warp's 4 KB block
      if (pix0 + (q >> 3) < total) dst[q] = s_rows[warp][q >> 3][q & 7];
    }
    __syncwarp();
  }
}

// Space-to-depth form of the image for the stem (default since round 2; CNL_STEM_IM2ROW=1 restores the im2row tensor):
//   S2D[pl][n][sy][sxp][c*4 + py*2 + px] = image[n][c][2*sy + py][2*(sxp - 2) + px]   (12 of 16 channels; 0 in the padding)
// with rows of SW + 3 pixels (2 zero pixels on the left, 1 on the right) of 16 fp16 channels = 32 B.  The horizontal im2col
// of the 4x4 conv - the K vector of output pixel x is the 4 consecutive pixels x .. x+3 of that padded row, 128 B - is NOT
// materialised: the conv's source tensor map has a 32-byte stride in its pixel dimension under a 128-byte row extent
// (overlapping rows; cuTensorMapEncodeTiled accepts it and the TMA unit delivers exactly these bytes, probed in
// tools/experiments/tma_overlap_probe.cu), so a [128 pixels][64] box load expands 4.1 KB of global memory into the 16 KB
// swizzled A tile on the fly.  HBM traffic of the stem's first two kernels: 637 + 1047 MB -> 236 + 632 MB.
template <int NPLANE>
__global__ void __launch_bounds__(256)
stem_s2d_kernel(const float* __restrict__ image, __half* __restrict__ S, int N, int H, int W, long long plane_elems) {
  const int SH = H / 2, SW = W / 2, PW = SW + 3;
  const long long total = (long long)N * SH * PW;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int sxp = (int)(t % PW);
  const int sy = (int)((t / PW) % SH);
  const int n = (int)(t / ((long long)PW * SH));
  const int sx = sxp - 2;
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = 0.0f;
  if (sx >= 0 && sx < SW) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int py = 0; py < 2; ++py) {
        const float2 f = __ldg(reinterpret_cast<const float2*>(image + (((size_t)n * 3 + c) * H + (2 * sy + py)) * W + 2 * sx));
        v[c * 4 + py * 2 + 0] = f.x;
        v[c * 4 + py * 2 + 1] = f.y;
      }
  }
#pragma unroll
  for (int pl = 0; pl < NPLANE; ++pl) {
    uint4 u[2];
    __half2* hh = reinterpret_cast<__half2*>(u);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float a = v[2 * q], b = v[2 * q + 1];
      const __half2 h2 = __floats2half2_rn(a, b);
      if (pl == 0) {
        hh[q] = h2;
      } else {
        const float2 back = __half22float2(h2);
        hh[q] = __floats2half2_rn(a - back.x, b - back.y);
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(S + (size_t)pl * plane_elems + (size_t)t * 16);
    dst[0] = u[0];
    dst[1] = u[1];
  }
}

// 3x3 stride-2 pad-1 max-pool over NHWC fp16 planes (value = hi + lo).  One thread = one output pixel x 8 channels.
template <int NPLANE>
__global__ void __launch_bounds__(256)
pool_planes_kernel(const __half* __restrict__ in, __half* __restrict__ out, int N, int IH, int IW, long long in_plane,
                   long long out_plane) {
  const int OH = IH / 2, OW = IW / 2;
  const long long total = (long long)N * OH * OW * 8;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c8 = (int)(t & 7);
  long long pix = t >> 3;
  const int ox = (int)(pix % OW);
  pix /= OW;
  const int oy = (int)(pix % OH);
  const int n = (int)(pix / OH);
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
  for (int dy = -1; dy <= 1; ++dy) {
    const int y = 2 * oy + dy;
    if (y < 0 || y >= IH) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int x = 2 * ox + dx;
      if (x < 0 || x >= IW) continue;
      const size_t o = (((size_t)n * IH + y) * IW + x) * 64 + c8 * 8;
      const uint4 uh = __ldg(reinterpret_cast<const uint4*>(in + o));
      const __half2* hh = reinterpret_cast<const __half2*>(&uh);
      float f[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) { const float2 a = __half22float2(hh[q]); f[2 * q] = a.x; f[2 * q + 1] = a.y; }
      if (NPLANE == 2) {
        const uint4 ul = __ldg(reinterpret_cast<const uint4*>(in + in_plane + o));
        const __half2* ll = reinterpret_cast<const __half2*>(&ul);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float2 a = __half22float2(ll[q]); f[2 * q] += a.x; f[2 * q + 1] += a.y; }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
    }
  }
  uint4 hi, lo;
  __half2* hh = reinterpret_cast<__half2*>(&hi);
  __half2* ll = reinterpret_cast<__half2*>(&lo);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half2 h2 = __floats2half2_rn(m[2 * j], m[2 * j + 1]);
    hh[j] = h2;
    const float2 back = __half22float2(h2);
    ll[j] = __floats2half2_rn(m[2 * j] - back.x, m[2 * j + 1] - back.y);
  }
  const size_t o = (((size_t)n * OH + oy) * OW + ox) * 64 + c8 * 8;
  *reinterpret_cast<uint4*>(out + o) = hi;
  if (NPLANE == 2) *reinterpret_cast<uint4*>(out + out_plane + o) = lo;
}


// ------------------------------------------------------------------------------------------------------------
// Bandwidth-bound ops of the separable-conv models (reference models/layers.py:40-79 `make_conv` separable branch,
// :138-177 `Fuse`; MobileNetV2's depthwise stages).  CUDA-core kernels on the same NHWC hi/lo planes: one thread = one
// output pixel x 8 channels (one 16-byte vector per plane), fp32 arithmetic on value = hi + lo.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load8(const __half* p, long long plane, bool two, float (&f)[8]) {
  const uint4 uh = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* hh = reinterpret_cast<const __half2*>(&uh);
#pragma unroll
  for (int q = 0; q < 4; ++q) { const float2 a = __half22float2(hh[q]); f[2 * q] = a.x; f[2 * q + 1] = a.y; }
  if (two) {
    const uint4 ul = __ldg(reinterpret_cast<const uint4*>(p + plane));
    const __half2* ll = reinterpret_cast<const __half2*>(&ul);
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float2 a = __half22float2(ll[q]); f[2 * q] += a.x; f[2 * q + 1] += a.y; }
  }
}
__device__ __forceinline__ void store8(__half* p, long long plane, bool two, const float (&m)[8]) {
  uint4 hi, lo;
  __half2* hh = reinterpret_cast<__half2*>(&hi);
  __half2* ll = reinterpret_cast<__half2*>(&lo);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half2 h2 = __floats2half2_rn(m[2 * j], m[2 * j + 1]);
    hh[j] = h2;
    const float2 back = __half22float2(h2);
    ll[j] = __floats2half2_rn(m[2 * j] - back.x, m[2 * j + 1] - back.y);
  }
  *reinterpret_cast<uint4*>(p) = hi;
  if (two) *reinterpret_cast<uint4*>(p + plane) = lo;
}
__device__ __forceinline__ float act_apply(float v, int act) {
  if (act) v = fmaxf(v, 0.0f);
  if (act == 2) v = fminf(v, 6.0f);
  return v;
}

// depthwise 3x3, pad 1, stride 1 or 2: out[n,y,x,c] = act(bias[c] + sum_{r,s} w[r*3+s][c] * in[n, y*stride+r-1, x*stride+s-1, c])
// (taps accumulated in the order r, s - the order ATen's direct depthwise kernel uses; fp32 FMAs)
template <int NPLANE>
__global__ void __launch_bounds__(256)
dw3x3_kernel(const __half* __restrict__ in, __half* __restrict__ out, const float* __restrict__ w9c, const float* __restrict__ bias,
             int N, int IH, int IW, int C, int stride, int act, long long in_plane, long long out_plane) {
  const int OH = IH / stride, OW = IW / stride, C8 = C / 8;
  const long long total = (long long)N * OH * OW * C8;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c8 = (int)(t % C8);
  long long pix = t / C8;
  const int ox = (int)(pix % OW);
  pix /= OW;
  const int oy = (int)(pix % OH);
  const int n = (int)(pix / OH);
  float acc[8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8 + 4));
    acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int y = oy * stride + r - 1;
    if (y < 0 || y >= IH) continue;
#pragma unroll
    for (int sx = 0; sx < 3; ++sx) {
      const int x = ox * stride + sx - 1;
      if (x < 0 || x >= IW) continue;
      float f[8];
      load8(in + (((size_t)n * IH + y) * IW + x) * C + c8 * 8, in_plane, NPLANE == 2, f);
      const float* wp = w9c + (size_t)(r * 3 + sx) * C + c8 * 8;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp)), w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
      acc[0] = fmaf(f[0], w0.x, acc[0]); acc[1] = fmaf(f[1], w0.y, acc[1]); acc[2] = fmaf(f[2], w0.z, acc[2]); acc[3] = fmaf(f[3], w0.w, acc[3]);
      acc[4] = fmaf(f[4], w1.x, acc[4]); acc[5] = fmaf(f[5], w1.y, acc[5]); acc[6] = fmaf(f[6], w1.z, acc[6]); acc[7] = fmaf(f[7], w1.w, acc[7]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = act_apply(acc[j], act);
  store8(out + (((size_t)n * OH + oy) * OW + ox) * C + c8 * 8, out_plane, NPLANE == 2, acc);
}

// Fuse node (reference models/layers.py:160-176): out = sum_i scale_i * x_i over up to three maps; the LAST one is resized
// first: resize 1 = nearest x2 up-sample (it has half the output resolution), 2 = MaxPool2d(2, 2) (twice the resolution).
template <int NPLANE>
__global__ void __launch_bounds__(256)
fuse_kernel(const __half* __restrict__ a, const __half* __restrict__ b, const __half* __restrict__ c, __half* __restrict__ out,
            float sa, float sb, float sc, int n_src, int resize, int N, int H, int W, int C, long long plane, long long last_plane) {
  const int C8 = C / 8;
  const long long total = (long long)N * H * W * C8;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c8 = (int)(t % C8);
  long long pix = t / C8;
  const int x = (int)(pix % W);
  pix /= W;
  const int y = (int)(pix % H);
  const int n = (int)(pix / H);
  const size_t o = (((size_t)n * H + y) * W + x) * C + c8 * 8;
  const __half* srcs[3] = {a, b, c};
  const float scales[3] = {sa, sb, sc};
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
  for (int i = 0; i < n_src; ++i) {
    float f[8];
    if (i < n_src - 1 || resize == 0) {
      load8(srcs[i] + o, plane, NPLANE == 2, f);
    } else if (resize == 1) {
      const int LH = H / 2, LW = W / 2;
      load8(srcs[i] + (((size_t)n * LH + (y >> 1)) * LW + (x >> 1)) * C + c8 * 8, last_plane, NPLANE == 2, f);
    } else {
      const int LH = H * 2, LW = W * 2;
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = -INFINITY;
      for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx) {
          float g[8];
          load8(srcs[i] + (((size_t)n * LH + (2 * y + dy)) * LW + (2 * x + dx)) * C + c8 * 8, last_plane, NPLANE == 2, g);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], g[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], scales[i], acc[j]);
  }
  store8(out + o, plane, NPLANE == 2, acc);
}

// MobileNetV2 stem: conv3x3 stride 2 pad 1 from the fp32 NCHW image (3 channels) + bias + activation -> NHWC planes with C
// (padded) channels.  w27c: [(c*3 + r)*3 + s][C] fp32.  One thread = one output pixel x 8 output channels.
template <int NPLANE>
__global__ void __launch_bounds__(256)
stem3x3_kernel(const float* __restrict__ image, __half* __restrict__ out, const float* __restrict__ w27c, const float* __restrict__ bias,
               int N, int H, int W, int C, int act, long long out_plane) {
  const int OH = H / 2, OW = W / 2, C8 = C / 8;
  const long long total = (long long)N * OH * OW * C8;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c8 = (int)(t % C8);
  long long pix = t / C8;
  const int ox = (int)(pix % OW);
  pix /= OW;
  const int oy = (int)(pix % OH);
  const int n = (int)(pix / OH);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = __ldg(bias + c8 * 8 + j);
  for (int ch = 0; ch < 3; ++ch)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int y = 2 * oy + r - 1;
      if (y < 0 || y >= H) continue;
#pragma unroll
      for (int sx = 0; sx < 3; ++sx) {
        const int x = 2 * ox + sx - 1;
        if (x < 0 || x >= W) continue;
        const float v = __ldg(image + (((size_t)n * 3 + ch) * H + y) * W + x);
        const float* wp = w27c + (size_t)((ch * 3 + r) * 3 + sx) * C + c8 * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, __ldg(wp + j), acc[j]);
      }
    }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = act_apply(acc[j], act);
  store8(out + (((size_t)n * OH + oy) * OW + ox) * C + c8 * 8, out_plane, NPLANE == 2, acc);
}

// ------------------------------------------------------------------------------------------------------------
// Layout converters (tests / debugging): NHWC fp16 planes <-> NCHW fp32
// ------------------------------------------------------------------------------------------------------------
__global__ void planes_to_nchw_kernel(const __half* __restrict__ in, float* __restrict__ out, int N, int C, int H, int W,
                                      int planes, long long plane_elems) {
  const long long total = (long long)N * C * H * W;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % C);
  long long rest = t / C;
  const int w = (int)(rest % W); rest /= W;
  const int h = (int)(rest % H);
  const int n = (int)(rest / H);
  float v = __half2float(in[t]);
  if (planes == 2) v += __half2float(in[plane_elems + t]);
  out[(((long long)n * C + c) * H + h) * W + w] = v;
}
__global__ void nchw_to_planes_kernel(const float* __restrict__ in, __half* __restrict__ out, int N, int C, int H, int W,
                                      int planes, long long plane_elems) {
  const long long total = (long long)N * C * H * W;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % C);
  long long rest = t / C;
  const int w = (int)(rest % W); rest /= W;
  const int h = (int)(rest % H);
  const int n = (int)(rest / H);
  const float v = in[(((long long)n * C + c) * H + h) * W + w];
  const __half hi = __float2half_rn(v);
  out[t] = hi;
  if (planes == 2) out[plane_elems + t] = __float2half_rn(v - __half2float(hi));
}

// ------------------------------------------------------------------------------------------------------------
// Host: engine
// ------------------------------------------------------------------------------------------------------------
struct BufferInfo {
  int channels, stride, fp32_nchw;
  int h, w;
  size_t offset, bytes;
  long long plane_elems;
};

struct OpInfo {
  cnl_conv_desc d;
  // conv
  int cout_pad, n_tile, n_tiles, tw, th, tiles_w, tiles_h, num_stages, store_w, store_h, cluster, acc_stages, corr_off, cat;
  int kh, kw, pad_h, pad_w, up;         // resolved kernel extent / padding, destination scale (1 or 2)
  int rows, a_slots, a_slot_bytes, box_w; // ROWS mode geometry (rows = 0: im2col tiles)
  int w_resident;                         // ROWS: weights loaded once per CTA (one stage per tap)
  int pair;                               // CTA-pair (cta_group::2) form
  int stage_depth;                        // epilogue staging ring depth (1..3)
  bool corr;                          // SPLIT precision: separate correction accumulator (long reductions) or fused
  float wscale;
  size_t w_offset, bias_offset, scratch_offset;
  std::vector<__half> w_packed;       // [plane][tap][cout_pad][cin]
  std::vector<float> bias_packed;     // [cout_pad]   (stem: [64]);
  std::vector<float> w_f32;           // kinds 2 / 4 (depthwise 3x3, 3x3/2 image stem): [tap][C] fp32
  size_t stem_t_offset, stem_s_offset;   // stem scratch: im2row tensor T (or the space-to-depth image) and the un-pooled conv output S
  int stem_s2d;                          // stem: 1 = space-to-depth image + overlapping-stride tensor map, 0 = materialised im2row tensor
  CUtensorMap src_map, w_map, dst_map;
};

}  // namespace cnl

struct cnl_engine {
  int batch, height, width, precision, planes, device, num_sms;
  std::vector<cnl::BufferInfo> bufs;
  std::vector<cnl::OpInfo> ops;
  size_t arena_bytes;
  void* uploaded_arena;
};

namespace cnl {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

static int encode_map(CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box, const cuuint32_t* estride, const char* what,
                      CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(CNL_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, base, dims, strides_bytes, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CNL_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return CNL_OK;
}

static int pow2_ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// CNL_DIRECT_STORE=1: epilogue writes 256-bit global stores from registers instead of smem staging + TMA stores (measured
// SLOWER: forward 13.8 -> 14.65 ms; kept for A/B timing)
static bool direct_store_enabled() {
  static const bool on = [] { const char* v = getenv("CNL_DIRECT_STORE"); return v && atoi(v) != 0; }();
  return on;
}

// CTAs per cluster that share one multicast weight tile.  Opt-in through CNL_CLUSTER=2 (or 4): measured on B200 the
// kernel is tensor/power-bound, not L2-bound, so halving the weight traffic changes nothing (1.229 vs 1.220 ms on the
// 256->256 tower conv; cluster 4 strands SMs and is slower); default 1.  Each CTA loads
// n_tile/cluster weight rows, which must stay a multiple of the 8-row swizzle atom.
static int cluster_env() {        // CNL_CLUSTER = 1 / 2 / 4 forces a cluster size for the single-CTA-MMA ops (0: the default rule)
  static const int env = [] { const char* v = getenv("CNL_CLUSTER"); int c = v ? atoi(v) : 0; return (c == 1 || c == 2 || c == 4) ? c : 0; }();
  return env;
}
static int choose_cluster(int n_tile, int m_tiles, int taps) {
  const int env = cluster_env();
  // Default: clusters of 4 for the 3x3 convs with Cout tiles of 256 on small maps (ResNet layers 3-4, the stride-16/32 FPN
  // output convs): every CTA streams the whole 2.4-4.7 MB weight tensor per pixel tile there, and sharing each tile four
  // ways measures 13-17 % faster (0.099 -> 0.085 ms on layer3, 0.097 -> 0.080 ms on layer4).  Everything else keeps 1:
  // the large ops are tensor / power bound (CTA pairs), and on the 64-channel / 128-channel layers clusters only strand SMs.
  int c = env ? env : ((n_tile == 256 && taps == 9) ? 4 : 1);
  while (c > 1 && (n_tile % (8 * c) != 0 || m_tiles < 2 * c)) c >>= 1;
  return c;
}

// CNL_CAT=0 switches the concatenated hi*[hi|lo] MMA off (A/B timing)
static bool cat_enabled() {
  static const bool on = [] { const char* v = getenv("CNL_CAT"); return !(v && atoi(v) == 0); }();
  return on;
}

// CNL_ROWS=0 switches the row-rolling kernel variant off (A/B timing)
static bool rows_enabled() {
  static const bool on = [] { const char* v = getenv("CNL_ROWS"); return !(v && atoi(v) == 0); }();
  return on;
}

// Row-rolling geometry for an op whose tiles are single output rows (see conv_tc_kernel<.., ROWS = true>).  Returns false
// (and leaves the im2col configuration in place) when the op does not qualify or the ring does not fit shared memory.
struct OpInfo;
static bool plan_rows_mode(OpInfo& op, int planes, int cin, int stride, bool nhwc_out, int precision);

static void split_half(float x, __half* hi, __half* lo) {
  *hi = __float2half_rn(x);
  *lo = __float2half_rn(x - __half2float(*hi));
}

static void pack_split_weights(OpInfo& op, const std::vector<float>& w, int cout, int cin, int taps, int planes);

static int prepare_conv(cnl_engine* e, OpInfo& op) {
  const cnl_conv_desc& d = op.d;
  const BufferInfo& src = e->bufs[d.src];
  const BufferInfo& dst = e->bufs[d.dst];
  if (src.fp32_nchw) return fail(CNL_ERR_UNSUPPORTED, "conv input must be an NHWC fp16 buffer");
  if (d.cin % 64 || d.src_c_off % 64) return fail(CNL_ERR_UNSUPPORTED, "conv Cin/src_c_off must be multiples of 64 (got %d/%d)", d.cin, d.src_c_off);
  if (d.src_c_off + d.cin > src.channels) return fail(CNL_ERR_INVALID_ARGUMENT, "conv reads past the source channels");
  // kernel extent: ksize x ksize with symmetric padding, or an explicit kh x kw window with its own top/left padding
  // (the sub-pixel phases of a stride-2 transposed conv are 1x1 / 1x2 / 2x1 / 2x2 windows)
  op.kh = d.kh > 0 ? d.kh : d.ksize; op.kw = d.kw > 0 ? d.kw : d.ksize;
  op.pad_h = d.kh > 0 ? d.pad_h : d.pad; op.pad_w = d.kw > 0 ? d.pad_w : d.pad;
  op.up = d.dst_up == 2 ? 2 : 1;
  if (op.kh < 1 || op.kh > 3 || op.kw < 1 || op.kw > 3) return fail(CNL_ERR_UNSUPPORTED, "conv window %dx%d (1..3 per side are implemented)", op.kh, op.kw);
  if (op.pad_h < 0 || op.pad_h >= 3 || op.pad_w < 0 || op.pad_w >= 3) return fail(CNL_ERR_INVALID_ARGUMENT, "conv padding %d/%d", op.pad_h, op.pad_w);
  if (d.stride != 1 && d.stride != 2) return fail(CNL_ERR_UNSUPPORTED, "conv stride %d", d.stride);
  if (d.dst_up != 0 && d.dst_up != 1 && d.dst_up != 2) return fail(CNL_ERR_INVALID_ARGUMENT, "conv dst_up %d", d.dst_up);
  if (op.up == 2 && (d.dst_phase < -1 || d.dst_phase > 3)) return fail(CNL_ERR_INVALID_ARGUMENT, "conv dst_phase %d", d.dst_phase);
  if (op.up == 2 && (dst.fp32_nchw || d.residual >= 0)) return fail(CNL_ERR_UNSUPPORTED, "an upsampling store needs an NHWC destination and no residual");
  if (src.h / d.stride * op.up != dst.h || src.w / d.stride * op.up != dst.w) return fail(CNL_ERR_INVALID_ARGUMENT, "conv output size mismatch");
  static const int env_ntile = [] { const char* v = getenv("CNL_NTILE"); return v ? atoi(v) : 0; }();
  if (env_ntile == 128 && d.cout % 128 == 0 && d.cout > 128 && !dst.fp32_nchw) { op.n_tile = 128; op.cout_pad = d.cout; }   // experiment
  else if (d.cout % 256 == 0) { op.n_tile = 256; op.cout_pad = d.cout; }
  else if (d.cout <= 256) { op.cout_pad = (d.cout + 15) / 16 * 16; op.n_tile = op.cout_pad; }
  // wider layers that are not multiples of 256 (MobileNetV2's 384 / 576 / 960 / 320 channels): the widest of 192 / 128 / 64 that divides
  else if (d.cout % 192 == 0) { op.n_tile = 192; op.cout_pad = d.cout; }
  else if (d.cout % 128 == 0) { op.n_tile = 128; op.cout_pad = d.cout; }
  else if (d.cout % 64 == 0) { op.n_tile = 64; op.cout_pad = d.cout; }
  else return fail(CNL_ERR_UNSUPPORTED, "conv Cout=%d (must be <= 256 or a multiple of 64)", d.cout);
  op.n_tiles = op.cout_pad / op.n_tile;
  if (!dst.fp32_nchw) {
    if (d.cout % 64 || d.dst_c_off % 64) return fail(CNL_ERR_UNSUPPORTED, "NHWC conv output needs Cout %% 64 == 0");
    if (d.dst_c_off + d.cout > dst.channels) return fail(CNL_ERR_INVALID_ARGUMENT, "conv writes past the destination channels");
  } else if (d.cout != dst.channels || d.dst_c_off != 0) {
    return fail(CNL_ERR_INVALID_ARGUMENT, "fp32 NCHW output must cover the whole buffer");
  }
  if (d.residual >= 0) {
    const BufferInfo& rb = e->bufs[d.residual];
    if (rb.fp32_nchw || dst.fp32_nchw) return fail(CNL_ERR_UNSUPPORTED, "residual needs NHWC buffers");
    if (rb.channels != d.cout || rb.h * d.residual_up != dst.h || rb.w * d.residual_up != dst.w)
      return fail(CNL_ERR_INVALID_ARGUMENT, "residual shape mismatch");
  }
  const int out_w = dst.w / op.up, out_h = dst.h / op.up;      // the conv's own output grid
  op.tw = std::max(8, std::min(128, pow2_ceil(out_w)));
  op.th = kBlockM / op.tw;
  op.tiles_w = (out_w + op.tw - 1) / op.tw;
  op.tiles_h = (out_h + op.th - 1) / op.th;
  op.store_w = std::min(op.tw, 32);
  op.store_h = 32 / op.store_w;
  const int planes = e->planes;
  const int stage_bytes = planes * (kATileBytes + op.n_tile * kBlockK * 2);
  const int staging = (dst.fp32_nchw || direct_store_enabled()) ? 0 : kEpiWarps * planes * kStageWarpBytes;
  op.num_stages = std::min(kMaxStages, (kSmemLimit - 1024 - staging) / stage_bytes);
  if (op.num_stages < 2) return fail(CNL_ERR_UNSUPPORTED, "conv tile does not fit shared memory");
  op.cluster = 1;                                    // decided below, after the row-rolling and CTA-pair forms
  // The separate correction accumulator pays off where the reduction is long (its rounding bias grows with the number
  // of accumulate steps); short reductions (K < 576: stem, 1x1 convs) keep the single one.
  op.corr = (op.kh * op.kw * (d.cin / 64)) >= 9;
  // TMEM: 512 columns = two accumulator stages of 256.  A correction accumulator sits 128 columns after the main one
  // when the Cout tile is at most 128 wide; wider tiles need the whole 512 columns for one (main, correction) pair.
  const bool split_corr = (e->precision == CNL_PRECISION_SPLIT) && op.corr;
  // Cout tiles of at most 128 ("cat"): hi x [hi|lo] is one MMA of N = 2*n_tile, so the correction accumulator starts
  // right after the main one and every such op runs with the separate accumulator, whatever its K.
  op.cat = (e->precision == CNL_PRECISION_SPLIT) && op.n_tile <= 128 && cat_enabled();
  if (op.cat) op.corr = true;
  op.corr_off = op.cat ? op.n_tile : ((op.n_tile <= 128) ? 128 : 256);
  op.acc_stages = (split_corr && op.n_tile > 128) ? 1 : 2;
  plan_rows_mode(op, planes, d.cin, d.stride, !dst.fp32_nchw, e->precision);
  // CTA pair (cta_group::2): Cout tiles of 256 in split precision with the separate correction accumulator
  static const bool pair_on = [] { const char* v = getenv("CNL_PAIR"); return !(v && atoi(v) == 0); }();
  static const int pair_min_tiles = [] { const char* v = getenv("CNL_PAIR_MIN_TILES"); return v ? atoi(v) : 1024; }();   // tests lower it
  op.pair = 0;
  op.stage_depth = 1;
  // Cout tiles of 128 as pairs (CNL_PAIR128=1) measure 4 % SLOWER than the single-CTA [hi|lo] form: there the A operand's
  // shared-memory read (streamed twice instead of three times) matters more than the fill traffic
  static const bool pair128 = [] { const char* v = getenv("CNL_PAIR128"); return v && atoi(v) != 0; }();
  if (pair_on && e->precision == CNL_PRECISION_SPLIT && op.corr && (op.n_tile == 256 || (op.n_tile == 128 && pair128)) && !op.rows && !dst.fp32_nchw && d.residual < 0 &&
      cluster_env() <= 1 && e->batch * op.tiles_w * op.tiles_h >= 2 &&
      e->batch * op.tiles_w * op.tiles_h * op.n_tiles >= pair_min_tiles) {     // short launches (layer3/4) measure 5-7 % slower as pairs
    op.pair = 1;
    op.cluster = 2;                                  // launch as clusters of 2; each CTA's weight box is n_tile / 2 rows
    if (op.n_tile == 128) {                          // the pair splits the weight rows, so the [hi|lo] concatenation is not used:
      op.cat = 0; op.corr_off = 128; op.acc_stages = 2;   // three N = 128 instructions, accumulators at columns 0 / 128, two stages
    }
    const int pair_stage = planes * (kATileBytes + op.n_tile / 2 * kBlockK * 2);
    // the pair's smaller stages leave room for three load stages + one staging buffer per warp (default) or two stages + a
    // 2-/3-deep staging ring (CNL_STAGE_DEPTH; measured equal: at the board's power cap removing idle time from the
    // epilogue only lowers the clock)
    static const int env_depth = [] { const char* v = getenv("CNL_STAGE_DEPTH"); return v ? atoi(v) : 1; }();
    op.stage_depth = std::max(1, std::min(3, env_depth));
    op.num_stages = std::min(kMaxStages, (kSmemLimit - 1024 - op.stage_depth * staging) / pair_stage);
    if (op.num_stages < 2) { op.stage_depth = 1; op.num_stages = std::min(kMaxStages, (kSmemLimit - 1024 - staging) / pair_stage); }
  }
  if (!op.pair && !op.rows) op.cluster = choose_cluster(op.n_tile, e->batch * op.tiles_w * op.tiles_h, op.kh * op.kw);

  // pack weights: [plane][tap][cout_pad][cin], scaled by a power of two (keeps the lo parts normal in fp16)
  const int taps = op.kh * op.kw;
  std::vector<float> wt((size_t)d.cout * taps * d.cin);
  for (int co = 0; co < d.cout; ++co)
    for (int ci = 0; ci < d.cin; ++ci)
      for (int t = 0; t < taps; ++t) wt[((size_t)co * taps + t) * d.cin + ci] = d.weight_host[((size_t)co * d.cin + ci) * taps + t];
  pack_split_weights(op, wt, d.cout, d.cin, taps, planes);
  op.bias_packed.assign(op.cout_pad, 0.f);
  for (int co = 0; co < d.cout; ++co) op.bias_packed[co] = d.bias_host[co];
  return CNL_OK;
}

static void pack_split_weights(OpInfo& op, const std::vector<float>& w /*[cout][taps][cin]*/, int cout, int cin, int taps,
                               int planes) {
  float wmax = 0.f;
  for (float x : w) wmax = std::max(wmax, std::fabs(x));
  int ex = 0;
  if (wmax > 0.f) std::frexp(wmax, &ex);
  op.wscale = std::ldexp(1.0f, 13 - ex);               // max |w| * wscale in [4096, 8192)
  op.w_packed.assign((size_t)planes * taps * op.cout_pad * cin, __float2half_rn(0.f));
  for (int co = 0; co < cout; ++co)
    for (int t = 0; t < taps; ++t)
      for (int ci = 0; ci < cin; ++ci) {
        __half hi, lo;
        split_half(w[((size_t)co * taps + t) * cin + ci] * op.wscale, &hi, &lo);
        op.w_packed[(((size_t)0 * taps + t) * op.cout_pad + co) * cin + ci] = hi;
        if (planes == 2) op.w_packed[(((size_t)1 * taps + t) * op.cout_pad + co) * cin + ci] = lo;
      }
}

static bool plan_rows_mode(OpInfo& op, int planes, int cin, int stride, bool nhwc_out, int precision) {
  op.rows = 0; op.a_slots = 0; op.a_slot_bytes = 0; op.box_w = 0; op.w_resident = 0;
  if (!rows_enabled() || precision != CNL_PRECISION_SPLIT || !op.cat || planes != 2) return false;
  if (cin != 64 || stride != 1 || op.th != 1 || op.n_tiles != 1 || op.kh * op.kw < 2 || op.cluster != 1 || !nhwc_out) return false;
  const int box_w = op.tw + op.kw - 1;
  if (box_w > 256) return false;
  const int slot_bytes = (box_w * kBlockK * 2 + 1023) / 1024 * 1024;
  const int b_stage = planes * op.n_tile * kBlockK * 2;
  const int staging = direct_store_enabled() ? 0 : kEpiWarps * planes * kStageWarpBytes;
  // Row slots: the oldest row's slot is released after the r = 0 taps (see the MMA warp), so kh slots already give the
  // producer two thirds of an output row of look-ahead; every further slot costs weight stages, and it is the number of
  // weight tiles in flight that hides their L2 latency.  Take the largest ring that still leaves kMinRowsStages stages
  // (CNL_ROWS_SLOTS = slots beyond kh, for A/B timing).
  static const int extra_env = [] { const char* v = getenv("CNL_ROWS_SLOTS"); return v ? atoi(v) : -1; }();
  constexpr int kMinRowsStages = 5;
  int slots = op.kh + 2, b_stages = 0;
  for (int extra = (extra_env >= 0 ? extra_env : 2); extra >= 0; --extra) {
    slots = op.kh + extra;
    b_stages = (kSmemLimit - 1024 - staging - slots * planes * slot_bytes) / b_stage;
    if (b_stages >= kMinRowsStages || extra_env >= 0) break;
  }
  // Resident weights: when every tap's [hi|lo] tile fits beside a ring of kh slots (the stem: 4 taps x 16 KB), each tap
  // keeps its own stage for the whole launch and is loaded once per CTA (CNL_ROWS_RESIDENT=0 for A/B timing).
  static const bool resident_on = [] { const char* v = getenv("CNL_ROWS_RESIDENT"); return !(v && atoi(v) == 0); }();
  const int taps = op.kh * op.kw;
  op.w_resident = 0;
  if (resident_on && taps <= kMaxStages) {
    for (int extra = 1; extra >= 0; --extra) {
      const int rs = op.kh + extra;
      if (rs <= kMaxStages && kSmemLimit - 1024 - staging - rs * planes * slot_bytes >= taps * b_stage) {
        op.rows = 1; op.a_slots = rs; op.a_slot_bytes = slot_bytes; op.box_w = box_w;
        op.num_stages = taps; op.w_resident = 1;
        return true;
      }
    }
  }
  // fewer than three weight stages in flight and the kernel waits on weight-tile latency instead
  if (b_stages < 3 || slots > kMaxStages) return false;
  op.rows = 1; op.a_slots = slots; op.a_slot_bytes = slot_bytes; op.box_w = box_w;
  op.num_stages = std::min(kMaxStages, b_stages);
  return true;
}


// kinds 2 (depthwise 3x3), 3 (fuse), 4 (3x3/2 stem from the image): CUDA-core kernels on the NHWC planes
static int prepare_elementwise(cnl_engine* e, OpInfo& op) {
  const cnl_conv_desc& d = op.d;
  const BufferInfo& dst = e->bufs[d.dst];
  op.rows = 0; op.pair = 0; op.corr = false; op.cat = 0; op.cluster = 1; op.num_stages = 0; op.stage_depth = 0; op.a_slots = 0; op.n_tile = 0;
  op.wscale = 1.0f; op.up = 1; op.kh = op.kw = 3; op.pad_h = op.pad_w = 1; op.acc_stages = 0; op.corr_off = 0; op.tw = op.th = 0; op.tiles_w = op.tiles_h = 0;
  op.store_w = op.store_h = 0; op.n_tiles = 0; op.cout_pad = 0; op.a_slot_bytes = 0; op.box_w = 0; op.w_resident = 0;
  if (dst.fp32_nchw || dst.channels % 8) return fail(CNL_ERR_UNSUPPORTED, "depthwise / fuse / stem3x3 ops write NHWC buffers with C %% 8 == 0");
  if (d.kind == 2) {
    const BufferInfo& src = e->bufs[d.src];
    if (src.fp32_nchw || src.channels != dst.channels || d.cin != dst.channels || d.cout != dst.channels || d.ksize != 3 || d.pad != 1 ||
        (d.stride != 1 && d.stride != 2) || src.h / d.stride != dst.h || src.w / d.stride != dst.w || d.residual >= 0)
      return fail(CNL_ERR_UNSUPPORTED, "depthwise op: 3x3, pad 1, stride 1 or 2, same channel count on both sides, no residual");
    const int C = dst.channels;
    op.w_f32.assign((size_t)9 * C, 0.f);
    for (int c = 0; c < C; ++c)
      for (int t = 0; t < 9; ++t) op.w_f32[(size_t)t * C + c] = d.weight_host[(size_t)c * 9 + t];
    op.bias_packed.assign(d.bias_host, d.bias_host + C);
    return CNL_OK;
  }
  if (d.kind == 3) {
    const int srcs[3] = {d.src, d.src2, d.src3};
    const int n_src = d.src3 >= 0 ? 3 : 2;
    if (d.src2 < 0 || d.src2 >= (int)e->bufs.size() || d.src3 >= (int)e->bufs.size()) return fail(CNL_ERR_INVALID_ARGUMENT, "fuse op: bad source buffer id");
    if (d.resize < 0 || d.resize > 2) return fail(CNL_ERR_INVALID_ARGUMENT, "fuse op: resize %d", d.resize);
    for (int i = 0; i < n_src; ++i) {
      const BufferInfo& sb = e->bufs[srcs[i]];
      const bool last = i == n_src - 1;
      const int eh = last && d.resize == 1 ? dst.h / 2 : (last && d.resize == 2 ? dst.h * 2 : dst.h);
      const int ew = last && d.resize == 1 ? dst.w / 2 : (last && d.resize == 2 ? dst.w * 2 : dst.w);
      if (sb.fp32_nchw || sb.channels != dst.channels || sb.h != eh || sb.w != ew) return fail(CNL_ERR_INVALID_ARGUMENT, "fuse op: source %d shape mismatch", i);
    }
    return CNL_OK;
  }
  // kind 4
  const BufferInfo& src = e->bufs[d.src];
  if (!src.fp32_nchw || src.channels != 3 || d.cin != 3 || d.ksize != 3 || d.stride != 2 || d.pad != 1 || d.cout != dst.channels ||
      dst.h * 2 != src.h || dst.w * 2 != src.w)
    return fail(CNL_ERR_UNSUPPORTED, "stem3x3 must be conv3x3/2 pad 1 from the fp32 image");
  const int C = dst.channels;
  op.w_f32.assign((size_t)27 * C, 0.f);
  for (int c = 0; c < C; ++c)
    for (int t = 0; t < 27; ++t) op.w_f32[(size_t)t * C + c] = d.weight_host[(size_t)c * 27 + t];
  op.bias_packed.assign(d.bias_host, d.bias_host + C);
  return CNL_OK;
}

static int prepare_stem(cnl_engine* e, OpInfo& op) {
  const cnl_conv_desc& d = op.d;
  const BufferInfo& src = e->bufs[d.src];
  const BufferInfo& dst = e->bufs[d.dst];
  if (!src.fp32_nchw || src.channels != 3 || d.cin != 3 || d.cout != 64 || d.ksize != 7 || d.stride != 2 || d.pad != 3 ||
      dst.fp32_nchw || dst.channels != 64 || dst.h * 4 != src.h || dst.w * 4 != src.w)
    return fail(CNL_ERR_UNSUPPORTED, "stem must be conv7x7/2 (3->64) + max-pool3x3/2 from the fp32 image");
  // inner conv: 4 vertical taps over the im2row tensor, Cin = 64 (48 used), Cout = 64, at half resolution
  const int sw = e->width / 2, sh = e->height / 2;
  op.cout_pad = 64; op.n_tile = 64; op.n_tiles = 1;
  op.kh = 4; op.kw = 1; op.pad_h = 2; op.pad_w = 0; op.up = 1;
  op.tw = std::max(8, std::min(128, pow2_ceil(sw)));
  op.th = kBlockM / op.tw;
  op.tiles_w = (sw + op.tw - 1) / op.tw;
  op.tiles_h = (sh + op.th - 1) / op.th;
  op.store_w = std::min(op.tw, 32);
  op.store_h = 32 / op.store_w;
  const int planes = e->planes;
  const int stage_bytes = planes * (kATileBytes + op.n_tile * kBlockK * 2);
  op.num_stages = std::min(kMaxStages, (kSmemLimit - 1024 - (direct_store_enabled() ? 0 : kEpiWarps * planes * kStageWarpBytes)) / stage_bytes);
  op.cluster = 1;
  op.cat = (e->precision == CNL_PRECISION_SPLIT) && cat_enabled();
  op.corr = op.cat;                        // K = 256: the correction accumulator only comes with the cat MMA
  op.corr_off = op.cat ? 64 : 128; op.acc_stages = 2;
  op.pair = 0;
  op.stage_depth = 1;
  plan_rows_mode(op, planes, 64, 1, true, e->precision);
  if (!op.rows) op.cluster = choose_cluster(op.n_tile, e->batch * op.tiles_w * op.tiles_h, 4);
  static const bool im2row_env = [] { const char* v = getenv("CNL_STEM_IM2ROW"); return v && atoi(v) != 0; }();
  op.stem_s2d = im2row_env ? 0 : 1;
  // W2[co][dyi][dxi*12 (16 in the space-to-depth form) + c*4 + py*2 + px] = w[co][c][ky][kx] with ky <-> (dyi, py), kx <-> (dxi, px):
  //   k - 3 = 2*(d - 2) + p  =>  k = 2*d + p - 1  (k = -1, i.e. d = 0 and p = 0, does not exist -> weight 0)
  std::vector<float> w2((size_t)64 * 4 * 64, 0.f);
  for (int co = 0; co < 64; ++co)
    for (int c = 0; c < 3; ++c)
      for (int dyi = 0; dyi < 4; ++dyi)
        for (int py = 0; py < 2; ++py) {
          const int ky = 2 * dyi + py - 1;
          if (ky < 0 || ky > 6) continue;
          for (int dxi = 0; dxi < 4; ++dxi)
            for (int px = 0; px < 2; ++px) {
              const int kx = 2 * dxi + px - 1;
              if (kx < 0 || kx > 6) continue;
              w2[((size_t)co * 4 + dyi) * 64 + dxi * (op.stem_s2d ? 16 : 12) + c * 4 + py * 2 + px] = d.weight_host[(((size_t)co * 3 + c) * 7 + ky) * 7 + kx];
            }
        }
  pack_split_weights(op, w2, 64, 64, 4, planes);
  op.bias_packed.assign(d.bias_host, d.bias_host + 64);
  return CNL_OK;
}

}  // namespace cnl

using namespace cnl;

extern "C" {

int cnl_engine_create(cnl_engine** out, const cnl_buffer_desc* buffers, int n_buffers, const cnl_conv_desc* ops, int n_ops,
                      int batch, int height, int width, int precision, int device) {
  if (!out || !buffers || !ops || n_buffers < 2 || n_ops < 1) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_engine_create: bad arguments");
  if (batch < 1 || height < 32 || width < 32 || height % 32 || width % 32)
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_engine_create: input must be a multiple of 32 in both dimensions (got %dx%d)", height, width);
  if (precision != CNL_PRECISION_SPLIT && precision != CNL_PRECISION_FAST && precision != CNL_PRECISION_SPLIT_FUSED) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_engine_create: precision");
  cnl_engine* e = new cnl_engine();
  e->batch = batch; e->height = height; e->width = width; e->precision = precision; e->device = device;
  e->planes = (precision == CNL_PRECISION_FAST) ? 1 : 2;
  e->uploaded_arena = nullptr;
  e->num_sms = 148;                  // B200; replaced by the device attribute in cnl_engine_upload (create runs without a GPU)
  { int sms = 0; if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) e->num_sms = sms; else (void)cudaGetLastError(); }
  size_t off = 0;
  for (int i = 0; i < n_buffers; ++i) {
    BufferInfo b;
    b.channels = buffers[i].channels; b.stride = buffers[i].stride; b.fp32_nchw = buffers[i].fp32_nchw;
    if (b.stride < 1 || height % b.stride || width % b.stride) { delete e; return fail(CNL_ERR_INVALID_ARGUMENT, "buffer %d: bad stride", i); }
    b.h = height / b.stride; b.w = width / b.stride;
    b.plane_elems = (long long)batch * b.h * b.w * b.channels;
    b.bytes = b.fp32_nchw ? (size_t)b.plane_elems * 4 : (size_t)b.plane_elems * 2 * e->planes;
    b.offset = off;
    if (i > 0) off += align_up(b.bytes, 1024);            // buffer 0 is the caller's image
    e->bufs.push_back(b);
  }
  for (int i = 0; i < n_ops; ++i) {
    OpInfo op;
    op.d = ops[i];
    if (op.d.src < 0 || op.d.src >= n_buffers || op.d.dst <= 0 || op.d.dst >= n_buffers || op.d.residual >= n_buffers) {
      delete e; return fail(CNL_ERR_INVALID_ARGUMENT, "op %d: bad buffer id", i);
    }
    if (op.d.kind < 0 || op.d.kind > 4) { delete e; return fail(CNL_ERR_INVALID_ARGUMENT, "op %d: kind %d", i, op.d.kind); }
    int st = (op.d.kind == 1) ? prepare_stem(e, op) : (op.d.kind >= 2 ? prepare_elementwise(e, op) : prepare_conv(e, op));
    if (st != CNL_OK) { delete e; return st; }
    op.d.weight_host = nullptr; op.d.bias_host = nullptr;
    if (op.d.kind == 1) {
      const size_t half_res = (size_t)batch * (height / 2) * (width / 2) * 64 * 2 * e->planes;
      op.w_offset = off; off += align_up(op.w_packed.size() * 2, 1024);
      op.bias_offset = off; off += align_up(64 * 4, 1024);
      const size_t s2d_bytes = (size_t)batch * (height / 2) * (width / 2 + 3) * 16 * 2 * e->planes;
      op.stem_t_offset = off; off += align_up(op.stem_s2d ? s2d_bytes : half_res, 1024);
      op.stem_s_offset = off; off += align_up(half_res, 1024);
      op.scratch_offset = 0;
    } else if (op.d.kind >= 2) {
      op.w_offset = off; off += align_up(op.w_f32.size() * 4 + 16, 1024);
      op.bias_offset = off; off += align_up(op.bias_packed.size() * 4 + 16, 1024);
      op.scratch_offset = 0;
    } else {
      op.w_offset = off; off += align_up(op.w_packed.size() * 2, 1024);
      op.bias_offset = off; off += align_up(op.bias_packed.size() * 4, 1024);
      op.scratch_offset = 0;
    }
    e->ops.push_back(std::move(op));
  }
  e->arena_bytes = off;
  *out = e;
  return CNL_OK;
}

void cnl_engine_destroy(cnl_engine* e) { delete e; }

size_t cnl_engine_arena_bytes(const cnl_engine* e) { return e ? e->arena_bytes : 0; }

size_t cnl_engine_buffer_offset(const cnl_engine* e, int buffer) {
  if (!e || buffer < 0 || buffer >= (int)e->bufs.size()) return (size_t)-1;
  return e->bufs[buffer].offset;
}

int cnl_engine_op_form(const cnl_engine* e, int op) {
  if (!e || op < 0 || op >= (int)e->ops.size()) return -1;
  const OpInfo& o = e->ops[op];
  if (o.d.kind >= 2) return 0;                       // CUDA-core op (depthwise / fuse / stem3x3)
  return (o.rows ? 1 : 0) | (o.pair ? 2 : 0) | (o.corr ? 4 : 0) | (o.cat ? 8 : 0) | (o.cluster << 4) | (o.num_stages << 8) | (o.stage_depth << 12) |
         (o.a_slots << 16) | ((o.n_tile / 16) << 20);
}

int cnl_engine_upload(cnl_engine* e, void* arena, void* stream) {
  if (!e || !arena) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_engine_upload: null argument");
  if (reinterpret_cast<uintptr_t>(arena) & 1023) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_engine_upload: arena must be 1024-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev_sms = 0;
  CNL_CUDA_CHECK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, e->device));
  int cc_major = 0;
  CNL_CUDA_CHECK(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, e->device));
  if (cc_major != 10) return fail(CNL_ERR_UNSUPPORTED, "cnl_b200 kernels are built for sm_100a only (device has compute capability %d.x)", cc_major);
  e->num_sms = dev_sms;
  CNL_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<1, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  CNL_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<2, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  CNL_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<2, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  CNL_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<2, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  CNL_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<2, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  uint8_t* base = static_cast<uint8_t*>(arena);
  const int planes = e->planes;
  for (OpInfo& op : e->ops) {
    const cnl_conv_desc& d = op.d;
    if (d.kind == 1) {
      CNL_CUDA_CHECK(cudaMemcpyAsync(base + op.w_offset, op.w_packed.data(), op.w_packed.size() * 2, cudaMemcpyHostToDevice, st));
      CNL_CUDA_CHECK(cudaMemcpyAsync(base + op.bias_offset, op.bias_packed.data(), 64 * 4, cudaMemcpyHostToDevice, st));
      const cuuint64_t sw = e->width / 2, sh = e->height / 2;
      cuuint64_t dims[4] = {64, sw, sh, (cuuint64_t)e->batch * planes};
      cuuint64_t str[3] = {128, sw * 128, sh * sw * 128};
      cuuint32_t es[4] = {1, 1, 1, 1};
      cuuint32_t box_in[4] = {64, (cuuint32_t)(op.rows ? op.box_w : op.tw), (cuuint32_t)op.th, 1};
      int r;
      if (op.stem_s2d) {
        // overlapping rows: pixel stride 32 B (16 channels) under a 128-byte row extent (4 pixels) - see stem_s2d_kernel
        const cuuint64_t pitch = (sw + 3) * 32;
        cuuint64_t str2[3] = {32, pitch, sh * pitch};
        r = encode_map(&op.src_map, base + op.stem_t_offset, 4, dims, str2, box_in, es, "stem space-to-depth (overlapping rows)");
      } else {
        r = encode_map(&op.src_map, base + op.stem_t_offset, 4, dims, str, box_in, es, "stem im2row");
      }
      if (r) return r;
      {
        cuuint64_t od[5] = {64, sw, sh, (cuuint64_t)e->batch, (cuuint64_t)planes};
        cuuint64_t os[4] = {128, sw * 128, sh * sw * 128, (cuuint64_t)e->batch * sh * sw * 128};
        cuuint32_t ob[5] = {32, (cuuint32_t)op.store_w, (cuuint32_t)op.store_h, 1, (cuuint32_t)planes};
        cuuint32_t oe[5] = {1, 1, 1, 1, 1};
        r = encode_map(&op.dst_map, base + op.stem_s_offset, 5, od, os, ob, oe, "stem out", CU_TENSOR_MAP_SWIZZLE_64B);
      }
      if (r) return r;
      cuuint64_t wd[3] = {64, 64, (cuuint64_t)4 * planes};
      cuuint64_t ws[2] = {128, 64 * 128};
      cuuint32_t wb[3] = {64, (cuuint32_t)(64 / op.cluster), 1};
      cuuint32_t we[3] = {1, 1, 1};
      r = encode_map(&op.w_map, base + op.w_offset, 3, wd, ws, wb, we, "stem weights");
      if (r) return r;
      continue;
    }
    if (d.kind >= 2) {
      if (!op.w_f32.empty()) CNL_CUDA_CHECK(cudaMemcpyAsync(base + op.w_offset, op.w_f32.data(), op.w_f32.size() * 4, cudaMemcpyHostToDevice, st));
      if (!op.bias_packed.empty()) CNL_CUDA_CHECK(cudaMemcpyAsync(base + op.bias_offset, op.bias_packed.data(), op.bias_packed.size() * 4, cudaMemcpyHostToDevice, st));
      continue;
    }
    CNL_CUDA_CHECK(cudaMemcpyAsync(base + op.w_offset, op.w_packed.data(), op.w_packed.size() * 2, cudaMemcpyHostToDevice, st));
    CNL_CUDA_CHECK(cudaMemcpyAsync(base + op.bias_offset, op.bias_packed.data(), op.bias_packed.size() * 4, cudaMemcpyHostToDevice, st));
    const BufferInfo& src = e->bufs[d.src];
    const BufferInfo& dst = e->bufs[d.dst];
    const int taps = op.kh * op.kw;
    {
      cuuint64_t dims[4] = {(cuuint64_t)src.channels, (cuuint64_t)src.w, (cuuint64_t)src.h, (cuuint64_t)e->batch * planes};
      cuuint64_t str[3] = {(cuuint64_t)src.channels * 2, (cuuint64_t)src.w * src.channels * 2, (cuuint64_t)src.h * src.w * src.channels * 2};
      cuuint32_t box[4] = {64, (cuuint32_t)(op.rows ? op.box_w : op.tw * d.stride), (cuuint32_t)(op.th * d.stride), 1};
      cuuint32_t es[4] = {1, (cuuint32_t)d.stride, (cuuint32_t)d.stride, 1};
      int r = encode_map(&op.src_map, base + src.offset, 4, dims, str, box, es, "src");
      if (r) return r;
    }
    {
      cuuint64_t dims[3] = {(cuuint64_t)d.cin, (cuuint64_t)op.cout_pad, (cuuint64_t)taps * planes};
      cuuint64_t str[2] = {(cuuint64_t)d.cin * 2, (cuuint64_t)op.cout_pad * d.cin * 2};
      cuuint32_t box[3] = {64, (cuuint32_t)(op.n_tile / op.cluster), 1};
      cuuint32_t es[3] = {1, 1, 1};
      int r = encode_map(&op.w_map, base + op.w_offset, 3, dims, str, box, es, "weights");
      if (r) return r;
    }
    if (!dst.fp32_nchw && op.up == 2) {
      // [n][y][py][x][px*C + c] view of the double-resolution NHWC destination (sub-pixel px folded into dim 0)
      const cuuint64_t C2 = (cuuint64_t)dst.channels, ow = dst.w / 2, oh = dst.h / 2;
      cuuint64_t dims[5] = {2 * C2, ow, 2, oh, (cuuint64_t)e->batch * planes};
      cuuint64_t str[4] = {2 * C2 * 2, 2 * ow * C2 * 2, 4 * ow * C2 * 2, 4 * oh * ow * C2 * 2};
      cuuint32_t box[5] = {32, (cuuint32_t)op.store_w, 1, (cuuint32_t)op.store_h, 1};
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      int r = encode_map(&op.dst_map, base + dst.offset, 5, dims, str, box, es, "dst (x2)", CU_TENSOR_MAP_SWIZZLE_64B);
      if (r) return r;
    } else if (!dst.fp32_nchw) {
      // [plane][n][h][w][c] with the plane as its own (5th) dimension: ONE store request writes a warp's chunk to both planes
      // (the staging buffers of the two planes are adjacent) - the TMA unit's cost is per request, not per byte
      cuuint64_t dims[5] = {(cuuint64_t)dst.channels, (cuuint64_t)dst.w, (cuuint64_t)dst.h, (cuuint64_t)e->batch, (cuuint64_t)planes};
      cuuint64_t str[4] = {(cuuint64_t)dst.channels * 2, (cuuint64_t)dst.w * dst.channels * 2, (cuuint64_t)dst.h * dst.w * dst.channels * 2,
                           (cuuint64_t)dst.plane_elems * 2};
      cuuint32_t box[5] = {32, (cuuint32_t)op.store_w, (cuuint32_t)op.store_h, 1, (cuuint32_t)planes};      // 32 channels = 64-byte rows
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      int r = encode_map(&op.dst_map, base + dst.offset, 5, dims, str, box, es, "dst", CU_TENSOR_MAP_SWIZZLE_64B);
      if (r) return r;
    } else {
      op.dst_map = op.src_map;
    }
  }
  CNL_CUDA_CHECK(cudaStreamSynchronize(st));       // packed host weights are pageable: make the copies land before returning
  e->uploaded_arena = arena;
  return CNL_OK;
}

int cnl_engine_forward(cnl_engine* e, void* arena, const float* image, int first_op, int last_op, void* stream, int* launches) {
  return cnl_engine_forward_act(e, arena, image, first_op, last_op, -1, stream, launches);
}

int cnl_engine_forward_act(cnl_engine* e, void* arena, const float* image, int first_op, int last_op, int sigmoid_buffer,
                           void* stream, int* launches) {
  if (!e || !arena) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_engine_forward: null argument");
  if (sigmoid_buffer >= (int)e->bufs.size() || (sigmoid_buffer >= 0 && !e->bufs[sigmoid_buffer].fp32_nchw))
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_engine_forward_act: sigmoid_buffer must be an fp32 NCHW output buffer");
  if (arena != e->uploaded_arena) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_engine_forward: call cnl_engine_upload for this arena first");
  if (first_op < 0 || last_op > (int)e->ops.size() || first_op > last_op) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_engine_forward: bad op range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(arena);
  const int planes = e->planes;
  int n_launch = 0;
  auto launch_conv = [&](const ConvParams& p, const OpInfo& op) -> cudaError_t {
    const int groups = ((p.m_tiles + p.cluster - 1) / p.cluster) * p.n_tiles;
    const int grid = p.cluster * std::min(groups, e->num_sms / p.cluster);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kConvThreads);
    cfg.dynamicSmemBytes = kSmemLimit;
    cfg.stream = st;
    static const bool pdl = [] { const char* v = getenv("CNL_PDL"); return !(v && atoi(v) == 0); }();
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see the griddepcontrol pair in the kernel
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 2 : 1;
    if (op.rows) {
      cfg.gridDim = dim3(std::min(p.m_tiles, e->num_sms));
      return cudaLaunchKernelEx(&cfg, conv_tc_kernel<2, true, true, false>, op.src_map, op.w_map, op.dst_map, p);
    }
    if (op.pair) return cudaLaunchKernelEx(&cfg, conv_tc_kernel<2, true, false, true>, op.src_map, op.w_map, op.dst_map, p);
    if (e->precision == CNL_PRECISION_SPLIT && op.corr) return cudaLaunchKernelEx(&cfg, conv_tc_kernel<2, true, false, false>, op.src_map, op.w_map, op.dst_map, p);
    else if (e->precision == CNL_PRECISION_SPLIT)       return cudaLaunchKernelEx(&cfg, conv_tc_kernel<2, false, false, false>, op.src_map, op.w_map, op.dst_map, p);
    else if (e->precision == CNL_PRECISION_SPLIT_FUSED) return cudaLaunchKernelEx(&cfg, conv_tc_kernel<2, false, false, false>, op.src_map, op.w_map, op.dst_map, p);
    return cudaLaunchKernelEx(&cfg, conv_tc_kernel<1, false, false, false>, op.src_map, op.w_map, op.dst_map, p);
  };
  for (int i = first_op; i < last_op; ++i) {
    OpInfo& op = e->ops[i];
    const cnl_conv_desc& d = op.d;
    const BufferInfo& dst = e->bufs[d.dst];
    ConvParams p;
    p.n_img = e->batch;
    p.tw = op.tw; p.th = op.th; p.tiles_w = op.tiles_w; p.tiles_h = op.tiles_h;
    p.m_tiles = e->batch * op.tiles_w * op.tiles_h;
    p.n_tiles = op.n_tiles; p.n_tile = op.n_tile;
    p.relu = d.relu;
    p.sigmoid = (sigmoid_buffer >= 0 && d.dst == sigmoid_buffer) ? 1 : 0;
    p.wscale_inv = 1.0f / op.wscale;
    p.bias = reinterpret_cast<const float*>(base + op.bias_offset);
    static const int dbg_env = [] { const char* v = getenv("CNL_DEBUG_EPI"); return v ? atoi(v) : 0; }();
    p.debug = dbg_env;
    p.bias_in_params = (d.kind <= 1 && op.bias_packed.size() <= (size_t)kParamBias) ? 1 : 0;
    if (p.bias_in_params) std::memcpy(p.bias_c, op.bias_packed.data(), op.bias_packed.size() * sizeof(float));
    p.res = nullptr; p.res_up = 1; p.res_c = 0; p.res_h = 0; p.res_w = 0; p.res_plane_elems = 0;
    p.num_stages = op.num_stages; p.store_w = op.store_w; p.store_h = op.store_h; p.cluster = op.cluster;
    p.acc_stages = op.acc_stages; p.corr_off = op.corr_off; p.cat = op.cat;
    p.dst_up = 1; p.dst_phase = -1; p.dst_c = dst.channels;
    p.direct_store = direct_store_enabled() ? 1 : 0;
    static const bool early_on = [] { const char* v = getenv("CNL_EARLY_RELEASE"); return !(v && atoi(v) == 0); }();
    p.early_release = (early_on && d.kind == 0 && op.pair && op.acc_stages == 1 && op.n_tile == 256 && !dst.fp32_nchw) ? 1 : 0;
    p.dst = dst.fp32_nchw ? nullptr : reinterpret_cast<__half*>(base + dst.offset);
    p.dst_plane_elems = dst.plane_elems; p.dst_h = dst.h; p.dst_w = dst.w;
    p.a_slots = op.a_slots; p.a_slot_bytes = op.a_slot_bytes; p.box_w = op.box_w; p.stage_depth = op.stage_depth;
    p.w_resident = op.rows ? op.w_resident : 0;
    if (d.kind == 1) {
      if (!image) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_engine_forward: image pointer required for the stem");
      const int SH = e->height / 2, SW = e->width / 2;
      const long long half_plane = (long long)e->batch * SH * SW * 64;
      __half* T = reinterpret_cast<__half*>(base + op.stem_t_offset);
      __half* S = reinterpret_cast<__half*>(base + op.stem_s_offset);
      const long long npix = (long long)e->batch * SH * SW;
      const int b1 = (int)((npix + 255) / 256);
      if (op.stem_s2d) {
        const long long s2d_plane = (long long)e->batch * SH * (SW + 3) * 16;
        const int b0 = (int)(((long long)e->batch * SH * (SW + 3) + 255) / 256);
        if (planes == 2) stem_s2d_kernel<2><<<b0, 256, 0, st>>>(image, T, e->batch, e->height, e->width, s2d_plane);
        else             stem_s2d_kernel<1><<<b0, 256, 0, st>>>(image, T, e->batch, e->height, e->width, s2d_plane);
      } else {
        if (planes == 2) stem_im2row_kernel<2><<<b1, 256, 0, st>>>(image, T, e->batch, e->height, e->width, half_plane);
        else             stem_im2row_kernel<1><<<b1, 256, 0, st>>>(image, T, e->batch, e->height, e->width, half_plane);
      }
      p.out_h = SH; p.out_w = SW;
      p.kh = 4; p.kw = 1; p.stride = 1; p.pad_h = 2; p.pad_w = 0; p.kblocks = 1;
      p.src_c_off = 0; p.dst_c_off = 0; p.out_mode = 0; p.cout_real = 64; p.out_nchw = nullptr;
      p.dst = S; p.dst_plane_elems = half_plane; p.dst_h = SH; p.dst_w = SW; p.dst_c = 64;      // the un-pooled conv output
      CNL_CUDA_CHECK(launch_conv(p, op));
      const long long total = (long long)e->batch * (SH / 2) * (SW / 2) * 8;
      const int b2 = (int)((total + 255) / 256);
      __half* o = reinterpret_cast<__half*>(base + dst.offset);
      if (planes == 2) pool_planes_kernel<2><<<b2, 256, 0, st>>>(S, o, e->batch, SH, SW, half_plane, dst.plane_elems);
      else             pool_planes_kernel<1><<<b2, 256, 0, st>>>(S, o, e->batch, SH, SW, half_plane, dst.plane_elems);
      n_launch += 3;
      CNL_CUDA_CHECK(cudaGetLastError());
      continue;
    }
    if (d.kind >= 2) {
      const long long total = (long long)e->batch * dst.h * dst.w * (dst.channels / 8);
      const int blocks = (int)((total + 255) / 256);
      __half* o = reinterpret_cast<__half*>(base + dst.offset);
      const float* wf = reinterpret_cast<const float*>(base + op.w_offset);
      if (d.kind == 2) {
        const BufferInfo& src = e->bufs[d.src];
        const __half* in = reinterpret_cast<const __half*>(base + src.offset);
        if (planes == 2) dw3x3_kernel<2><<<blocks, 256, 0, st>>>(in, o, wf, p.bias, e->batch, src.h, src.w, dst.channels, d.stride, d.relu, src.plane_elems, dst.plane_elems);
        else             dw3x3_kernel<1><<<blocks, 256, 0, st>>>(in, o, wf, p.bias, e->batch, src.h, src.w, dst.channels, d.stride, d.relu, src.plane_elems, dst.plane_elems);
      } else if (d.kind == 3) {
        const int n_src = d.src3 >= 0 ? 3 : 2;
        const int last = n_src == 3 ? d.src3 : d.src2;
        const __half* a = reinterpret_cast<const __half*>(base + e->bufs[d.src].offset);
        const __half* b = reinterpret_cast<const __half*>(base + e->bufs[d.src2].offset);
        const __half* c = d.src3 >= 0 ? reinterpret_cast<const __half*>(base + e->bufs[d.src3].offset) : nullptr;
        if (planes == 2) fuse_kernel<2><<<blocks, 256, 0, st>>>(a, b, c, o, d.scale0, d.scale1, d.scale2, n_src, d.resize, e->batch, dst.h, dst.w, dst.channels, dst.plane_elems, e->bufs[last].plane_elems);
        else             fuse_kernel<1><<<blocks, 256, 0, st>>>(a, b, c, o, d.scale0, d.scale1, d.scale2, n_src, d.resize, e->batch, dst.h, dst.w, dst.channels, dst.plane_elems, e->bufs[last].plane_elems);
      } else {
        if (!image) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_engine_forward: image pointer required for the stem");
        if (planes == 2) stem3x3_kernel<2><<<blocks, 256, 0, st>>>(image, o, wf, p.bias, e->batch, e->height, e->width, dst.channels, d.relu, dst.plane_elems);
        else             stem3x3_kernel<1><<<blocks, 256, 0, st>>>(image, o, wf, p.bias, e->batch, e->height, e->width, dst.channels, d.relu, dst.plane_elems);
      }
      ++n_launch;
      CNL_CUDA_CHECK(cudaGetLastError());
      continue;
    }
    p.out_h = dst.h / op.up; p.out_w = dst.w / op.up;
    p.dst_up = op.up; p.dst_phase = d.dst_phase;
    p.kh = op.kh; p.kw = op.kw; p.stride = d.stride; p.pad_h = op.pad_h; p.pad_w = op.pad_w; p.kblocks = d.cin / 64;
    p.src_c_off = d.src_c_off; p.dst_c_off = d.dst_c_off;
    p.out_mode = dst.fp32_nchw ? 1 : 0;
    p.cout_real = d.cout;
    p.out_nchw = dst.fp32_nchw ? reinterpret_cast<float*>(base + dst.offset) : nullptr;
    if (d.residual >= 0) {
      const BufferInfo& rb = e->bufs[d.residual];
      p.res = reinterpret_cast<const __half*>(base + rb.offset);
      p.res_up = d.residual_up; p.res_c = rb.channels; p.res_h = rb.h; p.res_w = rb.w; p.res_plane_elems = rb.plane_elems;
    }
    CNL_CUDA_CHECK(launch_conv(p, op));
    ++n_launch;
    CNL_CUDA_CHECK(cudaGetLastError());
  }
  if (launches) *launches = n_launch;
  return CNL_OK;
}

int cnl_engine_read_buffer(cnl_engine* e, void* arena, int buffer, float* out_nchw, void* stream) {
  if (!e || !arena || !out_nchw || buffer <= 0 || buffer >= (int)e->bufs.size()) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_engine_read_buffer: bad arguments");
  const BufferInfo& b = e->bufs[buffer];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(arena);
  if (b.fp32_nchw) {
    CNL_CUDA_CHECK(cudaMemcpyAsync(out_nchw, base + b.offset, b.bytes, cudaMemcpyDeviceToDevice, st));
    return CNL_OK;
  }
  const long long total = b.plane_elems;
  planes_to_nchw_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const __half*>(base + b.offset), out_nchw,
                                                                    e->batch, b.channels, b.h, b.w, e->planes, b.plane_elems);
  CNL_CUDA_CHECK(cudaGetLastError());
  return CNL_OK;
}

int cnl_engine_write_buffer(cnl_engine* e, void* arena, int buffer, const float* in_nchw, void* stream) {
  if (!e || !arena || !in_nchw || buffer <= 0 || buffer >= (int)e->bufs.size()) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_engine_write_buffer: bad arguments");
  const BufferInfo& b = e->bufs[buffer];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(arena);
  if (b.fp32_nchw) {
    CNL_CUDA_CHECK(cudaMemcpyAsync(base + b.offset, in_nchw, b.bytes, cudaMemcpyDeviceToDevice, st));
    return CNL_OK;
  }
  const long long total = b.plane_elems;
  nchw_to_planes_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(in_nchw, reinterpret_cast<__half*>(base + b.offset), e->batch,
                                                                    b.channels, b.h, b.w, e->planes, b.plane_elems);
  CNL_CUDA_CHECK(cudaGetLastError());
  return CNL_OK;
}

}  // extern "C"
