// CTA-pair (tcgen05 cta_group::2) variant of the persistent warp-specialised GEMM engine for sm_100a.
//
//   acc[b][m, n] = sum over operand pairs q, k:  A_q[b][m, k] * B_q[b][n, k]      (bf16 x bf16 -> fp32 in TMEM)
//
// Why a second engine: with one CTA per 128 x 256 tile every MMA cycle needs 96 B of operands from L2 per SM; ncu
// shows the MMA issuer of the single-CTA engine waiting on TMA data (tensor pipe 47-67 % active).  Here two CTAs of a
// cluster (the two SMs of a TPC) compute one 256 x BN tile: each CTA loads its own 128 rows of A and only HALF of
// B's rows (BN/2), the tensor cores of both SMs read both halves -> 64 B / MMA-cycle / SM for BN = 256, and a
// shared-memory stage shrinks from 48 KB to 32 KB (deeper ring).
//
//   warp 0      : TMA producer   (both CTAs; bytes are reported to the LEADER CTA's `full` barrier)
//   warp 1      : MMA issuer     (leader CTA only: tcgen05.mma.cta_group::2, M = 256, N = BN, K = 16; commits are
//                                 multicast to the `empty` / `tmem full` barriers of both CTAs)
//   warp 2      : TMEM allocator (512 columns per CTA = 2 accumulator buffers)
//   warps 4..11 : epilogue       (each CTA drains its own 128 accumulator rows; `tmem empty` arrives at the leader)
//
// Epilogue functor interface (shared with gemm_engine.cuh):
//   tile_begin(b, m, n0)                        per tile, before the accumulator is waited for
//   Side pre(b, m, n)              [optional]   issue the loads of chunk n's side data; the engine calls it one chunk
//                                               AHEAD (across tiles too) so their latency hides behind a chunk of math
//   chunk(b, m, n, v[32] [, side])              32 accumulator columns of row m
//   chunk2(b, m, n, x[32], d[32] [, side])      kDual: both accumulators
//   tile_end(b, m, n0, tile_n, half)
//   kTmaOut: v[] holds the output chunk afterwards; every epilogue warp stages its own [32 x 32] bf16 chunk in a
//            private double-buffered 2 KB area (SWIZZLE_64B) and writes it with its own TMA store: no cross-warp
//            barrier anywhere in the epilogue.
#pragma once
#include "gemm_engine.cuh"

namespace eng2 {

using eng::OperandMaps;
using eng::Problem;
using eng::OutDesc;
using eng::epi_dual;
using eng::epi_tma_out;
using eng::epi_has_side;
using eng::ksteps_of;

constexpr int BM = 128;            // accumulator rows per CTA; a CTA pair computes 2 * BM rows
constexpr int BK = 64;
constexpr int kThreads = 384;          // default: 4 control warps + 8 epilogue warps
constexpr int threads_for(int epi_warps) { return 128 + 32 * epi_warps; }
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr int kEpiBarId = 1;
constexpr int kSmemBudget = 227 * 1024;

template <int BN, bool DUAL = false, bool TMA_OUT = false, bool TMA_OUT2 = false, bool CHUNK_IN = false, int kEpiWarps = 8>
struct SmemLayout {
  static constexpr int kA1Bytes = BM * BK * 2;
  static constexpr int kABytes = kA1Bytes * (DUAL ? 2 : 1);
  static constexpr int kBBytes = (BN / 2) * BK * 2;               // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutChunkBytes = 32 * 64;                  // one epilogue warp's [32 rows][32 bf16] chunk
  static constexpr int kOut1Bytes = TMA_OUT ? kEpiWarps * 2 * kOutChunkBytes : 0;  // double-buffered per warp
  static constexpr int kOut2Bytes = TMA_OUT2 ? kEpiWarps * 2 * kOutChunkBytes : 0;
  static constexpr int kInBytes = CHUNK_IN ? kEpiWarps * 2 * kOutChunkBytes : 0;
  static constexpr int kOutBytes = kOut1Bytes + kOut2Bytes + kInBytes;
  static constexpr int kBarrierBytes = 1024;
  static constexpr int kAvail = kSmemBudget - 1024 /*align slack*/ - kBarrierBytes - kOutBytes;
  static constexpr int kStagesRaw = kAvail / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTotal = kStages * kStageBytes + kOutBytes + kBarrierBytes + 1024;
  static_assert(kStages >= 3, "smem ring too shallow");
  static_assert(kStageBytes % 1024 == 0, "stages must keep the 1 KB swizzle-atom alignment");
};

struct TileCoord {
  int b, m0, n0, tn;
  bool valid;
};

template <int BN>
__device__ __forceinline__ TileCoord tile_coord(const Problem& pb, int tile, int total_tiles, int tiles_per_batch,
                                                int rank) {
  TileCoord c;
  c.valid = tile < total_tiles;
  const int t = c.valid ? tile : 0;
  const int bl = t / tiles_per_batch;
  const int rem = t - bl * tiles_per_batch;
  c.b = pb.reverse ? pb.batches - 1 - bl : bl;
  c.m0 = (rem / pb.tiles_n) * (2 * BM) + rank * BM;
  c.tn = rem % pb.tiles_n;
  c.n0 = c.tn * BN;
  return c;
}

// columns the MMA of this tile computes: the N tail is rounded up to the cta_group::2 granularity of 16
template <int BN>
__device__ __forceinline__ int tile_ncols(const Problem& pb, int n0) {
  const int nrem = pb.N - n0;
  return nrem >= BN ? BN : ((nrem + 15) & ~15);
}

template <int BN, bool A_MN, bool B_MN, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(threads_for(eng::epi_warps<Epi>::value), 1)
gemm2_kernel(const __grid_constant__ OperandMaps maps, const Problem pb, const typename Epi::Params ep) {
  constexpr bool DUAL = epi_dual<Epi>::value;
  constexpr bool TMA_OUT = epi_tma_out<Epi>::value;
  constexpr bool HAS_SIDE = epi_has_side<Epi>::value;
  constexpr bool TMA_OUT2 = eng::epi_tma_out2<Epi>::value;
  constexpr bool CHUNK_IN = eng::epi_chunk_in<Epi>::value;
  constexpr bool TRANSPOSED = eng::epi_transposed<Epi>::value;
  static_assert(!TRANSPOSED || (CHUNK_IN && TMA_OUT && !TMA_OUT2), "transposed epilogues: chunk-in + one output");
  static_assert(!TMA_OUT2 || TMA_OUT, "a second output needs the first");
  static_assert(!CHUNK_IN || (HAS_SIDE && !DUAL), "chunk-in: uniform single-accumulator functors only");
  static_assert(!TMA_OUT2 || HAS_SIDE, "second output: functors with side data only");
  static_assert(BN % 32 == 0 && BN <= 256, "BN: multiple of 32, at most 256");
  static_assert(!B_MN || (BN % 128 == 0), "MN-major B: each CTA's half must be whole 64-wide swizzle groups");
  static_assert(!DUAL || (BN <= 128 && !A_MN), "dual accumulators: BN <= 128, K-major A operands");
  constexpr int kAccCols = DUAL ? 2 * BN : BN;
  constexpr int kEpiWarps = eng::epi_warps<Epi>::value;     // shadows the namespace default
  constexpr int NPH = kEpiWarps / 4;                        // column phases: warp (q4, phase) takes chunks phase, phase + NPH, ...
  static_assert(kEpiWarps == 8 || kEpiWarps == 12, "8 or 12 epilogue warps");
  using L = SmemLayout<BN, DUAL, TMA_OUT, TMA_OUT2, CHUNK_IN, kEpiWarps>;
  constexpr int kStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* out_smem = smem + kStages * L::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_smem + L::kOutBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;        // [2]  (the leader's copy is the live one)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* in_bar = tempty_bar + 4;           // [kEpiWarps][2]  per-warp "input chunk landed" barriers (kChunkIn)
  uint8_t* out2_smem = out_smem + L::kOut1Bytes;
  uint8_t* in_smem = out2_smem + L::kOut2Bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(ptx::cluster_ctarank());
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int total_tiles = pb.batches * pb.tiles_m * pb.tiles_n;
  const int tiles_per_batch = pb.tiles_m * pb.tiles_n;

  if (warp == 0 && lane == 0) {
    for (int q = 0; q < pb.num_pairs; ++q) {
      ptx::prefetch_tmap(&maps.a[q]);
      ptx::prefetch_tmap(&maps.b[q]);
    }
    if constexpr (DUAL) ptx::prefetch_tmap(&maps.a[1]);
    if constexpr (TMA_OUT) ptx::prefetch_tmap(&maps.out);
    if constexpr (TMA_OUT2) ptx::prefetch_tmap(&maps.out2);
    if constexpr (CHUNK_IN) ptx::prefetch_tmap(&maps.in);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);       // leader's producer arrives (expect_tx covers both CTAs' bytes)
      ptx::mbar_init(&empty_bar[s], 1);      // one multicast commit per use
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], 2 * kEpiWarps);   // every epilogue warp of both CTAs
    }
    if constexpr (CHUNK_IN) {
      for (int i = 0; i < 2 * kEpiWarps; ++i) ptx::mbar_init(&in_bar[i], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2sm(tmem_base_slot, 512);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();       // barriers initialised and TMEM allocated in BOTH CTAs before any remote signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, cluster sync) may overlap the tail
  // of the previous kernel in the stream; nothing below may start before that kernel has completed and flushed.
  // The dependents of THIS kernel may be scheduled as soon as every CTA has got here (they block at the same point).
  // Both instructions are no-ops when the launch carries no programmatic-serialisation attribute.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full_leader0 = ptx::mapa(ptx::smem_u32(&full_bar[0]), 0);   // shared::cluster address in the leader
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const TileCoord tc = tile_coord<BN>(pb, tile, total_tiles, tiles_per_batch, rank);
        const int nb0 = tc.n0 + rank * (tile_ncols<BN>(pb, tc.n0) >> 1);   // first B row of this CTA's half
        for (int q = 0; q < pb.num_pairs; ++q) {
          const int nks = ksteps_of(pb, q, tc.b);
          // No integer division in this loop: on this single thread each one costs ~150 cycles, which (with the barrier
          // wait and the TMA issues) bounded the ring at ~500 cycles per k-step -- above the 384 / 256 cycles an N = 192 /
          // 128 k-step of MMAs takes.  `subl` / `kin` split ks = subl * ksub + kin incrementally.
          int subl = 0, kin = 0;
          const int ksub = pb.ksub[q];
          const int kbase = tc.b * pb.k_boff[q];
          const int sub0 = tc.b * pb.sub_per_batch[q];
          const int ab0 = tc.b * pb.a_bmul[q], bb0 = tc.b * pb.b_bmul[q];
          const int asm_ = pb.a_smul[q], bsm_ = pb.b_smul[q];
          for (int ks = 0; ks < nks; ++ks) {
            const int k0 = kin * BK + kbase;
            const int sub = subl + sub0;
            const int ab = ab0 + sub * asm_;
            const int bb = bb0 + sub * bsm_;
            if (++kin == ksub) { kin = 0; ++subl; }
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * L::kStageBytes;
            uint8_t* sb = sa + L::kABytes;
            const uint32_t fb = full_leader0 + static_cast<uint32_t>(stage) * 8u;
            if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
            if constexpr (DUAL) {
              const int ab1 = tc.b * pb.a_bmul[1];
              ptx::tma_load_3d_2sm(sa, &maps.a[0], fb, k0, tc.m0, ab);
              ptx::tma_load_3d_2sm(sa + L::kA1Bytes, &maps.a[1], fb, k0, tc.m0, ab1);
            } else if constexpr (!A_MN) {
              ptx::tma_load_3d_2sm(sa, &maps.a[q], fb, k0, tc.m0, ab);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                ptx::tma_load_3d_2sm(sa + j * 8192, &maps.a[q], fb, tc.m0 + j * 64, k0, ab);
            }
            if constexpr (!B_MN) {
              ptx::tma_load_3d_2sm(sb, &maps.b[q], fb, k0, nb0, bb);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j)
                ptx::tma_load_3d_2sm(sb + j * 8192, &maps.b[q], fb, nb0 + j * 64, k0, bb);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
        const TileCoord tc = tile_coord<BN>(pb, tile, total_tiles, tiles_per_batch, 0);
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        ptx::mbar_wait(&tempty_bar[buf], bphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * kAccCols;
        const uint32_t idesc = ptx::umma_idesc_bf16(2 * BM, tile_ncols<BN>(pb, tc.n0), A_MN ? 1 : 0, B_MN ? 1 : 0);
        uint32_t accum = 0;
        for (int q = 0; q < pb.num_pairs; ++q) {
          const int nks = ksteps_of(pb, q, tc.b);
          for (int ks = 0; ks < nks; ++ks) {
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(smem + stage * L::kStageBytes);
            const uint32_t sb = sa + L::kABytes;
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              const uint64_t ad = A_MN ? ptx::umma_desc(sa + kk * 2048, 8192, 1024) : ptx::umma_desc(sa + kk * 32, 16, 1024);
              const uint64_t bd = B_MN ? ptx::umma_desc(sb + kk * 2048, 8192, 1024) : ptx::umma_desc(sb + kk * 32, 16, 1024);
              ptx::mma_bf16_ss_2sm(tmem_d, ad, bd, idesc, accum);
              if constexpr (DUAL) {
                const uint64_t ad1 = ptx::umma_desc(sa + L::kA1Bytes + kk * 32, 16, 1024);
                ptx::mma_bf16_ss_2sm(tmem_d + BN, ad1, bd, idesc, accum);
              }
              accum = 1;
            }
            ptx::mma_commit_2sm(&empty_bar[stage], 3);   // frees the stage in both CTAs once these MMAs retire
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
        ptx::mma_commit_2sm(&tfull_bar[buf], 3);         // accumulator (both CTAs' halves) ready
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows each)
    const int q4 = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;
    const uint32_t tempty_leader[2] = {ptx::mapa(ptx::smem_u32(&tempty_bar[0]), 0),
                                       ptx::mapa(ptx::smem_u32(&tempty_bar[1]), 0)};
    Epi epi(ep);
    int it = 0;
    int slab = 0;
#ifdef CLIPK_EPI_PROF
    long long pf_wait = 0, pf_ld = 0, pf_math = 0, pf_sts = 0, pf_tma = 0, pf_t0 = clock64(), pf_chunks = 0;
#define PF_MARK(acc) { const long long _t = clock64(); acc += _t - pf_t; pf_t = _t; }
#else
#define PF_MARK(acc)
#endif
    TileCoord tc = tile_coord<BN>(pb, cluster_id, total_tiles, tiles_per_batch, rank);
    [[maybe_unused]] typename eng::side_of<Epi>::type side{};
    const int ew = warp - kEpiWarp0;
    // kChunkIn: TMA-load this warp's [32 x 32] bf16 box of the input tensor for chunk number `g` into buffer g & 1
    auto issue_in = [&](int g, int bb, int mrow0, int ncol) {
      if constexpr (CHUNK_IN) {
        if (lane == 0) {
          uint64_t* bar = &in_bar[ew * 2 + (g & 1)];
          ptx::mbar_arrive_expect_tx(bar, L::kOutChunkBytes);
          if constexpr (TRANSPOSED)   // the input tensor is laid out like the output: [accumulator column][accumulator row]
            ptx::tma_load_3d(in_smem + (ew * 2 + (g & 1)) * L::kOutChunkBytes, &maps.in, bar, mrow0, ncol, bb);
          else
            ptx::tma_load_3d(in_smem + (ew * 2 + (g & 1)) * L::kOutChunkBytes, &maps.in, bar, ncol, mrow0, bb);
        }
      }
    };
    if (tc.valid) {
      if constexpr (HAS_SIDE) side = epi.pre(tc.b, tc.m0 + q4 * 32 + lane, tc.n0 + half * 32);
      issue_in(0, tc.b, tc.m0 + q4 * 32, tc.n0 + half * 32);
    }
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
      const TileCoord nx = tile_coord<BN>(pb, tile + num_clusters, total_tiles, tiles_per_batch, rank);
      const int b = tc.b, m0 = tc.m0, n0 = tc.n0;
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      const int m = m0 + q4 * 32 + lane;
      epi.tile_begin(b, m, n0);
#ifdef CLIPK_EPI_PROF
      long long pf_t = clock64();
#endif
      ptx::mbar_wait(&tfull_bar[buf], bphase);
      ptx::tc_fence_after();
      PF_MARK(pf_wait)
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + buf * kAccCols;
      // chunks (32 columns) of this tile that hold at least one existing column, and this warp's share of them
      const int nch_all = (pb.N - n0 + 31) >> 5;
      const int nch = nch_all < BN / 32 ? nch_all : BN / 32;
      const int jmax_raw = (nch - half + NPH - 1) / NPH;
      const int jmax = jmax_raw > 0 ? jmax_raw : 1;          // every warp walks at least one chunk (TMA clips it)
      [[maybe_unused]] typename eng::side_of<Epi>::type side_next{};
      // side data / input chunk of the chunk after (tile, j): next chunk of this tile, else the first chunk of this
      // CTA's next tile
      auto prefetch_next = [&](int j) {
        if (j + 1 < jmax) {
          if constexpr (HAS_SIDE) side_next = epi.pre(b, m, n0 + (NPH * (j + 1) + half) * 32);
          issue_in(slab + 1, b, m0 + q4 * 32, n0 + (NPH * (j + 1) + half) * 32);
        } else if (nx.valid) {
          if constexpr (HAS_SIDE) side_next = epi.pre(nx.b, nx.m0 + q4 * 32 + lane, nx.n0 + half * 32);
          issue_in(slab + 1, nx.b, nx.m0 + q4 * 32, nx.n0 + half * 32);
        }
      };
      auto release_tmem = [&]() {      // accumulator drained: hand the TMEM buffer back to the leader's MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader[buf]);
      };
      // Per-warp staging of the output chunk(s), no cross-warp barrier: this warp's [32 rows x 32 cols] bf16 chunk
      // goes through one of its two private 2 KB buffers (64-byte rows, SWIZZLE_64B: 16-byte unit u of row r sits
      // at u ^ ((r >> 1) & 3)) and leaves with its own TMA store.  Lane 0 waits (after committing the stores of chunk
      // g) until those of chunk g-1 have been read, so the other buffer is free before anybody writes chunk g+1.
      const uint32_t sw = (static_cast<uint32_t>(lane) >> 1) & 3u;
      auto stage_rows = [&](uint8_t* wbuf, const float* v) {
        const uint32_t rowbase = ptx::smem_u32(wbuf) + lane * 64;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t addr = rowbase + ((static_cast<uint32_t>(t) ^ sw) << 4);
          ptx::st_shared_v4(addr, ptx::pack_bf16x2(v[8 * t + 0], v[8 * t + 1]), ptx::pack_bf16x2(v[8 * t + 2], v[8 * t + 3]),
                            ptx::pack_bf16x2(v[8 * t + 4], v[8 * t + 5]), ptx::pack_bf16x2(v[8 * t + 6], v[8 * t + 7]));
        }
      };
      auto store_chunk = [&](const float* v, [[maybe_unused]] const float* v2, int c) {
        if constexpr (TMA_OUT) {
          uint8_t* wbuf = out_smem + (ew * 2 + (slab & 1)) * L::kOutChunkBytes;
          uint8_t* wbuf2 = out2_smem + (ew * 2 + (slab & 1)) * L::kOutChunkBytes;
          stage_rows(wbuf, v);
          if constexpr (TMA_OUT2) stage_rows(wbuf2, v2);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          PF_MARK(pf_sts)
          if (lane == 0) {
            ptx::tma_store_3d(&maps.out, wbuf, n0 + c * 32, m0 + q4 * 32, b);
            if constexpr (TMA_OUT2) ptx::tma_store_3d(&maps.out2, wbuf2, n0 + c * 32, m0 + q4 * 32, b);
            ptx::bulk_commit_group();
            ptx::bulk_wait_group_read<1>();
          }
          __syncwarp();
          PF_MARK(pf_tma)
#ifdef CLIPK_EPI_PROF
          ++pf_chunks;
#endif
        }
      };
      if constexpr (DUAL) {
#pragma unroll 1
        for (int j = 0; j < jmax; ++j) {
          const int c = NPH * j + half;
          float v[32], v1[32];
          ptx::tmem_ld_32x32(tacc + c * 32, v);
          ptx::tmem_ld_32x32(tacc + BN + c * 32, v1);
          prefetch_next(j);
          ptx::tmem_ld_wait();
          PF_MARK(pf_ld)
          if (j == jmax - 1) release_tmem();
          if constexpr (HAS_SIDE) epi.chunk2(b, m, n0 + c * 32, v, v1, side);
          else epi.chunk2(b, m, n0 + c * 32, v, v1);
          PF_MARK(pf_math)
          if constexpr (HAS_SIDE) side = side_next;
          store_chunk(v, v1, c);          // kTmaOut2: the functor left the second output chunk in v1
          ++slab;
        }
      } else {
        // single accumulator: the TMEM load of chunk j+1 is in flight while chunk j is processed (ping-pong registers)
        auto process = [&](float* v, int j) {
          const int c = NPH * j + half;
          if (c * 32 >= pb.N - n0) {       // (warp-uniform) chunk past the tile's last column: the MMA never wrote it
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
          prefetch_next(j);
          if constexpr (TRANSPOSED) {
            // the functor reads its input chunk from / writes its output chunk to the staging buffers itself
            // ([32 output rows][32 output cols] bf16, SWIZZLE_64B), one element per (row j, this lane's column)
            ptx::mbar_wait(&in_bar[ew * 2 + (slab & 1)], (slab >> 1) & 1);
            uint8_t* wbuf = out_smem + (ew * 2 + (slab & 1)) * L::kOutChunkBytes;
            epi.chunk_t(b, m, n0 + c * 32, v, side, ptx::smem_u32(in_smem + (ew * 2 + (slab & 1)) * L::kOutChunkBytes),
                        ptx::smem_u32(wbuf), lane);
            PF_MARK(pf_math)
            side = side_next;
            ptx::fence_proxy_async_smem();
            __syncwarp();                      // input buffer read by everybody, output chunk staged by everybody
            if (lane == 0) {
              ptx::tma_store_3d(&maps.out, wbuf, m0 + q4 * 32, n0 + c * 32, b);
              ptx::bulk_commit_group();
              ptx::bulk_wait_group_read<1>();
            }
            __syncwarp();
          } else if constexpr (TMA_OUT2 || CHUNK_IN) {
            [[maybe_unused]] uint32_t in[16];
            [[maybe_unused]] float v2[32];
            if constexpr (CHUNK_IN) {
              ptx::mbar_wait(&in_bar[ew * 2 + (slab & 1)], (slab >> 1) & 1);
              const uint32_t rowbase = ptx::smem_u32(in_smem + (ew * 2 + (slab & 1)) * L::kOutChunkBytes) + lane * 64;
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const uint4 q = ptx::ld_shared_v4(rowbase + ((static_cast<uint32_t>(t) ^ sw) << 4));
                in[4 * t] = q.x; in[4 * t + 1] = q.y; in[4 * t + 2] = q.z; in[4 * t + 3] = q.w;
              }
              __syncwarp();                    // everybody has read the buffer before lane 0 re-arms it (next chunk + 1)
            }
            epi.chunk(b, m, n0 + c * 32, v, side, in, v2);
            PF_MARK(pf_math)
            side = side_next;
            store_chunk(v, v2, c);
          } else {
            if constexpr (HAS_SIDE) epi.chunk(b, m, n0 + c * 32, v, side);
            else epi.chunk(b, m, n0 + c * 32, v);
            PF_MARK(pf_math)
            if constexpr (HAS_SIDE) side = side_next;
            store_chunk(v, nullptr, c);
          }
          ++slab;
        };
        float va[32], vb[32];
        ptx::tmem_ld_32x32(tacc + half * 32, va);
#pragma unroll 1
        for (int j = 0; j < jmax; j += 2) {
          ptx::tmem_ld_wait();
          PF_MARK(pf_ld)
          if (j + 1 < jmax) ptx::tmem_ld_32x32(tacc + (NPH * (j + 1) + half) * 32, vb);
          else release_tmem();
          process(va, j);
          if (j + 1 < jmax) {
            ptx::tmem_ld_wait();
            PF_MARK(pf_ld)
            if (j + 2 < jmax) ptx::tmem_ld_32x32(tacc + (NPH * (j + 2) + half) * 32, va);
            else release_tmem();
            process(vb, j + 1);
          }
        }
      }
      epi.tile_end(b, m, n0, tc.tn, half);
      tc = nx;
    }
    if constexpr (TMA_OUT) {
      if (lane == 0) ptx::bulk_wait_group<0>();
    }
#ifdef CLIPK_EPI_PROF
    if (blockIdx.x == 2 && lane == 0 && (warp == kEpiWarp0 || warp == kEpiWarp0 + 5))
      printf("epi prof warp %d: tiles %d chunks %lld total %lld cyc | per chunk: wait_tfull %lld ld %lld math %lld sts+fence %lld "
             "tma+waitread %lld\n", warp, it, pf_chunks, clock64() - pf_t0, pf_wait / (pf_chunks ? pf_chunks : 1),
             pf_ld / (pf_chunks ? pf_chunks : 1), pf_math / (pf_chunks ? pf_chunks : 1),
             pf_sts / (pf_chunks ? pf_chunks : 1), pf_tma / (pf_chunks ? pf_chunks : 1));
#endif
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();       // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace eng2
