// Persistent warp-specialised tcgen05 GEMM engine for sm_100a.
//
//   acc[b][m, n] = sum over operand pairs q, k:  A_q[b][m, k] * B_q[b][n, k]      (bf16 x bf16 -> fp32 in TMEM)
//
// then a fused EPILOGUE functor consumes the fp32 accumulator straight out of TMEM (one thread = one output row,
// 32 columns per tcgen05.ld) — nothing but what the functor writes reaches HBM.
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor.3d, SWIZZLE_128B tiles, mbarrier complete_tx)
//   warp 1      : MMA issuer     (one elected thread, tcgen05.mma cta_group::1, M=128, N=BN, K=16)
//   warp 2      : TMEM allocator (512 columns = 2 accumulator buffers of BN<=256 columns)
//   warps 4..11 : epilogue       (tcgen05.ld 32x32b.x32; warp w owns TMEM lanes 32*(w%4)..+31; the two warps of a
//                                 quadrant take alternate 32-column chunks, so each SMSP has two epilogue warps
//                                 in flight to hide TMEM / global-load latency)
//
// Operands may be K-major (k contiguous in global memory: tensor map dims (k, row, batch)) or MN-major (row index
// contiguous: tensor map dims (row, k, batch)); both land in shared memory as 128-byte swizzled rows, see
// ptx::umma_desc.  Tiles: BM=128 x BN x BK=64, kStages-deep ring of (A,B) stages.
//
// Two optional engine features are selected by traits of the epilogue functor:
//   Epi::kTmaOut = true : the functor leaves one bf16 output row-chunk per call in v[]; the engine packs it into a
//                         128B-swizzled [128 x 64] shared-memory slab (double-buffered) and one elected thread writes
//                         the slab with cp.async.bulk.tensor (TMA store: full-line coalesced writes, rows / columns
//                         outside the output extent are clipped by the tensor map).
//   Epi::kDual = true   : TWO A operands share one B operand and accumulate into two TMEM accumulators
//                         (acc0 = A0 B^T, acc1 = A1 B^T); the functor's chunk2() sees both (BN <= 128).
#pragma once
#include <type_traits>

#include "ptx.cuh"

namespace eng {

template <class E, class = void>
struct epi_tma_out : std::false_type {};
template <class E>
struct epi_tma_out<E, std::void_t<decltype(E::kTmaOut)>> : std::bool_constant<E::kTmaOut> {};
template <class E, class = void>
struct epi_dual : std::false_type {};
template <class E>
struct epi_dual<E, std::void_t<decltype(E::kDual)>> : std::bool_constant<E::kDual> {};

template <class E, class = void>
struct epi_tma_out2 : std::false_type {};
template <class E>
struct epi_tma_out2<E, std::void_t<decltype(E::kTmaOut2)>> : std::bool_constant<E::kTmaOut2> {};
template <class E, class = void>
struct epi_chunk_in : std::false_type {};
template <class E>
struct epi_chunk_in<E, std::void_t<decltype(E::kChunkIn)>> : std::bool_constant<E::kChunkIn> {};

// Epi::kTransposed: the accumulator tile is the TRANSPOSE of the stored output (rows of the accumulator = columns
// of the output tensor); input chunks and stores use swapped coordinates and the functor moves the data itself
template <class E, class = void>
struct epi_transposed : std::false_type {};
template <class E>
struct epi_transposed<E, std::void_t<decltype(E::kTransposed)>> : std::bool_constant<E::kTransposed> {};

// Epi::kEpiWarps (optional, CTA-pair engine): epilogue warps per CTA, 8 (default) or 12.  Twelve = three warps per
// SMSP, for latency-bound epilogues whose tiles are 192 columns wide (6 chunks: two per column phase); the register
// budget drops from 168 to 128 per thread.  (Measured on the all-pairs activation kernels: slower than 8 -- fewer
// smem stages and serialised math cost more than the extra warps hide -- so nothing selects 12 at present.)
template <class E, class = void>
struct epi_warps : std::integral_constant<int, 8> {};
template <class E>
struct epi_warps<E, std::void_t<decltype(E::kEpiWarps)>> : std::integral_constant<int, E::kEpiWarps> {};

// Epi::Side (optional): per-chunk side data that `Side pre(b, m, n)` loads ahead of time and chunk() consumes
template <class E, class = void>
struct epi_has_side : std::false_type {};
template <class E>
struct epi_has_side<E, std::void_t<typename E::Side>> : std::true_type {};
struct NoSide {};
template <class E, class = void>
struct side_of { using type = NoSide; };
template <class E>
struct side_of<E, std::void_t<typename E::Side>> { using type = typename E::Side; };

// description of the TMA-stored output (member `out` of Epi::Params when Epi::kTmaOut)
struct OutDesc {
  void* ptr;
  int64_t ld, stride;      // elements (2-byte): row stride, batch stride
  int rows, cols, batches; // extents: rows / cols beyond them are clipped
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 384;      // 4 control warps + 8 epilogue warps (2 per TMEM lane quadrant / SMSP)
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr int kEpiBarId = 1;       // named barrier of the 8 epilogue warps (TMA-out staging)
constexpr int kSmemBudget = 227 * 1024;
// L2-prefetch distance of the producer (k-steps ahead of the load cursor).  Measured on B200: prefetching doubles the
// TMA request rate and costs 20-25% on every shape (4096^3: 1290 -> 980 TFLOP/s), so it is disabled (0).
constexpr int kPrefetchDist = 0;

struct OperandMaps {
  CUtensorMap a[2];
  CUtensorMap b[2];
  CUtensorMap out;
  CUtensorMap out2;   // second bf16 output of the same extent   (Epi::kTmaOut2, CTA-pair engine only)
  CUtensorMap in;     // bf16 input of the output's extent, one [32 x 32] box per epilogue chunk (Epi::kChunkIn)
};

struct Problem {
  int M, N;            // output extent per batch item
  int batches;
  int tiles_m, tiles_n;
  int num_pairs;       // 1 or 2 operand pairs accumulated into the same tile
  int ksteps[2];       // BK-steps per pair
  int ksub[2];         // BK-steps per "sub-batch" (k index folds a second batch index; = ksteps when unused)
  int a_bmul[2], a_smul[2];  // A batch coordinate = b * a_bmul + sub * a_smul
  int b_bmul[2], b_smul[2];
  int sub_per_batch[2];      // split-K over sub-batches: sub index = b * sub_per_batch + ks / ksub
  int sub_total[2];          // >0: batch b only covers sub-batches [b*spb, min((b+1)*spb, sub_total)) (uneven split)
  int reverse;               // walk the batch index downwards: consecutive launches alternate direction so that a
                             // launch first touches what the previous one wrote last (still L2-resident)
  int k_boff[2];             // CTA-pair engine only: batch b starts at k = b * k_boff (plain split-K of one long K:
                             // the maps keep the TRUE k extent, so the last split's tail is zero-filled by TMA)
};

template <int BN, bool DUAL = false, bool TMA_OUT = false>
struct SmemLayout {
  static constexpr int kA1Bytes = BM * BK * 2;
  static constexpr int kABytes = kA1Bytes * (DUAL ? 2 : 1);
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutSlabBytes = BM * 128;                  // [128 rows][64 bf16]
  static constexpr int kOutBytes = TMA_OUT ? 2 * kOutSlabBytes : 0;
  static constexpr int kBarrierBytes = 1024;
  static constexpr int kAvail = kSmemBudget - 1024 /*align slack*/ - kBarrierBytes - kOutBytes;
  static constexpr int kStagesRaw = kAvail / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTotal = kStages * kStageBytes + kOutBytes + kBarrierBytes + 1024;
  static_assert(kStages >= 3, "smem ring too shallow");
};

__device__ __forceinline__ int ksteps_of(const Problem& pb, int q, int b) {
  if (pb.sub_total[q] <= 0) return pb.ksteps[q];
  int subs = pb.sub_total[q] - b * pb.sub_per_batch[q];
  subs = subs < pb.sub_per_batch[q] ? subs : pb.sub_per_batch[q];
  return (subs > 0 ? subs : 0) * pb.ksub[q];
}

template <int BN, bool A_MN, bool B_MN, class Epi>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ OperandMaps maps, const Problem pb, const typename Epi::Params ep) {
  constexpr bool DUAL = epi_dual<Epi>::value;
  constexpr bool TMA_OUT = epi_tma_out<Epi>::value;
  static_assert(!DUAL || (BN <= 128 && !A_MN), "dual accumulators: BN <= 128, K-major A operands");
  constexpr int kAccCols = DUAL ? 2 * BN : BN;          // TMEM columns of one accumulator buffer
  using L = SmemLayout<BN, DUAL, TMA_OUT>;
  constexpr int kStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* out_smem = smem + kStages * L::kStageBytes;  // 1024-aligned (stage sizes are multiples of 1 KB)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_smem + L::kOutBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;        // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = pb.batches * pb.tiles_m * pb.tiles_n;
  const int tiles_per_batch = pb.tiles_m * pb.tiles_n;

  if (warp == 0 && lane == 0) {
    for (int q = 0; q < pb.num_pairs; ++q) {
      ptx::prefetch_tmap(&maps.a[q]);
      ptx::prefetch_tmap(&maps.b[q]);
    }
    if constexpr (DUAL) ptx::prefetch_tmap(&maps.a[1]);
    if constexpr (TMA_OUT) ptx::prefetch_tmap(&maps.out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], kEpiWarps);   // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_base_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      // Cursor over the (tile, pair, k-step) sequence of this CTA.  Two cursors walk it: `ld` issues the loads into
      // the smem ring, `pf` runs kPrefetchDist k-steps ahead and only issues L2 prefetches, so that operands which
      // live in HBM (scratch larger than L2) arrive in L2 before the ring needs them.
      struct Cursor {
        int tile, q, ks, nks, b, m0, n0;
        bool valid;
        int subl, kin;      // ks = subl * ksub + kin, maintained incrementally: an integer division per k-step costs this
                            // single thread ~150 cycles, which is what bounded the ring at ~500 cycles per k-step
      };
      auto seek = [&](Cursor& c) {   // position on the first k-step of (tile, q), skipping empty ranges
        while (c.tile < total_tiles) {
          const int bl = c.tile / tiles_per_batch;
          const int rem = c.tile - bl * tiles_per_batch;
          c.b = pb.reverse ? pb.batches - 1 - bl : bl;
          c.m0 = (rem / pb.tiles_n) * BM;
          c.n0 = (rem % pb.tiles_n) * BN;
          while (c.q < pb.num_pairs) {
            c.nks = ksteps_of(pb, c.q, c.b);
            if (c.ks < c.nks) { c.valid = true; return; }
            ++c.q;
            c.ks = 0;
            c.subl = 0;
            c.kin = 0;
          }
          c.tile += gridDim.x;
          c.q = 0;
          c.ks = 0;
          c.subl = 0;
          c.kin = 0;
        }
        c.valid = false;
      };
      auto advance = [&](Cursor& c) {
        ++c.ks;
        if (++c.kin == pb.ksub[c.q]) { c.kin = 0; ++c.subl; }
        if (c.ks >= c.nks) seek(c);
      };
      auto coords = [&](const Cursor& c, int& k0, int& ab, int& bb) {
        const int subl = c.subl;
        k0 = c.kin * BK;
        const int sub = subl + c.b * pb.sub_per_batch[c.q];
        ab = c.b * pb.a_bmul[c.q] + sub * pb.a_smul[c.q];
        bb = c.b * pb.b_bmul[c.q] + sub * pb.b_smul[c.q];
      };
      auto prefetch = [&](const Cursor& c) {
        int k0, ab, bb;
        coords(c, k0, ab, bb);
        if constexpr (!A_MN) {
          ptx::tma_prefetch_3d(&maps.a[c.q], k0, c.m0, ab);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) ptx::tma_prefetch_3d(&maps.a[c.q], c.m0 + j * 64, k0, ab);
        }
        if constexpr (!B_MN) {
          ptx::tma_prefetch_3d(&maps.b[c.q], k0, c.n0, bb);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) ptx::tma_prefetch_3d(&maps.b[c.q], c.n0 + j * 64, k0, bb);
        }
      };
      Cursor ld{(int)blockIdx.x, 0, 0, 0, 0, 0, 0, false, 0, 0};
      seek(ld);
      Cursor pf = ld;
      for (int i = 0; i < kPrefetchDist && pf.valid; ++i) {
        if (i >= kStages) prefetch(pf);        // the first kStages steps are loaded right away
        advance(pf);
      }
      int stage = 0;
      uint32_t phase = 0;
      while (ld.valid) {
        if (kPrefetchDist > 0 && pf.valid) {
          prefetch(pf);
          advance(pf);
        }
        int k0, ab, bb;
        coords(ld, k0, ab, bb);
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * L::kStageBytes;
        uint8_t* sb = sa + L::kABytes;
        ptx::mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
        if constexpr (DUAL) {
          const int ab1 = ld.b * pb.a_bmul[1];
          ptx::tma_load_3d(sa, &maps.a[0], &full_bar[stage], k0, ld.m0, ab);
          ptx::tma_load_3d(sa + L::kA1Bytes, &maps.a[1], &full_bar[stage], k0, ld.m0, ab1);
        } else if constexpr (!A_MN) {
          ptx::tma_load_3d(sa, &maps.a[ld.q], &full_bar[stage], k0, ld.m0, ab);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j)
            ptx::tma_load_3d(sa + j * 8192, &maps.a[ld.q], &full_bar[stage], ld.m0 + j * 64, k0, ab);
        }
        if constexpr (!B_MN) {
          ptx::tma_load_3d(sb, &maps.b[ld.q], &full_bar[stage], k0, ld.n0, bb);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            ptx::tma_load_3d(sb + j * 8192, &maps.b[ld.q], &full_bar[stage], ld.n0 + j * 64, k0, bb);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
        advance(ld);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_full = ptx::umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        const int bl0 = tile / tiles_per_batch;
        const int b = pb.reverse ? pb.batches - 1 - bl0 : bl0;
        ptx::mbar_wait(&tempty_bar[buf], bphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * kAccCols;
        // N tail: the last n-tile only multiplies the columns that exist (rounded up to the MMA granularity of 16)
        const int n0t = ((tile - bl0 * tiles_per_batch) % pb.tiles_n) * BN;
        const int nrem = pb.N - n0t;
        const uint32_t idesc = nrem >= BN ? idesc_full
                                          : ptx::umma_idesc_bf16(BM, (nrem + 15) & ~15, A_MN ? 1 : 0, B_MN ? 1 : 0);
        uint32_t accum = 0;
        for (int q = 0; q < pb.num_pairs; ++q) {
          const int nks = ksteps_of(pb, q, b);
          for (int ks = 0; ks < nks; ++ks) {
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(smem + stage * L::kStageBytes);
            const uint32_t sb = sa + L::kABytes;
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              // K-major: 16 bf16 = 32 B further along the swizzled 128 B row; MN-major: 16 k-rows = 2048 B further.
              const uint64_t ad = A_MN ? ptx::umma_desc(sa + kk * 2048, 8192, 1024) : ptx::umma_desc(sa + kk * 32, 16, 1024);
              const uint64_t bd = B_MN ? ptx::umma_desc(sb + kk * 2048, 8192, 1024) : ptx::umma_desc(sb + kk * 32, 16, 1024);
              ptx::mma_bf16_ss(tmem_d, ad, bd, idesc, accum);
              if constexpr (DUAL) {
                const uint64_t ad1 = ptx::umma_desc(sa + L::kA1Bytes + kk * 32, 16, 1024);
                ptx::mma_bf16_ss(tmem_d + BN, ad1, bd, idesc, accum);
              }
              accum = 1;
            }
            ptx::mma_commit(&empty_bar[stage]);   // frees the smem stage once these MMAs retire
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
        ptx::mma_commit(&tfull_bar[buf]);         // accumulator ready for the epilogue
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue
    const int q4 = warp & 3;                       // TMEM lane quadrant this warp may read
    const int half = (warp - kEpiWarp0) >> 2;      // 0/1: which alternate chunks this warp takes
    const bool issuer = (warp == kEpiWarp0) && (lane == 0);   // the one thread that owns the TMA-store bulk groups
    Epi epi(ep);
    int it = 0;
    int slab = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int bl = tile / tiles_per_batch;
      const int rem = tile - bl * tiles_per_batch;
      const int b = pb.reverse ? pb.batches - 1 - bl : bl;
      const int m0 = (rem / pb.tiles_n) * BM;
      const int n0 = (rem % pb.tiles_n) * BN;
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      ptx::mbar_wait(&tfull_bar[buf], bphase);
      ptx::tc_fence_after();
      const int m = m0 + q4 * 32 + lane;
      epi.tile_begin(b, m, n0);
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + buf * kAccCols;
      // 64-column slabs that hold at least one existing output column (the MMA warp skips the rest of an N tail)
      const int jn = (pb.N - n0 + 63) >> 6;
      const int jmax = jn < BN / 64 ? jn : BN / 64;
#pragma unroll 1
      for (int j = 0; j < jmax; ++j) {
        const int c = 2 * j + half;
        float v[32];
        ptx::tmem_ld_32x32(tacc + c * 32, v);
        [[maybe_unused]] typename side_of<Epi>::type side{};
        if constexpr (epi_has_side<Epi>::value) side = epi.pre(b, m, n0 + c * 32);
        if constexpr (DUAL) {
          float v1[32];
          ptx::tmem_ld_32x32(tacc + BN + c * 32, v1);
          ptx::tmem_ld_wait();
          if (j == jmax - 1) {                  // accumulator drained: hand the TMEM buffer back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[buf]);
          }
          if constexpr (epi_has_side<Epi>::value) epi.chunk2(b, m, n0 + c * 32, v, v1, side);
          else epi.chunk2(b, m, n0 + c * 32, v, v1);
        } else {
          ptx::tmem_ld_wait();
          if (j == jmax - 1) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[buf]);
          }
          if constexpr (epi_has_side<Epi>::value) epi.chunk(b, m, n0 + c * 32, v, side);
          else epi.chunk(b, m, n0 + c * 32, v);
        }
        if constexpr (TMA_OUT) {
          uint8_t* sbuf = out_smem + (slab & 1) * L::kOutSlabBytes;
          if (issuer) ptx::bulk_wait_group_read<1>();        // the store that last used this slab has read it
          ptx::named_bar_sync(kEpiBarId, kEpiWarps * 32);
          const int r = q4 * 32 + lane;
          const uint32_t rowbase = ptx::smem_u32(sbuf) + r * 128;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint32_t addr = rowbase + ((static_cast<uint32_t>(half * 4 + t) ^ static_cast<uint32_t>(r & 7)) << 4);
            ptx::st_shared_v4(addr, ptx::pack_bf16x2(v[8 * t + 0], v[8 * t + 1]), ptx::pack_bf16x2(v[8 * t + 2], v[8 * t + 3]),
                              ptx::pack_bf16x2(v[8 * t + 4], v[8 * t + 5]), ptx::pack_bf16x2(v[8 * t + 6], v[8 * t + 7]));
          }
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(kEpiBarId, kEpiWarps * 32);
          if (issuer) {
            ptx::tma_store_3d(&maps.out, sbuf, n0 + j * 64, m0, b);
            ptx::bulk_commit_group();
          }
          ++slab;
        }
      }
      epi.tile_end(b, m, n0, rem % pb.tiles_n, half);
    }
    if constexpr (TMA_OUT) {
      if (issuer) ptx::bulk_wait_group<0>();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace eng
