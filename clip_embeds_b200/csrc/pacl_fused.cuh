// Fused flash-style forward of the PACL all-pairs scorer for sm_100a (BASELINE north_star (1)-(3); reference math
// PACL/model/pacl.py:120-145, eval loop PACL/eval_pacl.py:53-57):
//
//   per item = (image i, tile of 256 texts), one CTA pair (tcgen05 cta_group::2, 128 texts per CTA):
//     phase 1   S_c = T^ V_c^T            patch chunk c (<= 256 patches), K = D          -> TMEM
//               epilogue: a = bf16(act(s)), num += <a, x>; the activations go STRAIGHT INTO SHARED MEMORY as the K-major
//               SWIZZLE_128B A-operand of phase 2 ([128 texts x P] bf16 per CTA, resident: 9 x 16 KB)
//     phase 2   U_j = A V[:, 256 j ...]   feature tile j (256 columns), K = P            -> TMEM
//               epilogue: usq += |u|^2 and, when the backward wants them, the pooled vectors as bf16
//
// The [Bi, Bt, P] activation tensor never leaves the SM (no HBM, no L2 round trip), and the pooling GEMM reads its A
// operand from shared memory instead of L2.  D = 768 does not allow the textbook "one accumulator for the whole
// pooled row" (128 x 768 fp32 = 768 TMEM columns > 512), so the pooled vector is produced in feature tiles of 256
// columns from the resident activations; every accumulation job (S chunk or U tile) takes one 256-column half of
// TMEM, alternating, so the epilogue of job n overlaps the MMAs of job n + 1 across phase and item boundaries.
//
//   warp 0      : TMA producer (both CTAs)
//   warp 1      : MMA issuer   (leader CTA; phase 2 waits per patch chunk for the activations of BOTH CTAs)
//   warp 2      : TMEM allocator
//   warps 4..15 : epilogue     (TMEM lane quadrant = warp & 3; the three warps of a quadrant take every third 32-column
//                               chunk.  The epilogues are latency-bound chains (shuffle -> MUFU -> FMA -> pack), and with
//                               two warps per SMSP the activation epilogue of a 192-patch chunk took 6.2 k cycles against
//                               4.6 k of MMA: three warps per SMSP at 128 registers, no register ping-pong)
//
// Shared memory is what this design is short of (144 KB of resident activations leave 80 KB for the operand ring), and
// the measurements that shaped the pipeline are (tests/micro/tma_rate.cu, cycle counters of development builds):
//   * a load's round trip (slot released -> TMA issued -> data landed -> MMA issued -> MMA retired -> slot released) is
//     ~1300-1500 cycles in this kernel, so a phase-1 k-step (T^ slice 16 KB + V slice <= 16 KB, 384 MMA cycles at
//     N = 192) needs >= 4 k-steps in flight: phase 1 BORROWS the activation blocks of the item's last patch chunk as
//     three extra 16 KB slots -- they are written (by that chunk's epilogue) only after every phase-1 MMA of the item
//     has retired and were last read by phase 2 of the previous item;
//   * the producer and the MMA issuer are single threads: every mbarrier wait costs them ~100-240 cycles even when
//     the phase has already completed, and an integer division ~150.  A phase-1 k-step therefore uses ONE barrier
//     pair (both loads signal the first slot's barrier), and nothing in the loops divides.
#pragma once
#include "../../include/clipk.h"
#include "ptx.cuh"

namespace fz {

constexpr int kThreads = 512;     // 4 control warps + 12 epilogue warps (3 per SMSP / TMEM lane quadrant)
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 12;
constexpr int kEpiPerQuad = kEpiWarps / 4;   // warps sharing a lane quadrant take chunks c = t, t + 3, ...
constexpr int kMaxOwn = 6;         // 32-column activation chunks one warp owns per item: ceil(2 * kMaxKB / 3)
constexpr int kSlotBytes = 16384;
constexpr int kSlots = 5;          // ring slots R0..R4 (ids 0..4); borrowed activation blocks X0..X2 are ids 5..7
constexpr int kMaxX = 3;
constexpr int kAllSlots = kSlots + kMaxX;
constexpr int kMaxKB = 9;          // resident activation k-blocks of 64 patches: P <= 576
constexpr int kMaxChunks = 3;      // patch chunks per item (each <= 256 patches)
constexpr int kBarrierBytes = 1024;
constexpr int kSmemTotal = kMaxKB * 16384 + kSlots * kSlotBytes + kBarrierBytes + 1024 /*align*/;
static_assert(kSmemTotal <= 227 * 1024, "shared memory budget");

struct Maps {
  CUtensorMap T;     // T^ [Bt][D]        K-major, box (64, 128)
  CUtensorMap Vk;    // V  [Bi][P][D]     K-major (k = d), box (64, pc / 2)
  CUtensorMap Vmn;   // V  [Bi][P][D]     MN-major (mn = d, k = p), box (64, 64)
};

struct Params {
  int Bi, Bt, P, D, act;
  int nS, pc;          // patch chunks per item and their width (multiple of 64 when nS > 1, of 16 otherwise)
  int nU;              // feature tiles of 256 columns (the last one may be 128 wide)
  int kbA;             // ceil(P / 64): activation k-blocks
  int ksD;             // ceil(D / 64): k-steps of phase 1
  int tilesM;          // ceil(Bt / 256)
  int xslots;          // 1: phase 1 borrows the last chunk's activation blocks as extra ring slots
  int prefetch;        // 1: the epilogue warps prefetch the next item's image into L2 at the start of every item
  const void* V;               // [Bi][P][D] bf16 (prefetch address only; the operands arrive through the tensor maps)
  __nv_bfloat16* pooled;       // [Bi][Bt][D] bf16 or nullptr: the un-normalised pooled vectors (SAVE_U)
  const float* rnV;            // [Bi][P]  1 / max(|V_ip|, 1e-12)
  float* num;                  // [Bi][Bt] += <u, t^>      (zeroed by the caller)
  float* usq;                  // [Bi][Bt] += |u|^2        (zeroed by the caller)
};

__device__ __forceinline__ int chunk_cols(const Params& pr, int sc) {      // MMA N of patch chunk sc
  const int rem = pr.P - sc * pr.pc;
  const int r16 = (rem + 15) & ~15;
  return r16 < pr.pc ? r16 : pr.pc;
}
__device__ __forceinline__ int chunk_kb0(const Params& pr, int sc) { return (sc * pr.pc) >> 6; }
__device__ __forceinline__ int chunk_kbn(const Params& pr, int sc) {       // k-blocks covered by chunk sc
  const int end = (sc + 1 == pr.nS) ? pr.kbA : (((sc + 1) * pr.pc) >> 6);
  return end - chunk_kb0(pr, sc);
}
__device__ __forceinline__ int utile_cols(const Params& pr, int j) {
  const int rem = pr.D - j * 256;
  return rem < 256 ? rem : 256;
}

// Phase-1 k-steps take a PAIR of slots (a: T^ slice, b: V slice) guarded by slot a's barriers only.
//   pair 0 = (R0, R1)   pair 1 = (R2, R3)   pair 2 = (X0, X1)   pair 3 = (X2, R4)
// Phase-2 k-steps take single ring slots R0, R1, ..., R4, R0, ... (restarting at R0 every item) with their own barriers.
// R1 / R3 / R4 are therefore barrier-less in phase 1 and X* never appear in phase 2; the transition hazards are closed
// by three non-consuming waits, see the producer.
__device__ __forceinline__ int pair_a(int p) { return p == 0 ? 0 : (p == 1 ? 2 : (p == 2 ? 5 : 7)); }
__device__ __forceinline__ int pair_b(int p) { return p == 0 ? 1 : (p == 1 ? 3 : (p == 2 ? 6 : 4)); }

template <bool SAVE_U>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
pacl_fused_fwd_kernel(const __grid_constant__ Maps maps, const Params pr) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_res = smem;                                   // resident activations: kbA blocks of [128 rows][64] bf16
  uint8_t* ring = a_res + kMaxKB * 16384;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + kSlots * kSlotBytes);
  uint64_t* empty_bar = full_bar + kAllSlots;
  uint64_t* tfull_bar = empty_bar + kAllSlots;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]            (the leader's copy is the live one)
  uint64_t* aready_bar = tempty_bar + 2;        // [kMaxChunks]   (leader's copy: both CTAs' epilogue warps arrive)
  uint64_t* xfree_bar = aready_bar + 4;         // phase 2 of the previous item has finished reading the borrowed blocks
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(xfree_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(ptx::cluster_ctarank());
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int nitems = pr.Bi * pr.tilesM;
  const int kbX0 = chunk_kb0(pr, pr.nS - 1);               // first borrowed activation block
  const int nXraw = pr.xslots ? chunk_kbn(pr, pr.nS - 1) : 0;
  const int npairs = 2 + (nXraw >= 2 ? 1 : 0) + (nXraw >= 3 ? 1 : 0);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&maps.T);
    ptx::prefetch_tmap(&maps.Vk);
    ptx::prefetch_tmap(&maps.Vmn);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kAllSlots; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], 2 * kEpiWarps);
    }
    for (int i = 0; i < kMaxChunks; ++i) ptx::mbar_init(&aready_bar[i], 2 * kEpiWarps);
    ptx::mbar_init(xfree_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2sm(tmem_base_slot, 512);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (ptx::elect_one()) {
      uint32_t phm = 0;                // bit s: phase parity of slot s's barrier pair
      auto slot_ptr = [&](int s) -> uint8_t* {
        return s < kSlots ? ring + s * kSlotBytes : a_res + (kbX0 + s - kSlots) * 16384;
      };
      auto released = [&](int s) {     // non-consuming: the latest fill guarded by slot s's barriers has been consumed
        ptx::mbar_wait(&empty_bar[s], ((phm >> s) & 1u) ^ 1u);
      };
      auto acquire = [&](int s, uint32_t bytes_both) -> uint32_t {      // wait for the slot, arm the leader's full barrier
        ptx::mbar_wait(&empty_bar[s], ((phm >> s) & 1u) ^ 1u);
        phm ^= 1u << s;
        if (leader) ptx::mbar_arrive_expect_tx(&full_bar[s], bytes_both);
        return ptx::mapa(ptx::smem_u32(&full_bar[s]), 0);
      };
      const uint32_t vbytes = static_cast<uint32_t>(pr.pc >> 1) * 128u;     // this CTA's V k-slice (box rows = pc / 2)
      int iter = 0;
      for (int item = cluster_id; item < nitems; item += num_clusters, ++iter) {
        const int img = item / pr.tilesM;
        const int m0 = (item - img * pr.tilesM) * 256 + rank * 128;
        // ---- phase 1
        int pp = 0;
        uint32_t first = 0;              // pairs already used in this item
        bool xok = iter == 0;
        for (int sc = 0; sc < pr.nS; ++sc) {
          const int prow = sc * pr.pc + rank * (chunk_cols(pr, sc) >> 1);   // first patch row of this CTA's B half
          for (int ks = 0; ks < pr.ksD; ++ks) {
            const int p = pp;
            pp = pp + 1 == npairs ? 0 : pp + 1;
            const int a = pair_a(p), b = pair_b(p);
            if (!((first >> p) & 1u)) {
              first |= 1u << p;
              // borrowed blocks: phase 2 of the previous item must have finished reading them
              if (a >= kSlots && !xok) {
                ptx::mbar_wait(xfree_bar, static_cast<uint32_t>((iter - 1) & 1));
                xok = true;
              }
              // slot b carried phase-2 data under its own barrier: that fill must have been consumed
              if (b < kSlots) released(b);
            }
            const uint32_t fb = acquire(a, 2u * kSlotBytes + 2u * vbytes);
            ptx::tma_load_3d_2sm(slot_ptr(a), &maps.T, fb, ks * 64, m0, 0);
            ptx::tma_load_3d_2sm(slot_ptr(b), &maps.Vk, fb, ks * 64, prow, img);
          }
        }
        // ---- phase 2 (R0 before R1, R2 before R3: acquiring the pair's first slot orders the second one's reuse;
        //      R4 follows X2's barrier)
        int rp = 0;
        bool r4ok = npairs < 4;
        for (int j = 0; j < pr.nU; ++j) {
          const int tn = utile_cols(pr, j);
          const int nbox = tn >> 7;                                         // 64-wide feature groups of this CTA's half
          const int d0 = j * 256 + rank * (tn >> 1);
          for (int kb = 0; kb < pr.kbA; ++kb) {
            const int s2 = rp;
            rp = rp + 1 == kSlots ? 0 : rp + 1;
            if (s2 == 4 && !r4ok) {
              released(7);
              r4ok = true;
            }
            const uint32_t fb = acquire(s2, 2u * static_cast<uint32_t>(nbox) * 8192u);
            for (int g = 0; g < nbox; ++g)
              ptx::tma_load_3d_2sm(slot_ptr(s2) + g * 8192, &maps.Vmn, fb, d0 + g * 64, kb * 64, img);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && ptx::elect_one()) {
      uint32_t phm = 0;                // mirrors the producer's slot phases
      uint32_t job = 0;
      int iter = 0;
      auto slot_addr = [&](int s) -> uint32_t {
        return ptx::smem_u32(s < kSlots ? ring + s * kSlotBytes : a_res + (kbX0 + s - kSlots) * 16384);
      };
      auto wait_full = [&](int s) {
        ptx::mbar_wait(&full_bar[s], (phm >> s) & 1u);
        phm ^= 1u << s;
      };
      for (int item = cluster_id; item < nitems; item += num_clusters, ++iter) {
        // phase 1: S chunks
        int pp = 0;
        for (int sc = 0; sc < pr.nS; ++sc, ++job) {
          const uint32_t buf = job & 1u;
          ptx::mbar_wait(&tempty_bar[buf], ((job >> 1) & 1u) ^ 1u);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * 256u;
          const uint32_t idesc = ptx::umma_idesc_bf16(256, chunk_cols(pr, sc), 0, 0);
          uint32_t accum = 0;
          for (int ks = 0; ks < pr.ksD; ++ks) {
            const int p = pp;
            pp = pp + 1 == npairs ? 0 : pp + 1;
            const int a = pair_a(p);
            wait_full(a);
            ptx::tc_fence_after();
            const uint32_t sa = slot_addr(a);
            const uint32_t sb = slot_addr(pair_b(p));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              ptx::mma_bf16_ss_2sm(tmem_d, ptx::umma_desc(sa + kk * 32, 16, 1024), ptx::umma_desc(sb + kk * 32, 16, 1024),
                                   idesc, accum);
              accum = 1;
            }
            ptx::mma_commit_2sm(&empty_bar[a], 3);
          }
          ptx::mma_commit_2sm(&tfull_bar[buf], 3);
        }
        // phase 2: U tiles from the resident activations
        int rp = 0;
        for (int j = 0; j < pr.nU; ++j, ++job) {
          const uint32_t buf = job & 1u;
          ptx::mbar_wait(&tempty_bar[buf], ((job >> 1) & 1u) ^ 1u);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * 256u;
          const uint32_t idesc = ptx::umma_idesc_bf16(256, utile_cols(pr, j), 0, 1);
          uint32_t accum = 0;
          int sc = 0, kb_next = 0;       // next chunk boundary to wait for (first U tile of the item only)
          for (int kb = 0; kb < pr.kbA; ++kb) {
            if (j == 0 && kb == kb_next) {       // activations of chunk sc written (and proxy-fenced) by both CTAs
              ptx::mbar_wait_cluster(&aready_bar[sc], static_cast<uint32_t>(iter & 1));
              kb_next += chunk_kbn(pr, sc);
              ++sc;
            }
            const int s2 = rp;
            rp = rp + 1 == kSlots ? 0 : rp + 1;
            wait_full(s2);
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(a_res + kb * 16384);
            const uint32_t sb = slot_addr(s2);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              ptx::mma_bf16_ss_2sm(tmem_d, ptx::umma_desc(sa + kk * 32, 16, 1024),
                                   ptx::umma_desc(sb + kk * 2048, 8192, 1024), idesc, accum);
              accum = 1;
            }
            ptx::mma_commit_2sm(&empty_bar[s2], 3);
          }
          ptx::mma_commit_2sm(&tfull_bar[buf], 3);
        }
        // every MMA that reads this item's activations has been issued: when they retire, the borrowed blocks may take
        // the next item's phase-1 loads
        ptx::mma_commit_2sm(xfree_bar, 3);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 text rows each)
    const int q4 = warp & 3;
    const int third = (warp - kEpiWarp0) >> 2;             // 0 .. 2
    const int ew = warp - kEpiWarp0;
    const int rowl = q4 * 32 + lane;                       // row inside this CTA's 128-row slab
    const uint32_t tempty_leader[2] = {ptx::mapa(ptx::smem_u32(&tempty_bar[0]), 0),
                                       ptx::mapa(ptx::smem_u32(&tempty_bar[1]), 0)};
    const uint32_t a_row = ptx::smem_u32(a_res) + static_cast<uint32_t>(rowl) * 128u;
    const uint32_t rsw = static_cast<uint32_t>(rowl) & 7u;
    uint32_t job = 0;
#ifdef CLIPK_FZ_PROF
    long long pe_ts[24], pe_s = 0, pe_u = 0, pe_t = 0;
    for (int i = 0; i < 24; ++i) pe_ts[i] = 0;
#define FZ_TS() if (job < 24) pe_ts[job] = clock64();
#define FZ_E0() pe_t = clock64();
#define FZ_EACC(x) x += clock64() - pe_t;
#else
#define FZ_TS()
#define FZ_E0()
#define FZ_EACC(x)
#endif
    for (int item = cluster_id; item < nitems; item += num_clusters) {
      const int img = item / pr.tilesM;
      const int m0 = (item - img * pr.tilesM) * 256 + rank * 128;
      const int row = m0 + rowl;
      // rnV of the 32-patch chunks this warp owns, in the order it meets them (chunk sc, local chunks third, third + 3,
      // ...): lane l = column l of the chunk
      float rn[kMaxOwn];
      {
        int n = 0;
#pragma unroll
        for (int sc = 0; sc < kMaxChunks; ++sc) {
          if (sc < pr.nS) {
            const int gc0 = chunk_kb0(pr, sc) * 2, nch = chunk_kbn(pr, sc) * 2;
            for (int cl = third; cl < nch; cl += kEpiPerQuad) {
              const int p = (gc0 + cl) * 32 + lane;
              const float r = p < pr.P ? __ldg(pr.rnV + (int64_t)img * pr.P + p) : 0.f;
#pragma unroll
              for (int g = 0; g < kMaxOwn; ++g)
                if (g == n) rn[g] = r;
              ++n;
            }
          }
        }
#pragma unroll
        for (int g = 0; g < kMaxOwn; ++g)
          if (g >= n) rn[g] = 0.f;
      }
      // L2 prefetch of this pair's NEXT image (the V k-slices of phase 1 then hit L2: ~600 instead of 1400-2000 cycles).
      // The (up to 4) pairs that share an image and their two CTAs each take one eighth of its lines.
      if (pr.prefetch) {
        const int nitem = item + num_clusters;
        if (nitem < nitems) {
          const int nimg = nitem / pr.tilesM;
          const int part = ((nitem - nimg * pr.tilesM) * 2 + rank) & 7;
          const int64_t lines = ((int64_t)pr.P * pr.D * 2 + 127) >> 7;
          const int64_t per = (lines + 7) >> 3;
          const int64_t l1 = (part + 1) * per < lines ? (part + 1) * per : lines;
          const char* base = static_cast<const char*>(pr.V) + (int64_t)nimg * pr.P * pr.D * 2;
          for (int64_t l = part * per + ew * 32 + lane; l < l1; l += kEpiWarps * 32)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (l << 7)));
        }
      }
      // ---------------- phase 1 epilogues: activations -> resident shared memory, num += <a, x>
      float numacc = 0.f;
      int own = 0;                     // index into rn[]
      for (int sc = 0; sc < pr.nS; ++sc, ++job) {
        const uint32_t buf = job & 1u;
        ptx::mbar_wait(&tfull_bar[buf], (job >> 1) & 1u);
        FZ_TS()
        FZ_E0()
        ptx::tc_fence_after();
        const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + buf * 256u;
        const int gc0 = chunk_kb0(pr, sc) * 2;               // first global 32-column chunk of this patch chunk
        const int nch = chunk_kbn(pr, sc) * 2;               // chunks to write (whole k-blocks; tail columns are zeroed)
        const int ncols = chunk_cols(pr, sc);                // columns the MMA actually wrote
#pragma unroll 1
        for (int cl = third; cl < nch; cl += kEpiPerQuad, ++own) {
          float v[32];
          ptx::tmem_ld_32x32(tacc + cl * 32, v);
          const int gc = gc0 + cl;
          float rl = 0.f;
#pragma unroll
          for (int g = 0; g < kMaxOwn; ++g)
            if (g == own) rl = rn[g];
          ptx::tmem_ld_wait();
          if (cl + kEpiPerQuad >= nch) {                     // last chunk of this warp: hand the accumulator back
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader[buf]);
          }
          uint32_t pk[16];
          const bool pad = (cl * 32 + 32 > ncols) || (gc * 32 + 32 > pr.P);
          if (pad) {                                                   // (warp-uniform) chunk with padding columns:
#pragma unroll
            for (int j = 0; j < 32; ++j)                               // rnV = 0 there; zero the (undefined) accumulator too
              if (__shfl_sync(0xffffffffu, rl, j) == 0.f) v[j] = 0.f;
          }
          if (pr.act == CLIPK_ACT_ONES) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float r0 = __shfl_sync(0xffffffffu, rl, j), r1 = __shfl_sync(0xffffffffu, rl, j + 1);
              numacc += v[j] + v[j + 1];                               // x = 0 in padding columns
              pk[j >> 1] = ptx::pack_bf16x2(r0 != 0.f ? 1.f : 0.f, r1 != 0.f ? 1.f : 0.f);
            }
          } else if (pr.act == CLIPK_ACT_SOFTMAX10) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float r0 = __shfl_sync(0xffffffffu, rl, j), r1 = __shfl_sync(0xffffffffu, rl, j + 1);
              float a0 = ptx::ex2_approx(fmaf(14.426950408889634f * v[j], r0, -14.426950408889634f));
              float a1 = ptx::ex2_approx(fmaf(14.426950408889634f * v[j + 1], r1, -14.426950408889634f));
              a0 = r0 != 0.f ? a0 : 0.f;
              a1 = r1 != 0.f ? a1 : 0.f;
              const uint32_t p2 = ptx::pack_bf16x2(a0, a1);
              numacc = fmaf(__uint_as_float(p2 << 16), v[j], fmaf(__uint_as_float(p2 & 0xFFFF0000u), v[j + 1], numacc));
              pk[j >> 1] = p2;
            }
          } else {
            const float r5l = 5.f * rl;     // sigmoid(10 s) = 0.5 tanh(5 s) + 0.5
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 8) {
              float r[8], a[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) r[i] = __shfl_sync(0xffffffffu, r5l, j0 + i);
#pragma unroll
              for (int i = 0; i < 8; ++i) asm("tanh.approx.f32 %0, %1;" : "=f"(a[i]) : "f"(v[j0 + i] * r[i]));
#pragma unroll
              for (int i = 0; i < 8; ++i) a[i] = fmaf(0.5f, a[i], 0.5f);
#pragma unroll
              for (int i = 0; i < 8; i += 2) {
                const uint32_t p2 = ptx::pack_bf16x2(a[i], a[i + 1]);
                // (padding columns: a = sigma(0) = 0.5 but x = 0, so num is unaffected; the stored a is cleared below)
                numacc = fmaf(__uint_as_float(p2 << 16), v[j0 + i], fmaf(__uint_as_float(p2 & 0xFFFF0000u), v[j0 + i + 1], numacc));
                pk[(j0 + i) >> 1] = p2;
              }
            }
            if (pad) {                 // sigma(0) = 0.5 in the padding columns would pollute the pooling GEMM
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (__shfl_sync(0xffffffffu, rl, j) == 0.f) pk[j >> 1] &= (j & 1) ? 0x0000FFFFu : 0xFFFF0000u;
            }
          }
          // [row][32 patches] = 64 B = four 16-byte units of k-block gc / 2, units 4 (gc & 1) .. + 3, SWIZZLE_128B
          const uint32_t base = a_row + static_cast<uint32_t>(gc >> 1) * 16384u;
          const uint32_t u0 = static_cast<uint32_t>(gc & 1) * 4u;
#pragma unroll
          for (int t = 0; t < 4; ++t)
            ptx::st_shared_v4(base + (((u0 + t) ^ rsw) << 4), pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
        }
        if (third >= nch) {              // (cannot happen: every patch chunk has >= 2 column chunks... kept for safety)
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader[buf]);
        }
        // this warp's part of the chunk's activations is in shared memory: publish it to the tensor cores of the pair
        // (default-scope arrive, as for the TMEM hand-back: the data sits in THIS SM's shared memory and was handed to
        // the async proxy by the fence; a .release.cluster arrive costs a ~1900-cycle GPU-scope fence per chunk)
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&aready_bar[sc]), 0));
        FZ_EACC(pe_s)
      }
      // ---------------- phase 2 epilogues: |u|^2 (+ pooled vectors)
      float uacc = 0.f;
      const bool rowok = row < pr.Bt;
      __nv_bfloat16* urow = SAVE_U ? pr.pooled + ((int64_t)img * pr.Bt + (rowok ? row : 0)) * pr.D : nullptr;
      for (int j = 0; j < pr.nU; ++j, ++job) {
        const uint32_t buf = job & 1u;
        const int nch = utile_cols(pr, j) >> 5;
        ptx::mbar_wait(&tfull_bar[buf], (job >> 1) & 1u);
        FZ_TS()
        FZ_E0()
        ptx::tc_fence_after();
        const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + buf * 256u;
        bool released = false;
#pragma unroll 1
        for (int c = third; c < nch; c += kEpiPerQuad) {
          float v[32];
          ptx::tmem_ld_32x32(tacc + c * 32, v);
          ptx::tmem_ld_wait();
          if (c + kEpiPerQuad >= nch) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader[buf]);
            released = true;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) uacc = fmaf(v[i], v[i], uacc);
          if constexpr (SAVE_U) {
            if (rowok) {
              uint4* dst = reinterpret_cast<uint4*>(urow + j * 256 + c * 32);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                uint4 o;
                o.x = ptx::pack_bf16x2(v[8 * t + 0], v[8 * t + 1]);
                o.y = ptx::pack_bf16x2(v[8 * t + 2], v[8 * t + 3]);
                o.z = ptx::pack_bf16x2(v[8 * t + 4], v[8 * t + 5]);
                o.w = ptx::pack_bf16x2(v[8 * t + 6], v[8 * t + 7]);
                dst[t] = o;
              }
            }
          }
        }
        if (!released) {                 // narrow tile (128 columns = 4 chunks): the third warp of a quadrant owns none
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader[buf]);
        }
        FZ_EACC(pe_u)
      }
      if (rowok) {
        atomicAdd(pr.usq + (int64_t)img * pr.Bt + row, uacc);
        atomicAdd(pr.num + (int64_t)img * pr.Bt + row, numacc);
      }
    }
#ifdef CLIPK_FZ_PROF
    if (cluster_id == 3 && leader && warp == kEpiWarp0 && lane == 0) {
      printf("fz epilogue warp 4: busy cycles in S epilogues %lld, in U epilogues %lld (%u jobs)\n", pe_s, pe_u, job);
      printf("fz accumulator-ready deltas of jobs 1..23:");
      for (int i = 1; i < 24; ++i) printf(" %lld", pe_ts[i] - pe_ts[i - 1]);
      printf("\n");
    }
#endif
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace fz
