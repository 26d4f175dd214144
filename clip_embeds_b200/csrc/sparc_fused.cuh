// Fused per-sample forward of the SPARC token-to-patch alignment for sm_100a (sparc.forward,
// PACL/model/pacl.py:453-478):
//
//   one CTA per sample b (persistent over samples), tcgen05 cta_group::1:
//     GEMM 1   S^T[p, t] = <V_p, L_t>            patches on the MMA rows (576 = 4.5 x 128: no padding of 77 tokens to 128
//                                                rows), tokens on N = 80; all five [128 x 80] tiles of a sample stay in
//                                                TMEM (400 of 512 columns), so the whole score matrix of the sample is
//                                                on chip and is never written anywhere
//     epilogue min / max over the patches (with first-occurrence arg positions, pacl.py:463-465), threshold sigma,
//              row-normalise (:468-474) -> W as the K-major SWIZZLE_128B A operand of GEMM 2, straight into shared
//              memory (it takes over the buffer that held L)
//     GEMM 2   G^T[d, t] = sum_p V[p,d] W[t,p]   features on the MMA rows (tiles of 128), tokens on N = 80 again: a thread
//                                                owns one feature d, so the store of G[t][d] is one coalesced 128-byte
//                                                line per warp and token; row 77 of W is all ones, so G[77] is the patch
//                                                sum of V: SparcLoss's global image feature for free
//     epilogue raw G -> global; then one warp per token row: norm, g^ = n(G); l^ = n(L) (pacl.py:476-478)
//
// HBM traffic per sample: V once (the second pass of GEMM 2 hits L2), L once, the two fp32 outputs -- against S / W /
// G_raw round trips, a 128-row padding of the tokens and eight launches in the staged path.  Saved for the backward:
// per (b, t) the statistics (min, R, Z + eps) and the arg positions, 24 bytes.
#pragma once
#include "../../include/clipk.h"
#include "ptx.cuh"
#include "simt_util.cuh"

namespace sfz {

constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kTn = 80;              // tokens on the MMA N dimension (T <= 79: row T is the all-ones pooling row)
constexpr int kBlk = kTn * 128;      // one K-major k-block of [80 rows][64 bf16] = 10240 bytes
constexpr int kMaxKsD = 12;          // D <= 768
constexpr int kMaxMT = 5;            // P <= 640 (five 128-patch tiles: 400 TMEM columns)
constexpr int kSlots = 6;
constexpr int kSlotBytes = 16384;
constexpr int kLwBytes = kMaxKsD * kBlk;                       // 122880: L (12 k-blocks) or W (10 k-blocks + zero tail)
constexpr int kScratchBytes = 4 * kTn * 16 + kTn * 16 + 4 * kTn * 4 + 2 * 128 * 4;     // reductions
constexpr int kSmemTotal = kLwBytes + kSlots * kSlotBytes + kScratchBytes + 256 + 1024;
static_assert(kSmemTotal <= 227 * 1024, "shared memory budget");

struct Maps {
  CUtensorMap L;     // L [B][T][D]  K-major, box (64, 80)    (rows >= T zero-filled)
  CUtensorMap Vk;    // V [B][P][D]  K-major, box (64, 128)
  CUtensorMap Vmn;   // V [B][P][D]  MN-major (mn = d, k = p), box (64, 64)
};

struct Params {
  int B, T, P, D;
  int nMT, ksD, kbP, nU;        // ceil(P/128), ceil(D/64), ceil(P/64), D/128
  float sigma;
  const __nv_bfloat16* L;       // [B][T][D]
  float* l_hat;                 // [B][T][D]
  float* g_hat;                 // [B][T][D]
  float* lnorm;                 // [B][T]
  float* gnorm;                 // [B][T]
  float4* stats;                // [B][T] (min, R, Z + eps, 0)
  int2* arg;                    // [B][T] (argmin, argmax)
  float* pooled;                // [B][D] mean over patches of V, or nullptr
};

struct ColRed {
  float mn;
  int imn;
  float mx;
  int imx;
};

__global__ void __launch_bounds__(kThreads, 1) sparc_fused_fwd_kernel(const __grid_constant__ Maps maps, const Params pr) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* lw = smem;                                    // L k-blocks, later W k-blocks
  uint8_t* ring = lw + kLwBytes;
  ColRed* red = reinterpret_cast<ColRed*>(ring + kSlots * kSlotBytes);          // [4][kTn]
  float4* colstat = reinterpret_cast<float4*>(red + 4 * kTn);                   // [kTn] (min, R, Z + eps, 0)
  float* zred = reinterpret_cast<float*>(colstat + kTn);                        // [4][kTn]
  float* gnp = zred + 4 * kTn;                                                  // [2][128]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(gnp + 2 * 128);
  uint64_t* empty_bar = full_bar + kSlots;
  uint64_t* lfull = empty_bar + kSlots;
  uint64_t* sfull = lfull + 1;
  uint64_t* wready = sfull + 1;
  uint64_t* lwfree = wready + 1;
  uint64_t* gfull = lwfree + 1;      // [2]
  uint64_t* gempty = gfull + 2;      // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(gempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&maps.L);
    ptx::prefetch_tmap(&maps.Vk);
    ptx::prefetch_tmap(&maps.Vmn);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kSlots; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(lfull, 1);
    ptx::mbar_init(sfull, 1);
    ptx::mbar_init(wready, kEpiWarps);
    ptx::mbar_init(lwfree, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&gfull[i], 1);
      ptx::mbar_init(&gempty[i], kEpiWarps);
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
      int slot = 0;
      uint32_t ph = 0;
      auto acquire = [&](uint32_t bytes) {
        ptx::mbar_wait(&empty_bar[slot], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&full_bar[slot], bytes);
      };
      auto advance = [&]() {
        if (++slot == kSlots) { slot = 0; ph ^= 1; }
      };
      int iter = 0;
      for (int b = blockIdx.x; b < pr.B; b += gridDim.x, ++iter) {
        if (iter > 0) ptx::mbar_wait(lwfree, static_cast<uint32_t>((iter - 1) & 1));     // GEMM 2 of the previous sample read W
        ptx::mbar_arrive_expect_tx(lfull, static_cast<uint32_t>(pr.ksD) * kBlk);
        for (int ks = 0; ks < pr.ksD; ++ks) ptx::tma_load_3d(lw + ks * kBlk, &maps.L, lfull, ks * 64, 0, b);
        for (int mt = 0; mt < pr.nMT; ++mt)
          for (int ks = 0; ks < pr.ksD; ++ks) {
            acquire(kSlotBytes);
            ptx::tma_load_3d(ring + slot * kSlotBytes, &maps.Vk, &full_bar[slot], ks * 64, mt * 128, b);
            advance();
          }
        for (int j = 0; j < pr.nU; ++j)
          for (int kb = 0; kb < pr.kbP; ++kb) {
            acquire(kSlotBytes);
            ptx::tma_load_3d(ring + slot * kSlotBytes, &maps.Vmn, &full_bar[slot], j * 128, kb * 64, b);
            ptx::tma_load_3d(ring + slot * kSlotBytes + 8192, &maps.Vmn, &full_bar[slot], j * 128 + 64, kb * 64, b);
            advance();
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      int slot = 0;
      uint32_t ph = 0;
      uint32_t guse[2] = {0, 0};         // uses of each G accumulator so far
      auto advance = [&]() {
        if (++slot == kSlots) { slot = 0; ph ^= 1; }
      };
      const uint32_t lw_addr = ptx::smem_u32(lw);
      const uint32_t idesc1 = ptx::umma_idesc_bf16(128, kTn, 0, 0);
      const uint32_t idesc2 = ptx::umma_idesc_bf16(128, kTn, 1, 0);      // A = V^T (MN-major), B = W (K-major)
      int iter = 0;
      for (int b = blockIdx.x; b < pr.B; b += gridDim.x, ++iter) {
        ptx::mbar_wait(lfull, static_cast<uint32_t>(iter & 1));
        // the score accumulators overlap both G buffers: their last epilogues must have drained
        for (int i = 0; i < 2; ++i) ptx::mbar_wait(&gempty[i], (guse[i] & 1u) ^ 1u);
        ptx::tc_fence_after();
        for (int mt = 0; mt < pr.nMT; ++mt) {
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(mt * kTn);
          uint32_t accum = 0;
          for (int ks = 0; ks < pr.ksD; ++ks) {
            ptx::mbar_wait(&full_bar[slot], ph);
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(ring + slot * kSlotBytes);
            const uint32_t sb = lw_addr + ks * kBlk;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              ptx::mma_bf16_ss(tmem_d, ptx::umma_desc(sa + kk * 32, 16, 1024), ptx::umma_desc(sb + kk * 32, 16, 1024), idesc1, accum);
              accum = 1;
            }
            ptx::mma_commit(&empty_bar[slot]);
            advance();
          }
        }
        ptx::mma_commit(sfull);
        ptx::mbar_wait(wready, static_cast<uint32_t>(iter & 1));
        ptx::tc_fence_after();
        for (int j = 0; j < pr.nU; ++j) {
          const int buf = j & 1;
          ptx::mbar_wait(&gempty[buf], (guse[buf] & 1u) ^ 1u);
          ++guse[buf];
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * 128);
          uint32_t accum = 0;
          for (int kb = 0; kb < pr.kbP; ++kb) {
            ptx::mbar_wait(&full_bar[slot], ph);
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(ring + slot * kSlotBytes);      // V^T: [64 p][64 d] x 2 feature groups
            const uint32_t sb = lw_addr + kb * kBlk;                            // W: [80 t][64 p]
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              ptx::mma_bf16_ss(tmem_d, ptx::umma_desc(sa + kk * 2048, 8192, 1024), ptx::umma_desc(sb + kk * 32, 16, 1024), idesc2, accum);
              accum = 1;
            }
            ptx::mma_commit(&empty_bar[slot]);
            advance();
          }
          ptx::mma_commit(&gfull[buf]);
        }
        ptx::mma_commit(lwfree);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;                       // TMEM lane quadrant
    const int h = (warp - kEpiWarp0) >> 2;        // column half
    const int ew = warp - kEpiWarp0;
    const int tid = threadIdx.x - kEpiWarp0 * 32; // 0 .. 255
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t lw_addr = ptx::smem_u32(lw);
    uint32_t guse[2] = {0, 0};
    int iter = 0;
#ifdef CLIPK_SFZ_PROF
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt0 = clock64(), pq = 0;
#define SFZ_T(i) { const long long _n = clock64(); pt[i] += _n - pq; pq = _n; }
#define SFZ_T0() pq = clock64();
#else
#define SFZ_T(i)
#define SFZ_T0()
#endif
    for (int b = blockIdx.x; b < pr.B; b += gridDim.x, ++iter) {
      SFZ_T0()
      // ---- l^ = n(L) (overlaps GEMM 1): one warp per token row, straight from global memory
      for (int t = ew; t < pr.T; t += kEpiWarps) {
        const __nv_bfloat16* row = pr.L + ((int64_t)b * pr.T + t) * pr.D;
        float v[kMaxKsD / 4][8];
        float acc = 0.f;
#pragma unroll
        for (int it = 0; it < kMaxKsD / 4; ++it) {
          const int d0 = it * 256 + lane * 8;
          if (d0 < pr.D) simt::load8<__nv_bfloat16>(row + d0, v[it]);
          else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[it][j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) acc = fmaf(v[it][j], v[it][j], acc);
        }
        acc = ptx::warp_sum(acc);
        const float nr = sqrtf(acc);
        const float r = 1.f / fmaxf(nr, 1e-12f);
        float* orow = pr.l_hat + ((int64_t)b * pr.T + t) * pr.D;
#pragma unroll
        for (int it = 0; it < kMaxKsD / 4; ++it) {
          const int d0 = it * 256 + lane * 8;
          if (d0 < pr.D) {
            *reinterpret_cast<float4*>(orow + d0) = make_float4(v[it][0] * r, v[it][1] * r, v[it][2] * r, v[it][3] * r);
            *reinterpret_cast<float4*>(orow + d0 + 4) = make_float4(v[it][4] * r, v[it][5] * r, v[it][6] * r, v[it][7] * r);
          }
        }
        if (lane == 0) pr.lnorm[(int64_t)b * pr.T + t] = nr;
      }
      SFZ_T(0)
      // ---- scores of the whole sample are in TMEM
      ptx::mbar_wait(sfull, static_cast<uint32_t>(iter & 1));
      ptx::tc_fence_after();
      SFZ_T(1)
      // pass 1: min / max over the patches with first-occurrence positions.  All tiles of a column group are loaded
      // before the single wait (one TMEM round trip per group instead of one per tile).
      for (int g = 0; g < 5; ++g) {
        const int c0 = 40 * h + 8 * g;
        float v[kMaxMT][8];
#pragma unroll
        for (int mt = 0; mt < kMaxMT; ++mt)
          if (mt < pr.nMT) ptx::tmem_ld_32x8(lane_base + mt * kTn + c0, v[mt]);
        ptx::tmem_ld_wait();
        float mn[8], mx[8];
        int imn[8], imx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { mn[i] = INFINITY; mx[i] = -INFINITY; imn[i] = 0x7fffffff; imx[i] = 0x7fffffff; }
#pragma unroll
        for (int mt = 0; mt < kMaxMT; ++mt) {
          const int p = mt * 128 + q * 32 + lane;
          if (mt < pr.nMT && p < pr.P) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (v[mt][i] < mn[i]) { mn[i] = v[mt][i]; imn[i] = p; }
              if (v[mt][i] > mx[i]) { mx[i] = v[mt][i]; imx[i] = p; }
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float omn = __shfl_xor_sync(0xffffffffu, mn[i], o);
            const int oimn = __shfl_xor_sync(0xffffffffu, imn[i], o);
            if (omn < mn[i] || (omn == mn[i] && oimn < imn[i])) { mn[i] = omn; imn[i] = oimn; }
            const float omx = __shfl_xor_sync(0xffffffffu, mx[i], o);
            const int oimx = __shfl_xor_sync(0xffffffffu, imx[i], o);
            if (omx > mx[i] || (omx == mx[i] && oimx < imx[i])) { mx[i] = omx; imx[i] = oimx; }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (lane == i) red[q * kTn + c0 + i] = ColRed{mn[i], imn[i], mx[i], imx[i]};
      }
      ptx::named_bar_sync(1, kEpiThreads);
      if (tid < kTn) {
        ColRed a = red[tid];
        for (int qq = 1; qq < 4; ++qq) {
          const ColRed o = red[qq * kTn + tid];
          if (o.mn < a.mn || (o.mn == a.mn && o.imn < a.imn)) { a.mn = o.mn; a.imn = o.imn; }
          if (o.mx > a.mx || (o.mx == a.mx && o.imx < a.imx)) { a.mx = o.mx; a.imx = o.imx; }
        }
        // colstat = (min, 1 / R, 1 / (Z + eps) [pass 2], R): the passes multiply by the reciprocals (one IEEE division per
        // column instead of one per element)
        const float R = a.mx - a.mn + 1e-8f;
        colstat[tid] = make_float4(a.mn, 1.f / R, 0.f, R);
        if (tid < pr.T) pr.arg[(int64_t)b * pr.T + tid] = make_int2(a.imn, a.imx);
      }
      ptx::named_bar_sync(1, kEpiThreads);
      SFZ_T(2)
      // pass 2: Z = sum of the thresholded, min-max normalised scores
      for (int g = 0; g < 5; ++g) {
        const int c0 = 40 * h + 8 * g;
        float v[kMaxMT][8];
#pragma unroll
        for (int mt = 0; mt < kMaxMT; ++mt)
          if (mt < pr.nMT) ptx::tmem_ld_32x8(lane_base + mt * kTn + c0, v[mt]);
        ptx::tmem_ld_wait();
        float cmn[8], cR[8], zs[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 cs = colstat[c0 + i];
          cmn[i] = cs.x;
          cR[i] = cs.y;
          zs[i] = 0.f;
        }
#pragma unroll
        for (int mt = 0; mt < kMaxMT; ++mt) {
          const int p = mt * 128 + q * 32 + lane;
          if (mt < pr.nMT && p < pr.P) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float hh = (v[mt][i] - cmn[i]) * cR[i];
              zs[i] += (hh < pr.sigma) ? 0.f : hh;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float z = ptx::warp_sum(zs[i]);
          if (lane == i) zred[q * kTn + c0 + i] = z;
        }
      }
      ptx::named_bar_sync(1, kEpiThreads);
      if (tid < kTn) {
        const float z = zred[tid] + zred[kTn + tid] + zred[2 * kTn + tid] + zred[3 * kTn + tid];
        float4 cs = colstat[tid];
        const float zeps = z + 1e-8f;
        cs.z = 1.f / zeps;
        colstat[tid] = cs;
        if (tid < pr.T) pr.stats[(int64_t)b * pr.T + tid] = make_float4(cs.x, cs.w, zeps, 0.f);
      }
      ptx::named_bar_sync(1, kEpiThreads);
      SFZ_T(3)
      // pass 3: W[t][p] (bf16, K-major SWIZZLE_128B) into the buffer that held L; row T = ones (patch sum), rows > T = 0
      for (int g = 0; g < 5; ++g) {
        const int c0 = 40 * h + 8 * g;
        float v[kMaxMT][8];
#pragma unroll
        for (int mt = 0; mt < kMaxMT; ++mt)
          if (mt < pr.nMT) ptx::tmem_ld_32x8(lane_base + mt * kTn + c0, v[mt]);
        ptx::tmem_ld_wait();
        float cmn[8], cR[8], cZ[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 cs = colstat[c0 + i];
          cmn[i] = cs.x;
          cR[i] = cs.y;
          cZ[i] = cs.z;
        }
#pragma unroll
        for (int mt = 0; mt < kMaxMT; ++mt) {
          const int p = mt * 128 + q * 32 + lane;
          const int kb = p >> 6;
          if (mt < pr.nMT && kb < pr.kbP) {
            uint32_t wbits[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int t = c0 + i;
              float wv = 0.f;
              if (p < pr.P) {
                if (t < pr.T) {
                  const float hh = (v[mt][i] - cmn[i]) * cR[i];
                  wv = (hh < pr.sigma) ? 0.f : hh * cZ[i];
                } else if (t == pr.T) {
                  wv = 1.f;
                }
              }
              wbits[i] = ptx::pack_bf16x2(wv, 0.f);
            }
            const uint32_t unit = static_cast<uint32_t>(p & 63) >> 3;
            const uint32_t base = lw_addr + kb * kBlk + (static_cast<uint32_t>(p & 7) << 1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t t = static_cast<uint32_t>(c0 + i);
              ptx::st_shared_u16(base + t * 128u + ((unit ^ (t & 7u)) << 4), wbits[i]);
            }
          }
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(wready);
      SFZ_T(4)
      // ---- GEMM 2 epilogues: G^T tiles (lanes = features d, columns = tokens) -> raw G[t][d], one coalesced line per token
      for (int j = 0; j < pr.nU; ++j) {
        const int buf = j & 1;
        ptx::mbar_wait(&gfull[buf], guse[buf] & 1u);
        ++guse[buf];
        ptx::tc_fence_after();
        float v[40];
        ptx::tmem_ld_32x32(lane_base + buf * 128 + 40 * h, v);
        ptx::tmem_ld_32x8(lane_base + buf * 128 + 40 * h + 32, v + 32);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&gempty[buf]);
        const int d = j * 128 + q * 32 + lane;
        float* gcol = pr.g_hat + (int64_t)b * pr.T * pr.D + d;
#pragma unroll
        for (int i = 0; i < 40; ++i) {
          const int t = 40 * h + i;
          if (t < pr.T) gcol[(int64_t)t * pr.D] = v[i];
          else if (t == pr.T && pr.pooled != nullptr) pr.pooled[(int64_t)b * pr.D + d] = v[i] / static_cast<float>(pr.P);
        }
      }
      SFZ_T(5)
      ptx::named_bar_sync(1, kEpiThreads);          // the raw rows of every warp are visible
      // ---- g^ = n(G): one warp per token row, two rows in flight
      for (int t0 = ew; t0 < pr.T; t0 += 2 * kEpiWarps) {
        float4 x[2][kMaxKsD / 2];
        float acc[2] = {0.f, 0.f};
#pragma unroll
        for (int r2 = 0; r2 < 2; ++r2) {
          const int t = t0 + r2 * kEpiWarps;
          const float4* row = reinterpret_cast<const float4*>(pr.g_hat + ((int64_t)b * pr.T + (t < pr.T ? t : t0)) * pr.D);
#pragma unroll
          for (int i = 0; i < kMaxKsD / 2; ++i) {
            const int c = i * 32 + lane;
            x[r2][i] = (c < pr.D / 4) ? row[c] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int r2 = 0; r2 < 2; ++r2)
#pragma unroll
          for (int i = 0; i < kMaxKsD / 2; ++i)
            acc[r2] += x[r2][i].x * x[r2][i].x + x[r2][i].y * x[r2][i].y + x[r2][i].z * x[r2][i].z + x[r2][i].w * x[r2][i].w;
#pragma unroll
        for (int r2 = 0; r2 < 2; ++r2) {
          const int t = t0 + r2 * kEpiWarps;
          const float gn = sqrtf(ptx::warp_sum(acc[r2]));
          if (t < pr.T) {
            const float r = 1.f / fmaxf(gn, 1e-12f);
            float4* row = reinterpret_cast<float4*>(pr.g_hat + ((int64_t)b * pr.T + t) * pr.D);
#pragma unroll
            for (int i = 0; i < kMaxKsD / 2; ++i) {
              const int c = i * 32 + lane;
              if (c < pr.D / 4) row[c] = make_float4(x[r2][i].x * r, x[r2][i].y * r, x[r2][i].z * r, x[r2][i].w * r);
            }
            if (lane == 0) pr.gnorm[(int64_t)b * pr.T + t] = gn;
          }
        }
      }
      ptx::named_bar_sync(1, kEpiThreads);          // gnp / colstat are reused by the next sample
      SFZ_T(6)
    }
#ifdef CLIPK_SFZ_PROF
    if (blockIdx.x == 3 && warp == kEpiWarp0 && lane == 0)
      printf("sfz epilogue warp 4, %d samples, total %lld cyc: l_hat %lld | wait scores %lld | pass1 %lld | pass2 %lld | pass3+publish %lld | "
             "G tiles (incl. waiting) %lld | normalise %lld\n", iter, clock64() - pt0, pt[0], pt[1], pt[2], pt[3], pt[4], pt[5], pt[6]);
#endif
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace sfz
