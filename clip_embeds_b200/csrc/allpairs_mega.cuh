// PACL all-pairs forward / backward as ONE persistent, dependency-driven kernel per direction (sm_100a).
//
// Why: launched as separate GEMMs (pacl_allpairs.cu, "staged" path) the per-group scratch tensors A / X|E / G only
// amortise launch ramps and tails when a group is ~128 images, i.e. ~500 MB of scratch per group: every kernel then
// streams its operands through HBM (24 GB per step at C2, ncu: 2.5-4.4 TB/s next to the tensor work).  Here all GEMMs
// of all groups are tiles of one global, statically ordered tile sequence that 74 CTA pairs walk round-robin:
//
//   block of `depth` groups (lock-step):   for phase in phases:  for group in block:  tiles of (phase, group)
//
// so a group can be as small as a few images: its scratch (a ring of `slots` group-sized slots) lives in L2, there
// are no launch boundaries, and the tail of one GEMM overlaps the head of the next.  Dependencies are explicit:
// every epilogue warp publishes "tile done" on a per-(group, phase) counter in global memory (after its TMA stores
// have completed), and the TMA producer / epilogue warps of a consumer tile poll the counters of the phases it reads
// (same group) or overwrites (the group that used the scratch slot before).  All CTAs are co-resident (one per SM)
// and a tile only ever waits for tiles that are EARLIER in the global sequence, so the waits cannot deadlock.
//
// The pipeline inside a CTA pair is the one of gemm2_engine.cuh (TMA producer / tcgen05 cta_group::2 MMA issuer /
// 8 epilogue warps with per-warp TMA-store staging); the phase of a tile selects the operands, the tile shape and the
// fused epilogue functor at run time, each with compile-time constants (PhaseCfg).
//
//   forward :  ACT   A = act(T V_i^T), num += <u, t^>        USQ   usq += |A V_i|^2
//   backward:  ACTS  A, X = <t^, V>      GNEG  Gn = -beta A V_i      DS   d = Gn V_i^T, X -> E (in place), dsdot
//              DT    dt^ += E V  (red.add.v4.f32; K folds the images of the group)      DV   dV_i = A^T Gn + E^T T^ - ...
#pragma once
#ifndef CLIPK_MEGA_PROF
#define CLIPK_MEGA_PROF 0
#endif
#include "epilogues.cuh"
#include "gemm2_engine.cuh"

namespace mega {

enum Phase : int { PH_ACT = 0, PH_USQ = 1, PH_ACTS = 2, PH_GNEG = 3, PH_DS = 4, PH_DT = 5, PH_DV = 6 };
constexpr int kMaxPhases = 5;
constexpr int kBNP = 192;           // tile columns when N is the patch axis (N tails are trimmed per tile)
constexpr int kBND = 256;           // tile columns when N is the feature axis
constexpr int kStages = 5;
constexpr int kStageBytes = 32 * 1024;      // A: 16 KB at +0, B half: up to 16 KB at +16 KB
constexpr int kABytes = 16 * 1024;
constexpr int kChunkBytes = 32 * 64;        // one epilogue warp's [32 x 32] bf16 chunk
constexpr int kThreads = eng2::kThreads;
constexpr int kEpiWarp0 = eng2::kEpiWarp0;
constexpr int kEpiWarps = eng2::kEpiWarps;
constexpr int kStagingBytes = kEpiWarps * 2 * kChunkBytes;                  // 32 KB per staging area
// staging areas: [out] and [out2 | in] -- ACTS's second output and DS's input chunks share one (safe: an epilogue
// warp drains its bulk stores whenever it moves to a tile of another job)
constexpr int kSmemTotal = kStages * kStageBytes + 2 * kStagingBytes + 1024 /*barriers*/ + 1024 /*align*/;
static_assert(kSmemTotal <= 227 * 1024, "shared memory budget");

struct Maps {
  CUtensorMap T_k;      // T [Bt][D]            K-major, box 64 x 128   (A operand of ACT / ACTS)
  CUtensorMap V_k;      // V [Bi][P][D]         K-major, box 64 x 96    (B operand, N = patch axis)
  CUtensorMap V_mn;     // V                    MN-major (n = d), box 64 x 64   (B operand, K = patch axis)
  CUtensorMap A_k;      // A scratch [S][Bt][Ppad]  K-major, box 64 x 128
  CUtensorMap A_mn;     // A scratch            MN-major (m = patch), box 64 x 64
  CUtensorMap E_k, E_mn;   // E (X) scratch, same geometry as A
  CUtensorMap G_k;      // G scratch [S][Bt][D] K-major, box 64 x 128
  CUtensorMap G_mn;     // G scratch            MN-major (n = d), box 64 x 64
  CUtensorMap Th_mn;    // That [Bt][D]         MN-major (n = d), box 64 x 64
  CUtensorMap oA, oE, oG, odV;   // [32 x 32] SWIZZLE_64B boxes for the epilogue warps (oE also loads the X chunks)
};

struct Sched {
  int Bi, Bt, P, Ppad, D, act;
  int gs;            // images per group
  int ngroups;
  int depth;         // groups per lock-step block
  int slots;         // scratch ring slots (group g uses slot g % slots)
  int nph;           // phases of this launch
  int ph[kMaxPhases];
  int tm[kMaxPhases], tn[kMaxPhases];   // pair tiles per batch item (DT: per split)
  int dep_same[kMaxPhases];             // index (into ph[]) of the phase of the SAME group this phase reads, or -1
  int dep_ring[kMaxPhases][2];          // phases of group g - slots that must be complete before this phase writes
  int dt_spb;        // DT: images per K split
  unsigned need_full[kMaxPhases], need_last[kMaxPhases];   // completion count of a phase: full group / last (short) group
  int total_tiles;
  unsigned* done;    // [ngroups][nph] finished (tile, epilogue warp) counts
  // tensors the epilogues touch
  const float* rnV; const float* rnT;
  float* num; float* usq;
  const float* alpha; const float* beta;
  float* dsdot; float* dth;
  const __nv_bfloat16* V;
  float c;
  int flags;         // diagnostics (CLIPK_MEGA_FLAGS): 1 = publish without the gpu-scope fence, 2 = skip dependency waits
};

__host__ __device__ inline int group_images(const Sched& s, int g) {
  const int rem = s.Bi - g * s.gs;
  return rem < s.gs ? rem : s.gs;
}
__host__ __device__ inline int dt_splits(const Sched& s, int gi) { return (gi + s.dt_spb - 1) / s.dt_spb; }
__host__ __device__ inline int job_tiles(const Sched& s, int phi, int g) {
  const int gi = group_images(s, g);
  const int per = s.tm[phi] * s.tn[phi];
  return s.ph[phi] == PH_DT ? per * dt_splits(s, gi) : per * gi;
}

// Position in the global tile sequence.
struct Cursor {
  int blk, phi, gg;       // block, phase index, group within block
  int g, gi;              // group, images in it
  int job_begin, job_n;   // global index of the job's first tile, tiles in the job
  __device__ void init(const Sched& s) {
    blk = 0; phi = 0; gg = 0; g = 0;
    gi = group_images(s, 0);
    job_begin = 0;
    job_n = job_tiles(s, 0, 0);
  }
  __device__ void next_job(const Sched& s) {
    job_begin += job_n;
    const int nb = s.ngroups - blk * s.depth;
    const int in_blk = nb < s.depth ? nb : s.depth;
    if (++gg == in_blk) {
      gg = 0;
      if (++phi == s.nph) { phi = 0; ++blk; }
    }
    g = blk * s.depth + gg;
    gi = g < s.ngroups ? group_images(s, g) : 0;
    job_n = g < s.ngroups ? job_tiles(s, phi, g) : 0x3fffffff;
  }
  __device__ void seek(const Sched& s, int t) {
    while (t >= job_begin + job_n) next_job(s);
  }
};

struct Tile {
  int ph, phi, g, gi;
  int bl;            // image within the group (DT: split index)
  int bg;            // global image index (DT: first image of the split)
  int bs;            // scratch batch index  slot * gs + bl   (DT: of the first image of the split)
  int m0, n0, tn_idx;
  int nimg;          // DT: images folded into K
};

__device__ __forceinline__ Tile decode(const Sched& s, const Cursor& c, int t, int rank) {
  Tile x;
  x.ph = s.ph[c.phi]; x.phi = c.phi; x.g = c.g; x.gi = c.gi;
  const int tj = t - c.job_begin;
  const int per = s.tm[c.phi] * s.tn[c.phi];
  const int bl = tj / per;
  const int rem = tj - bl * per;
  const int mt = rem / s.tn[c.phi];
  x.tn_idx = rem - mt * s.tn[c.phi];
  x.m0 = mt * 256 + rank * 128;
  const int slot0 = (c.g % s.slots) * s.gs;
  if (x.ph == PH_DT) {
    const int i_first = bl * s.dt_spb;
    x.bl = bl;
    x.bg = c.g * s.gs + i_first;
    x.bs = slot0 + i_first;
    const int left = c.gi - i_first;
    x.nimg = left < s.dt_spb ? left : s.dt_spb;
  } else {
    x.bl = bl;
    x.bg = c.g * s.gs + bl;
    x.bs = slot0 + bl;
    x.nimg = 1;
  }
  const int bn = (x.ph == PH_ACT || x.ph == PH_ACTS || x.ph == PH_DS) ? kBNP : kBND;
  x.n0 = x.tn_idx * bn;
  return x;
}

// ------------------------------------------------------------------------------------------------ dependencies
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_count(const unsigned* p, unsigned need) {
  uint32_t spins = 0;
  while (ld_acquire(p) < need) {
    __nanosleep(64);
    if (++spins > (1u << 24)) {
      printf("clipk mega: dependency wait timed out (block %d thread %d, have %u need %u)\n", blockIdx.x, threadIdx.x,
             ld_acquire(p), need);
      __trap();
    }
  }
}
// block until everything job (phi, g) reads or overwrites is complete
__device__ __forceinline__ void wait_deps(const Sched& s, int phi, int g) {
  if (s.flags & 2) return;
  // one publication per CTA per tile: a job is complete at 2 x its tile count
  const int d = s.dep_same[phi];
  if (d >= 0) wait_count(s.done + (size_t)g * s.nph + d, g == s.ngroups - 1 ? s.need_last[d] : s.need_full[d]);
  const int gp = g - s.slots;
  if (gp >= 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = s.dep_ring[phi][i];
      if (r >= 0) wait_count(s.done + (size_t)gp * s.nph + r, s.need_full[r]);   // gp is never the last group
    }
  }
}

// one poll of every dependency of job (phi, g): true when none of them would block
__device__ __forceinline__ bool deps_ready(const Sched& s, int phi, int g) {
  if (s.flags & 2) return true;
  const int d = s.dep_same[phi];
  if (d >= 0 && ld_acquire(s.done + (size_t)g * s.nph + d) < (g == s.ngroups - 1 ? s.need_last[d] : s.need_full[d])) return false;
  const int gp = g - s.slots;
  if (gp >= 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = s.dep_ring[phi][i];
      if (r >= 0 && ld_acquire(s.done + (size_t)gp * s.nph + r) < s.need_full[r]) return false;
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------------ DT epilogue
// dt^[k][d] += acc   (fp32 vector reductions in L2: tiles of different groups / splits add into the same rows)
struct RedAdd {
  struct Params {
    float* C;
    int ldc, M, N;
  };
  Params p;
  __device__ explicit RedAdd(const Params& pp) : p(pp) {}
  __device__ void tile_begin(int, int, int) {}
  __device__ void chunk(int, int m, int n, float* v) {
    if (m >= p.M || n >= p.N) return;
    float* dst = p.C + (int64_t)m * p.ldc + n;
    if (n + 32 <= p.N && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                     "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                     : "memory");
    } else {
      for (int j = 0; j < 32; ++j)
        if (n + j < p.N) atomicAdd(dst + j, v[j]);
    }
  }
  __device__ void tile_end(int, int, int, int, int) {}
};

// ------------------------------------------------------------------------------------------------ phase configs
template <int PH>
struct Cfg;
template <>
struct Cfg<PH_ACT> {
  static constexpr bool PINGPONG = false;
  static constexpr int BN = kBNP; static constexpr bool A_MN = false, B_MN = false; static constexpr int NP = 1;
  using Epi = epi::PaclAct;
};
template <>
struct Cfg<PH_USQ> {
  static constexpr bool PINGPONG = true;
  static constexpr int BN = kBND; static constexpr bool A_MN = false, B_MN = true; static constexpr int NP = 1;
  using Epi = epi::Usq;
};
template <>
struct Cfg<PH_ACTS> {
  static constexpr bool PINGPONG = false;
  static constexpr int BN = kBNP; static constexpr bool A_MN = false, B_MN = false; static constexpr int NP = 1;
  using Epi = epi::PaclActS;
};
template <>
struct Cfg<PH_GNEG> {
  static constexpr bool PINGPONG = true;
  static constexpr int BN = kBND; static constexpr bool A_MN = false, B_MN = true; static constexpr int NP = 1;
  using Epi = epi::GNeg;
};
template <>
struct Cfg<PH_DS> {
  static constexpr bool PINGPONG = false;
  static constexpr int BN = kBNP; static constexpr bool A_MN = false, B_MN = false; static constexpr int NP = 1;
  using Epi = epi::DsIn;
};
template <>
struct Cfg<PH_DT> {
  static constexpr bool PINGPONG = true;
  static constexpr int BN = kBND; static constexpr bool A_MN = false, B_MN = true; static constexpr int NP = 1;
  using Epi = RedAdd;
};
template <>
struct Cfg<PH_DV> {
  static constexpr bool PINGPONG = false;
  static constexpr int BN = kBND; static constexpr bool A_MN = true, B_MN = true; static constexpr int NP = 2;
  using Epi = epi::DvOut;
};

// extents of the GEMM of a phase:  N (columns) and k-steps of operand pair q
__device__ __forceinline__ int phase_N(const Sched& s, int ph) {
  return (ph == PH_ACT || ph == PH_ACTS || ph == PH_DS) ? s.Ppad : s.D;
}
__device__ __forceinline__ int phase_ksteps(const Sched& s, const Tile& x) {
  switch (x.ph) {
    case PH_ACT: case PH_ACTS: case PH_DS: return (s.D + 63) >> 6;
    case PH_USQ: case PH_GNEG: return s.Ppad >> 6;
    case PH_DT: return x.nimg * (s.Ppad >> 6);
    default: return (s.Bt + 63) >> 6;   // PH_DV, per operand pair
  }
}
template <int BN>
__device__ __forceinline__ int ncols_of(int N, int n0) {
  const int nrem = N - n0;
  return nrem >= BN ? BN : ((nrem + 15) & ~15);
}

// epilogue functor parameters of a phase, built from the schedule (pointers are indexed by GLOBAL image)
template <int PH>
__device__ __forceinline__ typename Cfg<PH>::Epi::Params epi_params(const Sched& s) {
  const eng::OutDesc none{nullptr, 0, 0, 0, 0, 0};
  if constexpr (PH == PH_ACT) return {none, s.rnV, s.rnT, s.num, s.Bt, s.P, s.Ppad, s.act};
  else if constexpr (PH == PH_USQ) return {s.usq, s.Bt, s.D};
  else if constexpr (PH == PH_ACTS) return {none, none, s.rnV, s.rnT, s.Bt, s.P, s.Ppad, s.act};
  else if constexpr (PH == PH_GNEG) return {none, s.beta, s.Bt};
  else if constexpr (PH == PH_DS) return {none, none, s.rnV, s.alpha, s.dsdot, s.Bt, s.P, s.Ppad, s.act};
  else if constexpr (PH == PH_DT) return {s.dth, s.D, s.Bt, s.D};
  else return {none, s.V, s.rnV, s.dsdot, s.P, s.D};
}

// ------------------------------------------------------------------------------------------------ epilogue tile
struct EpiCtx {                   // state of one epilogue warp that persists across tiles
  const Maps* maps;
  const Sched* s;
  uint8_t *out_smem, *out2_smem, *in_smem;
  uint64_t *tfull_bar, *in_bar;
  uint32_t tempty_leader[2];
  uint32_t tmem_base;
  int q4, ew, half, lane;
  uint32_t sw;
  int it;                         // tiles processed by this CTA pair (TMEM buffer = it & 1)
  int slab;                       // chunks staged so far (staging buffer = slab & 1)
  int inq;                        // input chunks requested so far (in buffer = inq & 1)
  unsigned* pending;              // completion counter of the previous tile, published once its stores are complete
  int pending_g;                  // its group
  int pending_par;                // tile number & 3 (selects the CTA-level arrival counter: a warp that flushes can be
                                  // two tiles ahead of the slowest warp of its CTA, so two counters are not enough)
  unsigned* tile_done;            // [4] shared-memory arrival counters of the 8 epilogue warps
  long long pfc[8], pfn[8], pfw[8];   // per phase: cycles in chunks, chunk count, cycles in the TMA-store wait
  long long pfq[4];
  long long pf[6];                // diagnostics (flags & 4): cycles in dep wait / tfull wait / chunks / first-chunk publish
};

// lane 0 only; everything this warp wrote for that tile is complete
// The last of the 8 epilogue warps to get here publishes the tile for the whole CTA: one gpu-scope fence and one
// global reduction per CTA and tile instead of eight.  (cta-scope fence + shared-memory arrival: the publisher's
// gpu-scope fence is cumulative over what the other warps wrote before they arrived.)
__device__ __forceinline__ void publish_f(int flags, unsigned* ctr, unsigned* tile_done_slot) {
  __threadfence_block();
  const unsigned old = atomicAdd(tile_done_slot, 1u);
  if (old == kEpiWarps - 1) {
    *tile_done_slot = 0u;
    if (!(flags & 1)) asm volatile("fence.acq_rel.gpu;" ::: "memory");
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
  }
}
__device__ __forceinline__ void stage_rows(uint8_t* wbuf, const float* v, int lane, uint32_t sw) {
  const uint32_t rowbase = ptx::smem_u32(wbuf) + lane * 64;
#pragma unroll
  for (int tt = 0; tt < 4; ++tt) {
    const uint32_t addr = rowbase + ((static_cast<uint32_t>(tt) ^ sw) << 4);
    ptx::st_shared_v4(addr, ptx::pack_bf16x2(v[8 * tt + 0], v[8 * tt + 1]), ptx::pack_bf16x2(v[8 * tt + 2], v[8 * tt + 3]),
                      ptx::pack_bf16x2(v[8 * tt + 4], v[8 * tt + 5]), ptx::pack_bf16x2(v[8 * tt + 6], v[8 * tt + 7]));
  }
}

// One tile of phase PH for one epilogue warp.  (Measured: keeping these out of line costs more in call-ABI spills and
// context traffic than the shared register allocation does.)
template <int PH>
__device__ __forceinline__ void run_tile(EpiCtx& cx, const Tile xt) {
  using C = Cfg<PH>;
  using Epi = typename C::Epi;
  constexpr int BN = C::BN;
  constexpr bool TMA_OUT = eng::epi_tma_out<Epi>::value;
  constexpr bool TMA_OUT2 = eng::epi_tma_out2<Epi>::value;
  constexpr bool CHUNK_IN = eng::epi_chunk_in<Epi>::value;
  constexpr bool HAS_SIDE = eng::epi_has_side<Epi>::value;
  [[maybe_unused]] const long long pf_e = CLIPK_MEGA_PROF ? clock64() : 0;
  // Everything the chunk loop touches is copied into local values first: `cx`, `s` and the tile live in memory, and
  // every asm volatile with a "memory" clobber (mbarrier waits, shared stores, fences, TMA) would otherwise force
  // them to be re-loaded, which serialises the chunk math.
  const Sched& s = *cx.s;
  const Maps& maps = *cx.maps;
  const int lane = cx.lane, q4 = cx.q4, ew = cx.ew, half = cx.half;
  const uint32_t sw = cx.sw;
  int slab = cx.slab, inq = cx.inq;
  unsigned* pending = cx.pending;
  unsigned* const pending_slot = cx.tile_done + cx.pending_par;
  uint8_t* const c_out = cx.out_smem + ew * 2 * kChunkBytes;
  uint8_t* const c_out2 = cx.out2_smem + ew * 2 * kChunkBytes;
  uint8_t* const c_in = cx.in_smem + ew * 2 * kChunkBytes;
  uint64_t* const c_inbar = cx.in_bar + ew * 2;
  uint64_t* const c_tfull = cx.tfull_bar;
  const uint32_t c_tmem = cx.tmem_base;
  const int c_it = cx.it;
  const uint32_t c_tempty = cx.tempty_leader[c_it & 1];
  const int flags = s.flags;
  struct { int bg, bs, m0, n0, tn_idx, g, phi; } x = {xt.bg, xt.bs, xt.m0, xt.n0, xt.tn_idx, xt.g, xt.phi};
  unsigned* const my_counter = s.done + (size_t)xt.g * s.nph + xt.phi;
  const int N = phase_N(s, PH);
  const int b = x.bg;                                          // functor arrays are indexed by global image
  const int bo = PH == PH_DV ? x.bg : x.bs;                    // batch coordinate of the TMA-stored output
  const CUtensorMap* omap = PH == PH_DV ? &maps.odV : (PH == PH_GNEG ? &maps.oG : (PH == PH_DS ? &maps.oE : &maps.oA));
  const int m = x.m0 + q4 * 32 + lane;
  const int buf = c_it & 1;
  const uint32_t bphase = (c_it >> 1) & 1;
  Epi epi(epi_params<PH>(s));
  const int jn = (N - x.n0 + 63) >> 6;
  const int jmax = jn < BN / 64 ? jn : BN / 64;
  auto issue_in = [&](int ncol) {
    if constexpr (CHUNK_IN) {
      if (lane == 0) {
        uint64_t* bar = c_inbar + (inq & 1);
        ptx::mbar_arrive_expect_tx(bar, kChunkBytes);
        ptx::tma_load_3d(c_in + (inq & 1) * kChunkBytes, &maps.oE, bar, ncol, x.m0 + q4 * 32, x.bs);
      }
      ++inq;
    }
  };
  [[maybe_unused]] typename eng::side_of<Epi>::type side{}, side_next{};
  epi.tile_begin(b, m, x.n0);
  if constexpr (HAS_SIDE) side = epi.pre(b, m, x.n0 + half * 32);
  issue_in(x.n0 + half * 32);
  [[maybe_unused]] const long long pf_a = CLIPK_MEGA_PROF ? clock64() : 0;
  if (CLIPK_MEGA_PROF && (flags & 4)) cx.pfq[2] += pf_a - pf_e;
  ptx::mbar_wait(&c_tfull[buf], bphase);
  ptx::tc_fence_after();
  [[maybe_unused]] const long long pf_b = CLIPK_MEGA_PROF ? clock64() : 0;
  const uint32_t tacc = c_tmem + (static_cast<uint32_t>(q4 * 32) << 16) + buf * 256;
  auto release_tmem = [&]() {
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive_cluster(c_tempty);
  };
  long long pf_w = 0;
  auto process = [&](float* v, int j) {
    const int c = 2 * j + half;
    if (j + 1 < jmax) {                          // side data / input chunk of this warp's next chunk
      if constexpr (HAS_SIDE) side_next = epi.pre(b, m, x.n0 + (c + 2) * 32);
      issue_in(x.n0 + (c + 2) * 32);
    }
    [[maybe_unused]] float v2[32];
    if constexpr (TMA_OUT2 || CHUNK_IN) {
      [[maybe_unused]] uint32_t in[16];
      if constexpr (CHUNK_IN) {
        const int iq = inq - (j + 1 < jmax ? 2 : 1);           // request number of THIS chunk
        ptx::mbar_wait(c_inbar + (iq & 1), (iq >> 1) & 1);
        const uint32_t rowbase = ptx::smem_u32(c_in + (iq & 1) * kChunkBytes) + lane * 64;
#pragma unroll
        for (int tt = 0; tt < 4; ++tt) {
          const uint4 q = ptx::ld_shared_v4(rowbase + ((static_cast<uint32_t>(tt) ^ sw) << 4));
          in[4 * tt] = q.x; in[4 * tt + 1] = q.y; in[4 * tt + 2] = q.z; in[4 * tt + 3] = q.w;
        }
        __syncwarp();
      }
      epi.chunk(b, m, x.n0 + c * 32, v, side, in, v2);
    } else if constexpr (HAS_SIDE) {
      epi.chunk(b, m, x.n0 + c * 32, v, side);
    } else {
      epi.chunk(b, m, x.n0 + c * 32, v);
    }
    if constexpr (HAS_SIDE) side = side_next;
    if constexpr (TMA_OUT) {
      uint8_t* wbuf = c_out + (slab & 1) * kChunkBytes;
      uint8_t* wbuf2 = c_out2 + (slab & 1) * kChunkBytes;
      stage_rows(wbuf, v, lane, sw);
      if constexpr (TMA_OUT2) stage_rows(wbuf2, v2, lane, sw);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_3d(omap, wbuf, x.n0 + c * 32, x.m0 + q4 * 32, bo);
        if constexpr (TMA_OUT2) ptx::tma_store_3d(&maps.oE, wbuf2, x.n0 + c * 32, x.m0 + q4 * 32, bo);
        ptx::bulk_commit_group();
        if (pending != nullptr) {                // the previous tile's stores: all but this group are complete
          ptx::bulk_wait_group<1>();
          publish_f(flags, pending, pending_slot);
        } else {
          [[maybe_unused]] const long long pf_c = CLIPK_MEGA_PROF ? clock64() : 0;
          ptx::bulk_wait_group_read<1>();
          if (CLIPK_MEGA_PROF) pf_w += clock64() - pf_c;
        }
      }
      pending = nullptr;
      __syncwarp();
      ++slab;
    } else {
      if (pending != nullptr) {                  // (warp-uniform) no store in this phase: drain and publish now
        if (lane == 0) {
          ptx::bulk_wait_group<0>();
          publish_f(flags, pending, pending_slot);
        }
        pending = nullptr;
        __syncwarp();
      }
    }
  };
  if constexpr (!C::PINGPONG) {               // register-heavy functors: no TMEM ping-pong (keeps their math's ILP)
#pragma unroll 1
    for (int j = 0; j < jmax; ++j) {
      float va[32];
      ptx::tmem_ld_32x32(tacc + (2 * j + half) * 32, va);
      ptx::tmem_ld_wait();
      if (j + 1 == jmax) release_tmem();
      process(va, j);
    }
  } else {
    float va[32], vb[32];
    ptx::tmem_ld_32x32(tacc + half * 32, va);
#pragma unroll 1
    for (int j = 0; j < jmax; j += 2) {
      ptx::tmem_ld_wait();
      if (j + 1 < jmax) ptx::tmem_ld_32x32(tacc + (2 * (j + 1) + half) * 32, vb);
      else release_tmem();
      process(va, j);
      if (j + 1 < jmax) {
        ptx::tmem_ld_wait();
        if (j + 2 < jmax) ptx::tmem_ld_32x32(tacc + (2 * (j + 2) + half) * 32, va);
        else release_tmem();
        process(vb, j + 1);
      }
    }
  }
  epi.tile_end(b, m, x.n0, x.tn_idx, half);
  __syncwarp();
  if (CLIPK_MEGA_PROF && (flags & 4)) {
    cx.pf[1] += pf_b - pf_a;
    cx.pf[2] += clock64() - pf_b;
    cx.pf[3] += jmax;
    cx.pfc[PH] += clock64() - pf_b;
    cx.pfn[PH] += jmax;
    cx.pfw[PH] += pf_w;
  }
  cx.slab = slab;
  cx.inq = inq;
  cx.pending = my_counter;
  cx.pending_g = xt.g;
  cx.pending_par = c_it & 3;
}

// ------------------------------------------------------------------------------------------------ the kernel
struct Pipe {            // per-role ring / accumulator state that persists across tiles
  int stage;
  uint32_t phase;
  int it;                // tiles processed by this CTA pair (TMEM buffer = it & 1)
};

// All tiles of the current job that belong to this CTA pair, for one epilogue warp.  Not inlined: each phase gets its
// own register allocation (as in the stand-alone engine), the call happens once per job and the persistent state
// travels through `cx`.
template <int PH>
__device__ __forceinline__ int run_job(EpiCtx& cx, const Cursor& cur, int t, int rank, int num_clusters) {
  const Sched& s = *cx.s;
  const int job_end = cur.job_begin + cur.job_n;
  for (; t < job_end; t += num_clusters) {
    [[maybe_unused]] const long long q0 = CLIPK_MEGA_PROF ? clock64() : 0;
    const Tile x = decode(s, cur, t, rank);
    [[maybe_unused]] const long long q1 = CLIPK_MEGA_PROF ? clock64() : 0;
    run_tile<PH>(cx, x);
    if (CLIPK_MEGA_PROF) {
      cx.pfq[0] += q1 - q0;
      cx.pfq[1] += clock64() - q1;
    }
    ++cx.it;
  }
  return t;
}

// The three warp roles are separate, non-inlined functions: each gets its own register allocation (the TMA producer
// and the MMA issuer are single latency-critical threads and must not inherit the epilogue's spills).
template <bool BWD>
__device__ __forceinline__ void producer_role(const Maps& maps, const Sched& s, uint8_t* smem, uint64_t* full_bar,
                                           uint64_t* empty_bar, int rank, int cluster_id, int num_clusters) {
  const bool leader = rank == 0;
    if (ptx::elect_one()) {
      Cursor cur;
      cur.init(s);
      int stage = 0;
      uint32_t phase = 0;
      int checked_job = -1;
      const uint32_t full_leader = ptx::mapa(ptx::smem_u32(&full_bar[0]), 0);   // the leader's full barriers
      for (int t = cluster_id; t < s.total_tiles; t += num_clusters) {
        cur.seek(s, t);
        if (cur.job_begin != checked_job) {          // first tile of this job for this CTA: operands complete?
          wait_deps(s, cur.phi, cur.g);
          ptx::fence_proxy_async_all();              // the TMA loads below read what other SMs' TMA stores wrote
          checked_job = cur.job_begin;
        }
        const Tile x = decode(s, cur, t, rank);
        const int N = phase_N(s, x.ph);
        const bool np = (x.ph == PH_ACT || x.ph == PH_ACTS || x.ph == PH_DS);
        const int ncols = np ? ncols_of<kBNP>(N, x.n0) : ncols_of<kBND>(N, x.n0);
        const int nb0 = x.n0 + rank * (ncols >> 1);
        const int nks = phase_ksteps(s, x);
        const uint32_t tx = kABytes + (np ? (kBNP / 2) * 128 : (kBND / 2) * 128);
        const int kpi = s.Ppad >> 6;                 // k-steps per image (DT)
        // one pipeline step: wait for the stage, arm the leader's barrier, return the stage's A / B addresses
        auto step = [&](uint8_t*& sa, uint8_t*& sb, uint32_t& fb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          sa = smem + stage * kStageBytes;
          sb = sa + kABytes;
          fb = full_leader + stage * 8;
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * tx);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        };
        uint8_t *sa, *sb;
        uint32_t fb;
        switch (x.ph) {
          case PH_ACT: case PH_ACTS:
            for (int ks = 0; ks < nks; ++ks) {
              step(sa, sb, fb);
              ptx::tma_load_3d_2sm(sa, &maps.T_k, fb, ks << 6, x.m0, 0);
              ptx::tma_load_3d_2sm(sb, &maps.V_k, fb, ks << 6, nb0, x.bg);
            }
            break;
          case PH_DS:
            for (int ks = 0; ks < nks; ++ks) {
              step(sa, sb, fb);
              ptx::tma_load_3d_2sm(sa, &maps.G_k, fb, ks << 6, x.m0, x.bs);
              ptx::tma_load_3d_2sm(sb, &maps.V_k, fb, ks << 6, nb0, x.bg);
            }
            break;
          case PH_USQ: case PH_GNEG:
            for (int ks = 0; ks < nks; ++ks) {
              step(sa, sb, fb);
              ptx::tma_load_3d_2sm(sa, &maps.A_k, fb, ks << 6, x.m0, x.bs);
              ptx::tma_load_3d_2sm(sb, &maps.V_mn, fb, nb0, ks << 6, x.bg);
              ptx::tma_load_3d_2sm(sb + 8192, &maps.V_mn, fb, nb0 + 64, ks << 6, x.bg);
            }
            break;
          case PH_DT:
            for (int sub = 0; sub < x.nimg; ++sub)
              for (int kk = 0; kk < kpi; ++kk) {
                step(sa, sb, fb);
                ptx::tma_load_3d_2sm(sa, &maps.E_k, fb, kk << 6, x.m0, x.bs + sub);
                ptx::tma_load_3d_2sm(sb, &maps.V_mn, fb, nb0, kk << 6, x.bg + sub);
                ptx::tma_load_3d_2sm(sb + 8192, &maps.V_mn, fb, nb0 + 64, kk << 6, x.bg + sub);
              }
            break;
          default:   // PH_DV: two operand pairs
            for (int ks = 0; ks < nks; ++ks) {
              step(sa, sb, fb);
              ptx::tma_load_3d_2sm(sa, &maps.A_mn, fb, x.m0, ks << 6, x.bs);
              ptx::tma_load_3d_2sm(sa + 8192, &maps.A_mn, fb, x.m0 + 64, ks << 6, x.bs);
              ptx::tma_load_3d_2sm(sb, &maps.G_mn, fb, nb0, ks << 6, x.bs);
              ptx::tma_load_3d_2sm(sb + 8192, &maps.G_mn, fb, nb0 + 64, ks << 6, x.bs);
            }
            for (int ks = 0; ks < nks; ++ks) {
              step(sa, sb, fb);
              ptx::tma_load_3d_2sm(sa, &maps.E_mn, fb, x.m0, ks << 6, x.bs);
              ptx::tma_load_3d_2sm(sa + 8192, &maps.E_mn, fb, x.m0 + 64, ks << 6, x.bs);
              ptx::tma_load_3d_2sm(sb, &maps.Th_mn, fb, nb0, ks << 6, 0);
              ptx::tma_load_3d_2sm(sb + 8192, &maps.Th_mn, fb, nb0 + 64, ks << 6, 0);
            }
            break;
        }
      }
    }
}

__device__ __forceinline__ void mma_role(const Sched& s, uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar,
                                      uint64_t* tfull_bar, uint64_t* tempty_bar, uint32_t tmem_base, int cluster_id,
                                      int num_clusters) {
  const bool leader = true;
    if (leader && ptx::elect_one()) {
      Cursor cur;
      cur.init(s);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = cluster_id; t < s.total_tiles; t += num_clusters, ++it) {
        cur.seek(s, t);
        const Tile x = decode(s, cur, t, 0);
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        ptx::mbar_wait(&tempty_bar[buf], bphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * 256;
        const int N = phase_N(s, x.ph);
        const bool np = (x.ph == PH_ACT || x.ph == PH_ACTS || x.ph == PH_DS);
        const int ncols = np ? ncols_of<kBNP>(N, x.n0) : ncols_of<kBND>(N, x.n0);
        const bool a_mn = x.ph == PH_DV;
        const bool b_mn = !np;
        const uint32_t idesc = ptx::umma_idesc_bf16(256, ncols, a_mn ? 1 : 0, b_mn ? 1 : 0);
        const int nks = phase_ksteps(s, x) * (x.ph == PH_DV ? 2 : 1);
        uint32_t accum = 0;
        for (int ks = 0; ks < nks; ++ks) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ad = a_mn ? ptx::umma_desc(sa + kk * 2048, 8192, 1024) : ptx::umma_desc(sa + kk * 32, 16, 1024);
            const uint64_t bd = b_mn ? ptx::umma_desc(sb + kk * 2048, 8192, 1024) : ptx::umma_desc(sb + kk * 32, 16, 1024);
            ptx::mma_bf16_ss_2sm(tmem_d, ad, bd, idesc, accum);
            accum = 1;
          }
          ptx::mma_commit_2sm(&empty_bar[stage], 3);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit_2sm(&tfull_bar[buf], 3);
      }
    }
}

template <bool BWD>
__device__ __forceinline__ void epilogue_role(const Maps& maps, const Sched& s, uint8_t* out_smem, uint8_t* out2_smem,
                                           uint8_t* in_smem, uint64_t* tfull_bar, uint64_t* tempty_bar, uint64_t* in_bar,
                                           uint32_t tmem_base, int rank, int cluster_id, int num_clusters) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
    const int q4 = warp & 3;
    const int ew = warp - kEpiWarp0;
    const int half = ew >> 2;
    const uint32_t tempty_leader[2] = {ptx::mapa(ptx::smem_u32(&tempty_bar[0]), 0),
                                       ptx::mapa(ptx::smem_u32(&tempty_bar[1]), 0)};
    const uint32_t sw = (static_cast<uint32_t>(lane) >> 1) & 3u;
    Cursor cur;
    cur.init(s);
    int checked_job = -1;

    EpiCtx cx;
    cx.maps = &maps; cx.s = &s;
    cx.out_smem = out_smem; cx.out2_smem = out2_smem; cx.in_smem = in_smem;
    cx.tfull_bar = tfull_bar; cx.in_bar = in_bar;
    cx.tempty_leader[0] = tempty_leader[0]; cx.tempty_leader[1] = tempty_leader[1];
    cx.tmem_base = tmem_base;
    cx.q4 = q4; cx.ew = ew; cx.half = half; cx.lane = lane; cx.sw = sw;
    cx.it = 0; cx.slab = 0; cx.inq = 0; cx.pending = nullptr; cx.pending_g = -1; cx.pending_par = 0;
    cx.tile_done = reinterpret_cast<unsigned*>(in_bar + 2 * kEpiWarps);
    for (int i = 0; i < 6; ++i) cx.pf[i] = 0;
    for (int i = 0; i < 4; ++i) cx.pfq[i] = 0;
    for (int i = 0; i < 8; ++i) cx.pfc[i] = cx.pfn[i] = cx.pfw[i] = 0;
    [[maybe_unused]] const long long pf_start = CLIPK_MEGA_PROF ? clock64() : 0;

    for (int t = cluster_id; t < s.total_tiles;) {
      cur.seek(s, t);
      if (cur.job_begin != checked_job) {     // the epilogue reads in-kernel products too (X chunks, dsdot)
        [[maybe_unused]] const long long pf_d = CLIPK_MEGA_PROF ? clock64() : 0;
        if (CLIPK_MEGA_PROF) ++cx.pf[5];
        // Never block with an unpublished tile: whatever this job waits for may, through other CTA pairs that are
        // themselves waiting, depend on it.  When the dependencies are already complete (the common case) nothing is
        // flushed and the tile is published, as usual, behind the first chunk of the next tile.
        if (lane == 0) ptx::bulk_wait_group_read<0>();   // [out2 | in] staging is shared between phases
        const bool ready = __shfl_sync(0xffffffffu, lane == 0 ? (deps_ready(s, cur.phi, cur.g) ? 1 : 0) : 0, 0) != 0;
        if (!ready) {
          if (cx.pending != nullptr) {
            if (lane == 0) {
              ptx::bulk_wait_group<0>();
              publish_f(s.flags, cx.pending, cx.tile_done + cx.pending_par);
            }
            cx.pending = nullptr;
          }
          if (lane == 0) wait_deps(s, cur.phi, cur.g);
        }
        __syncwarp();
        ptx::fence_proxy_async_all();
        checked_job = cur.job_begin;
        if (CLIPK_MEGA_PROF) cx.pf[0] += clock64() - pf_d;
      }
      if constexpr (!BWD) {
        if (s.ph[cur.phi] == PH_ACT) t = run_job<PH_ACT>(cx, cur, t, rank, num_clusters);
        else t = run_job<PH_USQ>(cx, cur, t, rank, num_clusters);
      } else {
        switch (s.ph[cur.phi]) {
          case PH_ACTS: t = run_job<PH_ACTS>(cx, cur, t, rank, num_clusters); break;
          case PH_GNEG: t = run_job<PH_GNEG>(cx, cur, t, rank, num_clusters); break;
          case PH_DS: t = run_job<PH_DS>(cx, cur, t, rank, num_clusters); break;
          case PH_DT: t = run_job<PH_DT>(cx, cur, t, rank, num_clusters); break;
          default: t = run_job<PH_DV>(cx, cur, t, rank, num_clusters); break;
        }
      }
    }
    if (lane == 0) {
      ptx::bulk_wait_group<0>();
      if (cx.pending != nullptr) publish_f(s.flags, cx.pending, cx.tile_done + cx.pending_par);
    }
    if (CLIPK_MEGA_PROF && (s.flags & 4) && lane == 0 && (blockIdx.x == 2 || blockIdx.x == 77) && (warp == kEpiWarp0 || warp == kEpiWarp0 + 5))
      printf("mega prof blk %d warp %d: tiles %d total %lld cyc | jobchg %lld (%lld cyc) tfull-wait %lld  chunks %lld (%lld chunks) "
             "publish-in-chunk %lld\n", blockIdx.x, warp, cx.it, clock64() - pf_start, cx.pf[5], cx.pf[0], cx.pf[1], cx.pf[2],
             cx.pf[3], cx.pf[4]);
    if (CLIPK_MEGA_PROF && (s.flags & 4) && lane == 0 && blockIdx.x == 2 && warp == kEpiWarp0)
      printf("   decode %lld  run_tile %lld  (entry->tfull-wait %lld)\n", cx.pfq[0], cx.pfq[1], cx.pfq[2]);
    if (CLIPK_MEGA_PROF && (s.flags & 4) && lane == 0 && blockIdx.x == 2 && warp == kEpiWarp0)
      for (int i = 0; i < 7; ++i)
        if (cx.pfn[i] > 0)
          printf("   phase %d: %lld chunks, %lld cyc/chunk, of which store-read wait %lld\n", i, cx.pfn[i], cx.pfc[i] / cx.pfn[i],
                 cx.pfw[i] / cx.pfn[i]);
  }

template <bool BWD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
allpairs_mega_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Sched s_param) {
  // The role functions read the schedule through a pointer: keep it in shared memory (a generic load from the
  // kernel-parameter window costs a memory round trip, an LDS ~30 cycles).
  __shared__ Sched s;
  for (int i = threadIdx.x; i < (int)(sizeof(Sched) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(&s)[i] = reinterpret_cast<const uint32_t*>(&s_param)[i];
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* out_smem = smem + kStages * kStageBytes;
  uint8_t* out2_smem = out_smem + kStagingBytes;
  uint8_t* in_smem = out2_smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out2_smem + kStagingBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* in_bar = tempty_bar + 4;     // [kEpiWarps][2]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(ptx::cluster_ctarank());
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    const CUtensorMap* m = reinterpret_cast<const CUtensorMap*>(&maps);
    for (int i = 0; i < (int)(sizeof(Maps) / sizeof(CUtensorMap)); ++i) ptx::prefetch_tmap(m + i);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], 2 * kEpiWarps);
    }
    for (int i = 0; i < 2 * kEpiWarps; ++i) ptx::mbar_init(&in_bar[i], 1);
    for (int i = 0; i < 4; ++i) reinterpret_cast<unsigned*>(in_bar + 2 * kEpiWarps)[i] = 0u;   // CTA-level tile arrival counters
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

  if (warp == 0) {
    producer_role<BWD>(maps, s, smem, full_bar, empty_bar, rank, cluster_id, num_clusters);
  } else if (warp == 1) {
    if (leader) mma_role(s, smem, full_bar, empty_bar, tfull_bar, tempty_bar, tmem_base, cluster_id, num_clusters);
  } else if (warp >= kEpiWarp0) {
    epilogue_role<BWD>(maps, s, out_smem, out2_smem, in_smem, tfull_bar, tempty_bar, in_bar, tmem_base, rank, cluster_id,
                  num_clusters);
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace mega
