// PACL all-pairs text-conditioned scoring, forward and backward (BASELINE north_star (1)-(3)).
//
//   score[i,k] = c * < n(u_ik), n(t_k) >,   u_ik = sum_p sigmoid(10 <n(t_k), n(V_ip)>) V_ip
//
// Reference semantics: the eval call `model(img_i, texts)` of PACL/eval_pacl.py:53-57 / :303-309 (one image, all
// texts, diagonal of `c * image_features @ text_features.T`) looped over images; patch_alignment = pacl.py:120-133,
// pooling + normalise = pacl.py:143-145.
//
// Images are processed in GROUPS of `group` images.  Within a group every contraction is one launch of the
// tcgen05 engine with a fused epilogue; the [group, Bt, P] activation tiles live in a caller-provided scratch that
// is re-used by every group (sized to stay L2-resident), so the [Bi, Bt, P] tensor is never materialised in HBM and
// only O(Bi*Bt) + O(Bi*P) statistics are saved for backward (recompute, flash-style).
//
//  forward :  K1  A = act(T V_i^T)            (epilogue PaclAct: also num = <u, t^>; A leaves the SM by TMA store)
//             K2  u = A V_i  -> usq = |u|^2   (epilogue Usq: u never stored)
//  backward:  K1' (recompute A, also X = <t^,V>: epilogue PaclActS)   K2' Gn = -beta u   (epilogue GNeg, TMA store)
//             K4  d = Gn V_i^T  -> E, dsdot   (epilogue DsIn: the X chunk of each accumulator chunk arrives by TMA,
//                 E overwrites X in place)
//             K5  dt^ += E V                  (K folds (image, patch); fp32 accumulate)
//             K6  dV_i = A^T Gn + E^T T^ - rnV^2 dsdot V      (two operand pairs, epilogue DvOut, TMA store)
//  with G_ik = alpha_ik t^_k - beta_ik u_ik the gradient w.r.t. the pooled vector, E = ds rnV + alpha a and
//  T^ = bf16(T rnT): sum_k a G = E-part + A^T Gn because alpha a t^ is already inside E^T T^.
#include <stdlib.h>

#include "common.cuh"
#include "epilogues.cuh"
#include "allpairs_mega.cuh"
#include "pacl_fused.cuh"
#include "simt_util.cuh"

namespace clipk {

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static int env_int(const char* name, int dflt);
constexpr int kDthSplits = 6;   // split-K slabs of the text-gradient accumulator

// rn[row] = 1 / max(||x_row||, 1e-12)      (F.normalize denominator, pacl.py:122,125)
__global__ void rownorm_bf16_kernel(const __nv_bfloat16* __restrict__ X, int64_t rows, int D, float* __restrict__ rn) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const __nv_bfloat16* x = X + row * D;
  float acc = 0.f;
  for (int d = lane * 8; d < D; d += 256) {
    float v[8];
    simt::load8<__nv_bfloat16>(x + d, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(v[j], v[j], acc);
  }
  acc = ptx::warp_sum(acc);
  if (lane == 0) rn[row] = 1.f / fmaxf(sqrtf(acc), 1e-12f);
}

// T^ = bf16(T * rnT)   (operand of dV += E^T T^)
__global__ void that_bf16_kernel(const __nv_bfloat16* __restrict__ T, const float* __restrict__ rnT, int Bt, int D,
                                 __nv_bfloat16* __restrict__ That) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= (int64_t)Bt * D) return;
  const float r = rnT[i / D];
  float v[8];
  simt::load8<__nv_bfloat16>(T + i, v);
  __align__(16) __nv_bfloat16 o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = __float2bfloat16(v[j] * r);
  *reinterpret_cast<uint4*>(That + i) = *reinterpret_cast<const uint4*>(o);
}

__global__ void allpairs_scores_kernel(const float* __restrict__ num, const float* __restrict__ usq, int64_t n, float c,
                                       float* __restrict__ scores) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) scores[i] = c * num[i] / fmaxf(sqrtf(usq[i]), 1e-12f);
}

// alpha = g / |u|, beta = g <u,t^> / |u|^3 with g = c * dscore   (Jacobian of n(u) contracted with t^)
__global__ void allpairs_alpha_beta_kernel(const float* __restrict__ num, const float* __restrict__ usq,
                                           const float* __restrict__ dscore, int64_t n, float c,
                                           float* __restrict__ alpha, float* __restrict__ beta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float un = sqrtf(usq[i]);
  const float g = c * dscore[i];
  if (un < 1e-12f) {   // clamped normalisation: u / eps is linear, no projection term
    alpha[i] = g * 1e12f;
    beta[i] = 0.f;
  } else {
    const float r = 1.f / un;
    alpha[i] = g * r;
    beta[i] = g * num[i] * r * r * r;
  }
}

// dT_k = rnT_k (dth_k - t^_k <t^_k, dth_k>),  t^ = T rnT      (Jacobian of F.normalize on the text side)
struct DthSlabs {
  float* p[4];
};
__global__ void dtext_finalize_kernel(const __nv_bfloat16* __restrict__ T, const float* __restrict__ rnT,
                                      DthSlabs slabs, int lanes, int nslab, int Bt, int D, float* __restrict__ dT) {
  // one block per text row: the slab sums of a row are (lanes * nslab) independent coalesced loads per thread
  __shared__ float red[32];
  const int row = blockIdx.x;
  const float rt = rnT[row];
  float dot = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float g = 0.f;
    for (int l = 0; l < lanes; ++l)
      for (int s2 = 0; s2 < nslab; ++s2) g += slabs.p[l][((int64_t)s2 * Bt + row) * D + d];
    slabs.p[0][(int64_t)row * D + d] = g;   // lane 0 / slab 0 now holds the total (each thread touches only its own d)
    dot = fmaf(__bfloat162float(T[(int64_t)row * D + d]) * rt, g, dot);
  }
  dot = simt::block_sum(dot, red);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float th = __bfloat162float(T[(int64_t)row * D + d]) * rt;
    dT[(int64_t)row * D + d] = rt * (slabs.p[0][(int64_t)row * D + d] - th * dot);
  }
}

// Internal stream pool: image groups are issued round-robin on `lanes` side streams (forked from / joined to the
// caller's stream with events, no host synchronisation) so that the small per-group launches of different groups
// overlap and fill each other's tails while each group's scratch stays L2-resident.
constexpr int kMaxLanes = 4;
struct LanePool {
  int device = -1;
  cudaStream_t st[kMaxLanes] = {};
  cudaEvent_t fork = nullptr, join[kMaxLanes] = {};
};
static int get_lanes(LanePool** out) {
  constexpr int kMaxDevices = 64;
  static thread_local LanePool pools[kMaxDevices];
  int dev = 0;
  CLIPK_CHECK_CUDA(cudaGetDevice(&dev));
  CLIPK_REQUIRE(dev >= 0 && dev < kMaxDevices, "pacl_allpairs: device ordinal %d out of range", dev);
  LanePool& p = pools[dev];
  if (p.device != dev) {
    for (int i = 0; i < kMaxLanes; ++i) {
      CLIPK_CHECK_CUDA(cudaStreamCreateWithFlags(&p.st[i], cudaStreamNonBlocking));
      CLIPK_CHECK_CUDA(cudaEventCreateWithFlags(&p.join[i], cudaEventDisableTiming));
    }
    CLIPK_CHECK_CUDA(cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming));
    p.device = dev;
  }
  *out = &p;
  return 0;
}

struct ApWorkspace {      // per-lane scratch
  __nv_bfloat16 *A, *G, *E;
  float* dth;             // [kDthSplits][Bt][D] fp32 split-K slabs of this lane
};

struct ApShared {
  float *dsdot, *alpha, *beta, *dth;
  __nv_bfloat16* That;
};
// layout: [lanes x per-lane scratch][shared per-call arrays]
// `pooled`: the forward saves the pooled vectors u (bf16 [Bi,Bt,D], caller-owned) and the backward consumes them --
// no G scratch, and the forward needs T^ as well (both directions use T^ = bf16(T rnT) as the score operand).
static size_t ap_carve(ApWorkspace* w, ApShared* sh, void* base, int Bi, int Bt, int P, int D, int group, int lanes,
                       int backward, int pooled = 0, int fused_fwd = 0) {
  const int Ppad = round_up(P, 64);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + off : nullptr;
    off += (bytes + 1023) / 1024 * 1024;
    return p;
  };
  const size_t act = (size_t)group * Bt * Ppad * 2;
  for (int l = 0; l < lanes; ++l) {
    w[l].A = (!backward && fused_fwd) ? nullptr : static_cast<__nv_bfloat16*>(take(act));   // fused forward: activations stay on chip
    if (backward) {
      w[l].E = static_cast<__nv_bfloat16*>(take(act));
      w[l].G = pooled ? nullptr : static_cast<__nv_bfloat16*>(take((size_t)group * Bt * D * 2));
      w[l].dth = static_cast<float*>(take((size_t)kDthSplits * Bt * D * 4));
    }
  }
  if (backward) {
    sh->dsdot = static_cast<float*>(take((size_t)Bi * P * 4));
    sh->alpha = static_cast<float*>(take((size_t)Bi * Bt * 4));
    sh->beta = static_cast<float*>(take((size_t)Bi * Bt * 4));
    sh->dth = w[0].dth;
  }
  if (backward || pooled || fused_fwd) sh->That = static_cast<__nv_bfloat16*>(take((size_t)Bt * D * 2));
  return off;
}

// Engine choice per kernel: bit i of CLIPK_AP_ENGINE2 selects the CTA-pair engine for kernel i
// (0: K1 activations, 1: K2 pooling, 2: K4 dS/E, 3: K5 dT, 4: K6 dV).  Default: all on the CTA-pair engine.
static int engine2_mask() {
  static const int m = [] {
    const char* e = getenv("CLIPK_AP_ENGINE2");
    return e != nullptr ? atoi(e) : 0x1F;
  }();
  return m;
}
enum { kK1 = 1, kK2 = 2, kK4 = 4, kK5 = 8, kK6 = 16 };

template <int BN, bool A_MN, bool B_MN, class Epi>
static int gemm_on(bool pair, const OperandDesc* a, const OperandDesc* b, int npairs, const int* ks, const int* ksub,
                   int M, int N, int batches, const typename Epi::Params& ep, cudaStream_t st) {
  if constexpr (!B_MN || BN % 128 == 0) {
    if (pair) return launch_gemm2<BN, A_MN, B_MN, Epi>(a, b, npairs, ks, ksub, M, N, batches, ep, st);
  }
  return launch_gemm<BN, A_MN, B_MN, Epi>(a, b, npairs, ks, ksub, M, N, batches, ep, st);
}

static int pick_bn(int n) {
  // largest tile whose padding waste stays under ~12%, else the waste-free 64-multiple
  const int cands[3] = {256, 192, 128};
  for (int c : cands) {
    const int cover = round_up(n, c);
    if ((cover - n) * 8 <= n) return c;
  }
  return 64;
}

// K1: A = act(T V^T) for `gi` images starting at V0.
static int launch_k1(const __nv_bfloat16* T, const __nv_bfloat16* V0, int gi, int Bt, int P, int D, int act,
                     const float* rnV0, const float* rnT, const ApWorkspace& w, float* num0, cudaStream_t st) {
  const int Ppad = round_up(P, 64);
  OperandDesc a, b;
  a.ptr = T; a.rows = Bt; a.k = D; a.ld = D; a.batch = 1; a.bmul = 0;
  b.ptr = V0; b.rows = P; b.k = D; b.ld = D; b.batch = gi; b.batch_stride = (int64_t)P * D; b.bmul = 1;
  const int ks[1] = {(D + 63) / 64};
  epi::PaclAct::Params ep{{w.A, Ppad, (int64_t)Bt * Ppad, Bt, P, gi}, rnV0, rnT, num0, Bt, P, Ppad, act};   // extent P: pads clipped
  const bool pair = (engine2_mask() & kK1) != 0;
  switch (pick_bn(Ppad)) {
    case 256: return gemm_on<256, false, false, epi::PaclAct>(pair, &a, &b, 1, ks, ks, Bt, Ppad, gi, ep, st);
    case 192: return gemm_on<192, false, false, epi::PaclAct>(pair, &a, &b, 1, ks, ks, Bt, Ppad, gi, ep, st);
    case 128: return gemm_on<128, false, false, epi::PaclAct>(pair, &a, &b, 1, ks, ks, Bt, Ppad, gi, ep, st);
    default: return gemm_on<64, false, false, epi::PaclAct>(pair, &a, &b, 1, ks, ks, Bt, Ppad, gi, ep, st);
  }
}

// K2 operands: A = activations [gi][Bt][Ppad] (K-major), B = V [gi][P][D] viewed MN-major (N = d, K = p)
static void k2_operands(const ApWorkspace& w, const __nv_bfloat16* V0, int gi, int Bt, int P, int D, OperandDesc* a,
                        OperandDesc* b) {
  const int Ppad = round_up(P, 64);
  a->ptr = w.A; a->rows = Bt; a->k = P; a->ld = Ppad;   /* extent P: TMA zero-fills the pad columns */ a->batch = gi; a->batch_stride = (int64_t)Bt * Ppad; a->bmul = 1;
  a->reverse = 1;   // K1 walked the images upwards; start with the activations it wrote last
  b->ptr = V0; b->mn_major = true; b->rows = D; b->k = P; b->ld = D; b->batch = gi; b->batch_stride = (int64_t)P * D; b->bmul = 1;
}

template <class Epi, bool A_MN>
static int launch_nd(bool pair, const OperandDesc* a, const OperandDesc* b, int npairs, const int* ks, const int* ksub,
                     int M, int D, int batches, const typename Epi::Params& ep, cudaStream_t st) {
  // N = D output columns, B operand MN-major (d contiguous)
  switch (pick_bn(D)) {
    case 256: return gemm_on<256, A_MN, true, Epi>(pair, a, b, npairs, ks, ksub, M, D, batches, ep, st);
    case 192: return gemm_on<192, A_MN, true, Epi>(pair, a, b, npairs, ks, ksub, M, D, batches, ep, st);
    case 128: return gemm_on<128, A_MN, true, Epi>(pair, a, b, npairs, ks, ksub, M, D, batches, ep, st);
    default: return gemm_on<64, A_MN, true, Epi>(pair, a, b, npairs, ks, ksub, M, D, batches, ep, st);
  }
}

static int validate(int Bi, int Bt, int P, int D, int act, int group) {
  CLIPK_REQUIRE(Bi > 0 && Bt > 0 && P > 0 && D > 0, "pacl_allpairs: empty problem (Bi=%d Bt=%d P=%d D=%d)", Bi, Bt, P, D);
  CLIPK_REQUIRE(D % 8 == 0, "pacl_allpairs: D=%d must be a multiple of 8 (16-byte rows for TMA)", D);
  CLIPK_REQUIRE(act == CLIPK_ACT_SIGMOID10 || act == CLIPK_ACT_ONES || act == CLIPK_ACT_SOFTMAX10,
                "pacl_allpairs: bad activation %d", act);
  return 0;
}

static int fork_lanes(LanePool* lp, int lanes, cudaStream_t st) {
  CLIPK_CHECK_CUDA(cudaEventRecord(lp->fork, st));
  for (int l = 0; l < lanes; ++l) CLIPK_CHECK_CUDA(cudaStreamWaitEvent(lp->st[l], lp->fork, 0));
  return 0;
}
static int join_lanes(LanePool* lp, int lanes, cudaStream_t st) {
  for (int l = 0; l < lanes; ++l) {
    CLIPK_CHECK_CUDA(cudaEventRecord(lp->join[l], lp->st[l]));
    CLIPK_CHECK_CUDA(cudaStreamWaitEvent(st, lp->join[l], 0));
  }
  return 0;
}
// Joins the forked lanes back into the caller's stream exactly once: explicitly on the success path (join()), and
// from the destructor when a launch in between failed -- the caller's stream stays ordered after the lane streams
// and a stream capture in progress is not left forked.
struct LaneJoin {
  LanePool* lp;
  int lanes;
  cudaStream_t st;
  bool done = false;
  int join() {
    done = true;
    return (lp != nullptr && lanes > 1) ? join_lanes(lp, lanes, st) : 0;
  }
  ~LaneJoin() {
    if (!done && lp != nullptr && lanes > 1) (void)join_lanes(lp, lanes, st);
  }
};

// ------------------------------------------------------------------------------------------------ mega path
// One persistent kernel per direction (allpairs_mega.cuh).  `group` < 0: -group images per group; 0: automatic.
struct MegaPlan {
  int gs, depth, slots, ngroups;
};
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e != nullptr ? atoi(e) : dflt;
}
static MegaPlan mega_plan(int Bi, int Bt, int P, int D, int group, int lanes, int backward) {
  MegaPlan pl;
  const int Ppad = round_up(P, 64);
  int gs = group < 0 ? -group : env_int("CLIPK_MEGA_GS", 0);
  pl.depth = lanes > 0 && group < 0 ? lanes : env_int("CLIPK_MEGA_DEPTH", 2);
  if (gs <= 0) {
    // keep the live scratch (slots groups of A [+ E + G]) around 72 MB so that it stays in the 126 MB L2 next to V
    const double per_img = (double)Bt * (backward ? (2.0 * Ppad + D) : (double)Ppad) * 2.0;
    gs = (int)(72e6 / ((pl.depth + 1) * per_img));
    if (gs < 1) gs = 1;
    if (gs > 16) gs = 16;
  }
  if (gs > Bi) gs = Bi;
  pl.gs = gs;
  pl.ngroups = (Bi + gs - 1) / gs;
  if (pl.depth > pl.ngroups) pl.depth = pl.ngroups;
  if (pl.depth < 1) pl.depth = 1;
  pl.slots = env_int("CLIPK_MEGA_SLOTS", pl.depth + 1);
  if (pl.slots < pl.depth) pl.slots = pl.depth;
  return pl;
}

struct MegaWs {
  __nv_bfloat16 *A, *E, *G, *That;
  float *dsdot, *alpha, *beta, *dth;
  unsigned* done;
};
static size_t mega_carve(MegaWs* w, void* base, const MegaPlan& pl, int Bi, int Bt, int P, int D, int backward) {
  const int Ppad = round_up(P, 64);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + off : nullptr;
    off += (bytes + 1023) / 1024 * 1024;
    return p;
  };
  const size_t S = (size_t)pl.slots * pl.gs;
  w->A = static_cast<__nv_bfloat16*>(take(S * Bt * Ppad * 2));
  w->done = static_cast<unsigned*>(take((size_t)pl.ngroups * mega::kMaxPhases * 4));
  if (backward) {
    w->E = static_cast<__nv_bfloat16*>(take(S * Bt * Ppad * 2));
    w->G = static_cast<__nv_bfloat16*>(take(S * Bt * D * 2));
    w->That = static_cast<__nv_bfloat16*>(take((size_t)Bt * D * 2));
    w->dsdot = static_cast<float*>(take((size_t)Bi * P * 4));
    w->alpha = static_cast<float*>(take((size_t)Bi * Bt * 4));
    w->beta = static_cast<float*>(take((size_t)Bi * Bt * 4));
    w->dth = static_cast<float*>(take((size_t)Bt * D * 4));
  }
  return off;
}

static int mega_launch(const __nv_bfloat16* V, const __nv_bfloat16* T, int Bi, int Bt, int P, int D, int act, float c,
                       const float* rnV, const float* rnT, float* num, float* usq, __nv_bfloat16* dV,
                       const MegaWs& w, const MegaPlan& pl, int backward, cudaStream_t st) {
  const int Ppad = round_up(P, 64);
  const uint64_t S = (uint64_t)pl.slots * pl.gs;
  mega::Maps mp;
  memset(&mp, 0, sizeof(mp));
  const uint64_t ldP = (uint64_t)Ppad * 2, ldD = (uint64_t)D * 2;
  CLIPK_TRY(make_tmap_bf16(&mp.T_k, T, D, Bt, 1, ldD, 0, 128));
  CLIPK_TRY(make_tmap_bf16(&mp.V_k, V, D, P, Bi, ldD, (uint64_t)P * ldD, mega::kBNP / 2));
  CLIPK_TRY(make_tmap_bf16(&mp.V_mn, V, D, P, Bi, ldD, (uint64_t)P * ldD, 64));
  CLIPK_TRY(make_tmap_bf16(&mp.A_k, w.A, P, Bt, S, ldP, (uint64_t)Bt * ldP, 128));     // extent P: pads read as zero
  CLIPK_TRY(make_tmap_bf16_box(&mp.oA, w.A, P, Bt, S, ldP, (uint64_t)Bt * ldP, 32, 32));
  if (backward) {
    CLIPK_TRY(make_tmap_bf16(&mp.A_mn, w.A, P, Bt, S, ldP, (uint64_t)Bt * ldP, 64));
    CLIPK_TRY(make_tmap_bf16(&mp.E_k, w.E, P, Bt, S, ldP, (uint64_t)Bt * ldP, 128));
    CLIPK_TRY(make_tmap_bf16(&mp.E_mn, w.E, P, Bt, S, ldP, (uint64_t)Bt * ldP, 64));
    CLIPK_TRY(make_tmap_bf16(&mp.G_k, w.G, D, Bt, S, ldD, (uint64_t)Bt * ldD, 128));
    CLIPK_TRY(make_tmap_bf16(&mp.G_mn, w.G, D, Bt, S, ldD, (uint64_t)Bt * ldD, 64));
    CLIPK_TRY(make_tmap_bf16(&mp.Th_mn, w.That, D, Bt, 1, ldD, 0, 64));
    CLIPK_TRY(make_tmap_bf16_box(&mp.oE, w.E, P, Bt, S, ldP, (uint64_t)Bt * ldP, 32, 32));
    CLIPK_TRY(make_tmap_bf16_box(&mp.oG, w.G, D, Bt, S, ldD, (uint64_t)Bt * ldD, 32, 32));
    CLIPK_TRY(make_tmap_bf16_box(&mp.odV, dV, D, P, Bi, ldD, (uint64_t)P * ldD, 32, 32));
  }
  mega::Sched sc;
  memset(&sc, 0, sizeof(sc));
  sc.Bi = Bi; sc.Bt = Bt; sc.P = P; sc.Ppad = Ppad; sc.D = D; sc.act = act;
  sc.gs = pl.gs; sc.ngroups = pl.ngroups; sc.depth = pl.depth; sc.slots = pl.slots;
  const int tmT = (Bt + 255) / 256, tmP = (P + 255) / 256;
  const int tnP = (Ppad + mega::kBNP - 1) / mega::kBNP, tnD = (D + mega::kBND - 1) / mega::kBND;
  for (int i = 0; i < mega::kMaxPhases; ++i) { sc.dep_same[i] = -1; sc.dep_ring[i][0] = sc.dep_ring[i][1] = -1; }
  if (!backward) {
    sc.nph = 2;
    sc.ph[0] = mega::PH_ACT; sc.tm[0] = tmT; sc.tn[0] = tnP; sc.dep_ring[0][0] = 1;
    sc.ph[1] = mega::PH_USQ; sc.tm[1] = tmT; sc.tn[1] = tnD; sc.dep_same[1] = 0;
  } else {
    sc.nph = 5;
    sc.ph[0] = mega::PH_ACTS; sc.tm[0] = tmT; sc.tn[0] = tnP; sc.dep_ring[0][0] = 3; sc.dep_ring[0][1] = 4;
    sc.ph[1] = mega::PH_GNEG; sc.tm[1] = tmT; sc.tn[1] = tnD; sc.dep_same[1] = 0;
    sc.ph[2] = mega::PH_DS; sc.tm[2] = tmT; sc.tn[2] = tnP; sc.dep_same[2] = 1;
    // dV before dT: the dT tiles are long (K folds the images of the group) and nothing in the block depends on them,
    // so they go last, where the imbalance they cause overlaps the next block's first phase
    sc.ph[3] = mega::PH_DV; sc.tm[3] = tmP; sc.tn[3] = tnD; sc.dep_same[3] = 2;
    sc.ph[4] = mega::PH_DT; sc.tm[4] = tmT; sc.tn[4] = tnD; sc.dep_same[4] = 2;
  }
  sc.dt_spb = env_int("CLIPK_MEGA_DTSPB", pl.gs);      // images folded into one DT tile's K range
  if (sc.dt_spb < 1) sc.dt_spb = 1;
  long long total = 0;
  for (int g = 0; g < pl.ngroups; ++g)
    for (int i = 0; i < sc.nph; ++i) total += mega::job_tiles(sc, i, g);
  CLIPK_REQUIRE(total < 0x3fffffff, "pacl_allpairs: tile sequence too long (%lld)", total);
  sc.total_tiles = (int)total;
  for (int i = 0; i < sc.nph; ++i) {
    sc.need_full[i] = 2u * (unsigned)mega::job_tiles(sc, i, 0);
    sc.need_last[i] = 2u * (unsigned)mega::job_tiles(sc, i, pl.ngroups - 1);
  }
  sc.done = w.done;
  sc.rnV = rnV; sc.rnT = rnT; sc.num = num; sc.usq = usq;
  sc.alpha = w.alpha; sc.beta = w.beta; sc.dsdot = w.dsdot; sc.dth = w.dth;
  sc.V = V; sc.c = c;
  sc.flags = env_int("CLIPK_MEGA_FLAGS", 0);
  CLIPK_CHECK_CUDA(cudaMemsetAsync(w.done, 0, (size_t)pl.ngroups * mega::kMaxPhases * 4, st));
  static std::atomic<uint64_t> attr_mask{0};
  int dev = 0;
  CLIPK_CHECK_CUDA(cudaGetDevice(&dev));
  if (!(attr_mask.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
    CLIPK_CHECK_CUDA(cudaFuncSetAttribute(mega::allpairs_mega_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          mega::kSmemTotal));
    CLIPK_CHECK_CUDA(cudaFuncSetAttribute(mega::allpairs_mega_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          mega::kSmemTotal));
    attr_mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  const int pairs = sm_count() / 2;
  const int grid = 2 * (sc.total_tiles < pairs ? sc.total_tiles : pairs);
  const bool tr = trace_enabled();
  if (tr) trace_begin(backward ? "allpairs_mega_kernel (backward)" : "allpairs_mega_kernel (forward)", st);
  if (backward) mega::allpairs_mega_kernel<true><<<grid, mega::kThreads, mega::kSmemTotal, st>>>(mp, sc);
  else mega::allpairs_mega_kernel<false><<<grid, mega::kThreads, mega::kSmemTotal, st>>>(mp, sc);
  if (tr) trace_end(st);
  count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// dT_k = rnT_k (dth_k - t^_k <t^_k, dth_k>) from a single accumulated dt^ buffer
__global__ void dtext_finalize1_kernel(const __nv_bfloat16* __restrict__ T, const float* __restrict__ rnT,
                                       const float* __restrict__ dth, int Bt, int D, float* __restrict__ dT) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= Bt) return;
  const int lane = threadIdx.x & 31;
  const float rt = rnT[row];
  float dot = 0.f;
  for (int d = lane; d < D; d += 32)
    dot = fmaf(__bfloat162float(T[(int64_t)row * D + d]) * rt, dth[(int64_t)row * D + d], dot);
  dot = ptx::warp_sum(dot);
  for (int d = lane; d < D; d += 32) {
    const float th = __bfloat162float(T[(int64_t)row * D + d]) * rt;
    dT[(int64_t)row * D + d] = rt * (dth[(int64_t)row * D + d] - th * dot);
  }
}

static int mega_fwd(const __nv_bfloat16* V, const __nv_bfloat16* T, int Bi, int Bt, int P, int D, int act, float c,
                    float* rnV, float* rnT, float* num, float* usq, float* scores, void* ws, size_t ws_bytes, int group,
                    int lanes, cudaStream_t st) {
  const MegaPlan pl = mega_plan(Bi, Bt, P, D, group, lanes, 0);
  MegaWs w{};
  const size_t need = mega_carve(&w, ws, pl, Bi, Bt, P, D, 0);
  CLIPK_REQUIRE(ws != nullptr && ws_bytes >= need, "pacl_allpairs_fwd: workspace too small (%zu < %zu)", ws_bytes, need);
  rownorm_bf16_kernel<<<(unsigned)(((int64_t)Bi * P + 7) / 8), 256, 0, st>>>(V, (int64_t)Bi * P, D, rnV);
  rownorm_bf16_kernel<<<(Bt + 7) / 8, 256, 0, st>>>(T, Bt, D, rnT);
  count_launches(2);
  CLIPK_CHECK_CUDA(cudaMemsetAsync(num, 0, (size_t)Bi * Bt * 4, st));
  CLIPK_CHECK_CUDA(cudaMemsetAsync(usq, 0, (size_t)Bi * Bt * 4, st));
  CLIPK_TRY(mega_launch(V, T, Bi, Bt, P, D, act, c, rnV, rnT, num, usq, nullptr, w, pl, 0, st));
  const int64_t n = (int64_t)Bi * Bt;
  allpairs_scores_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(num, usq, n, c, scores);
  count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int mega_bwd(const __nv_bfloat16* V, const __nv_bfloat16* T, int Bi, int Bt, int P, int D, int act, float c,
                    const float* rnV, const float* rnT, const float* num, const float* usq, const float* dscores,
                    __nv_bfloat16* dV, float* dT, void* ws, size_t ws_bytes, int group, int lanes, cudaStream_t st) {
  const MegaPlan pl = mega_plan(Bi, Bt, P, D, group, lanes, 1);
  MegaWs w{};
  const size_t need = mega_carve(&w, ws, pl, Bi, Bt, P, D, 1);
  CLIPK_REQUIRE(ws != nullptr && ws_bytes >= need, "pacl_allpairs_bwd: workspace too small (%zu < %zu)", ws_bytes, need);
  const int64_t n = (int64_t)Bi * Bt;
  CLIPK_CHECK_CUDA(cudaMemsetAsync(w.dsdot, 0, (size_t)Bi * P * 4, st));
  CLIPK_CHECK_CUDA(cudaMemsetAsync(w.dth, 0, (size_t)Bt * D * 4, st));
  allpairs_alpha_beta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(num, usq, dscores, n, c, w.alpha, w.beta);
  that_bf16_kernel<<<(unsigned)(((int64_t)Bt * D / 8 + 255) / 256), 256, 0, st>>>(T, rnT, Bt, D, w.That);
  count_launches(2);
  CLIPK_TRY(mega_launch(V, T, Bi, Bt, P, D, act, c, rnV, rnT, const_cast<float*>(num), const_cast<float*>(usq), dV, w,
                        pl, 1, st));
  dtext_finalize1_kernel<<<(Bt + 7) / 8, 256, 0, st>>>(T, rnT, w.dth, Bt, D, dT);
  count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ fused forward
// One persistent launch for the whole forward (pacl_fused.cuh): activations never leave the SM.  Shapes it covers:
// P <= 576 (the activations of 128 texts x P patches are resident in shared memory) and D a multiple of 128; anything
// else takes the staged path below.  CLIPK_AP_FUSED=0 forces the staged path (A/B diagnostics).
static bool fused_fwd_eligible(int P, int D) {
  return P <= 64 * fz::kMaxKB && D >= 128 && D % 128 == 0 && env_int("CLIPK_AP_FUSED", 1) != 0;
}

template <bool SAVE_U>
static int fused_fwd_launch(const __nv_bfloat16* V, const __nv_bfloat16* That, int Bi, int Bt, int P, int D, int act,
                            const float* rnV, float* num, float* usq, __nv_bfloat16* pooled, cudaStream_t st) {
  fz::Params pr;
  memset(&pr, 0, sizeof(pr));
  pr.Bi = Bi; pr.Bt = Bt; pr.P = P; pr.D = D; pr.act = act;
  if (P <= 256) {
    pr.nS = 1;
    pr.pc = round_up(P, 16);
  } else {
    pr.nS = (P + 255) / 256;
    pr.pc = round_up((P + pr.nS - 1) / pr.nS, 64);
    pr.nS = (P + pr.pc - 1) / pr.pc;
  }
  pr.nU = (D + 255) / 256;
  pr.kbA = (P + 63) / 64;
  pr.ksD = (D + 63) / 64;
  pr.tilesM = (Bt + 255) / 256;
  pr.rnV = rnV; pr.num = num; pr.usq = usq;
  pr.pooled = pooled;
  pr.V = V;
  pr.prefetch = env_int("CLIPK_FZ_PREFETCH", 1);
  pr.xslots = env_int("CLIPK_FZ_XSLOTS", 1);
  CLIPK_REQUIRE(pr.nS <= fz::kMaxChunks && pr.pc <= 256 && pr.kbA <= fz::kMaxKB, "pacl fused forward: P=%d out of range", P);
  fz::Maps mp;
  memset(&mp, 0, sizeof(mp));
  const uint64_t ldD = (uint64_t)D * 2;
  CLIPK_TRY(make_tmap_bf16(&mp.T, That, D, Bt, 1, ldD, 0, 128));
  CLIPK_TRY(make_tmap_bf16(&mp.Vk, V, D, P, Bi, ldD, (uint64_t)P * ldD, (uint32_t)(pr.pc / 2)));
  CLIPK_TRY(make_tmap_bf16(&mp.Vmn, V, D, P, Bi, ldD, (uint64_t)P * ldD, 64));
  auto kern = fz::pacl_fused_fwd_kernel<SAVE_U>;
  constexpr int kSmem = fz::kSmemTotal;
  static std::atomic<uint64_t> attr_mask{0};
  int dev = 0;
  CLIPK_CHECK_CUDA(cudaGetDevice(&dev));
  if (!(attr_mask.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
    CLIPK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  const int nitems = Bi * pr.tilesM;
  const int pairs = sm_count() / 2;
  const int grid = 2 * (nitems < pairs ? nitems : pairs);
  const bool tr = trace_enabled();
  if (tr) trace_begin(SAVE_U ? "pacl_fused_fwd_kernel<save pooled>" : "pacl_fused_fwd_kernel", st);
  kern<<<grid, fz::kThreads, kSmem, st>>>(mp, pr);
  if (tr) trace_end(st);
  count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int allpairs_fwd(const __nv_bfloat16* V, const __nv_bfloat16* T, int Bi, int Bt, int P, int D, int act, float c,
                 float* rnV, float* rnT, float* num, float* usq, float* scores, __nv_bfloat16* pooled, void* ws,
                 size_t ws_bytes, int group, int lanes, int rnv_given, cudaStream_t st) {
  CLIPK_TRY(validate(Bi, Bt, P, D, act, group));
  if (group <= 0) {
    CLIPK_REQUIRE(pooled == nullptr && !rnv_given, "pacl_allpairs_fwd: the persistent kernel (group <= 0) does not save pooled vectors");
    return mega_fwd(V, T, Bi, Bt, P, D, act, c, rnV, rnT, num, usq, scores, ws, ws_bytes, group, lanes, st);
  }
  CLIPK_REQUIRE(lanes >= 1 && lanes <= kMaxLanes, "pacl_allpairs: lanes must be in [1, %d]", kMaxLanes);
  ApWorkspace w[kMaxLanes]{};
  ApShared sh{};
  const int save = pooled != nullptr;
  const bool fused = fused_fwd_eligible(P, D);
  const size_t need = ap_carve(w, &sh, ws, Bi, Bt, P, D, group, lanes, 0, save, fused);
  CLIPK_REQUIRE(ws != nullptr && ws_bytes >= need, "pacl_allpairs_fwd: workspace too small (%zu < %zu)", ws_bytes, need);
  const int Ppad = round_up(P, 64);
  // rnV given: the producer of V (the projection head's output GEMM) already emitted the row norms -- one pass over the
  // 906 MB patch tensor less
  if (!rnv_given) {
    rownorm_bf16_kernel<<<(unsigned)(((int64_t)Bi * P + 7) / 8), 256, 0, st>>>(V, (int64_t)Bi * P, D, rnV);
    count_launches(1);
  }
  rownorm_bf16_kernel<<<(Bt + 7) / 8, 256, 0, st>>>(T, Bt, D, rnT);
  count_launches(1);
  if (save || fused) {
    that_bf16_kernel<<<(unsigned)(((int64_t)Bt * D / 8 + 255) / 256), 256, 0, st>>>(T, rnT, Bt, D, sh.That);
    count_launches(1);
  }
  CLIPK_CHECK_CUDA(cudaMemsetAsync(num, 0, (size_t)Bi * Bt * 4, st));
  CLIPK_CHECK_CUDA(cudaMemsetAsync(usq, 0, (size_t)Bi * Bt * 4, st));
  if (fused) {
    if (save) CLIPK_TRY(fused_fwd_launch<true>(V, sh.That, Bi, Bt, P, D, act, rnV, num, usq, pooled, st));
    else CLIPK_TRY(fused_fwd_launch<false>(V, sh.That, Bi, Bt, P, D, act, rnV, num, usq, nullptr, st));
    const int64_t n = (int64_t)Bi * Bt;
    allpairs_scores_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(num, usq, n, c, scores);
    count_launches(1);
    CLIPK_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  LanePool* lp = nullptr;
  const PdlBlock no_pdl(lanes > 1);
  if (lanes > 1) {
    CLIPK_TRY(get_lanes(&lp));
    CLIPK_TRY(fork_lanes(lp, lanes, st));
  }
  LaneJoin joiner{lp, lanes, st};      // joins the lanes on every exit path (error returns included)
  int gidx = 0;
  for (int i0 = 0; i0 < Bi; i0 += group, ++gidx) {
    const int gi = (Bi - i0) < group ? (Bi - i0) : group;
    const int l = gidx % lanes;
    cudaStream_t ls = lanes > 1 ? lp->st[l] : st;
    const __nv_bfloat16* V0 = V + (int64_t)i0 * P * D;
    // pooled mode: the score operand is T^ (rnT folded in, as in the backward), so the epilogue's text norm is 1
    CLIPK_TRY(launch_k1(save ? sh.That : T, V0, gi, Bt, P, D, act, rnV + (int64_t)i0 * P, save ? nullptr : rnT, w[l],
                        num + (int64_t)i0 * Bt, ls));
    OperandDesc a, b;
    k2_operands(w[l], V0, gi, Bt, P, D, &a, &b);
    const int ks[1] = {Ppad / 64};
    if (save) {
      epi::UsqStore::Params ep{{pooled + (int64_t)i0 * Bt * D, D, (int64_t)Bt * D, Bt, D, gi}, usq + (int64_t)i0 * Bt, Bt, D};
      CLIPK_TRY(launch_nd<epi::UsqStore, false>(true, &a, &b, 1, ks, ks, Bt, D, gi, ep, ls));
    } else {
      epi::Usq::Params ep{usq + (int64_t)i0 * Bt, Bt, D};
      CLIPK_TRY(launch_nd<epi::Usq, false>((engine2_mask() & kK2) != 0, &a, &b, 1, ks, ks, Bt, D, gi, ep, ls));
    }
  }
  CLIPK_TRY(joiner.join());
  const int64_t n = (int64_t)Bi * Bt;
  allpairs_scores_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(num, usq, n, c, scores);
  count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Backward from the saved pooled vectors (7 GEMM units per step instead of 8: no recompute of the pooling GEMM, no
// separate recompute of the activations):
//   B1  [x | d] = [T^ ; U_i] V_i^T   (dual accumulators, epilogue DsDual)  ->  E, A2 = -beta a, dsdot
//   B2  dt^ += E V                   (K folds (image, patch), split-K slabs)
//   B3  dV_i^T = U_i^T A2 + T^^T E - rnV^2 dsdot V^T   (two operand pairs, epilogue DvOutT)
static int allpairs_bwd_pooled(const __nv_bfloat16* V, const __nv_bfloat16* T, int Bi, int Bt, int P, int D, int act,
                               float c, const float* rnV, const float* rnT, const float* num, const float* usq,
                               const float* dscores, const __nv_bfloat16* pooled, __nv_bfloat16* dV, float* dT, void* ws,
                               size_t ws_bytes, int group, int lanes, cudaStream_t st) {
  CLIPK_REQUIRE(group > 0, "pacl_allpairs_bwd: saved pooled vectors need the staged path (group > 0)");
  CLIPK_REQUIRE(lanes >= 1 && lanes <= kMaxLanes, "pacl_allpairs: lanes must be in [1, %d]", kMaxLanes);
  ApWorkspace wl[kMaxLanes]{};
  ApShared sh{};
  const size_t need = ap_carve(wl, &sh, ws, Bi, Bt, P, D, group, lanes, 1, 1);
  CLIPK_REQUIRE(ws != nullptr && ws_bytes >= need, "pacl_allpairs_bwd: workspace too small (%zu < %zu)", ws_bytes, need);
  const int Ppad = round_up(P, 64);
  const int64_t n = (int64_t)Bi * Bt;
  CLIPK_CHECK_CUDA(cudaMemsetAsync(sh.dsdot, 0, (size_t)Bi * P * 4, st));
  for (int l = 0; l < lanes; ++l)
    CLIPK_CHECK_CUDA(cudaMemsetAsync(wl[l].dth, 0, (size_t)kDthSplits * Bt * D * 4, st));
  allpairs_alpha_beta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(num, usq, dscores, n, c, sh.alpha, sh.beta);
  that_bf16_kernel<<<(unsigned)(((int64_t)Bt * D / 8 + 255) / 256), 256, 0, st>>>(T, rnT, Bt, D, sh.That);
  count_launches(2);
  LanePool* lp = nullptr;
  const PdlBlock no_pdl(lanes > 1);
  if (lanes > 1) {
    CLIPK_TRY(get_lanes(&lp));
    CLIPK_TRY(fork_lanes(lp, lanes, st));
  }
  LaneJoin joiner{lp, lanes, st};
  int gidx = 0;
  for (int i0 = 0; i0 < Bi; i0 += group, ++gidx) {
    const int gi = (Bi - i0) < group ? (Bi - i0) : group;
    const int l = gidx % lanes;
    const ApWorkspace& w = wl[l];
    cudaStream_t ls = lanes > 1 ? lp->st[l] : st;
    const __nv_bfloat16* V0 = V + (int64_t)i0 * P * D;
    const __nv_bfloat16* U0 = pooled + (int64_t)i0 * Bt * D;
    const float* rnV0 = rnV + (int64_t)i0 * P;
    const float* alpha0 = sh.alpha + (int64_t)i0 * Bt;
    const float* beta0 = sh.beta + (int64_t)i0 * Bt;
    float* dsdot0 = sh.dsdot + (int64_t)i0 * P;
    // B1: dual GEMM over V_i: x = T^ V^T, d = U V^T  ->  E, A2, dsdot
    {
      OperandDesc a[2], b;
      a[0].ptr = sh.That; a[0].rows = Bt; a[0].k = D; a[0].ld = D; a[0].batch = 1; a[0].bmul = 0;
      a[1].ptr = U0; a[1].rows = Bt; a[1].k = D; a[1].ld = D; a[1].batch = gi; a[1].batch_stride = (int64_t)Bt * D; a[1].bmul = 1;
      b.ptr = V0; b.rows = P; b.k = D; b.ld = D; b.batch = gi; b.batch_stride = (int64_t)P * D; b.bmul = 1;
      const int ks[1] = {(D + 63) / 64};
      const eng::OutDesc oe{w.E, Ppad, (int64_t)Bt * Ppad, Bt, P, gi};      // extent P: pad columns clipped on store,
      const eng::OutDesc oa{w.A, Ppad, (int64_t)Bt * Ppad, Bt, P, gi};      // zero-filled on load
      epi::DsDual::Params ep{oe, oa, rnV0, alpha0, beta0, dsdot0, Bt, P, env_int("CLIPK_DBG_ACT", act)};
      if (Ppad >= 128) CLIPK_TRY((launch_gemm2<128, false, false, epi::DsDual>(a, &b, 1, ks, ks, Bt, Ppad, gi, ep, ls)));
      else CLIPK_TRY((launch_gemm2<64, false, false, epi::DsDual>(a, &b, 1, ks, ks, Bt, Ppad, gi, ep, ls)));
    }
    // B2: dt^[k, :] += sum_{i,p} E[i,k,p] V[i,p,:]   (split-K over images into fp32 slabs, summed in the finalize)
    {
      const int want = gi < kDthSplits ? gi : kDthSplits;
      const int spb = (gi + want - 1) / want;
      const int nsplit = (gi + spb - 1) / spb;
      OperandDesc a, b;
      a.ptr = w.E; a.rows = Bt; a.k = P; a.ld = Ppad; a.batch = gi; a.batch_stride = (int64_t)Bt * Ppad; a.bmul = 0; a.smul = 1;
      a.sub_per_batch = spb;
      a.sub_total = gi;
      a.reverse = 1;   // B1 walked upwards: take the last-written E first
      b.ptr = V0; b.mn_major = true; b.rows = D; b.k = P; b.ld = D; b.batch = gi; b.batch_stride = (int64_t)P * D; b.bmul = 0; b.smul = 1;
      const int ks[1] = {spb * (Ppad / 64)};
      const int ksub[1] = {Ppad / 64};
      epi::Store<false>::Params ep{w.dth, D, (int64_t)Bt * D, Bt, D, 1.f, 1};
      CLIPK_TRY(launch_nd<epi::Store<false>, false>(true, &a, &b, 1, ks, ksub, Bt, D, nsplit, ep, ls));
    }
    // B3: dV_i = A2_i^T U_i + E_i^T T^ - rnV^2 dsdot V, in the orientation that wastes less of the 256-row tiles
    {
      const bool transposed = env_int("CLIPK_AP_K6T", 1) != 0 && D % 8 == 0 &&
                              (int64_t)round_up(D, 256) * round_up(P, 16) < (int64_t)round_up(P, 256) * round_up(D, 16);
      const int ks[2] = {(Bt + 63) / 64, (Bt + 63) / 64};
      if (transposed) {
        OperandDesc a[2], b[2];
        a[0].ptr = U0; a[0].mn_major = true; a[0].rows = D; a[0].k = Bt; a[0].ld = D; a[0].batch = gi;
        a[0].batch_stride = (int64_t)Bt * D; a[0].bmul = 1;
        a[0].reverse = 1;
        b[0].ptr = w.A; b[0].mn_major = true; b[0].rows = P; b[0].k = Bt; b[0].ld = Ppad; b[0].batch = gi;
        b[0].batch_stride = (int64_t)Bt * Ppad; b[0].bmul = 1;
        a[1].ptr = sh.That; a[1].mn_major = true; a[1].rows = D; a[1].k = Bt; a[1].ld = D; a[1].batch = 1; a[1].bmul = 0;
        b[1] = b[0]; b[1].ptr = w.E;
        const eng::OutDesc odv{dV + (int64_t)i0 * P * D, D, (int64_t)P * D, P, D, gi};
        const eng::OutDesc ov{const_cast<__nv_bfloat16*>(V0), D, (int64_t)P * D, P, D, gi};
        epi::DvOutT::Params ep{odv, ov, rnV0, dsdot0, P};
        CLIPK_TRY((launch_gemm2<256, true, true, epi::DvOutT>(a, b, 2, ks, ks, D, P, gi, ep, ls)));
      } else {
        OperandDesc a[2], b[2];
        a[0].ptr = w.A; a[0].mn_major = true; a[0].rows = P; a[0].k = Bt; a[0].ld = Ppad; a[0].batch = gi;
        a[0].batch_stride = (int64_t)Bt * Ppad; a[0].bmul = 1;
        a[0].reverse = 1;
        b[0].ptr = U0; b[0].mn_major = true; b[0].rows = D; b[0].k = Bt; b[0].ld = D; b[0].batch = gi;
        b[0].batch_stride = (int64_t)Bt * D; b[0].bmul = 1;
        a[1] = a[0]; a[1].ptr = w.E;
        b[1].ptr = sh.That; b[1].mn_major = true; b[1].rows = D; b[1].k = Bt; b[1].ld = D; b[1].batch = 1; b[1].bmul = 0;
        epi::DvOut::Params ep{{dV + (int64_t)i0 * P * D, D, (int64_t)P * D, P, D, gi}, V0, rnV0, dsdot0, P, D};
        CLIPK_TRY(launch_nd<epi::DvOut, true>(true, a, b, 2, ks, ks, P, D, gi, ep, ls));
      }
    }
  }
  CLIPK_TRY(joiner.join());
  DthSlabs slabs{};
  for (int l = 0; l < lanes; ++l) slabs.p[l] = wl[l].dth;
  dtext_finalize_kernel<<<Bt, 256, 0, st>>>(T, rnT, slabs, lanes, kDthSplits, Bt, D, dT);
  count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int allpairs_bwd(const __nv_bfloat16* V, const __nv_bfloat16* T, int Bi, int Bt, int P, int D, int act, float c,
                 const float* rnV, const float* rnT, const float* num, const float* usq, const float* dscores,
                 const __nv_bfloat16* pooled, __nv_bfloat16* dV, float* dT, void* ws, size_t ws_bytes, int group,
                 int lanes, cudaStream_t st) {
  CLIPK_TRY(validate(Bi, Bt, P, D, act, group));
  if (pooled != nullptr)
    return allpairs_bwd_pooled(V, T, Bi, Bt, P, D, act, c, rnV, rnT, num, usq, dscores, pooled, dV, dT, ws, ws_bytes,
                               group, lanes, st);
  if (group <= 0)
    return mega_bwd(V, T, Bi, Bt, P, D, act, c, rnV, rnT, num, usq, dscores, dV, dT, ws, ws_bytes, group, lanes, st);
  CLIPK_REQUIRE(lanes >= 1 && lanes <= kMaxLanes, "pacl_allpairs: lanes must be in [1, %d]", kMaxLanes);
  ApWorkspace wl[kMaxLanes]{};
  ApShared sh{};
  const size_t need = ap_carve(wl, &sh, ws, Bi, Bt, P, D, group, lanes, 1);
  CLIPK_REQUIRE(ws != nullptr && ws_bytes >= need, "pacl_allpairs_bwd: workspace too small (%zu < %zu)", ws_bytes, need);
  const int Ppad = round_up(P, 64);
  const int64_t n = (int64_t)Bi * Bt;
  CLIPK_CHECK_CUDA(cudaMemsetAsync(sh.dsdot, 0, (size_t)Bi * P * 4, st));
  for (int l = 0; l < lanes; ++l)
    CLIPK_CHECK_CUDA(cudaMemsetAsync(wl[l].dth, 0, (size_t)kDthSplits * Bt * D * 4, st));
  allpairs_alpha_beta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(num, usq, dscores, n, c, sh.alpha, sh.beta);
  that_bf16_kernel<<<(unsigned)(((int64_t)Bt * D / 8 + 255) / 256), 256, 0, st>>>(T, rnT, Bt, D, sh.That);
  count_launches(2);
  LanePool* lp = nullptr;
  const PdlBlock no_pdl(lanes > 1);
  if (lanes > 1) {
    CLIPK_TRY(get_lanes(&lp));
    CLIPK_TRY(fork_lanes(lp, lanes, st));
  }
  LaneJoin joiner{lp, lanes, st};
  int gidx = 0;
  for (int i0 = 0; i0 < Bi; i0 += group, ++gidx) {
    const int gi = (Bi - i0) < group ? (Bi - i0) : group;
    const int l = gidx % lanes;
    const ApWorkspace& w = wl[l];
    cudaStream_t ls = lanes > 1 ? lp->st[l] : st;
    const __nv_bfloat16* V0 = V + (int64_t)i0 * P * D;
    const float* rnV0 = rnV + (int64_t)i0 * P;
    const float* alpha0 = sh.alpha + (int64_t)i0 * Bt;
    const float* beta0 = sh.beta + (int64_t)i0 * Bt;
    float* dsdot0 = sh.dsdot + (int64_t)i0 * P;
    // K1' (recompute activations A; also X = bf16(<t^_k, V_ip>) into the E buffer, consumed in place by K4)
    {
      OperandDesc a, b;
      a.ptr = T; a.rows = Bt; a.k = D; a.ld = D; a.batch = 1; a.bmul = 0;
      b.ptr = V0; b.rows = P; b.k = D; b.ld = D; b.batch = gi; b.batch_stride = (int64_t)P * D; b.bmul = 1;
      const int ks[1] = {(D + 63) / 64};
      const eng::OutDesc oa{w.A, Ppad, (int64_t)Bt * Ppad, Bt, P, gi};      // extent P: pad columns are clipped on store
      const eng::OutDesc ox{w.E, Ppad, (int64_t)Bt * Ppad, Bt, P, gi};      // and zero-filled on load
      epi::PaclActS::Params ep{oa, ox, rnV0, rnT, Bt, P, Ppad, act};
      switch (pick_bn(Ppad)) {
        case 256: CLIPK_TRY((launch_gemm2<256, false, false, epi::PaclActS>(&a, &b, 1, ks, ks, Bt, Ppad, gi, ep, ls))); break;
        case 192: CLIPK_TRY((launch_gemm2<192, false, false, epi::PaclActS>(&a, &b, 1, ks, ks, Bt, Ppad, gi, ep, ls))); break;
        case 128: CLIPK_TRY((launch_gemm2<128, false, false, epi::PaclActS>(&a, &b, 1, ks, ks, Bt, Ppad, gi, ep, ls))); break;
        default: CLIPK_TRY((launch_gemm2<64, false, false, epi::PaclActS>(&a, &b, 1, ks, ks, Bt, Ppad, gi, ep, ls))); break;
      }
    }
    // K2': Gn = -beta u
    {
      OperandDesc a, b;
      k2_operands(w, V0, gi, Bt, P, D, &a, &b);
      const int ks[1] = {Ppad / 64};
      epi::GNeg::Params ep{{w.G, D, (int64_t)Bt * D, Bt, D, gi}, beta0, Bt};
      CLIPK_TRY(launch_nd<epi::GNeg, false>((engine2_mask() & kK2) != 0, &a, &b, 1, ks, ks, Bt, D, gi, ep, ls));
    }
    // K4: d = Gn V^T; with the X chunks read back through TMA  ->  E (in place over X), dsdot
    {
      OperandDesc a, b;
      a.ptr = w.G; a.rows = Bt; a.k = D; a.ld = D; a.batch = gi; a.batch_stride = (int64_t)Bt * D; a.bmul = 1;
      b.ptr = V0; b.rows = P; b.k = D; b.ld = D; b.batch = gi; b.batch_stride = (int64_t)P * D; b.bmul = 1;
      const int ks[1] = {(D + 63) / 64};
      const eng::OutDesc oe{w.E, Ppad, (int64_t)Bt * Ppad, Bt, P, gi};
      epi::DsIn::Params ep{oe, oe, rnV0, alpha0, dsdot0, Bt, P, Ppad, act};
      switch (pick_bn(Ppad)) {
        case 256: CLIPK_TRY((launch_gemm2<256, false, false, epi::DsIn>(&a, &b, 1, ks, ks, Bt, Ppad, gi, ep, ls))); break;
        case 192: CLIPK_TRY((launch_gemm2<192, false, false, epi::DsIn>(&a, &b, 1, ks, ks, Bt, Ppad, gi, ep, ls))); break;
        case 128: CLIPK_TRY((launch_gemm2<128, false, false, epi::DsIn>(&a, &b, 1, ks, ks, Bt, Ppad, gi, ep, ls))); break;
        default: CLIPK_TRY((launch_gemm2<64, false, false, epi::DsIn>(&a, &b, 1, ks, ks, Bt, Ppad, gi, ep, ls))); break;
      }
    }
    // K5: dt^[k, :] += sum_{i,p} E[i,k,p] V[i,p,:]   (K folds image and patch; split-K over images so that
    //     nsplit x tiles CTAs are busy; each split accumulates into its own fp32 slab, summed in the finalize)
    {
      const int want = gi < kDthSplits ? gi : kDthSplits;
      const int spb = (gi + want - 1) / want;                // images per split (last split may be shorter)
      const int nsplit = (gi + spb - 1) / spb;               // no empty split: an empty K range leaves TMEM undefined
      OperandDesc a, b;
      a.ptr = w.E; a.rows = Bt; a.k = P; a.ld = Ppad; a.batch = gi; a.batch_stride = (int64_t)Bt * Ppad; a.bmul = 0; a.smul = 1;
      a.sub_per_batch = spb;
      a.sub_total = gi;
      a.reverse = 1;   // K4 walked upwards: take the last-written E first
      b.ptr = V0; b.mn_major = true; b.rows = D; b.k = P; b.ld = D; b.batch = gi; b.batch_stride = (int64_t)P * D; b.bmul = 0; b.smul = 1;
      const int ks[1] = {spb * (Ppad / 64)};
      const int ksub[1] = {Ppad / 64};
      epi::Store<false>::Params ep{w.dth, D, (int64_t)Bt * D, Bt, D, 1.f, 1};
      CLIPK_TRY(launch_nd<epi::Store<false>, false>((engine2_mask() & kK5) != 0, &a, &b, 1, ks, ksub, Bt, D, nsplit, ep, ls));
    }
    // K6: dV_i = A_i^T Gn_i + E_i^T T^ - rnV^2 dsdot V.  Computed in whichever orientation wastes less of the
    // 256-row CTA-pair tiles: rows = patches (P) or, transposed, rows = features (D).
    {
      const bool transposed = (engine2_mask() & kK6) != 0 && env_int("CLIPK_AP_K6T", 1) != 0 && D % 8 == 0 &&
                              (int64_t)round_up(D, 256) * round_up(P, 16) < (int64_t)round_up(P, 256) * round_up(D, 16);
      const int ks[2] = {(Bt + 63) / 64, (Bt + 63) / 64};
      if (transposed) {
        OperandDesc a[2], b[2];
        a[0].ptr = w.G; a[0].mn_major = true; a[0].rows = D; a[0].k = Bt; a[0].ld = D; a[0].batch = gi;
        a[0].batch_stride = (int64_t)Bt * D; a[0].bmul = 1;
        a[0].reverse = 1;
        b[0].ptr = w.A; b[0].mn_major = true; b[0].rows = P; b[0].k = Bt; b[0].ld = Ppad; b[0].batch = gi;
        b[0].batch_stride = (int64_t)Bt * Ppad; b[0].bmul = 1;
        a[1].ptr = sh.That; a[1].mn_major = true; a[1].rows = D; a[1].k = Bt; a[1].ld = D; a[1].batch = 1; a[1].bmul = 0;
        b[1] = b[0]; b[1].ptr = w.E;
        const eng::OutDesc odv{dV + (int64_t)i0 * P * D, D, (int64_t)P * D, P, D, gi};
        const eng::OutDesc ov{const_cast<__nv_bfloat16*>(V0), D, (int64_t)P * D, P, D, gi};
        epi::DvOutT::Params ep{odv, ov, rnV0, dsdot0, P};
        CLIPK_TRY((launch_gemm2<256, true, true, epi::DvOutT>(a, b, 2, ks, ks, D, P, gi, ep, ls)));
      } else {
        OperandDesc a[2], b[2];
        a[0].ptr = w.A; a[0].mn_major = true; a[0].rows = P; a[0].k = Bt; a[0].ld = Ppad; a[0].batch = gi;
        a[0].batch_stride = (int64_t)Bt * Ppad; a[0].bmul = 1;
        a[0].reverse = 1;
        b[0].ptr = w.G; b[0].mn_major = true; b[0].rows = D; b[0].k = Bt; b[0].ld = D; b[0].batch = gi;
        b[0].batch_stride = (int64_t)Bt * D; b[0].bmul = 1;
        a[1] = a[0]; a[1].ptr = w.E;
        b[1].ptr = sh.That; b[1].mn_major = true; b[1].rows = D; b[1].k = Bt; b[1].ld = D; b[1].batch = 1; b[1].bmul = 0;
        epi::DvOut::Params ep{{dV + (int64_t)i0 * P * D, D, (int64_t)P * D, P, D, gi}, V0, rnV0, dsdot0, P, D};
        CLIPK_TRY(launch_nd<epi::DvOut, true>((engine2_mask() & kK6) != 0, a, b, 2, ks, ks, P, D, gi, ep, ls));
      }
    }
  }
  CLIPK_TRY(joiner.join());
  DthSlabs slabs{};
  for (int l = 0; l < lanes; ++l) slabs.p[l] = wl[l].dth;
  dtext_finalize_kernel<<<Bt, 256, 0, st>>>(T, rnT, slabs, lanes, kDthSplits, Bt, D, dT);
  count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace clipk

extern "C" {

size_t clipk_pacl_allpairs_workspace_bytes(int Bi, int Bt, int P, int D, int group, int lanes, int mode) {
  const int backward = mode & 1, pooled = (mode >> 1) & 1;
  if (group <= 0) {
    clipk::MegaWs mw{};
    return clipk::mega_carve(&mw, nullptr, clipk::mega_plan(Bi, Bt, P, D, group, lanes, backward), Bi, Bt, P, D, backward);
  }
  clipk::ApWorkspace w[clipk::kMaxLanes]{};
  clipk::ApShared sh{};
  if (lanes < 1 || lanes > clipk::kMaxLanes) return 0;
  return clipk::ap_carve(w, &sh, nullptr, Bi, Bt, P, D, group, lanes, backward, pooled,
                         !backward && clipk::fused_fwd_eligible(P, D));
}

int clipk_pacl_allpairs_fwd(const void* V, const void* T, int Bi, int Bt, int P, int D, int act, float c, float* rnV,
                            float* rnT, float* num, float* usq, float* scores, void* pooled, void* workspace,
                            size_t ws_bytes, int group, int lanes, int rnv_given, void* stream) {
  CLIPK_TRY(clipk::check_device());
  return clipk::allpairs_fwd(static_cast<const __nv_bfloat16*>(V), static_cast<const __nv_bfloat16*>(T), Bi, Bt, P, D,
                             act, c, rnV, rnT, num, usq, scores, static_cast<__nv_bfloat16*>(pooled), workspace,
                             ws_bytes, group, lanes, rnv_given, static_cast<cudaStream_t>(stream));
}

int clipk_pacl_allpairs_bwd(const void* V, const void* T, int Bi, int Bt, int P, int D, int act, float c,
                            const float* rnV, const float* rnT, const float* num, const float* usq,
                            const float* dscores, const void* pooled, void* dV, float* dT, void* workspace,
                            size_t ws_bytes, int group, int lanes, void* stream) {
  CLIPK_TRY(clipk::check_device());
  return clipk::allpairs_bwd(static_cast<const __nv_bfloat16*>(V), static_cast<const __nv_bfloat16*>(T), Bi, Bt, P, D,
                             act, c, rnV, rnT, num, usq, dscores, static_cast<const __nv_bfloat16*>(pooled),
                             static_cast<__nv_bfloat16*>(dV), dT, workspace, ws_bytes, group, lanes,
                             static_cast<cudaStream_t>(stream));
}
}
