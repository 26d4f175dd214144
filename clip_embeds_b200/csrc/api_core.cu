// Error reporting, device gate and TMA tensor-map construction.
#include <atomic>
#include <mutex>
#include <map>
#include <string>
#include <vector>
#include <stdlib.h>

#include <cudaTypedefs.h>

#include "common.cuh"

namespace clipk {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// ---- optional per-launch event trace (CLIPK_TRACE=1): diagnostic only, never enabled by the product path
struct TraceRec {
  const char* name;
  cudaEvent_t e0, e1;
};
static std::mutex g_trace_mu;
static std::vector<TraceRec> g_trace;
bool trace_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CLIPK_TRACE");
    on = (e != nullptr && e[0] != '0') ? 1 : 0;
  }
  return on == 1;
}
void trace_begin(const char* name, cudaStream_t st) {
  TraceRec r{name, nullptr, nullptr};
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
  std::lock_guard<std::mutex> lk(g_trace_mu);
  g_trace.push_back(r);
}
void trace_end(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_trace_mu);
  if (!g_trace.empty()) cudaEventRecord(g_trace.back().e1, st);
}

static thread_local int g_pdl_block = 0;
PdlBlock::PdlBlock(bool block) : active(block) { if (active) ++g_pdl_block; }
PdlBlock::~PdlBlock() { if (active) --g_pdl_block; }
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("CLIPK_PDL");
    return e == nullptr || e[0] != '0';
  }();
  return on && g_pdl_block == 0;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

struct DevInfo {
  std::atomic<int> state{0};  // 0 unknown, 1 ok, 2 wrong arch
  int sms = 0;
};
static DevInfo g_dev[64];

static int query_device(int* dev_out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    set_error("no CUDA device available (cudaGetDevice failed): the clipk kernels require a B200 (sm_100); "
              "there is no CPU fallback");
    return CLIPK_ERR_CUDA;
  }
  *dev_out = dev;
  if (g_dev[dev].state.load(std::memory_order_acquire) == 0) {
    int major = 0, minor = 0, sms = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      set_error("cudaDeviceGetAttribute failed");
      return CLIPK_ERR_CUDA;
    }
    g_dev[dev].sms = sms;
    g_dev[dev].state.store(major == 10 ? 1 : 2, std::memory_order_release);
  }
  return 0;
}

int check_device() {
  int dev = 0;
  int r = query_device(&dev);
  if (r != 0) return r;
  if (g_dev[dev].state.load(std::memory_order_acquire) != 1) {
    set_error("device %d is not sm_100 (B200): the clipk kernels are sm_100a-only and have no fallback", dev);
    return CLIPK_ERR_ARCH;
  }
  return 0;
}

int sm_count() {
  int dev = 0;
  if (query_device(&dev) != 0) return 1;
  return g_dev[dev].sms > 0 ? g_dev[dev].sms : 1;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static std::atomic<void*> cached{nullptr};
  void* p = cached.load(std::memory_order_acquire);
  if (p == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
      return nullptr;
    }
    cached.store(fn, std::memory_order_release);
    p = fn;
  }
  return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
}

int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batch,
                   uint64_t row_stride_bytes, uint64_t batch_stride_bytes, uint32_t box_rows) {
  return make_tmap_bf16_box(out, base, inner, rows, batch, row_stride_bytes, batch_stride_bytes, 64, box_rows);
}

int make_tmap_bf16_box(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batch,
                       uint64_t row_stride_bytes, uint64_t batch_stride_bytes, uint32_t box_inner, uint32_t box_rows) {
  auto encode = get_encode_fn();
  if (encode == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return CLIPK_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (row_stride_bytes & 15) != 0 ||
      (batch > 1 && (batch_stride_bytes & 15) != 0)) {
    set_error("TMA operand must be 16-byte aligned with 16-byte-multiple strides (base %p, row stride %llu B, "
              "batch stride %llu B)", base, (unsigned long long)row_stride_bytes,
              (unsigned long long)batch_stride_bytes);
    return CLIPK_ERR_INVALID;
  }
  if (batch < 1) batch = 1;
  if (batch == 1 || batch_stride_bytes == 0) batch_stride_bytes = row_stride_bytes * (rows > 0 ? rows : 1);
  cuuint64_t dims[3] = {inner, rows, batch};
  cuuint64_t strides[2] = {row_stride_bytes, batch_stride_bytes};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  // the swizzle span equals the box's inner extent in bytes: 64 elements -> 128 B, 32 elements -> 64 B
  const CUtensorMapSwizzle swz = box_inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): dims (%llu,%llu,%llu) strides (%llu,%llu) box rows %u", (int)r,
              (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)batch,
              (unsigned long long)row_stride_bytes, (unsigned long long)batch_stride_bytes, box_rows);
    return CLIPK_ERR_CUDA;
  }
  return 0;
}

}  // namespace clipk

extern "C" {
const char* clipk_last_error(void) { return clipk::g_err; }
int clipk_version(void) { return 100; }
unsigned long long clipk_launch_count(void) { return clipk::g_launches.load(std::memory_order_relaxed); }
int clipk_check_device(void) { return clipk::check_device(); }
int clipk_trace_dump(void) {
  if (cudaDeviceSynchronize() != cudaSuccess) return CLIPK_ERR_CUDA;
  std::lock_guard<std::mutex> lk(clipk::g_trace_mu);
  std::map<std::string, std::pair<int, double>> agg;
  double tot = 0;
  for (auto& r : clipk::g_trace) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    std::string n(r.name);
    size_t w = n.find("[with ");
    if (w != std::string::npos) n = n.substr(w + 6);
    if (n.size() > 100) n.resize(100);
    agg[n].first += 1;
    agg[n].second += ms;
    tot += ms;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  clipk::g_trace.clear();
  for (auto& kv : agg)
    printf("trace %9.3f ms %5d launches avg %8.1f us %5.1f%%  %s\n", kv.second.second, kv.second.first,
           1e3 * kv.second.second / kv.second.first, 100.0 * kv.second.second / (tot > 0 ? tot : 1), kv.first.c_str());
  printf("trace total %.3f ms\n", tot);
  fflush(stdout);
  return 0;
}
}
