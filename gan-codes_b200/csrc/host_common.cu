#include "host_common.h"

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace dfgan {

std::atomic<uint64_t> g_launches{0};

encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
  });
  return fn;
}

bool pdl_enabled() {
  static const bool on = getenv("DFGAN_NO_PDL") == nullptr;
  return on;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static CUtensorMapSwizzle swizzle_for(int bc) {
  return bc * 2 >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                       : (bc * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

int make_map_nhwc(CUtensorMap* m, const dfgan_view4& v, int bc, int bw, int bh, int bn) {
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return DFGAN_E_DRIVER;
  if ((reinterpret_cast<uintptr_t>(v.ptr) & 15) || (v.sw & 7) || (v.sh & 7) || (v.sn & 7))
    return DFGAN_E_ARG;
  cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
  cuuint64_t strides[3] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2};
  cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v.ptr), dims, strides,
                   box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bc),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DFGAN_OK : DFGAN_E_DRIVER;
}

int make_map_halo(CUtensorMap* m, const dfgan_view4& v, int bw, int bh) {
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return DFGAN_E_DRIVER;
  if ((reinterpret_cast<uintptr_t>(v.ptr) & 15) || (v.sw & 7) || (v.sh & 7) || (v.sn & 7) || (v.C & 7))
    return DFGAN_E_ARG;
  cuuint64_t dims[5] = {8, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)(v.C / 8), (cuuint64_t)v.N};
  cuuint64_t strides[4] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, 16, (cuuint64_t)v.sn * 2};
  cuuint32_t box[5] = {8, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)(v.C / 8), 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(v.ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DFGAN_OK : DFGAN_E_DRIVER;
}

int make_map_2d(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int bc, int brows) {
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return DFGAN_E_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (cols & 7)) return DFGAN_E_ARG;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)bc, (cuuint32_t)brows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                   box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bc),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DFGAN_OK : DFGAN_E_DRIVER;
}

struct ProfRec {
  cudaEvent_t e0, e1;
  int engine;
  double flops, bytes;
  char label[120];
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;

ProfileScope::ProfileScope(int engine, double flops, double bytes, cudaStream_t st_, const char* label)
    : slot(-1), st(st_) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{};
  r.engine = engine, r.flops = flops, r.bytes = bytes;
  snprintf(r.label, sizeof(r.label), "%s", label ? label : "");
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, st);
  g_prof.push_back(r);
  slot = static_cast<int>(g_prof.size()) - 1;
}
ProfileScope::~ProfileScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[slot].e1, st);
}

}  // namespace dfgan

extern "C" {
void dfgan_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(dfgan::g_prof_mu);
  dfgan::g_prof_on = on != 0;
}
// out[engine*4 + {0: ms, 1: flops, 2: bytes, 3: launches}] for engine 0 (conv fprop/dgrad) and 1 (wgrad);
// synchronises the device, then clears the records.
int dfgan_profile_read(double* out) {
  if (!out) return DFGAN_E_ARG;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return static_cast<int>(e);
  std::lock_guard<std::mutex> lk(dfgan::g_prof_mu);
  for (int i = 0; i < 8; ++i) out[i] = 0.0;
  // per-launch records (tools/prof_perlaunch.py): engine, label, ms, flops, algorithmic bytes
  const char* dump = getenv("DFGAN_PROFILE_DUMP");
  FILE* fd = dump ? fopen(dump, "a") : nullptr;
  for (auto& r : dfgan::g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      double* o = out + (r.engine ? 4 : 0);
      o[0] += ms, o[1] += r.flops, o[2] += r.bytes, o[3] += 1.0;
      if (fd) fprintf(fd, "%d\t%s\t%.4f\t%.0f\t%.0f\n", r.engine, r.label, ms, r.flops, r.bytes);
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  if (fd) fclose(fd);
  dfgan::g_prof.clear();
  return DFGAN_OK;
}
int dfgan_version(void) { return 100; }
const char* dfgan_arch(void) { return "sm_100a"; }
uint64_t dfgan_launch_count(void) { return dfgan::g_launches.load(); }
}
