// C ABI of libb200enc (see include/b200enc.h): argument checking, TMA tensor-map construction, launches.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>

#include "../../include/b200enc.h"
#include "attn_bwd.cuh"
#include "attn_bwd3.cuh"
#include "attn_fwd.cuh"
#include "attn_fwd3.cuh"
#include "gemm.cuh"
#include "gemm2.cuh"
#include "heads.cuh"
#include "optim.cuh"
#include "ponet.cuh"
#include "rowwise.cuh"

using namespace b200;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(B200_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return B200_OK;
}

// ---- cuTensorMapEncodeTiled is resolved at run time: the build box has no libcuda.so -------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t rows, cols, ld;
  uint32_t box_rows;   // bit 31 set: fp32 elements; bit 30 set: 64-byte boxes (SWIZZLE_64B)
  bool operator==(const MapKey& o) const { return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= k.rows * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= k.cols * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= (k.ld * 31 + k.box_rows) * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    return h;
  }
};
std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// 2D tensor [rows, cols] (fp16, or fp32 when f32 is set) with row pitch ld (elements); box = 128 bytes of columns
// (SWIZZLE_128B) x box_rows.
int get_tmap(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, CUtensorMap* out, bool f32 = false, bool box64 = false) {
  const uint64_t esz = f32 ? 4 : 2;
  const uint32_t box_bytes = box64 ? 64 : 128;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * esz) % 16) || cols == 0 || rows == 0 || box_rows == 0 || box_rows > 256)
    return fail(B200_ERR_SHAPE, "tensor map: ptr %p rows %llu cols %llu ld %llu box_rows %u violates 16B alignment / ld%%8", ptr,
                (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
  const MapKey key{ptr, rows, cols, ld, box_rows | (f32 ? 0x80000000u : 0u) | (box64 ? 0x40000000u : 0u)};
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) {
      *out = it->second;
      return B200_OK;
    }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(B200_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * esz};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_bytes / esz), box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  const CUresult r = enc(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, box64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", static_cast<int>(r));
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    if (g_maps.size() > 65536) g_maps.clear();
    g_maps.emplace(key, m);
  }
  *out = m;
  return B200_OK;
}

std::atomic<int> g_sm_limit{0};

// SMs the persistent kernels (one CTA / CTA pair per SM) may occupy: the device's count, minus what the caller reserved
// for concurrently running communication kernels (b200_set_sm_limit).
int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  const int lim = g_sm_limit.load(std::memory_order_relaxed);
  int m = (lim > 0 && lim < n) ? lim : n;
  if (m > 2) m &= ~1;                   // CTA pairs
  return m;
}

template <typename K>
int set_smem(K kern, int bytes) {
  const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return fail(B200_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", bytes, cudaGetErrorString(e));
  return B200_OK;
}


// grid of a streaming kernel over n8 8-element vectors
int stream_grid(size_t n8) {
  size_t g = (n8 + 255) / 256;
  const size_t cap = static_cast<size_t>(sm_count()) * 16;
  return static_cast<int>(g < cap ? (g ? g : 1) : cap);
}
const DropCfg kNoDrop{nullptr, 0, 0, 1.0f};
DropCfg make_drop(const uint32_t* seed, unsigned site, float p) {
  if (!seed || p <= 0.f) return kNoDrop;
  return DropCfg{seed, site, static_cast<uint32_t>(p * 32768.0f + 0.5f), 1.0f / (1.0f - p)};
}
// attention-probability sites (ptx.cuh: drop4_z): four 7-bit lane thresholds whose sum is round(512 p), packed one per byte
DropCfg make_drop_attn(const uint32_t* seed, unsigned site, float p) {
  if (!seed || p <= 0.f) return kNoDrop;
  const uint32_t t = static_cast<uint32_t>(p * 512.0f + 0.5f);
  uint32_t tt = 0;
  for (uint32_t l = 0; l < 4; ++l) tt |= (t / 4 + (l < t % 4 ? 1u : 0u)) << (8 * l);
  return DropCfg{seed, site, tt, 512.0f / (512.0f - static_cast<float>(t))};
}

template <int BN, int A_MN, int B_MN, int EPI, typename OutT, bool BF16IN = false>
int launch_gemm2(const Gemm2Maps& maps, const GemmArgs& g, cudaStream_t s) {
  auto kern = gemm2_f16_kernel<BN, A_MN, B_MN, EPI, OutT, BF16IN>;
  static int configured = set_smem(kern, Gemm2Smem<BN, EPI>::TOTAL);
  if (configured != B200_OK) return configured;
  const int m_tiles = (g.M + G2_BM - 1) / G2_BM, n_tiles = (g.N + BN - 1) / BN;
  const int units = m_tiles * n_tiles * (g.k_splits > 0 ? g.k_splits : 1);
  const int pairs = sm_count() / 2;
  const int grid = 2 * (units < pairs ? units : pairs);
  kern<<<grid, G2_THREADS, Gemm2Smem<BN, EPI>::TOTAL, s>>>(maps, g);
  return check_launch("gemm2_f16_kernel");
}

}  // namespace

extern "C" {

void b200_set_sm_limit(int sms) { g_sm_limit.store(sms); }

const char* b200_last_error(void) { return g_err.c_str(); }
int b200_version(void) { return 100; }
#ifndef B200_SRC_HASH
#define B200_SRC_HASH "unknown"
#endif
static const char g_src_tag[] = "B200SRC:" B200_SRC_HASH;        // the binding also finds this tag by scanning the file
const char* b200_source_hash(void) { return g_src_tag + 8; }
long long b200_launch_count(void) { return g_launches.load(); }
#ifdef B200_ATT_TRACE
// trace build only (tools/attn_trace.py): copy out the timeline records of CTA 0 (dst: [16 warps][4096], counts: [16])
int b200_att_trace_read(unsigned long long* dst, unsigned int* counts) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(counts, b200::g_att_trace_n, 16 * sizeof(unsigned int));
  cudaMemcpyFromSymbol(dst, b200::g_att_trace, 16 * b200::kAttTraceCap * sizeof(unsigned long long));
  return 0;
}
#endif

static int gemm_impl(const void* A, int lda, int a_layout, const void* B, int ldb, int b_layout, int M, int N, int K, int epilogue,
                     const float* bias, const void* aux, int ld_aux, void* out, int ld_out, int out_dtype, void* out2, int ld_out2,
                     const float* alpha, int k_splits, DropCfg drop, void* stream, float* colsum = nullptr, const float* col_alpha = nullptr);

int b200_gemm_f16(const void* A, int lda, int a_layout, const void* B, int ldb, int b_layout, int M, int N, int K, int epilogue,
                  const float* bias, const void* aux, int ld_aux, void* out, int ld_out, int out_dtype, void* out2, int ld_out2,
                  const float* alpha, int k_splits, void* stream) {
  return gemm_impl(A, lda, a_layout, B, ldb, b_layout, M, N, K, epilogue, bias, aux, ld_aux, out, ld_out, out_dtype, out2, ld_out2, alpha,
                   k_splits, DropCfg{nullptr, 0, 0, 1.0f}, stream);
}

int b200_gemm_f16_drop(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int epilogue, const float* bias, const void* aux,
                       int ld_aux, void* out, int ld_out, int out_dtype, const uint32_t* seed, unsigned site, float p, void* stream) {
  if (epilogue != EPI_BIAS_RES32 && epilogue != EPI_BIAS_RES) return fail(B200_ERR_SHAPE, "gemm_drop: dropout is fused into the residual epilogues only");
  if (p < 0.f || p >= 1.f) return fail(B200_ERR_SHAPE, "gemm_drop: p must be in [0,1)");
  return gemm_impl(A, lda, 0, B, ldb, 0, M, N, K, epilogue, bias, aux, ld_aux, out, ld_out, out_dtype, nullptr, 0, nullptr, 1,
                   make_drop(seed, site, p), stream);
}

int b200_gemm_f16_resadd(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias, float* out, int ld_out,
                         const uint32_t* seed, unsigned site, float p, void* stream) {
  if (p < 0.f || p >= 1.f) return fail(B200_ERR_SHAPE, "gemm_resadd: p must be in [0,1)");
  return gemm_impl(A, lda, 0, B, ldb, 0, M, N, K, EPI_RESADD, bias, nullptr, 0, out, ld_out, B200_DT_F32, nullptr, 0, nullptr, 1,
                   make_drop(seed, site, p), stream);
}

/* bf16 operands (A, B bf16; fp32 accumulate): forward Linear (K-major x K-major: STORE / BIAS, result bf16 or fp32), dgrad
 * (b_layout = 1, STORE, bf16) and wgrad (a_layout = b_layout = 1, ATOMIC split-K, fp32). */
int b200_gemm_bf16(const void* A, int lda, int a_layout, const void* B, int ldb, int b_layout, int M, int N, int K, int epilogue, const float* bias,
                   void* out, int ld_out, int out_dtype, const float* alpha, int k_splits, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || !A || !B || !out) return fail(B200_ERR_SHAPE, "gemm_bf16: empty problem or null operand");
  if ((N % 8) || (ld_out % 4)) return fail(B200_ERR_SHAPE, "gemm_bf16: N %% 8 and ld_out %% 4 required (N=%d ld_out=%d)", N, ld_out);
  if (epilogue == EPI_BIAS && !bias) return fail(B200_ERR_SHAPE, "gemm_bf16: EPI_BIAS needs bias");
  if (k_splits > 1 && epilogue != EPI_ATOMIC) return fail(B200_ERR_SHAPE, "gemm_bf16: split-K only with the atomic epilogue");
  constexpr int BN = 256;
  Gemm2Maps mp;
  int rc = a_layout == 0 ? get_tmap(A, M, K, lda, 128, &mp.a) : get_tmap(A, K, M, lda, GEMM_BK, &mp.a);      // (2-byte elements: the fp16 map moves bf16 bits alike)
  if (rc) return rc;
  rc = b_layout == 0 ? get_tmap(B, N, K, ldb, BN / 2, &mp.b) : get_tmap(B, K, N, ldb, GEMM_BK, &mp.b);
  if (rc) return rc;
  if ((rc = get_tmap(out, M, N, ld_out, 32, &mp.out, out_dtype == B200_DT_F32))) return rc;
  mp.aux = mp.out;
  mp.out2 = mp.out;
  GemmArgs g{M, N, K, k_splits > 0 ? k_splits : 1, bias, nullptr, 0, out, ld_out, nullptr, 0, alpha, kNoDrop, nullptr, nullptr};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (a_layout * 1000 + b_layout * 100 + epilogue * 10 + out_dtype) {
    case 0 * 1000 + 0 * 100 + EPI_STORE * 10 + B200_DT_BF16: return launch_gemm2<BN, 0, 0, EPI_STORE, __nv_bfloat16>(mp, g, s);
    case 0 * 1000 + 0 * 100 + EPI_BIAS * 10 + B200_DT_BF16: return launch_gemm2<BN, 0, 0, EPI_BIAS, __nv_bfloat16>(mp, g, s);
    case 0 * 1000 + 0 * 100 + EPI_STORE * 10 + B200_DT_F32: return launch_gemm2<BN, 0, 0, EPI_STORE, float, true>(mp, g, s);
    case 0 * 1000 + 0 * 100 + EPI_BIAS * 10 + B200_DT_F32: return launch_gemm2<BN, 0, 0, EPI_BIAS, float, true>(mp, g, s);
    case 0 * 1000 + 1 * 100 + EPI_STORE * 10 + B200_DT_BF16: return launch_gemm2<BN, 0, 1, EPI_STORE, __nv_bfloat16>(mp, g, s);
    case 1 * 1000 + 1 * 100 + EPI_ATOMIC * 10 + B200_DT_F32: return launch_gemm2<BN, 1, 1, EPI_ATOMIC, float, true>(mp, g, s);
    default:
      return fail(B200_ERR_SHAPE, "gemm_bf16: unsupported (a_layout=%d, b_layout=%d, epilogue=%d, out_dtype=%d)", a_layout, b_layout, epilogue, out_dtype);
  }
}

int b200_cast_f32_to_bf16(const float* src, void* dst, size_t n, void* stream) {
  if (n % 8) return fail(B200_ERR_SHAPE, "cast: n %% 8 != 0");
  cast_f32_bf16_kernel<<<stream_grid(n / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, static_cast<__nv_bfloat16*>(dst), n / 8);
  return check_launch("cast_f32_bf16_kernel");
}

int b200_gemm_f16_dgelu_colsum(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const void* dact, int ld_dact, void* out,
                               int ld_out, float* colsum, const float* col_alpha, void* stream) {
  if (!colsum) return fail(B200_ERR_SHAPE, "gemm_dgelu_colsum: colsum is required (use b200_gemm_f16 for the plain epilogue)");
  return gemm_impl(A, lda, 0, B, ldb, 1, M, N, K, EPI_DGELU, nullptr, dact, ld_dact, out, ld_out, B200_DT_F16, nullptr, 0, nullptr, 1, kNoDrop,
                   stream, colsum, col_alpha);
}

int b200_gemm_f16_dgrad_delta(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const void* ctx, int ld_ctx, void* out,
                              int ld_out, float* delta, int heads, int Sq, void* stream) {
  if (!delta || heads <= 0 || Sq <= 0 || N != heads * 64 || (M % Sq))
    return fail(B200_ERR_SHAPE, "gemm_dgrad_delta: need delta, N == heads*64 and M %% Sq == 0 (M=%d N=%d heads=%d Sq=%d)", M, N, heads, Sq);
  // delta and Sq travel in the out2 / ld_out2 slots (GemmArgs)
  return gemm_impl(A, lda, 0, B, ldb, 1, M, N, K, EPI_STORE_DELTA, nullptr, ctx, ld_ctx, out, ld_out, B200_DT_F16, delta, Sq, nullptr, 1,
                   kNoDrop, stream);
}

static int gemm_impl(const void* A, int lda, int a_layout, const void* B, int ldb, int b_layout, int M, int N, int K, int epilogue,
                     const float* bias, const void* aux, int ld_aux, void* out, int ld_out, int out_dtype, void* out2, int ld_out2,
                     const float* alpha, int k_splits, DropCfg drop, void* stream, float* colsum, const float* col_alpha) {
  if (colsum && epilogue != EPI_DGELU) return fail(B200_ERR_SHAPE, "gemm: the fused column sum rides on the EPI_DGELU epilogue only");
  if (M <= 0 || N <= 0 || K <= 0) return fail(B200_ERR_SHAPE, "gemm: empty problem %dx%dx%d", M, N, K);
  if ((N % 4) || (ld_out % 4)) return fail(B200_ERR_SHAPE, "gemm: N and ld_out must be multiples of 4 (N=%d ld_out=%d)", N, ld_out);
  if (!A || !B || !out) return fail(B200_ERR_SHAPE, "gemm: null operand");
  const bool needs_bias = epilogue == EPI_BIAS || epilogue == EPI_BIAS_GELU || epilogue == EPI_BIAS_RES || epilogue == EPI_BIAS_RES32 ||
                          epilogue == EPI_RESADD;
  const bool needs_aux = epilogue == EPI_BIAS_RES || epilogue == EPI_DGELU || epilogue == EPI_ADD || epilogue == EPI_BIAS_RES32 ||
                         epilogue == EPI_STORE_DELTA;
  if (needs_bias && !bias) return fail(B200_ERR_SHAPE, "gemm: epilogue %d needs bias", epilogue);
  if (needs_aux && (!aux || (ld_aux % 4))) return fail(B200_ERR_SHAPE, "gemm: epilogue %d needs aux with ld%%4==0", epilogue);
  if (k_splits > 1 && epilogue != EPI_ATOMIC) return fail(B200_ERR_SHAPE, "gemm: split-K only with the atomic epilogue");
  constexpr int BN = 256;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc;
  // 2-CTA path: per-CTA operand boxes are 128 rows (K-major) / 64x64 (MN-major); epilogue slabs are 32 rows x 128 B
  Gemm2Maps mp;
  rc = a_layout == 0 ? get_tmap(A, M, K, lda, 128, &mp.a) : get_tmap(A, K, M, lda, GEMM_BK, &mp.a);
  if (rc) return rc;
  rc = b_layout == 0 ? get_tmap(B, N, K, ldb, BN / 2, &mp.b) : get_tmap(B, K, N, ldb, GEMM_BK, &mp.b);
  if (rc) return rc;
  const bool o32 = out_dtype == B200_DT_F32;
  if ((rc = get_tmap(out, M, N, ld_out, 32, &mp.out, o32))) return rc;
  mp.aux = mp.out;
  mp.out2 = mp.out;
  if (epilogue == EPI_BIAS_GELU && out2 && (rc = get_tmap(out2, M, N, ld_out2, 32, &mp.out2))) return rc;
  const bool has_out2 = out2 && epilogue != EPI_STORE_DELTA;       // (EPI_STORE_DELTA keeps its fp32 row statistic in the out2 slot)
  if ((needs_aux || has_out2) && ((N % 8) || (needs_aux && (ld_aux * (epilogue == EPI_BIAS_RES32 ? 4 : 2)) % 16) || (has_out2 && (ld_out2 % 8)) ||
                              (needs_aux && (reinterpret_cast<uintptr_t>(aux) & 15)) || (has_out2 && (reinterpret_cast<uintptr_t>(out2) & 15))))
    return fail(B200_ERR_SHAPE, "gemm: aux / out2 need N %% 8 == 0 and 16-byte aligned rows");
  GemmArgs g2{M, N, K, k_splits > 0 ? k_splits : 1, bias, static_cast<const __half*>(aux), ld_aux, out, ld_out,
              static_cast<__half*>(out2), ld_out2, alpha, drop, colsum, col_alpha};
  switch (a_layout * 1000 + b_layout * 100 + epilogue * 10 + out_dtype) {
    case 0 * 1000 + 0 * 100 + EPI_STORE * 10 + B200_DT_F16: return launch_gemm2<BN, 0, 0, EPI_STORE, __half>(mp, g2, s);
    case 0 * 1000 + 0 * 100 + EPI_STORE * 10 + B200_DT_F32: return launch_gemm2<BN, 0, 0, EPI_STORE, float>(mp, g2, s);
    case 0 * 1000 + 0 * 100 + EPI_BIAS * 10 + B200_DT_F16: return launch_gemm2<BN, 0, 0, EPI_BIAS, __half>(mp, g2, s);
    case 0 * 1000 + 0 * 100 + EPI_BIAS * 10 + B200_DT_F32: return launch_gemm2<BN, 0, 0, EPI_BIAS, float>(mp, g2, s);
    case 0 * 1000 + 0 * 100 + EPI_BIAS_GELU * 10 + B200_DT_F16: return launch_gemm2<BN, 0, 0, EPI_BIAS_GELU, __half>(mp, g2, s);
    case 0 * 1000 + 0 * 100 + EPI_BIAS_RES * 10 + B200_DT_F16: return launch_gemm2<BN, 0, 0, EPI_BIAS_RES, __half>(mp, g2, s);
    case 0 * 1000 + 0 * 100 + EPI_BIAS_RES32 * 10 + B200_DT_F32: return launch_gemm2<BN, 0, 0, EPI_BIAS_RES32, float>(mp, g2, s);
    case 0 * 1000 + 1 * 100 + EPI_STORE * 10 + B200_DT_F16: return launch_gemm2<BN, 0, 1, EPI_STORE, __half>(mp, g2, s);
    case 0 * 1000 + 1 * 100 + EPI_ADD * 10 + B200_DT_F16: return launch_gemm2<BN, 0, 1, EPI_ADD, __half>(mp, g2, s);
    case 0 * 1000 + 1 * 100 + EPI_DGELU * 10 + B200_DT_F16: return launch_gemm2<BN, 0, 1, EPI_DGELU, __half>(mp, g2, s);
    case 1 * 1000 + 1 * 100 + EPI_ATOMIC * 10 + B200_DT_F32: return launch_gemm2<BN, 1, 1, EPI_ATOMIC, float>(mp, g2, s);
    case 0 * 1000 + 0 * 100 + EPI_RESADD * 10 + B200_DT_F32: return launch_gemm2<BN, 0, 0, EPI_RESADD, float>(mp, g2, s);
    case 0 * 1000 + 1 * 100 + EPI_STORE_DELTA * 10 + B200_DT_F16: return launch_gemm2<BN, 0, 1, EPI_STORE_DELTA, __half>(mp, g2, s);
    default:
      return fail(B200_ERR_SHAPE, "gemm: unsupported (a_layout=%d, b_layout=%d, epilogue=%d, out_dtype=%d)", a_layout, b_layout,
                  epilogue, out_dtype);
  }
}

static int attn_fwd_impl(const void* q, int ldq, int q_col0, const void* kv, int ldkv, int k_col0, int v_col0, const float* key_bias,
                         const int32_t* kv_len, void* ctx, int ld_out, float* lse2, int B, int heads, int Sq, int Sk, const uint32_t* seed,
                         unsigned site, float p, const int32_t* cu_seqlens, long long rows, void* stream) {
  if (B <= 0 || heads <= 0 || Sq <= 0 || Sk <= 0 || rows <= 0) return fail(B200_ERR_SHAPE, "attn_fwd: empty problem");
  if ((q_col0 % 8) || (k_col0 % 8) || (v_col0 % 8) || (ld_out % 8)) return fail(B200_ERR_SHAPE, "attn_fwd: column offsets / ld_out must be multiples of 8");
  if (p < 0.f || p > 0.9f) return fail(B200_ERR_SHAPE, "attn_fwd: dropout probability %g outside [0, 0.9]", p);
  const DropCfg drop = make_drop_attn(seed, site, p);
  const uint64_t rq = cu_seqlens ? static_cast<uint64_t>(rows) : static_cast<uint64_t>(B) * Sq;
  const uint64_t rk = cu_seqlens ? static_cast<uint64_t>(rows) : static_cast<uint64_t>(B) * Sk;
  CUtensorMap tq, tkv, to;
  int rc = get_tmap(q, rq, ldq, ldq, ATT_BQ, &tq);
  if (rc) return rc;
  rc = get_tmap(kv, rk, ldkv, ldkv, ATT_BK, &tkv);
  if (rc) return rc;
  // context output fp16 [rows, ld_out], 64 x 32 patches (one per softmax warp)
  if ((rc = get_tmap(ctx, rq, ld_out, ld_out, 32, &to))) return rc;
  static int e0 = set_smem(attn_fwd3_kernel<false>, AttnFwd3Smem::TOTAL);
  static int e1 = set_smem(attn_fwd3_kernel<true>, AttnFwd3Smem::TOTAL);
  if (e0 != B200_OK || e1 != B200_OK) return e0 ? e0 : e1;
  AttnFwdArgs a{B, heads, Sq, Sk, q_col0, k_col0, v_col0, key_bias, kv_len, static_cast<__half*>(ctx), ld_out, lse2, kAttScaleLog2, drop, cu_seqlens};
  // persistent kernel, one CTA per SM walking (batch, head, 256-query pair) items
  const long long items = static_cast<long long>((Sq + 2 * ATT_BQ - 1) / (2 * ATT_BQ)) * heads * B;
  const int ctas = items < sm_count() ? static_cast<int>(items) : sm_count();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (drop.seed_base) attn_fwd3_kernel<true><<<ctas, ATTP_THREADS, AttnFwd3Smem::TOTAL, st>>>(tq, tkv, to, a);
  else attn_fwd3_kernel<false><<<ctas, ATTP_THREADS, AttnFwd3Smem::TOTAL, st>>>(tq, tkv, to, a);
  return check_launch("attn_fwd3_kernel");
}

int b200_attn_fwd_drop(const void* q, int ldq, int q_col0, const void* kv, int ldkv, int k_col0, int v_col0, const float* key_bias,
                       const int32_t* kv_len, void* ctx, int ld_out, float* lse2, int B, int heads, int Sq, int Sk, const uint32_t* seed,
                       unsigned site, float p, void* stream) {
  return attn_fwd_impl(q, ldq, q_col0, kv, ldkv, k_col0, v_col0, key_bias, kv_len, ctx, ld_out, lse2, B, heads, Sq, Sk, seed, site, p, nullptr,
                       static_cast<long long>(B) * Sq, stream);
}

int b200_attn_fwd_varlen(const void* qkv, int ld, int q_col0, int k_col0, int v_col0, const int32_t* cu_seqlens, long long rows, void* ctx, int ld_out,
                         float* lse2, int B, int heads, int S_max, const uint32_t* seed, unsigned site, float p, void* stream) {
  if (!cu_seqlens) return fail(B200_ERR_SHAPE, "attn_fwd_varlen: cu_seqlens is required");
  return attn_fwd_impl(qkv, ld, q_col0, qkv, ld, k_col0, v_col0, nullptr, nullptr, ctx, ld_out, lse2, B, heads, S_max, S_max, seed, site, p,
                       cu_seqlens, rows, stream);
}

int b200_attn_fwd(const void* q, int ldq, int q_col0, const void* kv, int ldkv, int k_col0, int v_col0, const float* key_bias,
                  const int32_t* kv_len, void* ctx, int ld_out, float* lse2, int B, int heads, int Sq, int Sk, void* stream) {
  return b200_attn_fwd_drop(q, ldq, q_col0, kv, ldkv, k_col0, v_col0, key_bias, kv_len, ctx, ld_out, lse2, B, heads, Sq, Sk, nullptr, 0, 0.f, stream);
}

int b200_attn_probs(const void* q, int ldq, int q_col0, const void* k, int ldk, int k_col0, const float* key_bias, const float* lse2,
                    float* probs, int B, int heads, int Sq, int Sk, void* stream) {
  if (!lse2 || !probs) return fail(B200_ERR_SHAPE, "attn_probs: lse2 and probs are required");
  dim3 grid((Sk + 31) / 32, (Sq + 31) / 32, B * heads);
  attn_probs_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(q) + q_col0, ldq,
                                                                        static_cast<const __half*>(k) + k_col0, ldk, key_bias, lse2,
                                                                        probs, B, heads, Sq, Sk, 1.4426950408889634f / 8.0f);
  return check_launch("attn_probs_kernel");
}

}  // extern "C"

namespace {
template <typename T>
__global__ void mask_to_bias_kernel(const T* __restrict__ mask, float* __restrict__ bias, int32_t* __restrict__ kv_len, int S) {
  // kv_len[b] = +(last kept key + 1) when the kept keys form a prefix (right padding: the only masks the reference
  // builds), -(last kept key + 1) when there are holes inside the kept range (kernels then read the per-key bias).
  const int b = blockIdx.x;
  int last = 0, cnt = 0;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const bool keep = mask[static_cast<size_t>(b) * S + s] != T(0);
    bias[static_cast<size_t>(b) * S + s] = keep ? 0.f : -INFINITY;
    if (keep) {
      last = s + 1;
      ++cnt;
    }
  }
  __shared__ int red[32], redc[32];
  for (int o = 16; o > 0; o >>= 1) {
    last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = last;
    redc[threadIdx.x >> 5] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int m = 0, c = 0;
    for (int i = 0; i < (blockDim.x + 31) / 32; ++i) {
      m = max(m, red[i]);
      c += redc[i];
    }
    if (kv_len) kv_len[b] = (c == m) ? m : -m;
  }
}
}  // namespace

extern "C" {

int b200_mask_to_bias(const void* mask, int mask_dtype, float* key_bias, int32_t* kv_len, int B, int S, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (mask_dtype == 0) mask_to_bias_kernel<int64_t><<<B, 256, 0, s>>>(static_cast<const int64_t*>(mask), key_bias, kv_len, S);
  else if (mask_dtype == 1) mask_to_bias_kernel<float><<<B, 256, 0, s>>>(static_cast<const float*>(mask), key_bias, kv_len, S);
  else if (mask_dtype == 2) mask_to_bias_kernel<int32_t><<<B, 256, 0, s>>>(static_cast<const int32_t*>(mask), key_bias, kv_len, S);
  else return fail(B200_ERR_DTYPE, "mask_to_bias: dtype %d", mask_dtype);
  return check_launch("mask_to_bias_kernel");
}

static int check_row_shape(const char* who, int rows, int H) {
  if (rows <= 0 || H <= 0 || (H % 8) || H > ROW_MAXV * 256) return fail(B200_ERR_SHAPE, "%s: rows=%d H=%d (need H%%8==0, H<=%d)", who, rows, H, ROW_MAXV * 256);
  return B200_OK;
}

int b200_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y, float* y32, float* mean, float* rstd,
                       int rows, int H, float eps, void* stream) {
  if (int rc = check_row_shape("layernorm_fwd", rows, H)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = (rows + ROW_WARPS - 1) / ROW_WARPS;
  if (x_dtype == B200_DT_F32)
    ln_fwd_kernel<float><<<grid, ROW_WARPS * 32, 0, s>>>(static_cast<const float*>(x), gamma, beta, static_cast<__half*>(y), y32, mean, rstd, rows, H, eps);
  else
    ln_fwd_kernel<__half><<<grid, ROW_WARPS * 32, 0, s>>>(static_cast<const __half*>(x), gamma, beta, static_cast<__half*>(y), y32, mean, rstd, rows, H, eps);
  return check_launch("ln_fwd_kernel");
}

static int layernorm_bwd_impl(const void* dy, const void* dy2, const void* x, int x_dtype, const float* mean, const float* rstd, const float* gamma,
                              void* dx, float* dgamma, float* dbeta, float* dbias, const float* alpha, int rows, int H, void* dx_drop, DropCfg drop,
                              void* stream);

int b200_layernorm_bwd(const void* dy, const void* dy2, const void* x, int x_dtype, const float* mean, const float* rstd, const float* gamma,
                       void* dx, float* dgamma, float* dbeta, float* dbias, const float* alpha, int rows, int H, void* stream) {
  return layernorm_bwd_impl(dy, dy2, x, x_dtype, mean, rstd, gamma, dx, dgamma, dbeta, dbias, alpha, rows, H, nullptr, kNoDrop, stream);
}

int b200_layernorm_bwd_drop(const void* dy, const void* dy2, const void* x, int x_dtype, const float* mean, const float* rstd, const float* gamma,
                            void* dx, void* dx_drop, float* dgamma, float* dbeta, float* dbias, const float* alpha, int rows, int H,
                            const uint32_t* seed, unsigned site, float p, void* stream) {
  if (!dx_drop) return fail(B200_ERR_SHAPE, "layernorm_bwd_drop: dx_drop is required");
  return layernorm_bwd_impl(dy, dy2, x, x_dtype, mean, rstd, gamma, dx, dgamma, dbeta, dbias, alpha, rows, H, dx_drop, make_drop(seed, site, p), stream);
}

static int layernorm_bwd_impl(const void* dy, const void* dy2, const void* x, int x_dtype, const float* mean, const float* rstd, const float* gamma,
                              void* dx, float* dgamma, float* dbeta, float* dbias, const float* alpha, int rows, int H, void* dx_drop, DropCfg drop,
                              void* stream) {
  if (int rc = check_row_shape("layernorm_bwd", rows, H)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  {
    const int wps = (H / 8 + 31) / 32, tpr = 32 * wps, threads = lnb_threads(wps), slots = threads / tpr;     // wps in 1..4 (H <= 1024)
    int grid2 = (rows + slots * LNB_R - 1) / (slots * LNB_R);
    if (grid2 > sm_count() * lnb_ctas(wps)) grid2 = sm_count() * lnb_ctas(wps);
    const size_t smem2 = 3 * slots * H * sizeof(float) + 2 * slots * wps * 16;
    const __half* dyh = static_cast<const __half*>(dy);
    const __half* dy2h = static_cast<const __half*>(dy2);
    __half* dxh = static_cast<__half*>(dx);
    __half* dxd = static_cast<__half*>(dx_drop);
#define B200_LNB2(T, W)                                                                                                          \
  do {                                                                                                                           \
    static int c = set_smem(ln_bwd2_kernel<T, W>, 64 * 1024);                                                                    \
    if (c) return c;                                                                                                             \
    ln_bwd2_kernel<T, W><<<grid2, threads, smem2, s>>>(dyh, dy2h, static_cast<const T*>(x), mean, rstd, gamma, dxh, dgamma, \
                                                          dbeta, dbias, alpha, rows, H, dxd, drop);                              \
  } while (0)
    if (x_dtype == B200_DT_F32) {
      if (wps == 1) B200_LNB2(float, 1); else if (wps == 2) B200_LNB2(float, 2); else if (wps == 3) B200_LNB2(float, 3); else B200_LNB2(float, 4);
    } else {
      if (wps == 1) B200_LNB2(__half, 1); else if (wps == 2) B200_LNB2(__half, 2); else if (wps == 3) B200_LNB2(__half, 3); else B200_LNB2(__half, 4);
    }
#undef B200_LNB2
    return check_launch("ln_bwd2_kernel");
  }
}

int b200_embed_ln_fwd_drop(const int64_t* ids, const int64_t* tt, const int64_t* pos, const float* inputs_embeds, const float* word,
                           const float* pos_tab, const float* type_tab, const float* gamma, const float* beta, void* y, float* y32, int rows, int S,
                           int H, float eps, const uint32_t* seed, unsigned site, float p, void* stream);

int b200_embed_ln_fwd(const int64_t* ids, const int64_t* tt, const int64_t* pos, const float* inputs_embeds, const float* word,
                      const float* pos_tab, const float* type_tab, const float* gamma, const float* beta, void* y, float* y32, int rows, int S,
                      int H, float eps, void* stream) {
  return b200_embed_ln_fwd_drop(ids, tt, pos, inputs_embeds, word, pos_tab, type_tab, gamma, beta, y, y32, rows, S, H, eps, nullptr, 0, 0.f, stream);
}

int b200_embed_ln_fwd_drop(const int64_t* ids, const int64_t* tt, const int64_t* pos, const float* inputs_embeds, const float* word,
                           const float* pos_tab, const float* type_tab, const float* gamma, const float* beta, void* y, float* y32, int rows, int S,
                           int H, float eps, const uint32_t* seed, unsigned site, float p, void* stream) {
  if (int rc = check_row_shape("embed_ln_fwd", rows, H)) return rc;
  if (!ids && !inputs_embeds) return fail(B200_ERR_SHAPE, "embed_ln_fwd: need input_ids or inputs_embeds");
  const int grid = (rows + ROW_WARPS - 1) / ROW_WARPS;
  embed_ln_fwd_kernel<<<grid, ROW_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(ids, tt, pos, inputs_embeds, word, pos_tab, type_tab, gamma,
                                                                                      beta, static_cast<__half*>(y), y32, rows, S, H, eps, make_drop(seed, site, p));
  return check_launch("embed_ln_fwd_kernel");
}

int b200_embed_ln_bwd_drop(const void* dy, const void* dy2, const int64_t* ids, const int64_t* tt, const int64_t* pos, const float* word,
                           const float* pos_tab, const float* type_tab, const float* gamma, float* dword, float* dpos, float* dtype_tab,
                           float* dgamma, float* dbeta, const float* alpha, int rows, int S, int H, float eps, long long pad_id,
                           const float* inputs_embeds, float* d_inputs_embeds, const uint32_t* seed, unsigned site, float p, void* stream);

int b200_embed_ln_bwd(const void* dy, const void* dy2, const int64_t* ids, const int64_t* tt, const int64_t* pos, const float* word,
                      const float* pos_tab, const float* type_tab, const float* gamma, float* dword, float* dpos, float* dtype_tab,
                      float* dgamma, float* dbeta, const float* alpha, int rows, int S, int H, float eps, long long pad_id,
                      const float* inputs_embeds, float* d_inputs_embeds, void* stream) {
  return b200_embed_ln_bwd_drop(dy, dy2, ids, tt, pos, word, pos_tab, type_tab, gamma, dword, dpos, dtype_tab, dgamma, dbeta, alpha, rows, S, H,
                                eps, pad_id, inputs_embeds, d_inputs_embeds, nullptr, 0, 0.f, stream);
}

int b200_embed_ln_bwd_drop(const void* dy, const void* dy2, const int64_t* ids, const int64_t* tt, const int64_t* pos, const float* word,
                           const float* pos_tab, const float* type_tab, const float* gamma, float* dword, float* dpos, float* dtype_tab,
                           float* dgamma, float* dbeta, const float* alpha, int rows, int S, int H, float eps, long long pad_id,
                           const float* inputs_embeds, float* d_inputs_embeds, const uint32_t* seed, unsigned site, float p, void* stream) {
  if (int rc = check_row_shape("embed_ln_bwd", rows, H)) return rc;
  if (!ids && !inputs_embeds) return fail(B200_ERR_SHAPE, "embed_ln_bwd: need input_ids or inputs_embeds (whichever the forward used)");
  if (ids && !word) return fail(B200_ERR_SHAPE, "embed_ln_bwd: input_ids need the word table");
  if (!dpos || !dtype_tab || !dgamma || !dbeta) return fail(B200_ERR_SHAPE, "embed_ln_bwd: null gradient table");
  int grid = (rows + ROW_WARPS - 1) / ROW_WARPS;
  if (grid > sm_count() * 4) grid = sm_count() * 4;
  static int c = set_smem(embed_ln_bwd_kernel, 3 * ROW_WARPS * ROW_MAXV * 256 * 4);
  if (c) return c;
  embed_ln_bwd_kernel<<<grid, ROW_WARPS * 32, 3 * ROW_WARPS * H * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(dy), static_cast<const __half*>(dy2), ids, tt, pos, word, pos_tab, type_tab, gamma, dword, dpos, dtype_tab,
      dgamma, dbeta, alpha, rows, S, H, eps, make_drop(seed, site, p), pad_id, inputs_embeds, d_inputs_embeds);
  return check_launch("embed_ln_bwd_kernel");
}

int b200_cls_head_fwd_drop(const void* h, const float* W, const float* b, float* logits, int32_t* argmax, int rows, int H, int C,
                           const uint32_t* seed, unsigned site, float p, void* stream);

int b200_cls_head_fwd(const void* h, const float* W, const float* b, float* logits, int32_t* argmax, int rows, int H, int C, void* stream) {
  return b200_cls_head_fwd_drop(h, W, b, logits, argmax, rows, H, C, nullptr, 0, 0.f, stream);
}

int b200_cls_head_fwd_drop(const void* h, const float* W, const float* b, float* logits, int32_t* argmax, int rows, int H, int C,
                           const uint32_t* seed, unsigned site, float p, void* stream) {
  if (int rc = check_row_shape("cls_head_fwd", rows, H)) return rc;
  const DropCfg drop = make_drop(seed, site, p);
  int grid = (rows + CLS_WARPS * CLS_RPW - 1) / (CLS_WARPS * CLS_RPW);
  if (grid > sm_count()) grid = sm_count();      // one resident CTA per SM (~240 registers x 256 threads), contiguous rows per warp
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (C != 2 && C != 3) return fail(B200_ERR_SHAPE, "cls_head_fwd: C=%d (2 or 3)", C);
  const int nvec = (H + 255) / 256;
  const size_t wsm = static_cast<size_t>(C) * H * sizeof(float);
#define B200_CLS(CC, NV) cls_head_fwd_kernel<CC, __half, NV><<<grid, CLS_WARPS * 32, wsm, s>>>(static_cast<const __half*>(h), W, b, logits, argmax, rows, H, drop)
  if (C == 2) { if (nvec == 1) B200_CLS(2, 1); else if (nvec == 2) B200_CLS(2, 2); else if (nvec == 3) B200_CLS(2, 3); else B200_CLS(2, 4); }
  else { if (nvec == 1) B200_CLS(3, 1); else if (nvec == 2) B200_CLS(3, 2); else if (nvec == 3) B200_CLS(3, 3); else B200_CLS(3, 4); }
#undef B200_CLS
  return check_launch("cls_head_fwd_kernel");
}

int b200_ce_stats(const float* logits, const int64_t* labels, const float* class_weight, float* stats, int rows, int C, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int grid = (rows + 255) / 256;
  if (grid > 592) grid = 592;
  if (C == 2) ce_stats_kernel<2><<<grid, 256, 0, s>>>(logits, labels, class_weight, stats, rows);
  else if (C == 3) ce_stats_kernel<3><<<grid, 256, 0, s>>>(logits, labels, class_weight, stats, rows);
  else return fail(B200_ERR_SHAPE, "ce_stats: C=%d (2 or 3)", C);
  return check_launch("ce_stats_kernel");
}

int b200_cls_head_bwd_drop(const void* h, const float* logits, const int64_t* labels, const float* class_weight, const float* stats, const float* W,
                           const float* scale, void* dh, float* dW, float* db, int rows, int H, int C, const uint32_t* seed, unsigned site,
                           float p, void* stream);

int b200_cls_head_bwd(const void* h, const float* logits, const int64_t* labels, const float* class_weight, const float* stats, const float* W,
                      const float* scale, void* dh, float* dW, float* db, int rows, int H, int C, void* stream) {
  return b200_cls_head_bwd_drop(h, logits, labels, class_weight, stats, W, scale, dh, dW, db, rows, H, C, nullptr, 0, 0.f, stream);
}

int b200_cls_head_bwd_drop(const void* h, const float* logits, const int64_t* labels, const float* class_weight, const float* stats, const float* W,
                           const float* scale, void* dh, float* dW, float* db, int rows, int H, int C, const uint32_t* seed, unsigned site,
                           float p, void* stream) {
  if (int rc = check_row_shape("cls_head_bwd", rows, H)) return rc;
  const DropCfg drop = make_drop(seed, site, p);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int grid = (rows + ROW_WARPS - 1) / ROW_WARPS;
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  const size_t smem = static_cast<size_t>(ROW_WARPS) * C * H * sizeof(float);
  if (C == 2) {
    static int c = set_smem(cls_head_bwd_kernel<2>, ROW_WARPS * 2 * ROW_MAXV * 256 * 4);
    if (c) return c;
    cls_head_bwd_kernel<2><<<grid, ROW_WARPS * 32, smem, s>>>(static_cast<const __half*>(h), logits, labels, class_weight, stats, W, scale,
                                                             static_cast<__half*>(dh), dW, db, rows, H, drop);
  } else if (C == 3) {
    static int c = set_smem(cls_head_bwd_kernel<3>, ROW_WARPS * 3 * ROW_MAXV * 256 * 4);
    if (c) return c;
    cls_head_bwd_kernel<3><<<grid, ROW_WARPS * 32, smem, s>>>(static_cast<const __half*>(h), logits, labels, class_weight, stats, W, scale,
                                                             static_cast<__half*>(dh), dW, db, rows, H, drop);
  } else {
    return fail(B200_ERR_SHAPE, "cls_head_bwd: C=%d (2 or 3)", C);
  }
  return check_launch("cls_head_bwd_kernel");
}

int b200_colsum(const void* dy, int ld, float* db, const float* alpha, int rows, int cols, void* stream) {
  if ((cols % 8) || (ld % 8)) return fail(B200_ERR_SHAPE, "colsum: cols/ld must be multiples of 8");
  int slabs = (rows + 511) / 512;
  if (slabs < 1) slabs = 1;
  const int rpb = (rows + slabs - 1) / slabs;
  dim3 grid((cols + 63) / 64, slabs);
  colsum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(dy), ld, db, alpha, rows, cols, rpb);
  return check_launch("colsum_kernel");
}

int b200_cast_f32_to_f16(const float* src, void* dst, size_t n, void* stream) {
  if (n % 8) return fail(B200_ERR_SHAPE, "cast: n %% 8 != 0");
  cast_f32_f16_kernel<<<stream_grid(n / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, static_cast<__half*>(dst), n / 8);
  return check_launch("cast_f32_f16_kernel");
}
int b200_cast_f16_to_f32(const void* src, float* dst, size_t n, void* stream) {
  if (n % 8) return fail(B200_ERR_SHAPE, "cast: n %% 8 != 0");
  cast_f16_f32_kernel<<<stream_grid(n / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(src), dst, n / 8);
  return check_launch("cast_f16_f32_kernel");
}
int b200_unscale_cast_grad(const void* src, float* dst, size_t n, const float* scale, void* stream) {
  if (n % 8) return fail(B200_ERR_SHAPE, "unscale_cast: n %% 8 != 0");
  unscale_cast_f16_f32_kernel<<<stream_grid(n / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(src), scale, dst, n / 8);
  return check_launch("unscale_cast_f16_f32_kernel");
}
int b200_scale_cast_grad(const float* src, void* dst, size_t n, float target, float* scale, void* amax_slot, void* stream) {
  if (n % 8) return fail(B200_ERR_SHAPE, "scale_cast: n %% 8 != 0");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaMemsetAsync(amax_slot, 0, 4, s);
  amax_f32_kernel<<<stream_grid(n / 8), 256, 0, s>>>(src, n, static_cast<unsigned int*>(amax_slot));
  if (int rc = check_launch("amax_f32_kernel")) return rc;
  pick_scale_kernel<<<1, 1, 0, s>>>(static_cast<const unsigned int*>(amax_slot), target, scale);
  if (int rc = check_launch("pick_scale_kernel")) return rc;
  scale_cast_f32_f16_kernel<<<stream_grid(n / 8), 256, 0, s>>>(src, scale, static_cast<__half*>(dst), n / 8);
  return check_launch("scale_cast_f32_f16_kernel");
}

}  // extern "C"

namespace {
__global__ void fill_f32_kernel(float* p, float v, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) p[i] = v;
}
}  // namespace

extern "C" {

size_t b200_ponet_workspace(int B, int S, int H, int heads, int nseg) {
  const size_t nchunks = (S + 127) / 128;
  return (static_cast<size_t>(B) * H * 2 + B + static_cast<size_t>(B) * heads * nchunks * 66 + static_cast<size_t>(B) * nseg * H) * sizeof(float) + 256;
}

int b200_ponet_mix_fwd(const void* proj, int ld, const float* key_bias, const int64_t* segment_ids, void* workspace, void* out, int B, int S,
                       int H, int heads, int nseg, void* stream) {
  if (B <= 0 || S <= 0 || H != heads * 64 || H > PONET_MAX_H || (ld % 8) || nseg <= 0) return fail(B200_ERR_SHAPE, "ponet_mix_fwd: bad shape");
  if (!proj || !segment_ids || !workspace || !out) return fail(B200_ERR_SHAPE, "ponet_mix_fwd: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nchunks = (S + 127) / 128;
  float* qsum = static_cast<float*>(workspace);
  float* cnt = qsum + static_cast<size_t>(B) * H;
  float* g = cnt + ((B + 3) / 4) * 4;
  float* part = g + static_cast<size_t>(B) * H;
  float* segmax = part + static_cast<size_t>(B) * heads * nchunks * 66;
  segmax += (4 - (reinterpret_cast<uintptr_t>(segmax) / 4) % 4) % 4;      // 16-byte alignment for 128-bit loads
  cudaMemsetAsync(qsum, 0, (static_cast<size_t>(B) * H + ((B + 3) / 4) * 4) * sizeof(float), s);
  const size_t nsm = static_cast<size_t>(B) * nseg * H;
  fill_f32_kernel<<<static_cast<int>((nsm + 255) / 256 < 2368 ? (nsm + 255) / 256 : 2368), 256, 0, s>>>(segmax, -INFINITY, nsm);
  int rc;
  if ((rc = check_launch("fill_f32_kernel"))) return rc;
  const __half* p = static_cast<const __half*>(proj);
  constexpr int qsum_pos = PONET_GROUPS * PONET_QSUM_ROUNDS * PONET_QSUM_POS;
  ponet_qsum_kernel<<<dim3((S + qsum_pos - 1) / qsum_pos, B), dim3(H / 8, PONET_GROUPS), 0, s>>>(p, ld, key_bias, qsum, cnt, S, H);
  if ((rc = check_launch("ponet_qsum_kernel"))) return rc;
  ponet_global_part_kernel<<<dim3(nchunks, heads, B), 128, 0, s>>>(p, ld, key_bias, qsum, cnt, part, S, H, heads);
  if ((rc = check_launch("ponet_global_part_kernel"))) return rc;
  ponet_global_comb_kernel<<<dim3(heads, B), 64, 0, s>>>(part, g, nchunks, H, heads);
  if ((rc = check_launch("ponet_global_comb_kernel"))) return rc;
  ponet_segmax_kernel<<<dim3((S + PONET_RUN_POS - 1) / PONET_RUN_POS, B), H / 8, 0, s>>>(p, ld, key_bias, segment_ids, segmax, S, H, nseg);
  if ((rc = check_launch("ponet_segmax_kernel"))) return rc;
  ponet_mix_kernel<<<(B * S + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, s>>>(p, ld, key_bias, segment_ids, g, segmax,
                                                                                 static_cast<__half*>(out), B, S, H, nseg);
  return check_launch("ponet_mix_kernel");
}

size_t b200_ponet_bwd_workspace(int B, int S, int H, int heads, int nseg) {
  return (static_cast<size_t>(B) * H * 2 + static_cast<size_t>(B) * heads + 2 * static_cast<size_t>(B) * nseg * H) * sizeof(float) + 256;
}

int b200_ponet_mix_bwd(const void* proj, int ld, const void* dout, const float* key_bias, const int64_t* segment_ids, const void* fwd_workspace,
                       void* bwd_workspace, void* dproj, int ld_d, int B, int S, int H, int heads, int nseg, void* stream) {
  if (B <= 0 || S <= 0 || H != heads * 64 || H > 1024 || (ld % 8) || (ld_d % 8) || nseg <= 0) return fail(B200_ERR_SHAPE, "ponet_mix_bwd: bad shape");
  if (!proj || !dout || !segment_ids || !fwd_workspace || !bwd_workspace || !dproj) return fail(B200_ERR_SHAPE, "ponet_mix_bwd: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nchunks = (S + 127) / 128;
  // forward workspace layout (see b200_ponet_mix_fwd)
  const float* qsum = static_cast<const float*>(fwd_workspace);
  const float* cnt = qsum + static_cast<size_t>(B) * H;
  const float* g = cnt + ((B + 3) / 4) * 4;
  const float* part = g + static_cast<size_t>(B) * H;
  const float* segmax = part + static_cast<size_t>(B) * heads * nchunks * 66;
  segmax += (4 - (reinterpret_cast<uintptr_t>(segmax) / 4) % 4) % 4;
  float* dg = static_cast<float*>(bwd_workspace);
  float* dqbar = dg + static_cast<size_t>(B) * H;
  float* segsum = dqbar + static_cast<size_t>(B) * H;
  float* segties = segsum + static_cast<size_t>(B) * nseg * H;
  float* lse = segties + static_cast<size_t>(B) * nseg * H;
  cudaMemsetAsync(dg, 0, (static_cast<size_t>(B) * H * 2 + 2 * static_cast<size_t>(B) * nseg * H) * sizeof(float), s);
  const __half* p = static_cast<const __half*>(proj);
  const __half* d = static_cast<const __half*>(dout);
  int rc;
  ponet_bwd_sums_kernel<<<dim3((S + PONET_GROUPS * PONET_RUN_POS - 1) / (PONET_GROUPS * PONET_RUN_POS), B), dim3(H / 8, PONET_GROUPS), 0, s>>>(p, ld, d, key_bias, segment_ids, segmax, dg, segsum, segties, S, H, nseg);
  if ((rc = check_launch("ponet_bwd_sums_kernel"))) return rc;
  ponet_global_lse_kernel<<<dim3(heads, B), 32, 0, s>>>(part, lse, nchunks, heads);
  if ((rc = check_launch("ponet_global_lse_kernel"))) return rc;
  ponet_bwd_global_kernel<<<dim3(nchunks, heads, B), 128, 0, s>>>(p, ld, key_bias, qsum, cnt, g, lse, dg, dqbar, static_cast<__half*>(dproj), ld_d,
                                                                  S, H, heads);
  if ((rc = check_launch("ponet_bwd_global_kernel"))) return rc;
  ponet_bwd_rows_kernel<<<dim3((S + PONET_ROWS_POS - 1) / PONET_ROWS_POS, B), H / 8, 0, s>>>(p, ld, d, key_bias, segment_ids, g, segmax, segsum, segties, dqbar, cnt,
                                                                                      static_cast<__half*>(dproj), ld_d, B, S, H, nseg);
  return check_launch("ponet_bwd_rows_kernel");
}

}  // extern "C"

#include "api_train.inc"
#include "api_heads.inc"
