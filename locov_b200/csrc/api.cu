// api.cu — library-level pieces of the C ABI: version, thread-local error string, device checks and
// the TMA tensor-map factory (driver entry point resolved at run time, so the .so links against
// libcudart only and loads on a machine without a GPU for the symbol/ABI tests).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <mutex>

#include "common.cuh"

namespace loco {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode_fn() {
    static std::once_flag once;
    static encode_tiled_fn fn = nullptr;
    std::call_once(once, []() {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_tiled_fn>(p);
    });
    return fn;
}

static int make_tmap_2d(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows,
                        CUtensorMapDataType dtype, uint32_t elem_bytes) {
    encode_tiled_fn fn = get_encode_fn();
    LOCO_REQUIRE(fn != nullptr, LOCO_E_DRIVER, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    LOCO_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, LOCO_E_ALIGN, "tensor-core operand base %p is not 16-byte aligned", base);
    LOCO_REQUIRE((ld_elems * elem_bytes) % 16 == 0, LOCO_E_ALIGN, "tensor-core operand row stride (%llu elements of %u bytes) is not a multiple of 16 bytes",
                 (unsigned long long)ld_elems, elem_bytes);
    LOCO_REQUIRE(box_rows >= 1 && box_rows <= 256, LOCO_E_BADARG, "TMA box rows %u out of range", box_rows);
    LOCO_REQUIRE(rows >= 1 && cols >= 1, LOCO_E_BADARG, "empty TMA tensor");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_elems * elem_bytes};   // bytes, dimension 1
    cuuint32_t box[2] = {128u / elem_bytes, box_rows};  // inner box = 128 B = one swizzle span
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dtype, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LOCO_REQUIRE(r == CUDA_SUCCESS, LOCO_E_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu box_rows=%u)",
                 (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows);
    return LOCO_OK;
}

int make_tmap_bf16_2d(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows) {
    return make_tmap_2d(map, base, rows, cols, ld_elems, box_rows, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
}

int make_tmap_f32_2d(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                     uint32_t box_rows) {
    return make_tmap_2d(map, base, rows, cols, ld_elems, box_rows, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4);
}

// bf16 row-major OUTPUT of a GEMM epilogue as 32 x 32 element boxes (64-byte rows) with the 64-byte swizzle: the epilogue warp owning
// 32 accumulator rows stages a 32-column block in shared memory (lane = row, 16-byte chunk c of row r at chunk position
// c ^ ((r >> 1) & 3): the 32 lanes of a store instruction cover all 32 banks) and one thread hands it to the copy engine.
int make_tmap_bf16_out(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t ld_elems) {
    encode_tiled_fn fn = get_encode_fn();
    LOCO_REQUIRE(fn != nullptr, LOCO_E_DRIVER, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    LOCO_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld_elems * 2) % 16 == 0 && rows >= 1 && cols >= 1, LOCO_E_ALIGN,
                 "bf16 GEMM output is not addressable by a TMA store (base %p, ld %llu)", base, (unsigned long long)ld_elems);
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_elems * 2};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LOCO_REQUIRE(r == CUDA_SUCCESS, LOCO_E_DRIVER, "cuTensorMapEncodeTiled (bf16 output) failed with CUresult %d (rows=%llu cols=%llu ld=%llu)", (int)r,
                 (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems);
    return LOCO_OK;
}

// Developer probe: with LOCOV_B200_TIMELINE=1 every tensor-core kernel writes up to 16 globaltimer stamps per CTA (entry, setup
// done, first stage landed, last MMA issued, last accumulator complete, epilogue done, exit) into this buffer; the most
// recent launch overwrites it.  Read back with loco_debug_timeline_read().
static unsigned long long *g_timeline = nullptr;
static int g_timeline_ctas = 0;
constexpr int kTimelineMaxCtas = 8192;
unsigned long long *debug_timeline_buffer(int ctas) {
    static int enabled = -1;
    if (enabled < 0) {
        const char *e = getenv("LOCOV_B200_TIMELINE");
        enabled = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    if (!enabled || ctas > kTimelineMaxCtas) return nullptr;
    if (g_timeline == nullptr && cudaMalloc(&g_timeline, (size_t)kTimelineMaxCtas * 16 * sizeof(unsigned long long)) != cudaSuccess) return nullptr;
    g_timeline_ctas = ctas;
    return g_timeline;
}

int make_tmap_roi_out(CUtensorMap *map, const float *out, uint64_t R, uint64_t C, uint64_t PHW) {
    encode_tiled_fn fn = get_encode_fn();
    LOCO_REQUIRE(fn != nullptr, LOCO_E_DRIVER, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    LOCO_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && C % 4 == 0 && PHW >= 1 && PHW <= 256 && (PHW * 4) % 16 == 0 && R >= 1, LOCO_E_ALIGN,
                 "pooled output is not addressable by a TMA store (base %p, C=%llu, PH*PW=%llu)", (const void *)out, (unsigned long long)C,
                 (unsigned long long)PHW);
    cuuint64_t dims[4] = {PHW, 4, C / 4, R};
    cuuint64_t strides[3] = {PHW * 4, 4 * PHW * 4, C * PHW * 4};        // bytes, dimensions 1..3
    cuuint32_t box[4] = {(cuuint32_t)PHW, 1, 32, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(out), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LOCO_REQUIRE(r == CUDA_SUCCESS, LOCO_E_DRIVER, "cuTensorMapEncodeTiled (pooled output) failed with CUresult %d (R=%llu C=%llu PHW=%llu)", (int)r,
                 (unsigned long long)R, (unsigned long long)C, (unsigned long long)PHW);
    return LOCO_OK;
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("LOCOV_B200_PDL");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

int current_device_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace loco

extern "C" {

long long loco_launch_count(void) { return loco::g_launches.load(std::memory_order_relaxed); }

int loco_version(void) { return 10000 * 0 + 100 * 1 + 0; }

const char *loco_last_error(void) { return loco::g_err; }

int loco_device_check(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        loco::set_error("no CUDA device visible (%s)", cudaGetErrorString(e));
        return LOCO_E_DEVICE;
    }
    LOCO_REQUIRE(device >= 0 && device < n, LOCO_E_DEVICE, "device ordinal %d out of range (%d devices)", device, n);
    int major = 0, minor = 0;
    LOCO_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    LOCO_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    LOCO_REQUIRE(major == 10, LOCO_E_DEVICE, "device %d is compute capability %d.%d; liblocov_b200 is built for sm_100a only",
                 device, major, minor);
    return LOCO_OK;
}

int loco_debug_timeline_read(unsigned long long *host, int max_ctas) {
    if (loco::g_timeline == nullptr || host == nullptr) return 0;
    const int n = loco::g_timeline_ctas < max_ctas ? loco::g_timeline_ctas : max_ctas;
    if (cudaMemcpy(host, loco::g_timeline, (size_t)n * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return n;
}

int loco_sm_count(int device) {
    int n = 0;
    cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return -loco::cuda_fail(e, "cudaDeviceGetAttribute(MultiProcessorCount)");
    return n;
}

}  // extern "C"
