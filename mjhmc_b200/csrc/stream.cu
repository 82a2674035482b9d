// Host side of the streaming separable-energy kernels: TMA tensor maps and dispatch.
#include <cudaTypedefs.h>
#include <mutex>
#include "stream_separable.cuh"
#include "stream.h"

namespace mjhmc {

#define MJ_STREAM_UNITS(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6)
#define MJ_DECL(g) stream_launch_fn find_stream_f64_g##g(int, int); stream_launch_fn find_stream_f32_g##g(int, int);
MJ_STREAM_UNITS(MJ_DECL)

static stream_launch_fn find_stream(int dtype, int kind, int DT) {
    stream_launch_fn f = nullptr;
#define MJ_TRY64(g) if (!f) f = find_stream_f64_g##g(kind, DT);
#define MJ_TRY32(g) if (!f) f = find_stream_f32_g##g(kind, DT);
    if (dtype == MJHMC_F64) { MJ_STREAM_UNITS(MJ_TRY64) } else { MJ_STREAM_UNITS(MJ_TRY32) }
    return f;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link dependency on libcuda)
static PFN_cuTensorMapEncodeTiled encode_fn() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled)f;
    });
    return fn;
}

static bool make_map(CUtensorMap* m, const void* base, int dtype, long long n, long long ld, int d, int P, int rows) {
    PFN_cuTensorMapEncodeTiled enc = encode_fn();
    if (!enc) return false;
    const size_t S = dtype == MJHMC_F64 ? 8 : 4;
    if (((uintptr_t)base & 15u) || ((size_t)ld * S) % 16u || n <= 0 || n > 0x7fffffffLL || rows > 256 || rows < d)
        return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)d};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * S};
    // the box is taller than the array when ndims is not a multiple of the dims per thread: rows >= ndims (and
    // columns >= n in the last tile) are out of range and arrive as zeros
    const cuuint32_t box[2] = {(cuuint32_t)P, (cuuint32_t)rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(m, dtype == MJHMC_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                           2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

static int g_force_no_tma = 0;

// fp32 array of `rank` dims (dim 0 contiguous), strides in elements for dims 1.., box extents per dim; false when the
// array cannot be described (alignment, sizes) or TMA is switched off for the tests
bool make_tensor_map_f32(CUtensorMap* m, const void* base, int rank, const long long* dims, const long long* strides,
                         const int* box) {
    memset(m, 0, sizeof *m);
    PFN_cuTensorMapEncodeTiled enc = encode_fn();
    if (!enc || g_force_no_tma || rank < 2 || rank > 3 || ((uintptr_t)base & 15u)) return false;
    cuuint64_t gdim[3];
    cuuint64_t gstride[2];
    cuuint32_t bx[3], estr[3] = {1, 1, 1};
    for (int k = 0; k < rank; ++k) {
        if (dims[k] <= 0 || dims[k] > 0x7fffffffLL || box[k] < 1 || box[k] > 256) return false;
        gdim[k] = (cuuint64_t)dims[k];
        bx[k] = (cuuint32_t)box[k];
        if (k) {
            if (strides[k - 1] <= 0 || (strides[k - 1] * 4) % 16) return false;
            gstride[k - 1] = (cuuint64_t)strides[k - 1] * 4u;
        }
    }
    if ((box[0] * 4) % 16) return false;
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstride, bx, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t stream_make_maps(const LaunchParams& p, int dtype, int P, int rows, CUtensorMap* mx, CUtensorMap* mv,
                             int* use_tma) {
    memset(mx, 0, sizeof *mx);
    memset(mv, 0, sizeof *mv);
    *use_tma = !g_force_no_tma && make_map(mx, p.Xin, dtype, p.n, p.ld, p.d, P, rows) &&
               make_map(mv, p.Vin, dtype, p.n, p.ld, p.d, P, rows);
    return cudaSuccess;
}

static thread_local long long g_last[7] = {0, 0, 0, 0, 0, 0, 0};
void stream_note_launch(int use_tma, int stages, long long grid, int per_sm, int G, int DT, size_t smem) {
    g_last[0] = use_tma; g_last[1] = stages; g_last[2] = grid; g_last[3] = per_sm; g_last[4] = G; g_last[5] = DT;
    g_last[6] = (long long)smem;
}
void stream_last_launch(long long* out7) { for (int k = 0; k < 7; ++k) out7[k] = g_last[k]; }

// Occupancy probe (developer tool): resident 256-thread CTAs of a trivial kernel with `smem` dynamic bytes.
__global__ void __launch_bounds__(256) stream_probe_kernel(int* out) {
    extern __shared__ int probe_smem[];
    if (out) out[threadIdx.x] = probe_smem[threadIdx.x];
}
int stream_probe_blocks(long long smem) {
    int nb = 0;
    if (cudaFuncSetAttribute(stream_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, stream_probe_kernel, 256, (size_t)smem) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return nb;
}

int stream_sm_count() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
    return sms;
}

bool stream_supported(int dtype, int kind, int ndims) {
    const StreamPlan pl = stream_plan(ndims);
    return pl.DT && find_stream(dtype, kind, pl.DT) != nullptr;
}

cudaError_t launch_stream_kernel(int dtype, int kind, const LaunchParams& p, cudaStream_t stream) {
    const StreamPlan pl = stream_plan(p.d);
    stream_launch_fn fn = pl.DT ? find_stream(dtype, kind, pl.DT) : nullptr;
    if (!fn) return cudaErrorInvalidValue;
    return fn(p, pl, dtype, stream);
}

void stream_set_tma(int enabled) { g_force_no_tma = !enabled; }

}  // namespace mjhmc
