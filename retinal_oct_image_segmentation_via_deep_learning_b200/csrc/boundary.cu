// Boundary-position error on CONTINUOUS (soft) layer positions, and topology violations.
//
// The only producer of boundary positions in the reference is LayerEngine.get_layer_positions
// (SOTAS/Layers_Segment/SD_Layer_Net/layer_engine.py:46-47): soft-argmax rows (B, K-1, W) of floats.
// Scoring them against ground-truth positions applies the reference's array metrics to float rows:
//   mean_squared_error / root_mean_squared_error   Metrics/PixelError_based_metrics.py:14-17, 32-35
//   mad                                            Metrics/Contour_based_metrics.py:68-71
// i.e. mean((a.astype(float) - b.astype(float)) ** 2) and mean(|a - b|) in float64.  The kernel forms
// the two float64 sums per row in a fixed order (numpy's pairwise order differs: <= 1e-13 relative).
// get_topology_violations (layer_engine.py:74-76) = relu(pos[k] - pos[k+1]): per adjacent pair of
// boundaries the kernel returns the summed violation and the number of violating columns.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace octm {

__device__ __forceinline__ double ld_f64(const float* p, long long i) { return static_cast<double>(p[i]); }
__device__ __forceinline__ double ld_f64(const double* p, long long i) { return p[i]; }
__device__ __forceinline__ double ld_f64(const __half* p, long long i) { return static_cast<double>(__half2float(p[i])); }
__device__ __forceinline__ double ld_f64(const __nv_bfloat16* p, long long i) { return static_cast<double>(__bfloat162float(p[i])); }

constexpr int kRowThreads = 128;

__device__ __forceinline__ double block_sum_fixed(double v, double* s_part) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < kRowThreads / 32; ++w) r = __dadd_rn(r, s_part[w]);
    __syncthreads();
    return r;
}

// one CTA per row of W positions: sum (a-b)^2 and sum |a-b| in float64
template <class T>
__global__ void __launch_bounds__(kRowThreads) boundary_error_float_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                                           long long n_rows, long long W, double* sum_sq,
                                                                           double* sum_abs) {
    __shared__ double s_part[kRowThreads / 32];
    for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
        const T* ra = a + row * W;
        const T* rb = b + row * W;
        double sq = 0.0, ab = 0.0;
        for (long long x = threadIdx.x; x < W; x += kRowThreads) {
            const double d = __dsub_rn(ld_f64(ra, x), ld_f64(rb, x));
            sq = __dadd_rn(sq, __dmul_rn(d, d));
            ab = __dadd_rn(ab, fabs(d));
        }
        sq = block_sum_fixed(sq, s_part);
        ab = block_sum_fixed(ab, s_part);
        if (threadIdx.x == 0) {
            sum_sq[row] = sq;
            sum_abs[row] = ab;
        }
    }
}

// one CTA per (item, adjacent boundary pair): relu(pos[k] - pos[k+1]) summed, and the count of columns > 0
template <class T>
__global__ void __launch_bounds__(kRowThreads) topology_kernel(const T* __restrict__ pos, long long n_items, int Kb, long long W,
                                                               double* sum_viol, unsigned int* n_viol) {
    __shared__ double s_part[kRowThreads / 32];
    __shared__ unsigned int s_cnt;
    const long long pairs = n_items * (Kb - 1);
    for (long long pr = blockIdx.x; pr < pairs; pr += gridDim.x) {
        const long long item = pr / (Kb - 1), k = pr - item * (Kb - 1);
        const T* up = pos + (item * Kb + k) * W;
        const T* dn = up + W;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        double s = 0.0;
        unsigned int c = 0;
        for (long long x = threadIdx.x; x < W; x += kRowThreads) {
            const double d = __dsub_rn(ld_f64(up, x), ld_f64(dn, x));
            if (d > 0.0) { s = __dadd_rn(s, d); ++c; }
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, c);
        s = block_sum_fixed(s, s_part);
        if (threadIdx.x == 0) {
            sum_viol[pr] = s;
            n_viol[pr] = s_cnt;
        }
        __syncthreads();
    }
}

template <class T>
static int launch_boundary_float(const void* a, const void* b, long long rows, long long W, double* sq, double* ab, cudaStream_t st) {
    long long grid = rows < 148ll * 16 ? rows : 148ll * 16;
    OCTM_TIMED("boundary_error_float_kernel", st) boundary_error_float_kernel<T><<<static_cast<unsigned>(grid), kRowThreads, 0, st>>>(static_cast<const T*>(a), static_cast<const T*>(b),
                                                                                      rows, W, sq, ab);
    return check_launch("boundary_error_float_kernel");
}

template <class T>
static int launch_topology(const void* pos, long long n, int Kb, long long W, double* sv, unsigned int* nv, cudaStream_t st) {
    const long long pairs = n * (Kb - 1);
    long long grid = pairs < 148ll * 16 ? pairs : 148ll * 16;
    OCTM_TIMED("topology_kernel", st) topology_kernel<T><<<static_cast<unsigned>(grid), kRowThreads, 0, st>>>(static_cast<const T*>(pos), n, Kb, W, sv, nv);
    return check_launch("topology_kernel");
}

}  // namespace octm

extern "C" int octm_boundary_error_float(const void* bnd_true, const void* bnd_pred, int dtype, int64_t n_items,
                                         int num_boundaries, int64_t W, double* sum_sq, double* sum_abs, void* stream) {
    if (n_items < 0 || num_boundaries < 1 || W < 1) return octm::fail(OCTM_ERR_INVALID, "bad shape");
    if (n_items == 0) return OCTM_OK;
    if (!bnd_true || !bnd_pred || !sum_sq || !sum_abs) return octm::fail(OCTM_ERR_INVALID, "null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long rows = n_items * num_boundaries;
    switch (dtype) {
        case OCTM_DTYPE_F32: return octm::launch_boundary_float<float>(bnd_true, bnd_pred, rows, W, sum_sq, sum_abs, st);
        case OCTM_DTYPE_F64: return octm::launch_boundary_float<double>(bnd_true, bnd_pred, rows, W, sum_sq, sum_abs, st);
        case OCTM_DTYPE_F16: return octm::launch_boundary_float<__half>(bnd_true, bnd_pred, rows, W, sum_sq, sum_abs, st);
        case OCTM_DTYPE_BF16: return octm::launch_boundary_float<__nv_bfloat16>(bnd_true, bnd_pred, rows, W, sum_sq, sum_abs, st);
        default: return octm::fail(OCTM_ERR_INVALID, "dtype %d: expected OCTM_DTYPE_F32 / F16 / BF16 / F64", dtype);
    }
}

extern "C" int octm_topology_violations_float(const void* positions, int dtype, int64_t n_items, int num_boundaries, int64_t W,
                                              double* sum_violation, uint32_t* n_violations, void* stream) {
    if (n_items < 0 || num_boundaries < 2 || W < 1) return octm::fail(OCTM_ERR_INVALID, "bad shape (need >= 2 boundaries)");
    if (n_items == 0) return OCTM_OK;
    if (!positions || !sum_violation || !n_violations) return octm::fail(OCTM_ERR_INVALID, "null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (dtype) {
        case OCTM_DTYPE_F32: return octm::launch_topology<float>(positions, n_items, num_boundaries, W, sum_violation, n_violations, st);
        case OCTM_DTYPE_F64: return octm::launch_topology<double>(positions, n_items, num_boundaries, W, sum_violation, n_violations, st);
        case OCTM_DTYPE_F16: return octm::launch_topology<__half>(positions, n_items, num_boundaries, W, sum_violation, n_violations, st);
        case OCTM_DTYPE_BF16: return octm::launch_topology<__nv_bfloat16>(positions, n_items, num_boundaries, W, sum_violation, n_violations, st);
        default: return octm::fail(OCTM_ERR_INVALID, "dtype %d: expected OCTM_DTYPE_F32 / F16 / BF16 / F64", dtype);
    }
}
