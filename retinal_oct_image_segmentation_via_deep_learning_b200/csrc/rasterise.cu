// Boundary rows -> label maps (annotation formats of the catalogued layer datasets store
// boundary curves, reference Datasets.md:3-26; the metrics of Metrics/*.py take masks).
//
//   labels[i][y][x] = #{ k : b_k(i, x) <= y }            (0 .. num_boundaries)
//
// the inverse of the label pass's boundary rows (#{label < k} per column) on layered maps.  Boundaries may be
// int32 or float32 rows, in any order, outside [0, H] or NaN (NaN = boundary absent: counted for no row).
// One CTA per (item, strip of columns): the strip's thresholds ceil(b_k) are sorted per column in shared
// memory, then every thread runs down its columns with one compare per pixel (the next threshold sits in a
// register) and stores whole rows of the strip: the kernel is bound by the label bytes it writes.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.cuh"

namespace octm {

constexpr int kRasterThreads = 128;
constexpr int kRasterMaxB = 15;                     // num_classes <= 16

struct RasterParams {
    const void* bnd;        // [n][nb][W] int32 or float32
    int is_float;
    long long n_items;
    int nb, H, W;
    uint8_t* labels;        // [n][H][W]
};

__device__ __forceinline__ int raster_threshold(const RasterParams& p, long long idx) {
    if (!p.is_float) return static_cast<const int*>(p.bnd)[idx];
    const float b = static_cast<const float*>(p.bnd)[idx];
    if (!(b == b)) return 0x7fffffff;               // NaN: absent
    const float c = ceilf(b);                       // y >= b  <=>  y >= ceil(b) for integer rows y
    return c >= 2147483000.f ? 0x7fffffff : (c <= -2147483000.f ? -0x7fffffff : static_cast<int>(c));
}

// V = 4: a thread owns 4 adjacent columns and stores one 32-bit word per row; V = 1: one column, byte stores.
template <int V>
__global__ void __launch_bounds__(kRasterThreads) rasterise_kernel(const RasterParams prm) {
    extern __shared__ int s_thr[];                  // [nb][strip] sorted ascending per column
    constexpr int strip = kRasterThreads * V;
    const int nb = prm.nb, H = prm.H, W = prm.W;
    const int strips = (W + strip - 1) / strip;
    const long long total = prm.n_items * strips;
    for (long long job = blockIdx.x; job < total; job += gridDim.x) {
        const long long item = job / strips;
        const int x0 = static_cast<int>(job % strips) * strip;
        __syncthreads();
        for (int i = threadIdx.x; i < nb * strip; i += kRasterThreads) {
            const int k = i / strip, c = i % strip;
            s_thr[i] = x0 + c < W ? raster_threshold(prm, (item * nb + k) * static_cast<long long>(W) + x0 + c) : 0x7fffffff;
        }
        __syncthreads();
        int nxt[V], ptr[V];
        uint32_t lab = 0;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int c = threadIdx.x * V + v;
            for (int a = 1; a < nb; ++a) {           // insertion sort of this column's thresholds
                const int key = s_thr[a * strip + c];
                int b = a - 1;
                while (b >= 0 && s_thr[b * strip + c] > key) {
                    s_thr[(b + 1) * strip + c] = s_thr[b * strip + c];
                    --b;
                }
                s_thr[(b + 1) * strip + c] = key;
            }
            ptr[v] = 0;
            nxt[v] = nb > 0 ? s_thr[c] : 0x7fffffff;
        }
        const int xc = x0 + threadIdx.x * V;
        uint8_t* out = prm.labels + item * H * static_cast<long long>(W) + xc;
        for (int y = 0; y < H; ++y, out += W) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                while (y >= nxt[v]) {                // rare: a boundary is crossed
                    lab += 1u << (8 * v);
                    ++ptr[v];
                    nxt[v] = ptr[v] < nb ? s_thr[ptr[v] * strip + threadIdx.x * V + v] : 0x7fffffff;
                }
            }
            if (V == 4) {
                if (xc < W) *reinterpret_cast<uint32_t*>(out) = lab;
            } else {
                if (xc < W) *out = static_cast<uint8_t>(lab);
            }
        }
    }
}

}  // namespace octm

extern "C" int octm_labels_from_boundaries(const void* boundaries, int dtype, int64_t n_items, int num_boundaries, int H,
                                           int W, uint8_t* labels, void* stream) {
    if (n_items < 0 || H < 1 || W < 1) return octm::fail(OCTM_ERR_INVALID, "bad shape");
    if (num_boundaries < 0 || num_boundaries > octm::kRasterMaxB)
        return octm::fail(OCTM_ERR_INVALID, "num_boundaries %d outside [0, 15]", num_boundaries);
    if (dtype != OCTM_DTYPE_I32 && dtype != OCTM_DTYPE_F32)
        return octm::fail(OCTM_ERR_UNSUPPORTED, "boundaries must be int32 (OCTM_DTYPE_I32) or float32 (OCTM_DTYPE_F32)");
    if (n_items == 0) return OCTM_OK;
    if (!labels || (num_boundaries > 0 && !boundaries)) return octm::fail(OCTM_ERR_INVALID, "null pointer");
    octm::RasterParams p{boundaries, dtype == OCTM_DTYPE_F32 ? 1 : 0, n_items, num_boundaries, H, W, labels};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool wide = W % 4 == 0 && reinterpret_cast<uintptr_t>(labels) % 4 == 0;
    const int strip = octm::kRasterThreads * (wide ? 4 : 1);
    const long long jobs = n_items * ((W + strip - 1) / strip);
    long long grid = jobs;
    const long long cap = static_cast<long long>(octm::sm_count()) * 16;
    if (grid > cap) grid = cap;
    const size_t smem = static_cast<size_t>(num_boundaries > 0 ? num_boundaries : 1) * strip * sizeof(int);
    if (wide) OCTM_TIMED("rasterise_kernel", st) octm::rasterise_kernel<4><<<static_cast<unsigned>(grid), octm::kRasterThreads, smem, st>>>(p);
    else OCTM_TIMED("rasterise_kernel", st) octm::rasterise_kernel<1><<<static_cast<unsigned>(grid), octm::kRasterThreads, smem, st>>>(p);
    return octm::check_launch("rasterise_kernel");
}
