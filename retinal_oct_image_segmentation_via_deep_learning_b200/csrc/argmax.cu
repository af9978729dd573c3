// Front end of the metric path: model scores -> uint8 label map (argmax over the class dimension).
//
// Every model of the reference returns (B, num_classes, H, W) scores and a user turns them into the
// label map the metrics take with an argmax over dim 1 (e.g.
// SOTAS/Lesions_Segment/ReLayNet_2017.py:106-108, RetiFluidNet_pytorch_2022.py:130-134).  This step
// moves K x (2..4) bytes per pixel where the whole metric suite moves 2, so it is the HBM-bound part
// of an evaluation that starts from scores.  The labels are written once (1 B per pixel): the
// contour walk needs them in memory anyway.
//
// Semantics = numpy / torch argmax: the FIRST maximal class wins ties; a NaN counts as maximal.
//
//   argmax_planes_kernel   class-major scores [n][K][P] (NCHW): a thread owns 32 bytes of pixels and
//                          streams the K planes with independent 256-bit loads (L1-bypassing, read
//                          once), keeping a running (max, argmax) per pixel.
//   argmax_generic_kernel  any K / alignment, and channels-last scores [n][P][K]: one pixel per thread.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace octm {

template <class T> struct Vec;          // VPT = scores per 32-byte vector (sm_100 has 256-bit global loads)
template <> struct Vec<float> { static constexpr int VPT = 8; };
template <> struct Vec<__half> { static constexpr int VPT = 16; };
template <> struct Vec<__nv_bfloat16> { static constexpr int VPT = 16; };

struct alignas(32) Raw256 { uint32_t w[8]; };

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

// streaming 256-bit load: the scores are read exactly once (no L1 allocation, first out of L2)
__device__ __forceinline__ Raw256 ld_stream(const void* p) {
    Raw256 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]), "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]), "=r"(v.w[7])
                 : "l"(p));
    return v;
}

// `better(v, best)`: v replaces best (strictly greater, or v is the first NaN)
__device__ __forceinline__ bool better(float v, float best) { return v > best || (v != v && best == best); }

template <class T, int KU /* planes loaded per batch */>
__global__ void __launch_bounds__(256) argmax_planes_kernel(const T* __restrict__ scores, long long n_items, int K,
                                                            long long P /* pixels per plane, multiple of VPT */,
                                                            uint8_t* __restrict__ labels) {
    constexpr int VPT = Vec<T>::VPT;
    const long long vec_per_item = P / VPT;
    const long long total = n_items * vec_per_item;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; g < total; g += stride) {
        const long long item = g / vec_per_item, v = g - item * vec_per_item;
        const T* base = scores + (item * K) * P + v * VPT;
        float best[VPT];
        uint32_t arg[VPT];
#pragma unroll
        for (int i = 0; i < VPT; ++i) { best[i] = -INFINITY; arg[i] = 0; }
        for (int c0 = 0; c0 < K; c0 += KU) {
            Raw256 raw[KU];
#pragma unroll
            for (int u = 0; u < KU; ++u)
                if (c0 + u < K) raw[u] = ld_stream(base + static_cast<long long>(c0 + u) * P);
#pragma unroll
            for (int u = 0; u < KU; ++u) {
                if (c0 + u < K) {
                    const T* vals = reinterpret_cast<const T*>(&raw[u]);
#pragma unroll
                    for (int i = 0; i < VPT; ++i) {
                        const float x = to_f32(vals[i]);
                        if (better(x, best[i])) { best[i] = x; arg[i] = c0 + u; }
                    }
                }
            }
        }
        uint8_t* out = labels + item * P + v * VPT;
        uint32_t packed[VPT / 4];
#pragma unroll
        for (int i = 0; i < VPT / 4; ++i)
            packed[i] = arg[4 * i] | (arg[4 * i + 1] << 8) | (arg[4 * i + 2] << 16) | (arg[4 * i + 3] << 24);
        if (VPT == 8) *reinterpret_cast<uint2*>(out) = make_uint2(packed[0], packed[1]);
        else *reinterpret_cast<uint4*>(out) = make_uint4(packed[0], packed[1], packed[2], packed[VPT / 4 - 1]);
    }
}

// any K / P / alignment: one pixel per thread, plane-major (coalesced across threads) or interleaved
template <class T>
__global__ void __launch_bounds__(256) argmax_generic_kernel(const T* __restrict__ scores, long long n_items, int K, long long P,
                                                             long long class_stride, long long pixel_stride,
                                                             uint8_t* __restrict__ labels) {
    const long long total = n_items * P;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; g < total; g += stride) {
        const long long item = g / P, px = g - item * P;
        const T* base = scores + item * K * P + px * pixel_stride;
        float best = to_f32(base[0]);
        uint32_t arg = 0;
        for (int c = 1; c < K; ++c) {
            const float x = to_f32(base[c * class_stride]);
            if (better(x, best)) { best = x; arg = c; }
        }
        labels[g] = static_cast<uint8_t>(arg);
    }
}

template <class T>
static int launch_argmax(const void* scores, int64_t n_items, int K, int64_t P, int channels_last, uint8_t* labels,
                         cudaStream_t st) {
    const T* s = static_cast<const T*>(scores);
    constexpr int VPT = Vec<T>::VPT;
    const int sms = sm_count();
    if (!channels_last && P % VPT == 0 && reinterpret_cast<uintptr_t>(scores) % 32 == 0 &&
        reinterpret_cast<uintptr_t>(labels) % 16 == 0) {
        long long blocks = (n_items * (P / VPT) + 255) / 256;
        const long long cap = static_cast<long long>(sms) * 8;
        if (blocks > cap) blocks = cap;
        OCTM_TIMED("argmax_planes_kernel", st) argmax_planes_kernel<T, 4><<<static_cast<unsigned>(blocks), 256, 0, st>>>(s, n_items, K, P, labels);
        return check_launch("argmax_planes_kernel");
    }
    long long blocks = (n_items * P + 255) / 256;
    const long long cap = static_cast<long long>(sms) * 16;
    if (blocks > cap) blocks = cap;
    OCTM_TIMED("argmax_generic_kernel", st) argmax_generic_kernel<T><<<static_cast<unsigned>(blocks), 256, 0, st>>>(s, n_items, K, P, channels_last ? 1 : P,
                                                                           channels_last ? K : 1, labels);
    return check_launch("argmax_generic_kernel");
}

}  // namespace octm

extern "C" int octm_argmax_labels(const void* scores, int dtype, int64_t n_items, int num_classes, int64_t plane_elems,
                                  int channels_last, uint8_t* labels, void* stream) {
    if (n_items < 0 || plane_elems < 1) return octm::fail(OCTM_ERR_INVALID, "bad shape");
    if (num_classes < 1 || num_classes > 256) return octm::fail(OCTM_ERR_INVALID, "num_classes %d outside [1, 256]", num_classes);
    if (n_items == 0) return OCTM_OK;
    if (!scores || !labels) return octm::fail(OCTM_ERR_INVALID, "null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (dtype) {
        case OCTM_DTYPE_F32: return octm::launch_argmax<float>(scores, n_items, num_classes, plane_elems, channels_last, labels, st);
        case OCTM_DTYPE_F16: return octm::launch_argmax<__half>(scores, n_items, num_classes, plane_elems, channels_last, labels, st);
        case OCTM_DTYPE_BF16: return octm::launch_argmax<__nv_bfloat16>(scores, n_items, num_classes, plane_elems, channels_last, labels, st);
        default: return octm::fail(OCTM_ERR_INVALID, "dtype %d: expected OCTM_DTYPE_F32 / F16 / BF16", dtype);
    }
}
