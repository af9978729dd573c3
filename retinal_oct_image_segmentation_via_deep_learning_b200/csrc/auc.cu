// auc_score on the GPU: area under the ROC curve of a score map against a binary truth mask.
//
// Reference: Metrics/ConfusionMatrix_based_metrics.py:65-84
//     return roc_auc_score(y_true.flatten(), y_pred.flatten())     # ValueError -> 0.0
// scikit-learn (unpinned by the reference; 1.9.0 in this image) computes the trapezoidal area of
// the ROC curve over the DISTINCT score values, which equals the Mann-Whitney statistic with ties
// counted one half:
//     AUC = sum_{p in pos} ( #{n in neg : s_n < s_p} + 0.5 #{n in neg : s_n == s_p} ) / (n_pos n_neg)
// Here 2 * numerator is formed as an exact 64-bit integer: after sorting by score,
//     2 * num = sum_{p in pos} ( cneg[first of p's tie run] + cneg[one past the last of p's tie run] )
// with cneg[i] = number of negatives among sorted positions < i.  One float64 division ends it
// (the reference's own float result differs from that quotient by a few ulp; tolerance 1e-6).
//
// One CTA (1024 threads) per item, persistent over items; per item
//   1. label census (min / max label and their counts; any other value => "multiclass"), keys =
//      order-preserving integer image of the scores (-0.0 == +0.0), non-finite scores flagged;
//   2. stable LSD radix sort by key, 8 bits per pass, payload = the 0/1 label.  Each warp owns a
//      contiguous chunk; equal digits inside a 32-element slice are ranked with MATCH.ANY, so a pass
//      needs two CTA barriers and no atomics;
//   3. a forward and a backward sweep over the sorted sequence (each thread a contiguous chunk,
//      carries joined by CTA-wide max / min scans) accumulate the sum above.
// Outcomes that make scikit-learn raise ValueError (more than two label values, NaN / inf scores)
// give 0.0 like the reference's except-branch.  A single-class y_true gives `single_class_value`
// (NaN by default = scikit-learn >= 1.6, which warns instead of raising; pass 0.0 for older ones).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace octm {

constexpr int kAucThreads = 1024;
constexpr int kAucWarps = kAucThreads / 32;

template <class KeyT> struct KeyOf;
template <> struct KeyOf<uint32_t> {
    __device__ static uint32_t make(float x, bool& finite) {
        finite = isfinite(x);
        if (x == 0.0f) x = 0.0f;                      // -0.0 and +0.0 compare equal
        const uint32_t u = __float_as_uint(x);
        return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    }
};
template <> struct KeyOf<uint64_t> {
    __device__ static uint64_t make(double x, bool& finite) {
        finite = isfinite(x);
        if (x == 0.0) x = 0.0;
        const uint64_t u = static_cast<uint64_t>(__double_as_longlong(x));
        return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    }
};

__device__ __forceinline__ float load_score(const float* p, long long i) { return p[i]; }
__device__ __forceinline__ float load_score(const __half* p, long long i) { return __half2float(p[i]); }
__device__ __forceinline__ float load_score(const __nv_bfloat16* p, long long i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ double load_score(const double* p, long long i) { return p[i]; }

struct AucParams {
    const uint8_t* y_true;     // [n][P]
    const void* scores;        // [n][P]
    long long n_items;
    long long P;
    double single_class_value;
    double* auc;               // [n]
    uint8_t* workspace;        // per CTA: keys[2][P], labs[2][P]
    size_t ws_per_cta;
};

// exclusive prefix MAX over threads in thread order (identity 0)
__device__ __forceinline__ uint32_t block_excl_scan_max(uint32_t v, uint32_t* s_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = max(incl, n);
    }
    uint32_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 0;
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi = max(wi, n);
        }
        uint32_t we = __shfl_up_sync(0xffffffffu, wi, 1);
        if (lane == 0) we = 0;
        s_warp[lane] = we;
    }
    __syncthreads();
    const uint32_t res = max(s_warp[warp], excl);
    __syncthreads();
    return res;
}

// exclusive SUFFIX MIN over threads (thread t gets the min over threads > t; identity 0xffffffff)
__device__ __forceinline__ uint32_t block_excl_suffix_min(uint32_t v, uint32_t* s_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl = min(incl, n);
    }
    uint32_t excl = __shfl_down_sync(0xffffffffu, incl, 1);
    if (lane == 31) excl = 0xffffffffu;
    if (lane == 0) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_down_sync(0xffffffffu, wi, o);
            if (lane + o < 32) wi = min(wi, n);
        }
        uint32_t we = __shfl_down_sync(0xffffffffu, wi, 1);
        if (lane == 31) we = 0xffffffffu;
        s_warp[lane] = we;
    }
    __syncthreads();
    const uint32_t res = min(s_warp[warp], excl);
    __syncthreads();
    return res;
}

template <class ScoreT, class KeyT>
__global__ void __launch_bounds__(kAucThreads, 1) auc_kernel(const AucParams prm) {
    __shared__ uint32_t s_hist[kAucWarps][256];     // per-warp digit counts, then per-warp write offsets
    __shared__ uint32_t s_tot[256];
    __shared__ unsigned long long s_scan64[33];
    __shared__ uint32_t s_scan32[33];
    __shared__ uint32_t s_stat[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long P = prm.P;
    const uint32_t n = static_cast<uint32_t>(P);
    uint8_t* ws = prm.workspace + static_cast<size_t>(blockIdx.x) * prm.ws_per_cta;
    KeyT* keys[2] = {reinterpret_cast<KeyT*>(ws), reinterpret_cast<KeyT*>(ws) + P};
    uint8_t* labs[2] = {ws + 2 * P * sizeof(KeyT), ws + 2 * P * sizeof(KeyT) + P};
    const ScoreT* all_scores = static_cast<const ScoreT*>(prm.scores);
    // sort chunks: one contiguous, 32-aligned chunk per warp
    const uint32_t wchunk = ((n + kAucWarps - 1) / kAucWarps + 31u) & ~31u;
    const uint32_t wb = min(n, warp * wchunk), we = min(n, wb + wchunk);
    // sweep chunks: one contiguous chunk per thread
    const uint32_t tchunk = (n + kAucThreads - 1) / kAucThreads;
    const uint32_t tb = min(n, tid * tchunk), te = min(n, tb + tchunk);

    for (long long item = blockIdx.x; item < prm.n_items; item += gridDim.x) {
        const uint8_t* yt = prm.y_true + item * P;
        const ScoreT* sc = all_scores + item * P;
        // ---------------------------------------------------------------- 1. census + keys
        if (tid < 8) s_stat[tid] = tid == 0 ? 255u : 0u;      // [0] min label, [1] max label, [2] #min, [3] #max, [4] non-finite
        __syncthreads();
        uint32_t lmin = 255, lmax = 0;
        for (uint32_t i = tid; i < n; i += kAucThreads) {
            const uint32_t l = yt[i];
            lmin = min(lmin, l);
            lmax = max(lmax, l);
        }
        lmin = __reduce_min_sync(0xffffffffu, lmin);
        lmax = __reduce_max_sync(0xffffffffu, lmax);
        if (lane == 0) { atomicMin(&s_stat[0], lmin); atomicMax(&s_stat[1], lmax); }
        __syncthreads();
        lmin = s_stat[0];
        lmax = s_stat[1];
        uint32_t cmin = 0, cmax = 0, bad = 0;
        for (uint32_t i = tid; i < n; i += kAucThreads) {
            const uint32_t l = yt[i];
            cmin += l == lmin;
            cmax += l == lmax;
            bool finite;
            keys[0][i] = KeyOf<KeyT>::make(load_score(sc, i), finite);
            labs[0][i] = l == lmax ? 1 : 0;
            bad |= finite ? 0u : 1u;
        }
        cmin = __reduce_add_sync(0xffffffffu, cmin);
        cmax = __reduce_add_sync(0xffffffffu, cmax);
        bad = __reduce_or_sync(0xffffffffu, bad);
        if (lane == 0) { atomicAdd(&s_stat[2], cmin); atomicAdd(&s_stat[3], cmax); atomicOr(&s_stat[4], bad); }
        __syncthreads();
        const uint32_t n_neg = s_stat[2], n_pos = s_stat[3];
        const bool single = lmin == lmax, multi = !single && n_neg + n_pos != n, nonfinite = s_stat[4] != 0;
        __syncthreads();
        if (n == 0 || single || multi || nonfinite) {
            // scikit-learn: non-finite scores and more than two label values raise ValueError (-> 0.0 in the
            // reference); an empty or single-class y_true is the version-dependent case
            if (tid == 0) prm.auc[item] = (multi || nonfinite) ? 0.0 : prm.single_class_value;
            continue;
        }
        // ---------------------------------------------------------------- 2. LSD radix sort
        int cur = 0;
        for (int shift = 0; shift < static_cast<int>(sizeof(KeyT)) * 8; shift += 8) {
            const KeyT* kin = keys[cur];
            const uint8_t* lin = labs[cur];
            KeyT* kout = keys[cur ^ 1];
            uint8_t* lout = labs[cur ^ 1];
#pragma unroll
            for (int i = 0; i < 8; ++i) s_hist[warp][i * 32 + lane] = 0;
            if (tid == 0) s_stat[5] = 0;          // set below when every key shares this pass's digit
            __syncwarp();
            for (uint32_t i0 = wb; i0 < we; i0 += 32) {
                const uint32_t i = i0 + lane;
                const bool ok = i < we;
                const uint32_t d = ok ? static_cast<uint32_t>(kin[i] >> shift) & 255u : 256u + lane;
                const uint32_t peers = __match_any_sync(0xffffffffu, d);
                if (ok && lane == __ffs(peers) - 1) s_hist[warp][d] += __popc(peers);
                __syncwarp();
            }
            __syncthreads();
            // digit totals; per-warp counts -> exclusive prefix over warps
            if (tid < 256) {
                uint32_t run = 0;
                for (int w = 0; w < kAucWarps; ++w) {
                    const uint32_t c = s_hist[w][tid];
                    s_hist[w][tid] = run;
                    run += c;
                }
                s_tot[tid] = run;
                if (run == n) s_stat[5] = 1;
            }
            __syncthreads();
            const bool trivial = s_stat[5] != 0;
            if (warp == 0) {
                uint32_t c[8], tot = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) { c[i] = s_tot[lane * 8 + i]; tot += c[i]; }
                uint32_t incl = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t nb = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += nb;
                }
                uint32_t run = incl - tot;
#pragma unroll
                for (int i = 0; i < 8; ++i) { s_tot[lane * 8 + i] = run; run += c[i]; }
            }
            __syncthreads();
            if (trivial) continue;            // (uniform) nothing moves in this pass
            for (uint32_t i0 = wb; i0 < we; i0 += 32) {
                const uint32_t i = i0 + lane;
                const bool ok = i < we;
                const KeyT k = ok ? kin[i] : 0;
                const uint8_t l = ok ? lin[i] : 0;
                const uint32_t d = ok ? static_cast<uint32_t>(k >> shift) & 255u : 256u + lane;
                const uint32_t peers = __match_any_sync(0xffffffffu, d);
                uint32_t base = 0;
                if (ok) base = s_hist[warp][d] + s_tot[d];
                __syncwarp();
                if (ok) {
                    const uint32_t pos = base + __popc(peers & lanemask_lt());
                    kout[pos] = k;
                    lout[pos] = l;
                    if (lane == __ffs(peers) - 1) s_hist[warp][d] += __popc(peers);
                }
                __syncwarp();
            }
            __syncthreads();
            cur ^= 1;
        }
        // ---------------------------------------------------------------- 3. sweeps over the sorted sequence
        const KeyT* ks = keys[cur];
        const uint8_t* ls = labs[cur];
        // forward: A[i] = cneg at the first element of i's tie run
        uint32_t negs = 0, last_start_local = 0xffffffffu;       // negatives before the last run start inside the chunk
        {
            KeyT prev = tb > 0 && tb < n ? ks[tb - 1] : 0;
            for (uint32_t i = tb; i < te; ++i) {
                const KeyT k = ks[i];
                if (i == 0 || k != prev) last_start_local = negs;
                negs += ls[i] ? 0u : 1u;
                prev = k;
            }
        }
        unsigned long long dummy;
        const unsigned long long neg_before = block_excl_scan_sum<unsigned long long>(negs, s_scan64, dummy);
        const uint32_t off = static_cast<uint32_t>(neg_before);
        // run starts carry cneg + 1 so that 0 can be the identity of the max scan
        const uint32_t carry_a = block_excl_scan_max(last_start_local == 0xffffffffu ? 0u : off + last_start_local + 1u, s_scan32);
        unsigned long long acc = 0;
        {
            uint32_t cn = off, a = carry_a;       // a = cneg(run start) + 1
            KeyT prev = tb > 0 && tb < n ? ks[tb - 1] : 0;
            for (uint32_t i = tb; i < te; ++i) {
                const KeyT k = ks[i];
                if (i == 0 || k != prev) a = cn + 1u;
                if (ls[i]) acc += a - 1u;
                else ++cn;
                prev = k;
            }
        }
        // backward: B[i] = cneg one past the last element of i's tie run = negatives at positions <= run end
        uint32_t first_end_incl = 0xffffffffu;        // inclusive negative count at the FIRST run end inside the chunk
        {
            uint32_t cn = off;                        // negatives before position i
            for (uint32_t i = tb; i < te; ++i) {
                cn += ls[i] ? 0u : 1u;                // now inclusive of i
                const bool is_end = i + 1 == n || ks[i + 1] != ks[i];
                if (is_end) { first_end_incl = cn; break; }
            }
        }
        const uint32_t carry_b = block_excl_suffix_min(first_end_incl, s_scan32);
        {
            uint32_t cn = off + negs, b = carry_b;    // cn = negatives at positions < te
            for (uint32_t i = te; i > tb; --i) {
                const uint32_t j = i - 1;
                const bool is_end = j + 1 == n || ks[j + 1] != ks[j];
                if (is_end) b = cn;                   // negatives at positions <= j
                if (ls[j]) acc += b;
                else --cn;
            }
        }
        unsigned long long total;
        block_excl_scan_sum<unsigned long long>(acc, s_scan64, total);
        if (tid == 0)
            prm.auc[item] = static_cast<double>(total) / (2.0 * static_cast<double>(n_pos) * static_cast<double>(n_neg));
        __syncthreads();
    }
}

template <class ScoreT, class KeyT>
static int launch_auc(AucParams p, int ctas, cudaStream_t st) {
    OCTM_TIMED("auc_kernel", st) auc_kernel<ScoreT, KeyT><<<ctas, kAucThreads, 0, st>>>(p);
    return check_launch("auc_kernel");
}

static size_t auc_ws_per_cta(int64_t P, int dtype) {
    const size_t key = dtype == OCTM_DTYPE_F64 ? 8 : 4;
    return ((static_cast<size_t>(P) * (2 * key + 2)) + 255) & ~static_cast<size_t>(255);
}

static int auc_ctas(int64_t n_items) {
    const long long cap = sm_count();
    return static_cast<int>(n_items < cap ? n_items : cap);
}

}  // namespace octm

extern "C" size_t octm_auc_workspace_bytes(int64_t n_items, int64_t item_elems, int dtype) {
    if (n_items <= 0 || item_elems <= 0) return 0;
    return octm::auc_ws_per_cta(item_elems, dtype) * static_cast<size_t>(octm::auc_ctas(n_items));
}

extern "C" int octm_auc_u8(const uint8_t* y_true, const void* scores, int dtype, int64_t n_items, int64_t item_elems,
                           double single_class_value, double* auc, void* workspace, size_t workspace_bytes, void* stream) {
    if (n_items < 0 || item_elems < 0) return octm::fail(OCTM_ERR_INVALID, "bad shape");
    if (item_elems >= (1ll << 31)) return octm::fail(OCTM_ERR_UNSUPPORTED, "item_elems >= 2^31");
    if (n_items == 0) return OCTM_OK;
    if (!auc || (item_elems > 0 && (!y_true || !scores))) return octm::fail(OCTM_ERR_INVALID, "null pointer");
    const size_t need = octm_auc_workspace_bytes(n_items, item_elems, dtype);
    if (item_elems > 0 && (workspace == nullptr || workspace_bytes < need))
        return octm::fail(OCTM_ERR_WORKSPACE, "workspace too small: need %zu B", need);
    if (item_elems > 0 && reinterpret_cast<uintptr_t>(workspace) % 16 != 0)
        return octm::fail(OCTM_ERR_INVALID, "workspace must be 16-byte aligned");
    octm::AucParams p{y_true, scores, n_items, item_elems, single_class_value, auc, static_cast<uint8_t*>(workspace),
                      octm::auc_ws_per_cta(item_elems, dtype)};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int ctas = octm::auc_ctas(n_items);
    switch (dtype) {
        case OCTM_DTYPE_F32: return octm::launch_auc<float, uint32_t>(p, ctas, st);
        case OCTM_DTYPE_F16: return octm::launch_auc<__half, uint32_t>(p, ctas, st);
        case OCTM_DTYPE_BF16: return octm::launch_auc<__nv_bfloat16, uint32_t>(p, ctas, st);
        case OCTM_DTYPE_F64: return octm::launch_auc<double, uint64_t>(p, ctas, st);
        default: return octm::fail(OCTM_ERR_INVALID, "dtype %d: expected OCTM_DTYPE_F32 / F16 / BF16 / F64", dtype);
    }
}
