// Contour metrics (hausdorff_distance, hausdorff_distance_95, assd) -- sm_100a.
//
// Reference: Metrics/Contour_based_metrics.py:5-56.  Each function takes
//     A = skimage.measure.find_contours(y_true, 0.5)[0],  B = find_contours(y_pred, 0.5)[0]
// (lines 15-16, 33-34, 50-51) and brute-forces  d1 = [min_a |a - p| for p in B],
// d2 = [min_b |b - p| for p in A]  (lines 19-20, 36-37, 53-54).
//
// What contour [0] is for a binary mask (restated in
// oracle/contours_oracle.py): marching squares over 2x2 pixel squares in raster order; vertices are
// the midpoints of "cracks" between unequal 4-neighbours; each square joins its cracks pairwise
// (saddles keep the two high pixels apart); [0] is the polyline through the FIRST emitted segment.
// If it closes on itself one vertex is repeated: the `to` end of its raster-last segment.
//
//   K5  trace_kernel     one thread per (item, class, map): locate the raster-first mixed square
//                        from the label pass's first-occurrence table, walk the polyline forward
//                        (and backward if it is open), emit doubled-lattice vertices.
//   K6  distance_search_kernel  one CTA per (item, class, direction): source contour staged in shared
//                        memory tile by tile; exact integer squared distances with the expansion
//                        |a|^2 - 2 a.q + |q|^2 (2 IMAD + 1 IMNMX per pair), pruned by bounding boxes.
//   K7  distance_select_kernel  one warp per (item, class, direction): max, sum of sqrt(D2/4) in
//                        float64, and a radix select of the two order statistics numpy's linear 95th
//                        percentile interpolates between.
//       distance_kernel  single-kernel brute-force / per-lane-pruned variants (OCTM_DISTANCE_MODE),
//                        kept as independent checks of the default path.
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <type_traits>

#include <mutex>

#include "common.cuh"
#include "trace_core.h"

namespace octm {

struct TraceParams {
    const uint8_t* yt;
    const uint8_t* yp;
    long long n_items;
    int H, W, K, max_pts;
    const uint32_t* first_pos;   // [n][2][K]
    uint32_t* verts;             // [n][K][2][max_pts]
    uint32_t* n_pts;             // [n][K][2]
    uint32_t* flags;             // [n][K]
    const int* bnd_t;            // [n][K-1][W] #{label < k} per column of y_true (label pass), or null
    const int* bnd_p;
    bool only_todo;              // walk only the contours trace_layered_kernel left (n_pts == kTraceTodo)
    const uint32_t* todo_count;  // device word: contours / units left for the fallback kernels; 0 = nothing to do (or null)
    uint32_t* walk_list;         // compacted ids of the contours left to the walk (written by trace_layered_kernel), or null
    uint32_t* walk_count;        // its length (device word, zeroed by the caller)
    int walk_when;               // list mode: 0 = walk, 1 = only a list of >= kWalkMany contours, 2 = only a shorter one
    const uint32_t* unsorted;    // [n] layering certificate (bit m: map m rejected), or null
    int take;                    // trace_layered_kernel: 0 = every contour / only_todo, 1 = the contours of REJECTED maps (before
                                 // layered_distance_kernel<3>), 2 = only_todo contours of CERTIFIED maps (after it)
};
constexpr uint32_t kWalkMany = 65536;

constexpr uint32_t kTraceTodo = 0xffffffffu;
constexpr uint32_t kLayeredBit = 0x80000000u;    // n_pts: contour verified as a height function, vertices not emitted
constexpr uint32_t kRowBit = 0x40000000u;        // n_pts (with kLayeredBit): the height row h(x) is in the contour's verts slot (int32 [W])
constexpr uint32_t kNeedsSearch = 0xfffffffeu;   // max_sq: unit left to the general (vertex-list) distance search

// Both row caches of the walk in one go: 16 aligned label bytes for the even-row cache and 16 for the odd-row
// cache, each loaded only when its predicate is set (lanes of a warp hit and miss independently: predicated
// loads, not branches).  The two loads are issued back to back, so a step that needs both rows waits for one
// memory latency, not two.  (Asking L2 for the line ahead in the direction of travel, or for the rows above
// and below, was measured slower both times: the prefetches cost more LSU slots than the latency they hide.)
__device__ __forceinline__ void load_rows16(uint4& re, uint4& ro, bool pe, bool po, const uint4* ae, const uint4* ao) {
    asm volatile(
        "{\n"
        ".reg .pred pe, po;\n"
        "setp.ne.u32 pe, %8, 0;\n"
        "setp.ne.u32 po, %9, 0;\n"
        "@pe ld.global.nc.L2::128B.v4.u32 {%0, %1, %2, %3}, [%10];\n"
        "@po ld.global.nc.L2::128B.v4.u32 {%4, %5, %6, %7}, [%11];\n"
        "}\n"
        : "+r"(re.x), "+r"(re.y), "+r"(re.z), "+r"(re.w), "+r"(ro.x), "+r"(ro.y), "+r"(ro.z), "+r"(ro.w)
        : "r"(static_cast<uint32_t>(pe)), "r"(static_cast<uint32_t>(po)), "l"(ae), "l"(ao));
}

// 16 aligned label bytes through the non-coherent path
__device__ __forceinline__ uint4 ldg_labels16(const uint4* p) {
    uint4 w;
    asm volatile("ld.global.nc.L2::128B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(p));
    return w;
}

// label byte (c & 15) of a 16-byte group
__device__ __forceinline__ uint32_t label_of(const uint4& w, uint32_t c) {
    const uint32_t lo = __byte_perm(w.x, w.y, c & 7u), hi = __byte_perm(w.z, w.w, c & 7u);
    return ((c & 8u) ? hi : lo) & 0xffu;
}

// WORDS: W % 16 == 0 and 16-byte aligned maps -> the walk reads aligned 16-byte label groups and keeps the last
// group of the even and of the odd row in registers (load_rows16).
#ifndef OCTM_TRACE_MINB
#define OCTM_TRACE_MINB 8
#endif
#ifdef OCTM_WALK_STATS
__device__ unsigned long long g_walk_stats[8];
#endif
template <bool WORDS>
__global__ void __launch_bounds__(128, OCTM_TRACE_MINB) trace_kernel(const TraceParams prm) {
    __shared__ uint32_t s_step[64];      // step_word table: (case, entry edge) -> exit edge, vertex offset, order
    if (prm.todo_count != nullptr && *prm.todo_count == 0) return;
    if (threadIdx.x < 64) s_step[threadIdx.x] = step_word(threadIdx.x);
    __syncthreads();
    const int K = prm.K, H = prm.H, W = prm.W;
    // With a compacted list (the fused path) consecutive THREADS take consecutive listed contours: the few contours
    // that need a walk are spread thinly over the id space, and a warp that meets them one lane at a time, iteration
    // after iteration, walks them one after the other (measured 0.98 ms against 0.3 ms for 16,384 lightly noisy items).
    const long long total = prm.walk_list != nullptr ? static_cast<long long>(*prm.walk_count) : prm.n_items * K * 2;
    if (prm.walk_list != nullptr && prm.walk_when != 0 && (total >= kWalkMany) != (prm.walk_when == 1)) return;
    // persistent CTAs (grid = SMs x a tuned number of CTAs per SM): the walks of the resident warps must
    // keep their few label rows in L1, so occupancy is capped by the launch, not by registers
    for (long long base = static_cast<long long>(blockIdx.x) * blockDim.x; base < total;
         base += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (base + threadIdx.x >= total) continue;
    const long long gid = prm.walk_list != nullptr ? static_cast<long long>(prm.walk_list[base + threadIdx.x]) : base + threadIdx.x;
    // warps are homogeneous in the map (contours of y_true and of y_pred differ in length, and a warp
    // runs as long as its longest walk): consecutive threads = same map, consecutive (item, class)
    const long long per_map = prm.n_items * K;
    const int m = gid >= per_map ? 1 : 0;
    const long long rem = gid - m * per_map;
    const long long item = rem / K;
    const int cls = static_cast<int>(rem % K);
    const uint8_t* L = (m ? prm.yp : prm.yt) + item * H * static_cast<long long>(W);
    uint32_t* out = prm.verts + ((item * K + cls) * 2 + m) * static_cast<long long>(prm.max_pts);
    const uint32_t* fp = prm.first_pos + (item * 2 + m) * K;
    if (prm.only_todo && prm.n_pts[(item * K + cls) * 2 + m] != kTraceTodo) continue;
    uint32_t npts = 0;
    bool closed = false, overflow = false;

    // seed: first pixel (raster order) whose mask value differs from pixel (0, 0)
    uint32_t seed = OCTM_NO_SEED;
    if (H >= 2 && W >= 2) {
        const int c00 = L[0];
        if (cls == c00) {
            for (int c = 0; c < K; ++c)
                if (c != c00) seed = min(seed, fp[c]);
        } else {
            seed = fp[cls];
        }
    }
#ifdef OCTM_WALK_STATS
    const long long t_begin = clock64();
#endif
    if (seed != OCTM_NO_SEED) {
        const uint32_t cap = static_cast<uint32_t>(prm.max_pts);
        uint32_t g0 = 0, g1 = 0, g2 = 0;
        // register cache of the last aligned 16-byte label group of an even (0) and an odd (1) row
        const uint4* L16 = reinterpret_cast<const uint4*>(L);
        const uint32_t w16 = static_cast<uint32_t>(W) >> 4;
        uint32_t tag0 = 0xffffffffu, tag1 = 0xffffffffu;
        uint4 row0 = make_uint4(0, 0, 0, 0), row1 = row0;
        const uint32_t c4 = 0x01010101u * static_cast<uint32_t>(cls);
        const TraceResult r = trace_first_contour(
            H, W, seed,
            [&](int r0, int c0) -> int {
                const uint8_t* p = L + (static_cast<uint32_t>(r0) * static_cast<uint32_t>(W) + static_cast<uint32_t>(c0));
                // four labels -> one word -> exact zero-byte test against the class (labels < 16)
                const uint32_t w = __ldg(p) | (__ldg(p + 1) << 8) | (__ldg(p + W) << 16) | (__ldg(p + W + 1) << 24);
                const uint32_t z = ~((w ^ c4) + 0x7f7f7f7fu) & 0x80808080u;
                return static_cast<int>(((z >> 7) * 0x01020408u) >> 24);
            },
            [&](int r0, int c0, int e) -> int {
                // the two pixels not shared with the previous square (trace_core.h: carried_bits)
                if (WORDS) {
                    const bool horiz = e >= 2;
                    const uint32_t ra = r0 + (e == 1), ca = c0 + (e == 3);
                    const uint32_t rb = ra + (horiz ? 1u : 0u), cb = ca + (horiz ? 0u : 1u);
                    const uint32_t wa = ra * w16 + (ca >> 4);
                    const uint32_t wb = horiz ? wa + w16 : ra * w16 + (cb >> 4);
                    const bool oa = ra & 1u;
                    // a horizontal move needs one pixel of an even and one of an odd row; a vertical move two
                    // neighbours of one row (the second is in another group once in 16 moves: third load below)
                    const uint32_t want0 = oa ? wb : wa, want1 = oa ? wa : wb;
                    const bool p0 = (horiz || !oa) && tag0 != want0, p1 = (horiz || oa) && tag1 != want1;
                    load_rows16(row0, row1, p0, p1, L16 + want0, L16 + want1);
                    tag0 = p0 ? want0 : tag0;
                    tag1 = p1 ? want1 : tag1;
                    const uint32_t la = oa ? label_of(row1, ca) : label_of(row0, ca);
                    if (!horiz && wb != wa) {                               // rare: the pair straddles two groups
                        const uint4 w = __ldg(L16 + wb);
                        if (oa) { row1 = w; tag1 = wb; } else { row0 = w; tag0 = wb; }
                    }
                    const bool ob = rb & 1u;
                    const uint32_t lb = ob ? label_of(row1, cb) : label_of(row0, cb);
                    const int bits = (la == static_cast<uint32_t>(cls) ? 1 : 0) | (lb == static_cast<uint32_t>(cls) ? 2 : 0);
                    return bits;
                }
                const uint8_t* p = L + (static_cast<uint32_t>(r0) * static_cast<uint32_t>(W) + static_cast<uint32_t>(c0));
                const int oa = (e == 1 ? W : 0) + (e == 3 ? 1 : 0), ob = oa + ((e & 2) ? W : 1);
                return (__ldg(p + oa) == cls ? 1 : 0) | (__ldg(p + ob) == cls ? 2 : 0);
            },
            [&](int rr, int cc) -> int { return __ldg(L + static_cast<long long>(rr) * W + cc) == cls ? 1 : 0; },
            [&](int idx6) -> uint32_t { return s_step[idx6]; },
            [&](uint32_t i, uint32_t v) {
                // vertices leave in groups of four (one 16-byte store): the walk is bound by the number of
                // per-lane memory transactions, not by bytes.  cap % 4 == 0, so a group is in or out as a whole.
                const uint32_t ph = i & 3u;
                if (ph == 3u) {
                    if (i < cap) *reinterpret_cast<uint4*>(out + (i - 3u)) = make_uint4(g0, g1, g2, v);
                } else {
                    g0 = ph == 0u ? v : g0;
                    g1 = ph == 1u ? v : g1;
                    g2 = ph == 2u ? v : g2;
                }
            });
        npts = r.npts;
        for (uint32_t i = npts & ~3u; i < npts; ++i)       // the last, incomplete group
            if (i < cap) out[i] = (i & 3u) == 0u ? g0 : ((i & 3u) == 1u ? g1 : g2);
        closed = r.closed;
        overflow = npts > cap;
    }
#ifdef OCTM_WALK_STATS
    {
        const unsigned long long dt = clock64() - t_begin;
        atomicAdd(&g_walk_stats[0], 1ull);
        atomicAdd(&g_walk_stats[1], static_cast<unsigned long long>(npts));
        atomicMax(&g_walk_stats[2], static_cast<unsigned long long>(npts));
        atomicAdd(&g_walk_stats[3], npts < 64 ? 1ull : 0ull);
        atomicAdd(&g_walk_stats[4], npts >= 512 ? 1ull : 0ull);
        atomicAdd(&g_walk_stats[5], npts >= 512 ? dt : 0ull);
        atomicAdd(&g_walk_stats[6], npts >= 512 ? static_cast<unsigned long long>(npts) : 0ull);
        atomicMax(&g_walk_stats[7], dt);
    }
#endif
    prm.n_pts[(item * K + cls) * 2 + m] = overflow ? static_cast<uint32_t>(prm.max_pts) : npts;
    uint32_t f = 0;
    if (closed) f |= m ? OCTM_CF_PRED_CLOSED : OCTM_CF_TRUE_CLOSED;
    if (overflow) f |= m ? OCTM_CF_PRED_OVERFLOW : OCTM_CF_TRUE_OVERFLOW;
    if (f) atomicOr(&prm.flags[item * K + cls], f);
    }
}

// ------------------------------------------------------------------------------ layered fast path
// On a layered B-scan contour [0] of a class mask is the boundary between "mask as at pixel (0, 0)" above and
// the rest below, running from the left to the right image border: a height function h(x).  The label pass
// already knows a candidate: h(x) = #{label < k} in column x (its boundary rows).  This kernel checks the
// candidate instead of walking it: one WARP per (item, class, map), 64 adjacent columns at a time, two per lane:
//   * 1 <= h(x) <= H - 1 everywhere;
//   * every pixel of every 2x2 square the walk would visit has the value the step function predicts: in column
//     x the rows [min(h(x-1), h(x), h(x+1)) - 1, max(..)] hold "mask(0,0)" above h(x) and the opposite from h(x)
//     down (16-bit loads, adjacent lanes = adjacent column pairs, four rows in flight);
//   * the raster-first pixel of the path is the seed the label pass found (no pixel of the class above it).
// The walk is a deterministic function of exactly these pixels and starts at the seed's square, so when all
// three hold it would trace this very polyline.  Its vertices -- (2 h(x) - 1, 2 x) per column and (2 r, 2 x + 1)
// for the rows r between h(x) and h(x+1) -- are written straight to the vertex list, left to right, at
// positions from a warp prefix sum (adjacent lanes write adjacent words).  Everything else (blobs, broken or
// touching layers, too many vertices) gets n_pts = kTraceTodo and is walked by trace_kernel afterwards, which
// overwrites whatever part of the list was written before the check failed.
#ifndef OCTM_LAYERED_MINB
#define OCTM_LAYERED_MINB 8
#endif
// EMIT = false: verification only.  A verified contour gets n_pts = (vertex count | kLayeredBit) and no vertices
// are written: layered_distance_kernel works from the boundary rows and emits vertices only for the pairs it
// cannot finish itself.
template <bool EMIT>
__global__ void __launch_bounds__(128, OCTM_LAYERED_MINB) trace_layered_kernel(const TraceParams prm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = prm.K, H = prm.H, W = prm.W;
    const uint32_t cap = static_cast<uint32_t>(prm.max_pts);
    const long long total = prm.n_items * K * 2, per_map = prm.n_items * K;
    if (prm.todo_count != nullptr && *prm.todo_count == 0) return;
    for (long long gid = static_cast<long long>(blockIdx.x) * 4 + warp; gid < total; gid += static_cast<long long>(gridDim.x) * 4) {
        const int m = gid >= per_map ? 1 : 0;
        const long long rem = gid - m * per_map;
        const long long item = rem / K;
        const int cls = static_cast<int>(rem % K);
        const uint8_t* L = (m ? prm.yp : prm.yt) + item * H * static_cast<long long>(W);
        uint32_t* out = prm.verts + ((item * K + cls) * 2 + m) * static_cast<long long>(cap);
        uint32_t* np = prm.n_pts + (item * K + cls) * 2 + m;
        if (prm.take != 0) {
            const bool rejected = (prm.unsorted[item] >> m) & 1u;
            if (rejected != (prm.take == 1)) continue;
            if (prm.take == 2 && *np != kTraceTodo) continue;
        } else if (prm.only_todo && *np != kTraceTodo) {
            continue;
        }
        const uint32_t* fp = prm.first_pos + (item * 2 + m) * K;
        // one load for all seeds; the class at pixel (0, 0) is the one whose first occurrence is index 0
        const uint32_t myfp = lane < K ? fp[lane] : OCTM_NO_SEED;
        const int c00 = __ffs(__ballot_sync(0xffffffffu, myfp == 0u)) - 1;
        const bool inv = cls == c00;                       // the class is the region above the path
        const uint32_t others = __reduce_min_sync(0xffffffffu, lane == c00 ? OCTM_NO_SEED : myfp);
        const uint32_t seed = inv ? others : __shfl_sync(0xffffffffu, myfp, cls);
        if (seed == OCTM_NO_SEED) {                        // empty or full mask: no contour
            if (lane == 0) *np = 0;
            continue;
        }
        const int brow = inv ? cls : cls - 1;              // h = #{label < brow + 1}
        bool ok = brow >= 0 && brow < K - 1;
        const int* hrow = (m ? prm.bnd_p : prm.bnd_t) + (item * (K - 1) + (ok ? brow : 0)) * static_cast<long long>(W);
        // The path's raster-first pixel must be the seed (checked in full at the end): unless the seed sits exactly on
        // the candidate in its own column there is nothing to verify -- the usual fate of a noisy prediction, whose
        // contour [0] is the blob around a stray pixel far above the layer.
        // The boundary rows are COUNTS (#{label < k} per column): a stray pixel elsewhere in a column moves the count by
        // one without touching the path.  The estimate is therefore re-centred on the step actually present in rows
        // h - 2 .. h + 1 of the column; the window check below then validates the re-centred candidate like any other.
        auto refine = [&](int h, int col) -> int {
            if (h < 2 || h > H - 2) return h;
            const uint8_t* p = L + static_cast<uint32_t>((h - 2) * W + col);
            uint32_t pat = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r) pat |= ((__ldg(p + r * W) == cls) != inv ? 1u : 0u) << r;
            return pat == 14u ? h - 1 : (pat == 8u ? h + 1 : h);       // 0111 / 0001 / (0011 or anything else)
        };
        if (ok) {
            const int xs = static_cast<int>(seed % static_cast<uint32_t>(W));
            if (refine(hrow[xs], xs) != static_cast<int>(seed / static_cast<uint32_t>(W))) ok = false;
        }
        uint32_t base = 0, minkey = 0xffffffffu;
        bool bad = false;
        // 64 columns per step, two adjacent columns (2 l, 2 l + 1) per lane: W is even on this path
        const int2* hrow2 = reinterpret_cast<const int2*>(hrow);
        const int W2 = W >> 1;
        int hleft = 0;                                     // h of the column left of this block
        int2 hcur = ok ? hrow2[min(lane, W2 - 1)] : make_int2(1, 1);
        for (int x0 = 0; ok && x0 < W; x0 += 64) {
            const int x = x0 + 2 * lane;                   // this lane's first column
            const bool valid = x < W;
            const int xc = min(x, W - 2);
            const int ha = refine(hcur.x, xc), hb = refine(hcur.y, xc + 1);
            hcur = hrow2[min((x0 >> 1) + 32 + lane, W2 - 1)];      // next block's heights, in flight during the checks
            int hl = __shfl_up_sync(0xffffffffu, hb, 1);
            if (lane == 0) hl = x0 > 0 ? hleft : ha;
            int hr = __shfl_down_sync(0xffffffffu, ha, 1);
            const int hfirst_next = __shfl_sync(0xffffffffu, hcur.x, 0);
            if (lane == 31 && x + 2 < W) hr = refine(hfirst_next, x + 2);
            if (x + 2 >= W) hr = hb;
            // the re-centred row is what layered_distance_kernel<2> builds this side's table from
            if (!EMIT && valid && static_cast<uint32_t>(W) <= cap) reinterpret_cast<int2*>(out)[x >> 1] = make_int2(ha, hb);
            hleft = __shfl_sync(0xffffffffu, hb, 31);
            // all heights inside [1, H - 1] before any pixel is addressed through them (the right neighbour of
            // lane 31 belongs to the next block and has not been looked at yet)
            const unsigned hm = static_cast<unsigned>(H - 2);
            const bool inside = static_cast<unsigned>(ha - 1) <= hm && static_cast<unsigned>(hb - 1) <= hm &&
                                static_cast<unsigned>(hl - 1) <= hm && static_cast<unsigned>(hr - 1) <= hm;
            if (!__all_sync(0xffffffffu, !valid || inside)) { ok = false; break; }
            // windows of the two columns as 32-bit pixel offsets (H * W < 2^31 on this path)
            const int rla = min(hl, min(ha, hb)) - 1, rha = max(hl, max(ha, hb));
            const int rlb = min(ha, min(hb, hr)) - 1, rhb = max(ha, max(hb, hr));
            const int loa = rla * W, hia = rha * W, lob = rlb * W, hib = rhb * W;
            const int olo = min(loa, lob), ohi = max(hia, hib), oha = ha * W, ohb = hb * W;
            const uint8_t* colp = L + xc;
            const int nrows = __reduce_max_sync(0xffffffffu, valid ? max(rha, rhb) - min(rla, rlb) + 1 : 0);
            for (int u0 = 0; u0 < nrows; u0 += 4) {        // four rows in flight, two pixels each
                int off[4];
                uint32_t px[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {              // past the window: its last row again (harmless)
                    off[u] = min(olo + (u0 + u) * W, ohi);
                    px[u] = __ldg(reinterpret_cast<const unsigned short*>(colp + static_cast<uint32_t>(off[u])));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool ea = (px[u] & 0xffu) == static_cast<uint32_t>(cls), eb = (px[u] >> 8) == static_cast<uint32_t>(cls);
                    bad |= off[u] >= loa && off[u] <= hia && ((ea != (off[u] >= oha)) != inv);
                    bad |= off[u] >= lob && off[u] <= hib && ((eb != (off[u] >= ohb)) != inv);
                }
            }
            // a window pixel off: not this polyline -- stop here (a ragged contour fails in its first block; running the
            // remaining blocks only to say "no" cost 0.9 ms per 16,384 ragged predictions)
            if (__any_sync(0xffffffffu, bad)) { ok = false; break; }
            // vertices of these columns at their left-to-right positions
            const int da = abs(hb - ha), db = abs(hr - hb);
            const uint32_t cnt = valid ? 2u + static_cast<uint32_t>(da + db) : 0u;
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            const uint32_t pos = base + incl - cnt;
            base += __shfl_sync(0xffffffffu, incl, 31);
            if (base > cap) { ok = false; break; }         // too long: the walk flags the overflow
            if (valid) {
                const uint32_t x2 = 2u * static_cast<uint32_t>(x);
                const uint32_t ua = static_cast<uint32_t>(ha), ub = static_cast<uint32_t>(hb), uw = static_cast<uint32_t>(W);
                minkey = min(minkey, min(ua * uw + static_cast<uint32_t>(x), ub * uw + static_cast<uint32_t>(x) + 1u));
            }
            if (EMIT && valid) {
                const uint32_t x2 = 2u * static_cast<uint32_t>(x);
                uint32_t* o = out + pos;
                *o++ = (static_cast<uint32_t>(2 * ha - 1) << 16) | x2;
                uint32_t va = (static_cast<uint32_t>(2 * min(ha, hb)) << 16) | (x2 + 1u);
                for (int t = 0; t < da; ++t, va += 2u << 16) *o++ = va;
                *o++ = (static_cast<uint32_t>(2 * hb - 1) << 16) | (x2 + 2u);
                uint32_t vb = (static_cast<uint32_t>(2 * min(hb, hr)) << 16) | (x2 + 3u);
                for (int t = 0; t < db; ++t, vb += 2u << 16) *o++ = vb;
            }
        }
        if (ok) ok = !__any_sync(0xffffffffu, bad) && __reduce_min_sync(0xffffffffu, minkey) == seed;
        if (ok && !EMIT && static_cast<uint32_t>(W) > cap) ok = false;      // no room for the row: walk it
        if (lane == 0) {
            *np = ok ? (EMIT ? base : (base | kLayeredBit | kRowBit)) : kTraceTodo;
            if (!ok && prm.walk_list != nullptr) prm.walk_list[atomicAdd(prm.walk_count, 1u)] = static_cast<uint32_t>(gid);
        }
    }
}

// ------------------------------------------------------------------------------ first occurrence
// first_pos[i][c] = smallest flat index with label c (OCTM_NO_SEED when absent); CTA per item.
__global__ void __launch_bounds__(256) first_pos_kernel(const uint8_t* __restrict__ labels, long long n_items,
                                                        long long item_elems, int K, uint32_t* first_pos, long long out_stride) {
    __shared__ uint32_t s_first[16];
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        if (threadIdx.x < 16) s_first[threadIdx.x] = OCTM_NO_SEED;
        __syncthreads();
        const uint8_t* L = labels + item * item_elems;
        // contiguous chunk per thread so its own indices ascend
        const long long chunk = (item_elems + blockDim.x - 1) / blockDim.x;
        const long long b = threadIdx.x * chunk, e = min(item_elems, b + chunk);
        uint32_t seen = 0;
        for (long long i = b; i < e; ++i) {
            const uint32_t v = L[i] & 15u;
            if (!((seen >> v) & 1u)) {
                seen |= 1u << v;
                atomicMin(&s_first[v], static_cast<uint32_t>(i));
            }
        }
        __syncthreads();
        if (threadIdx.x < K) first_pos[item * out_stride + threadIdx.x] = s_first[threadIdx.x];
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------ distances
struct DistParams {
    const uint32_t* verts;   // [n][K][2][max_pts]
    const uint32_t* n_pts;   // [n][K][2]
    long long n_pairs;       // n * K
    int max_pts;
    uint32_t* max_sq;        // [n][K][2]
    uint32_t* p95_sq;        // [n][K][2][2]
    double* sum_dist;        // [n][K][2]
    uint32_t* sq_out;        // [n][K][2][max_pts] or null
};

constexpr int kDistThreads = 128;
constexpr int kQ = 4;   // queries per thread per sweep

template <int T>
__device__ __forceinline__ uint32_t block_reduce_max(uint32_t v, uint32_t* scratch) {
    v = __reduce_max_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t r = 0;
    for (int w = 0; w < T / 32; ++w) r = max(r, scratch[w]);
    __syncthreads();
    return r;
}
template <int T>
__device__ __forceinline__ uint32_t block_reduce_min(uint32_t v, uint32_t* scratch) {
    v = __reduce_min_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t r = 0xffffffffu;
    for (int w = 0; w < T / 32; ++w) r = min(r, scratch[w]);
    __syncthreads();
    return r;
}
template <int T>
__device__ __forceinline__ uint32_t block_reduce_add(uint32_t v, uint32_t* scratch) {
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t r = 0;
    for (int w = 0; w < T / 32; ++w) r += scratch[w];
    __syncthreads();
    return r;
}
template <int T>
__device__ __forceinline__ double block_reduce_add(double v, double* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0;
    for (int w = 0; w < T / 32; ++w) r += scratch[w];
    __syncthreads();
    return r;
}

// k-th smallest (0-based) of vals[0..m) by 8-bit radix passes; all threads return the value.
template <int T>
__device__ uint32_t block_select(const uint32_t* vals, int m, uint32_t k, uint32_t vmax, uint32_t* hist /*256*/,
                                 uint32_t* bcast /*2*/) {
    uint32_t prefix = 0, maskbits = 0;
    int shift = vmax >= (1u << 24) ? 24 : (vmax >= (1u << 16) ? 16 : (vmax >= (1u << 8) ? 8 : 0));
    for (; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += T) hist[i] = 0;
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += T) {
            const uint32_t v = vals[j];
            if ((v & maskbits) == prefix) atomicAdd(&hist[(v >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            uint32_t c[8], tot = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { c[i] = hist[lane * 8 + i]; tot += c[i]; }
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            const uint32_t excl = incl - tot;
            if (k >= excl && k < incl) {   // exactly one lane
                uint32_t run = excl;
                int d = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (k >= run + c[i]) { run += c[i]; d = i + 1; }
                    else break;
                }
                bcast[0] = static_cast<uint32_t>(lane * 8 + d);
                bcast[1] = k - run;
            }
        }
        __syncthreads();
        prefix |= bcast[0] << shift;
        maskbits |= 255u << shift;
        k = bcast[1];
        __syncthreads();
    }
    return prefix;
}

constexpr int kBox = 16;      // vertices per bounding box (consecutive in trace order => compact)
constexpr int kSuper = 8;     // boxes per super-box

// squared distance from q to an axis-aligned box {ymin, ymax, xmin, xmax}: exact lower bound of the
// squared distance to every vertex inside it
__device__ __forceinline__ int box_lb2(int4 bx, int qy, int qx) {
    const int dy = max(max(bx.x - qy, qy - bx.y), 0);
    const int dx = max(max(bx.z - qx, qx - bx.w), 0);
    return dy * dy + dx * dx;
}

// PRUNE = false: brute force (every query against every source vertex).
// PRUNE = true : the same minimum, skipping boxes whose lower bound cannot beat the current best.
template <bool PRUNE>
__global__ void __launch_bounds__(kDistThreads) distance_kernel(const DistParams prm) {
    extern __shared__ __align__(16) uint8_t dsm[];
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_scr[8];
    __shared__ double s_dscr[8];
    __shared__ uint32_t s_bc[2];
    const int cap = prm.max_pts;
    const int capb = (cap + kBox - 1) / kBox, caps = (capb + kSuper - 1) / kSuper;
    // shared-memory carve-up (offsets, so every access stays an LDS/STS rather than a generic load)
    int2* const pts0 = reinterpret_cast<int2*>(dsm);                                     // {y<<16|x, y^2+x^2}
    uint32_t* const d2 = reinterpret_cast<uint32_t*>(dsm + static_cast<size_t>(cap) * 16);
    int4* const boxes0 = reinterpret_cast<int4*>(dsm + static_cast<size_t>(cap) * 20);
    int4* const sboxes0 = boxes0 + 2 * capb;
#define PTS(mm) (pts0 + (mm) * cap)
#define BOXES(mm) (boxes0 + (mm) * capb)
#define SBOXES(mm) (sboxes0 + (mm) * caps)

    for (long long pair = blockIdx.x; pair < prm.n_pairs; pair += gridDim.x) {
        const uint32_t n0 = prm.n_pts[pair * 2 + 0], n1 = prm.n_pts[pair * 2 + 1];
        const int n[2] = {static_cast<int>(min(n0, (uint32_t)cap)), static_cast<int>(min(n1, (uint32_t)cap))};
        if (n[0] == 0 || n[1] == 0) {
            if (threadIdx.x < 2) {
                prm.max_sq[pair * 2 + threadIdx.x] = 0;
                prm.p95_sq[pair * 4 + threadIdx.x * 2] = prm.p95_sq[pair * 4 + threadIdx.x * 2 + 1] = 0;
                prm.sum_dist[pair * 2 + threadIdx.x] = 0.0;
            }
            continue;
        }
        __syncthreads();
        for (int mm = 0; mm < 2; ++mm) {
            const uint32_t* src = prm.verts + (pair * 2 + mm) * static_cast<long long>(cap);
            for (int i = threadIdx.x; i < n[mm]; i += kDistThreads) {
                const uint32_t v = src[i];
                const int y = v >> 16, x = v & 0xffff;
                PTS(mm)[i] = make_int2(static_cast<int>(v), y * y + x * x);
            }
        }
        __syncthreads();
        if (PRUNE) {
            for (int mm = 0; mm < 2; ++mm) {
                const int nb = (n[mm] + kBox - 1) / kBox;
                for (int b = threadIdx.x; b < nb; b += kDistThreads) {
                    int ymin = 0x7fffffff, ymax = -1, xmin = 0x7fffffff, xmax = -1;
                    const int e = min(n[mm], (b + 1) * kBox);
                    for (int i = b * kBox; i < e; ++i) {
                        const uint32_t v = static_cast<uint32_t>(PTS(mm)[i].x);
                        const int y = v >> 16, x = v & 0xffff;
                        ymin = min(ymin, y); ymax = max(ymax, y); xmin = min(xmin, x); xmax = max(xmax, x);
                    }
                    BOXES(mm)[b] = make_int4(ymin, ymax, xmin, xmax);
                }
            }
            __syncthreads();
            for (int mm = 0; mm < 2; ++mm) {
                const int nb = (n[mm] + kBox - 1) / kBox, nsb = (nb + kSuper - 1) / kSuper;
                for (int sb = threadIdx.x; sb < nsb; sb += kDistThreads) {
                    int4 acc = BOXES(mm)[sb * kSuper];
                    const int e = min(nb, (sb + 1) * kSuper);
                    for (int b = sb * kSuper + 1; b < e; ++b) {
                        const int4 c = BOXES(mm)[b];
                        acc.x = min(acc.x, c.x); acc.y = max(acc.y, c.y); acc.z = min(acc.z, c.z); acc.w = max(acc.w, c.w);
                    }
                    SBOXES(mm)[sb] = acc;
                }
            }
            __syncthreads();
        }
        // direction 0: queries = pred vertices (map 1), sources = true vertices (map 0); direction 1 swapped
        for (int dir = 0; dir < 2; ++dir) {
            const int2* qs = PTS(1 - dir);
            const int2* ss = PTS(dir);
            const int nq = n[1 - dir], ns = n[dir];
            if (PRUNE) {
                const int4* bxs = BOXES(dir);
                const int4* sbs = SBOXES(dir);
                const float ratio = static_cast<float>(ns) / static_cast<float>(nq);
                const int nb = (ns + kBox - 1) / kBox, nsb = (nb + kSuper - 1) / kSuper;
                for (int j = threadIdx.x; j < nq; j += kDistThreads) {
                    const int2 q = qs[j];
                    const int qy = static_cast<uint32_t>(q.x) >> 16, qx = q.x & 0xffff;
                    const int cy = -2 * qy, cx = -2 * qx;
                    // prime the bound with the vertex at the same relative position along the other contour
                    const int2 g = ss[min(ns - 1, static_cast<int>(static_cast<float>(j) * ratio))];
                    int bm = (g.x & 0xffff) * cx + (static_cast<int>(static_cast<uint32_t>(g.x) >> 16) * cy + g.y);
                    int bestd = bm + q.y;
                    for (int sb = 0; sb < nsb; ++sb) {
                        if (box_lb2(sbs[sb], qy, qx) >= bestd) continue;
                        const int be = min(nb, (sb + 1) * kSuper);
                        for (int b = sb * kSuper; b < be; ++b) {
                            if (box_lb2(bxs[b], qy, qx) >= bestd) continue;
                            const int e = min(ns, (b + 1) * kBox);
#pragma unroll 4
                            for (int i = b * kBox; i < e; ++i) {
                                const int2 s = ss[i];
                                const int ay = static_cast<uint32_t>(s.x) >> 16, ax = s.x & 0xffff;
                                bm = min(bm, ax * cx + (ay * cy + s.y));
                            }
                            bestd = bm + q.y;
                        }
                    }
                    d2[j] = static_cast<uint32_t>(bestd);
                }
            } else {
            for (int base = 0; base < nq; base += kDistThreads * kQ) {
                int cy[kQ], cx[kQ], best[kQ], qn[kQ];
#pragma unroll
                for (int t = 0; t < kQ; ++t) {
                    const int j = min(base + t * kDistThreads + static_cast<int>(threadIdx.x), nq - 1);
                    const int2 q = qs[j];
                    cy[t] = -2 * static_cast<int>(static_cast<uint32_t>(q.x) >> 16);
                    cx[t] = -2 * (q.x & 0xffff);
                    qn[t] = q.y;
                    best[t] = 0x7fffffff;
                }
#pragma unroll 4
                for (int i = 0; i < ns; ++i) {
                    const int2 s = ss[i];
                    const int ay = static_cast<uint32_t>(s.x) >> 16, ax = s.x & 0xffff;
#pragma unroll
                    for (int t = 0; t < kQ; ++t) best[t] = min(best[t], ax * cx[t] + (ay * cy[t] + s.y));
                }
#pragma unroll
                for (int t = 0; t < kQ; ++t) {
                    const int j = base + t * kDistThreads + static_cast<int>(threadIdx.x);
                    if (j < nq) d2[j] = static_cast<uint32_t>(best[t] + qn[t]);
                }
            }
            }
            __syncthreads();
            // K7: max, sum of sqrt, two order statistics
            uint32_t vmax = 0;
            double dsum = 0.0;
            for (int j = threadIdx.x; j < nq; j += kDistThreads) {
                const uint32_t v = d2[j];
                vmax = max(vmax, v);
                dsum += sqrt(static_cast<double>(v) / 4.0);
            }
            vmax = block_reduce_max<kDistThreads>(vmax, s_scr);
            dsum = block_reduce_add<kDistThreads>(dsum, s_dscr);
            // numpy linear percentile: virtual index (m - 1) * 0.95, neighbours floor and floor + 1
            const double pos = __dmul_rn(static_cast<double>(nq - 1), 0.95);
            const uint32_t lo = static_cast<uint32_t>(floor(pos));
            const uint32_t v_lo = block_select<kDistThreads>(d2, nq, lo, vmax, s_hist, s_bc);
            uint32_t cnt_le = 0, next_gt = 0xffffffffu;
            for (int j = threadIdx.x; j < nq; j += kDistThreads) {
                const uint32_t v = d2[j];
                cnt_le += v <= v_lo ? 1u : 0u;
                if (v > v_lo) next_gt = min(next_gt, v);
            }
            cnt_le = block_reduce_add<kDistThreads>(cnt_le, s_scr);
            next_gt = block_reduce_min<kDistThreads>(next_gt, s_scr);
            const uint32_t v_hi = (lo + 1 >= static_cast<uint32_t>(nq) || cnt_le >= lo + 2) ? v_lo : next_gt;
            if (prm.sq_out != nullptr) {
                uint32_t* o = prm.sq_out + (pair * 2 + dir) * static_cast<long long>(cap);
                for (int j = threadIdx.x; j < nq; j += kDistThreads) o[j] = d2[j];
            }
            if (threadIdx.x == 0) {
                prm.max_sq[pair * 2 + dir] = vmax;
                prm.p95_sq[pair * 4 + dir * 2 + 0] = v_lo;
                prm.p95_sq[pair * 4 + dir * 2 + 1] = v_hi;
                prm.sum_dist[pair * 2 + dir] = dsum;
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------ tiled search + select
// Default path.  Two kernels, neither of which waits at a CTA barrier inside its hot loop:
//
//   distance_search_kernel   one CTA per (item, class, direction).  The source contour is staged in
//       shared memory in tiles of `tile` vertices as int4 {y, x, y^2+x^2, -} with one bounding box per
//       16 consecutive vertices (consecutive in trace order => compact).  A warp owns 32 CONSECUTIVE
//       query vertices (adjacent on their polyline, so they share almost the same neighbourhood of
//       the other contour):
//         1. box phase, shared by the warp: each lane tests a different source box against the
//            bounding box of the 32 queries; a ballot yields the candidate boxes;
//         2. each candidate is re-tested per lane against the lane's own best (one vote decides for
//            the warp), then scanned by all lanes together: the source vertex is one broadcast LDS.128
//            and every lane folds it into its own minimum -- 2 IMAD + 1 IMNMX per (query, vertex).
//       The minimum is exact: a box is skipped only when its lower bound cannot beat any lane's best.
//       Squared distances go to the d2 scratch in global memory (4 B per query vertex); contours of
//       any length are handled tile by tile, the running minimum carried in that scratch.
//   distance_select_kernel   one WARP per (item, class, direction): max, sum of sqrt(D2/4) in float64
//       (fixed order), and the two order statistics of numpy's linear 95th percentile by an 8-bit
//       radix select whose histogram updates are aggregated with MATCH.ANY (no atomics, no barriers).
constexpr int kSearchThreads = 256;
#ifndef OCTM_SEARCH_MINB
#define OCTM_SEARCH_MINB 6
#endif

struct SearchParams {
    const uint32_t* verts;   // [n][K][2][max_pts]
    const uint32_t* n_pts;   // [n][K][2]
    long long n_units;       // n * K * 2
    int max_pts;
    int tile;                // source vertices staged per pass, multiple of kBox
    uint32_t* d2;            // [n][K][2][max_pts]
    uint32_t* max_sq;        // [n][K][2]      kNeedsSelect when the unit is left to distance_select_kernel
    uint32_t* p95_sq;        // [n][K][2][2]
    double* sum_dist;        // [n][K][2]
    bool keep_d2;            // store every unit's squared distances as well
    bool only_marked;        // search only the units whose max_sq is kNeedsSearch
    const uint32_t* todo_count;
};

constexpr int kCountBins = 1024;                    // squared distances 0, 2, .. 2046 are counted in shared memory
constexpr uint32_t kNeedsSelect = 0xffffffffu;      // not a squared distance (coordinates are below 2^14)

// Counting form of the per-unit statistics (shared by both search kernels).  Squared distances between contour
// vertices are even (a vertex has exactly one odd doubled-lattice coordinate), so value 2 h is counted in 16-bit
// counter h, two counters per word.  Equal values of a warp are merged with MATCH.ANY: one shared-memory atomic
// per distinct value.  The running maximum and a flag for values the counters cannot hold stay in registers.
__device__ __forceinline__ void count_minima(int bestd, bool valid, int lane, uint32_t* s_bins, uint32_t& run_max,
                                             bool& run_bad) {
    const uint32_t dv = static_cast<uint32_t>(bestd), h = dv >> 1;
    const bool ok = valid && (dv & 1u) == 0 && h < static_cast<uint32_t>(kCountBins);
    const uint32_t peers = __match_any_sync(0xffffffffu, ok ? h : static_cast<uint32_t>(kCountBins) + lane);
    if (ok && lane == __ffs(peers) - 1) atomicAdd(&s_bins[h >> 1], static_cast<uint32_t>(__popc(peers)) << ((h & 1u) * 16));
    if (valid) run_max = max(run_max, dv);
    run_bad |= valid && !ok;
}

// A warp publishes what count_minima gathered over its chunks (before the CTA barrier that completes the counters).
__device__ __forceinline__ void publish_counts(uint32_t run_max, bool run_bad, int lane, uint32_t* s_vmax, uint32_t* s_big) {
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, run_max);
    const bool bad = __any_sync(0xffffffffu, run_bad);
    if (lane == 0) {
        atomicMax(s_vmax, wmax);
        if (bad) *s_big = 1;
    }
}

// One warp: maximum, the two neighbours of numpy's linear 95th percentile and sum of sqrt(D2 / 4) (float64, one
// square root per distinct value, fixed order) of the nq counted values; the counters are left cleared.
__device__ __forceinline__ void stats_from_counters(uint32_t* s_bins, uint32_t vmax, int nq, int lane, uint32_t* max_sq,
                                                    uint32_t* p95_sq, double* sum_dist) {
    const double pos = __dmul_rn(static_cast<double>(nq - 1), 0.95);     // virtual index (m - 1) * 0.95
    const uint32_t lo = static_cast<uint32_t>(floor(pos));
    const uint32_t hi = lo + 1 < static_cast<uint32_t>(nq) ? lo + 1 : lo;
    uint32_t v_lo = 0, v_hi = 0, base = 0;
    double dsum = 0.0;
    for (uint32_t h0 = 0; h0 <= (vmax >> 1); h0 += 32) {
        const uint32_t h = h0 + lane;
        const uint32_t cnt = (s_bins[h >> 1] >> ((h & 1u) * 16)) & 0xffffu;
        __syncwarp();
        if ((lane & 1) == 0) s_bins[h >> 1] = 0;
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        const uint32_t first = base + incl - cnt;                    // ranks [first, first + cnt) hold value 2 h
        const uint32_t m_lo = __ballot_sync(0xffffffffu, cnt != 0 && first <= lo && lo < first + cnt);
        const uint32_t m_hi = __ballot_sync(0xffffffffu, cnt != 0 && first <= hi && hi < first + cnt);
        if (m_lo) v_lo = 2 * (h0 + __ffs(m_lo) - 1);
        if (m_hi) v_hi = 2 * (h0 + __ffs(m_hi) - 1);
        if (cnt) dsum = __dadd_rn(dsum, __dmul_rn(static_cast<double>(cnt), sqrt(static_cast<double>(2 * h) / 4.0)));
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dsum = __dadd_rn(dsum, __shfl_xor_sync(0xffffffffu, dsum, o));
    if (lane == 0) {
        *max_sq = vmax;
        p95_sq[0] = v_lo;
        p95_sq[1] = v_hi;
        *sum_dist = dsum;
    }
}

__device__ __forceinline__ void coop_scan_box(const int4* src4, int bb, int ns, int cy, int cx, int& bm) {
    const int i0 = bb * kBox;
    if (i0 + kBox <= ns) {
#pragma unroll
        for (int i = 0; i < kBox; i += 2) {
            const int4 s0 = src4[i0 + i], s1 = src4[i0 + i + 1];
            bm = min(bm, min(s0.y * cx + (s0.x * cy + s0.z), s1.y * cx + (s1.x * cy + s1.z)));
        }
    } else {
        for (int i = i0; i < ns; ++i) {
            const int4 s0 = src4[i];
            bm = min(bm, s0.y * cx + (s0.x * cy + s0.z));
        }
    }
}

__global__ void __launch_bounds__(kSearchThreads, OCTM_SEARCH_MINB) distance_search_kernel(const SearchParams prm) {
    extern __shared__ __align__(16) uint8_t dsm[];
    __shared__ int s_next;
    __shared__ uint32_t s_bins[kCountBins / 2];      // two 16-bit counters per word: count of squared distance 2 * h
    __shared__ uint32_t s_vmax, s_big;
    if (prm.todo_count != nullptr && *prm.todo_count == 0) return;
    for (int i = threadIdx.x; i < kCountBins / 2; i += kSearchThreads) s_bins[i] = 0;
    if (threadIdx.x == 0) s_vmax = s_big = 0;
    const int cap = prm.max_pts, tile = prm.tile;
    int4* const src4 = reinterpret_cast<int4*>(dsm);                                          // {y, x, y^2+x^2, -}
    int4* const boxes = reinterpret_cast<int4*>(dsm + static_cast<size_t>(tile) * 16);       // {ymin, ymax, xmin, xmax}
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (long long unit = blockIdx.x; unit < prm.n_units; unit += gridDim.x) {
        const long long pair = unit >> 1;
        const int dir = static_cast<int>(unit & 1);
        // direction 0: queries = pred vertices (map 1), sources = true vertices (map 0); direction 1 swapped
        if (prm.only_marked && prm.max_sq[unit] != kNeedsSearch) continue;      // CTA-uniform
        const int ns_all = static_cast<int>(min(prm.n_pts[pair * 2 + dir], static_cast<uint32_t>(cap)));
        const int nq = static_cast<int>(min(prm.n_pts[pair * 2 + 1 - dir], static_cast<uint32_t>(cap)));
        if (ns_all == 0 || nq == 0) {
            if (threadIdx.x == 0) {
                prm.max_sq[unit] = 0;
                prm.p95_sq[unit * 2] = prm.p95_sq[unit * 2 + 1] = 0;
                prm.sum_dist[unit] = 0.0;
            }
            continue;
        }
        const uint32_t* vs = prm.verts + (pair * 2 + dir) * static_cast<long long>(cap);
        const uint32_t* vq = prm.verts + (pair * 2 + 1 - dir) * static_cast<long long>(cap);
        uint32_t* dq = prm.d2 + unit * static_cast<long long>(cap);
        const float ratio = static_cast<float>(ns_all) / static_cast<float>(nq);
        const int nchunks = (nq + 31) >> 5;
        {   // the next unit of this CTA: pull its vertex counts and both vertex lists towards L2 while this one is
            // searched (the per-unit prologue otherwise waits on DRAM: 28 % of the stall samples)
            const long long nu = unit + gridDim.x;
            if (nu < prm.n_units) {
                const char* nv = reinterpret_cast<const char*>(prm.verts + (nu >> 1) * 2 * static_cast<long long>(cap));
                const int lines = (2 * cap * 4 + 127) / 128;                 // both maps of the pair
                for (int i = threadIdx.x; i < lines; i += kSearchThreads)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nv + static_cast<long long>(i) * 128));
                if (threadIdx.x == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(prm.n_pts + (nu >> 1) * 2));
            }
        }

        for (int t0 = 0; t0 < ns_all; t0 += tile) {
            const int ns = min(tile, ns_all - t0);
            const int nb = (ns + kBox - 1) / kBox;
            __syncthreads();                       // every warp is done with the previous tile
            for (int i = threadIdx.x; i < ns; i += kSearchThreads) {
                const uint32_t v = vs[t0 + i];
                const int y = v >> 16, x = v & 0xffff;
                src4[i] = make_int4(y, x, y * y + x * x, 0);
            }
            __syncthreads();
            if (threadIdx.x == 0) { s_next = 0; s_vmax = 0; s_big = 0; }
            for (int b = threadIdx.x; b < nb; b += kSearchThreads) {
                int ymin = 0x7fffffff, ymax = -1, xmin = 0x7fffffff, xmax = -1;
                const int e = min(ns, (b + 1) * kBox);
                for (int i = b * kBox; i < e; ++i) {
                    const int4 sv = src4[i];
                    ymin = min(ymin, sv.x); ymax = max(ymax, sv.x); xmin = min(xmin, sv.y); xmax = max(xmax, sv.y);
                }
                boxes[b] = make_int4(ymin, ymax, xmin, xmax);
            }
            __syncthreads();
            // On the last tile the minima are final: they are counted (16-bit counters, one per even value) instead
            // of stored, and the unit's statistics come from the counters.  A unit with a value the counters cannot
            // hold repeats the tile's search in store mode and is left to distance_select_kernel.
            const bool last = t0 + tile >= ns_all;
            bool count = last && nq <= 0xffff;
            uint32_t run_max = 0;
            bool run_bad = false;
          for (;;) {
            // chunks are handed out dynamically (shared counter): their cost varies with the local geometry
            for (;;) {
                int c = 0;
                if (lane == 0) c = atomicAdd(&s_next, 1);
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c >= nchunks) break;
                const int j = c * 32 + lane;
                const uint32_t v = vq[min(j, nq - 1)];          // tail lanes repeat the last query
                const int qy = v >> 16, qx = v & 0xffff;
                const int cy = -2 * qy, cx = -2 * qx, qn = qy * qy + qx * qx;
                const int q_ymin = __reduce_min_sync(0xffffffffu, qy), q_ymax = __reduce_max_sync(0xffffffffu, qy);
                const int q_xmin = __reduce_min_sync(0xffffffffu, qx), q_xmax = __reduce_max_sync(0xffffffffu, qx);
                // prime every lane's bound: first tile -> the box at the same relative position along the other
                // contour (clamped into the tile); later tiles -> the minimum over the tiles before
                int bg = -1, bm, bestd;
                if (t0 == 0) {
                    bg = min(min(ns_all - 1, static_cast<int>(static_cast<float>(c * 32 + 16) * ratio)) / kBox, nb - 1);
                    bm = 0x7fffffff;
                    coop_scan_box(src4, bg, ns, cy, cx, bm);
                    bestd = bm + qn;
                } else {
                    bestd = static_cast<int>(dq[min(j, nq - 1)]);
                    bm = bestd - qn;
                }
                int bmax = __reduce_max_sync(0xffffffffu, bestd);
                // box phase, shared by the warp: each lane tests a different source box against the bounding
                // box of the 32 queries; candidates are then re-tested per lane and scanned by all lanes
                for (int b0 = 0; b0 < nb; b0 += 32) {
                    const int b = b0 + lane;
                    bool cand = false;
                    if (b < nb && b != bg) {
                        const int4 bx = boxes[b];
                        const int dy = max(max(bx.x - q_ymax, q_ymin - bx.y), 0);
                        const int dx = max(max(bx.z - q_xmax, q_xmin - bx.w), 0);
                        cand = dy * dy + dx * dx < bmax;
                    }
                    uint32_t mask = __ballot_sync(0xffffffffu, cand);
                    if (mask == 0) continue;
                    while (mask) {
                        const int bb = b0 + __ffs(mask) - 1;
                        mask &= mask - 1;
                        if (!__any_sync(0xffffffffu, box_lb2(boxes[bb], qy, qx) < bestd)) continue;
                        coop_scan_box(src4, bb, ns, cy, cx, bm);
                        bestd = bm + qn;
                    }
                    bmax = __reduce_max_sync(0xffffffffu, bestd);
                }
                if ((!count || prm.keep_d2) && j < nq) dq[j] = static_cast<uint32_t>(bestd);
                if (!count) continue;
                count_minima(bestd, j < nq, lane, s_bins, run_max, run_bad);
            }
            if (!count) break;
            publish_counts(run_max, run_bad, lane, &s_vmax, &s_big);
            __syncthreads();                       // counters, s_vmax, s_big complete
            if (s_big == 0) break;
            count = false;                         // CTA-uniform: search the tile again, storing
            if (threadIdx.x == 0) s_next = 0;
            __syncthreads();
          }
            if (!last) continue;
            if (warp != 0) continue;
            if (!count) {                          // stored: hand the unit to the select kernel
                if (lane == 0) prm.max_sq[unit] = kNeedsSelect;
                if (nq <= 0xffff)
                    for (int i = lane; i < kCountBins / 2; i += 32) s_bins[i] = 0;
                continue;
            }
            // warp 0: statistics from the counters (cleared on the way); the next unit's first counter update is
            // two CTA barriers away
            stats_from_counters(s_bins, s_vmax, nq, lane, prm.max_sq + unit, prm.p95_sq + unit * 2, prm.sum_dist + unit);
        }
    }
}

// ------------------------------------------------------------------------------ column-sorted search
// Default search.  One CTA per (item, class, direction).  The source contour is counting-sorted by its doubled
// lattice column x2 into shared memory (int4 {y, x, y^2 + x^2, -}; col[c] .. col[c + 1] delimit column c).  A
// warp owns 32 consecutive query vertices spanning the columns [x0, x1]:
//   1. all lanes scan the sources of the columns [x0, x1] together (one broadcast LDS.128 per source vertex,
//      2 IMAD + half a 3-input minimum per lane and vertex);
//   2. r = isqrt(max over lanes of the best squared distance - 1) bounds every lane's remaining search: a
//      nearer vertex lies less than sqrt(best) columns from the query, hence within [x0 - r, x1 + r].  The two
//      flanks are scanned the same way.
// The minimum is exact (every vertex that can beat a lane's best lies in the scanned columns; scanning a few
// vertices more never hurts: loops run in steps of four over the sorted array, which is padded with copies
// of a real vertex).  No boxes, no per-candidate tests: ~40 vertex evaluations per query on layered B-scans
// against ~70 for the tiled search, and the worst case (contours far apart) is a brute-force scan.
constexpr int kColThreads = 256;
#ifndef OCTM_COL_MINB
#define OCTM_COL_MINB 5
#endif

struct ColumnParams {
    const uint32_t* verts;   // [n][K][2][max_pts]
    const uint32_t* n_pts;   // [n][K][2]
    long long n_units;       // n * K * 2
    int max_pts;
    int ncol;                // doubled-lattice columns: x2 < ncol (2 W)
    uint32_t* d2;            // [n][K][2][max_pts]
    uint32_t* max_sq;        // [n][K][2]      kNeedsSelect when the unit is left to distance_select_kernel
    uint32_t* p95_sq;        // [n][K][2][2]
    double* sum_dist;        // [n][K][2]
    bool keep_d2;
    bool only_marked;        // search only the units whose max_sq is kNeedsSearch (the rest is already finished)
    const uint32_t* todo_count;   // device word: 0 = no unit is marked, leave at once (or null)
};

// Query groups: kColGroup consecutive lanes share one column span (their sources are read from one address per
// group: 32 / kColGroup distinct LDS.128 addresses per instruction, mostly in different banks).  Smaller groups
// have narrower spans and a smaller flank radius (fewer vertex evaluations per query), at the price of shuffle
// reductions inside the group.
#ifndef OCTM_COL_GROUP
#define OCTM_COL_GROUP 2
#endif
constexpr int kColGroup = OCTM_COL_GROUP;

__device__ __forceinline__ int group_min(int v) {
    if (kColGroup == 32) return __reduce_min_sync(0xffffffffu, v);
#pragma unroll
    for (int o = kColGroup / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int group_max(int v) {
    if (kColGroup == 32) return __reduce_max_sync(0xffffffffu, v);
#pragma unroll
    for (int o = kColGroup / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Queries per lane: a lane holds kColQ consecutive query vertices, so one broadcast LDS.128 of a source vertex
// feeds kColQ evaluations and the per-chunk bookkeeping (spans, look-ups, loops) is shared by 32 * kColQ queries.
#ifndef OCTM_COL_Q
#define OCTM_COL_Q 2
#endif
constexpr int kColQ = OCTM_COL_Q;

// All lanes fold the sorted source vertices [a, a + len) of THEIR group into bm[]; the trip count is the warp's
// maximum, so a group with a shorter range runs on into vertices it does not need (harmless: every entry is a
// real vertex).  nsp = padded size of the sorted array; the start moves left when the 4-wide steps would
// run past it.
__device__ __forceinline__ void coop_scan_range(const int4* src, int a, int len, int nsp, const int (&cy)[kColQ],
                                                const int (&cx)[kColQ], int (&bm)[kColQ]) {
    const int n = (__reduce_max_sync(0xffffffffu, len) + 3) & ~3;
    const int4* p = src + min(a, nsp - n);
#pragma unroll 1
    for (int i = 0; i < n; i += 4) {
        const int4 s0 = p[i], s1 = p[i + 1], s2 = p[i + 2], s3 = p[i + 3];
#pragma unroll
        for (int q = 0; q < kColQ; ++q) {
            const int m01 = min(s0.y * cx[q] + (s0.x * cy[q] + s0.z), s1.y * cx[q] + (s1.x * cy[q] + s1.z));
            const int m23 = min(s2.y * cx[q] + (s2.x * cy[q] + s2.z), s3.y * cx[q] + (s3.x * cy[q] + s3.z));
            bm[q] = min(bm[q], min(m01, m23));
        }
    }
}

// The search proper: every warp takes the query chunks c = warp, warp + 8, ... of the unit.  COUNT: the minima go
// to the shared-memory counters; STORE: they are written to the d2 scratch.  Instantiated three times so that the
// usual pass (count, do not store) carries neither the store's address arithmetic nor its registers.
// Chunks are dealt out statically: a shared counter balances the warps slightly better but costs more than it
// saves (measured 3.96 vs 3.84 ms).
template <bool COUNT, bool STORE>
__device__ __forceinline__ void column_chunks(const int4* src, const uint32_t* col, const uint32_t* vq, uint32_t* dq,
                                              int ns, int nq, int ncol, int nchunks, int warp, int lane, uint32_t* s_bins,
                                              uint32_t& run_max, bool& run_bad) {
    constexpr int kWarps = kColThreads / 32;
#pragma unroll 1
    for (int c = warp; c < nchunks; c += kWarps) {
        // lane l holds the consecutive queries c * 32 Q + l Q .. + Q - 1 (tail lanes repeat the last query)
        const int j0 = (c * 32 + lane) * kColQ;
        int qx[kColQ], cy[kColQ], cx[kColQ], qn[kColQ], bm[kColQ];
#pragma unroll
        for (int q = 0; q < kColQ; ++q) {
            const uint32_t v = vq[min(j0 + q, nq - 1)];
            const int qy = v >> 16;
            qx[q] = v & 0xffff;
            cy[q] = -2 * qy; cx[q] = -2 * qx[q]; qn[q] = qy * qy + qx[q] * qx[q];
            bm[q] = 0x3fffffff;                // best of s.y * cy + s.x * cx + |s|^2 ( = d^2 - |q|^2 )
        }
        // Consecutive polyline vertices are at most 2 columns apart; a larger step is the seam between the
        // forward and the backward run of a walked open contour (or the closing repeat): the queries before
        // and after it get their own column span, so that no span covers the columns in between.
        // first = index (in this chunk's 32 Q queries) of the first query after the seam.
        int first = 32 * kColQ;
        {
            const int qprev = __shfl_up_sync(0xffffffffu, qx[kColQ - 1], 1);
#pragma unroll
            for (int q = 0; q < kColQ; ++q) {
                const bool jump = q == 0 ? (lane > 0 && abs(qx[0] - qprev) > 2) : abs(qx[q] - qx[q - 1]) > 2;
                const uint32_t m = __ballot_sync(0xffffffffu, jump);
                if (m) first = min(first, (__ffs(m) - 1) * kColQ + q);
            }
        }
        for (int part = 0;; ++part) {          // part 0: queries before `first`, part 1: the rest
            int x0 = 0x7fffffff, x1 = -1, bmine = 0;
#pragma unroll
            for (int q = 0; q < kColQ; ++q) {
                const bool mine = ((lane * kColQ + q) < first) == (part == 0);
                x0 = min(x0, mine ? qx[q] : 0x7fffffff);
                x1 = max(x1, mine ? qx[q] : -1);
            }
            x0 = min(group_min(x0), ncol - 1);
            x1 = min(group_max(x1), ncol - 1);
            const bool has = x1 >= 0;          // this group has queries in this part
            const int lo = has ? static_cast<int>(col[x0]) : 0, hi = has ? static_cast<int>(col[x1 + 1]) : 0;
            coop_scan_range(src, lo, hi - lo, ns + 3, cy, cx, bm);
#pragma unroll
            for (int q = 0; q < kColQ; ++q) {
                const bool mine = ((lane * kColQ + q) < first) == (part == 0);
                bmine = max(bmine, mine ? bm[q] + qn[q] : 0);
            }
            const int bmax = group_max(bmine);
            // columns that can still hold a nearer vertex: dx^2 < best  =>  dx <= isqrt(best - 1) <= r
            // (the approximate square root is exact enough below 2^20; real distances are below 2^29)
            float rf;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(static_cast<float>(bmax)));
            const int r = bmax >= 0x3fffffff ? ncol : static_cast<int>(rf) + (bmax >= (1 << 20) ? 1 : 0);
            const int fl = has ? static_cast<int>(col[max(x0 - r, 0)]) : 0;
            const int fr = has ? static_cast<int>(col[min(x1 + r, ncol - 1) + 1]) : 0;
            coop_scan_range(src, fl, lo - fl, ns + 3, cy, cx, bm);
            coop_scan_range(src, hi, fr - hi, ns + 3, cy, cx, bm);
            if (part == 1 || first == 32 * kColQ) break;
        }
        #pragma unroll
        for (int q = 0; q < kColQ; ++q) {
            const int bestd = bm[q] + qn[q];
            const bool valid = j0 + q < nq;
            if (STORE && valid) dq[j0 + q] = static_cast<uint32_t>(bestd);
            if (COUNT) count_minima(bestd, valid, lane, s_bins, run_max, run_bad);
        }
    }
}

__global__ void __launch_bounds__(kColThreads, OCTM_COL_MINB) distance_column_kernel(const ColumnParams prm) {
    extern __shared__ __align__(16) uint8_t dsm[];
    __shared__ uint32_t s_bins[kCountBins / 2];
    __shared__ uint32_t s_vmax, s_big;
    __shared__ uint32_t s_wtot[kColThreads / 32];
    const int cap = prm.max_pts, ncol = prm.ncol;
    if (prm.todo_count != nullptr && *prm.todo_count == 0) return;
    int4* const src = reinterpret_cast<int4*>(dsm);                                                 // cap + 4 entries
    uint32_t* const col = reinterpret_cast<uint32_t*>(dsm + (static_cast<size_t>(cap) + 4) * 16);   // ncol + 1 entries
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kColThreads / 32;
    const int per = (ncol + kColThreads - 1) / kColThreads;      // columns per lane in the prefix sum
    for (int i = threadIdx.x; i < kCountBins / 2; i += kColThreads) s_bins[i] = 0;
    bool synced = false;                               // CTA-uniform: the previous unit ended with a CTA barrier

    for (long long unit = blockIdx.x; unit < prm.n_units; unit += gridDim.x) {
        const long long pair = unit >> 1;
        const int dir = static_cast<int>(unit & 1);
        // direction 0: queries = pred vertices (map 1), sources = true vertices (map 0); direction 1 swapped
        if (prm.only_marked && prm.max_sq[unit] != kNeedsSearch) continue;      // CTA-uniform
        const int ns = static_cast<int>(min(prm.n_pts[pair * 2 + dir], static_cast<uint32_t>(cap)));
        const int nq = static_cast<int>(min(prm.n_pts[pair * 2 + 1 - dir], static_cast<uint32_t>(cap)));
        if (ns == 0 || nq == 0) {
            if (threadIdx.x == 0) {
                prm.max_sq[unit] = 0;
                prm.p95_sq[unit * 2] = prm.p95_sq[unit * 2 + 1] = 0;
                prm.sum_dist[unit] = 0.0;
            }
            continue;
        }
        const uint32_t* vs = prm.verts + (pair * 2 + dir) * static_cast<long long>(cap);
        const uint32_t* vq = prm.verts + (pair * 2 + 1 - dir) * static_cast<long long>(cap);
        uint32_t* dq = prm.d2 + unit * static_cast<long long>(cap);
        const int nchunks = (nq + 32 * kColQ - 1) / (32 * kColQ);
        {   // pull the next unit's vertex lists towards L2 while this one is searched
            const long long nu = unit + gridDim.x;
            if (nu < prm.n_units) {
                const char* nv = reinterpret_cast<const char*>(prm.verts + (nu >> 1) * 2 * static_cast<long long>(cap));
                const int lines = (2 * cap * 4 + 127) / 128;
                for (int i = threadIdx.x; i < lines; i += kColThreads)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nv + static_cast<long long>(i) * 128));
                if (threadIdx.x == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(prm.n_pts + (nu >> 1) * 2));
            }
        }
        if (!synced) __syncthreads();                  // every warp is done with the previous unit's tables
        if (threadIdx.x == 0) { s_vmax = 0; s_big = 0; }
        // A contour that is already in column order (what trace_layered_kernel emits) needs no sort: one pass copies
        // the vertices and fills col[] from the places where the column changes -- col[c] = number of vertices left
        // of column c.  Done optimistically; the vote below tells whether the order held.
        bool sorted = true;
        {
            const int xfirst = min(static_cast<int>(vs[0] & 0xffffu), ncol - 1);
            const int xlast = min(static_cast<int>(vs[ns - 1] & 0xffffu), ncol - 1);
            for (int i = threadIdx.x; i < ns; i += kColThreads) {
                const uint32_t v = vs[i];
                const int y = v >> 16, xr = v & 0xffff, x = min(xr, ncol - 1);
                const int xp = i > 0 ? min(static_cast<int>(vs[i - 1] & 0xffffu), ncol - 1) : x;
                sorted &= x >= xp;
                src[i] = make_int4(y, xr, y * y + xr * xr, 0);
                if (i < 3) src[ns + i] = make_int4(y, xr, y * y + xr * xr, 0);   // padding for the 4-wide scans
                for (int c = xp + 1; c <= x; ++c) col[c] = static_cast<uint32_t>(i);
            }
            if (threadIdx.x < 3 && threadIdx.x >= ns) {                      // fewer than three vertices
                const uint32_t v = vs[0];
                const int y = v >> 16, x = v & 0xffff;
                src[ns + threadIdx.x] = make_int4(y, x, y * y + x * x, 0);
            }
            for (int c = threadIdx.x; c <= xfirst; c += kColThreads) col[c] = 0;
            for (int c = xlast + 1 + threadIdx.x; c <= ncol; c += kColThreads) col[c] = static_cast<uint32_t>(ns);
        }
        if (!__syncthreads_and(sorted)) {
        for (int i = threadIdx.x; i <= ncol; i += kColThreads) col[i] = 0;
        __syncthreads();
        // counting sort by column: count into col[x + 1] ...
        for (int i = threadIdx.x; i < ns; i += kColThreads) {
            const int x = min(static_cast<int>(vs[i] & 0xffffu), ncol - 1);
            atomicAdd(&col[x + 1], 1u);
        }
        __syncthreads();
        // ... exclusive prefix sum in place (col[x + 1] = first slot of column x): every warp scans a contiguous
        // block of 32 * per columns, then adds the totals of the warps before it ...
        {
            const int base = 1 + warp * 32 * per;
            uint32_t carry = 0;
            for (int k = 0; k < per; ++k) {
                const int i = base + k * 32 + lane;
                const uint32_t c = i <= ncol ? col[i] : 0u;
                uint32_t incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                if (i <= ncol) col[i] = carry + incl - c;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) s_wtot[warp] = carry;
            __syncthreads();
            uint32_t off = 0;
            for (int w = 0; w < warp; ++w) off += s_wtot[w];
            if (off)
                for (int k = 0; k < per; ++k) {
                    const int i = base + k * 32 + lane;
                    if (i <= ncol) col[i] += off;
                }
        }
        __syncthreads();
        // ... and scatter: afterwards col[x + 1] is the END of column x, so column x is col[x] .. col[x + 1]
        for (int i = threadIdx.x; i < ns; i += kColThreads) {
            const uint32_t v = vs[i];
            const int y = v >> 16, x = v & 0xffff;
            const uint32_t slot = atomicAdd(&col[min(x, ncol - 1) + 1], 1u);
            src[slot] = make_int4(y, x, y * y + x * x, 0);
            if (i < 3) src[ns + i] = make_int4(y, x, y * y + x * x, 0);      // padding for the 4-wide scans
        }
        if (threadIdx.x < 3 && threadIdx.x >= ns) {                          // fewer than three vertices
            const uint32_t v = vs[0];
            const int y = v >> 16, x = v & 0xffff;
            src[ns + threadIdx.x] = make_int4(y, x, y * y + x * x, 0);
        }
        __syncthreads();
        }

        // The final minima are counted (16-bit counters, one per even value) instead of stored; a unit with a value
        // the counters cannot hold repeats the search in store mode and is left to distance_select_kernel.
        bool count = nq <= 0xffff;
        if (count) {
            uint32_t run_max = 0;
            bool run_bad = false;
            if (prm.keep_d2) column_chunks<true, true>(src, col, vq, dq, ns, nq, ncol, nchunks, warp, lane, s_bins, run_max, run_bad);
            else column_chunks<true, false>(src, col, vq, dq, ns, nq, ncol, nchunks, warp, lane, s_bins, run_max, run_bad);
            publish_counts(run_max, run_bad, lane, &s_vmax, &s_big);
            __syncthreads();                           // counters, s_vmax, s_big complete
            count = s_big == 0;                        // CTA-uniform
        }
        if (!count) {                                  // search (again), storing
            uint32_t run_max = 0;
            bool run_bad = false;
            column_chunks<false, true>(src, col, vq, dq, ns, nq, ncol, nchunks, warp, lane, s_bins, run_max, run_bad);
        }
        synced = count;                                // the barrier above also ends every warp's use of the tables
        if (warp != 0) continue;
        if (!count) {                                  // stored: hand the unit to the select kernel
            if (lane == 0) prm.max_sq[unit] = kNeedsSelect;
            if (nq <= 0xffff)
                for (int i = lane; i < kCountBins / 2; i += 32) s_bins[i] = 0;
            continue;
        }
        stats_from_counters(s_bins, s_vmax, nq, lane, prm.max_sq + unit, prm.p95_sq + unit * 2, prm.sum_dist + unit);
    }
}

struct SelectParams {
    const uint32_t* n_pts;   // [n][K][2]
    const uint32_t* d2;      // [n][K][2][max_pts]
    long long n_units;
    int max_pts;
    uint32_t* max_sq;        // [n][K][2]
    uint32_t* p95_sq;        // [n][K][2][2]
    double* sum_dist;        // [n][K][2]
    const uint32_t* todo_count;
};

// k-th smallest (0-based) of vals[0..m): 8-bit radix passes over a warp-private histogram; every lane
// returns the value.  Equal digits inside one 32-value slice are merged with MATCH.ANY, so the
// histogram update is one plain read-modify-write per distinct digit.
__device__ uint32_t warp_select(const uint32_t* vals, int m, uint32_t k, uint32_t vmax, uint32_t* hist /*256*/, int lane) {
    uint32_t prefix = 0, maskbits = 0;
    int shift = vmax >= (1u << 24) ? 24 : (vmax >= (1u << 16) ? 16 : (vmax >= (1u << 8) ? 8 : 0));
    for (; shift >= 0; shift -= 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) hist[i * 32 + lane] = 0;
        __syncwarp();
        for (int j0 = 0; j0 < m; j0 += 32) {
            const int j = j0 + lane;
            const uint32_t v = j < m ? vals[j] : 0u;
            const bool ok = j < m && (v & maskbits) == prefix;
            const uint32_t d = (v >> shift) & 255u;
            const uint32_t peers = __match_any_sync(0xffffffffu, ok ? d : 256u + lane);
            if (ok && lane == __ffs(peers) - 1) hist[d] += __popc(peers);
            __syncwarp();
        }
        uint32_t c[8], tot = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { c[i] = hist[lane * 8 + i]; tot += c[i]; }
        uint32_t incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t nb = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += nb;
        }
        const uint32_t excl = incl - tot;
        const bool mine = k >= excl && k < incl;                 // exactly one lane
        uint32_t digit = 0, knew = 0;
        if (mine) {
            uint32_t run = excl;
            int d = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (k >= run + c[i]) { run += c[i]; d = i + 1; }
                else break;
            }
            digit = static_cast<uint32_t>(lane * 8 + d);
            knew = k - run;
        }
        const int src = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
        digit = __shfl_sync(0xffffffffu, digit, src);
        k = __shfl_sync(0xffffffffu, knew, src);
        prefix |= digit << shift;
        maskbits |= 255u << shift;
        __syncwarp();
    }
    return prefix;
}

__global__ void __launch_bounds__(128) distance_select_kernel(const SelectParams prm) {
    __shared__ uint32_t s_hist[4][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cap = prm.max_pts;
    if (prm.todo_count != nullptr && *prm.todo_count == 0) return;
    for (long long unit = static_cast<long long>(blockIdx.x) * 4 + warp; unit < prm.n_units;
         unit += static_cast<long long>(gridDim.x) * 4) {
        const long long pair = unit >> 1;
        const int dir = static_cast<int>(unit & 1);
        if (prm.max_sq[unit] != kNeedsSelect) continue;         // finished by the search kernel (the usual case)
        const int nq = static_cast<int>(min(prm.n_pts[pair * 2 + 1 - dir], static_cast<uint32_t>(cap)));
        const uint32_t* dq = prm.d2 + unit * static_cast<long long>(cap);
        uint32_t vmax = 0;
        double dsum = 0.0;
        {   // four loads in flight per lane; the partial sums are folded in a fixed order
            double ds[4] = {0.0, 0.0, 0.0, 0.0};
            int j = lane;
            for (; j + 96 < nq; j += 128) {
                uint32_t v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = dq[j + 32 * u];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    vmax = max(vmax, v[u]);
                    ds[u] += sqrt(static_cast<double>(v[u]) / 4.0);
                }
            }
            double dt = 0.0;
            for (; j < nq; j += 32) {
                const uint32_t v = dq[j];
                vmax = max(vmax, v);
                dt += sqrt(static_cast<double>(v) / 4.0);
            }
            dsum = ((ds[0] + ds[1]) + (ds[2] + ds[3])) + dt;
        }
        vmax = __reduce_max_sync(0xffffffffu, vmax);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
        // numpy linear percentile: virtual index (m - 1) * 0.95, neighbours floor and floor + 1
        const double pos = __dmul_rn(static_cast<double>(nq - 1), 0.95);
        const uint32_t lo = static_cast<uint32_t>(floor(pos));
        const uint32_t v_lo = warp_select(dq, nq, lo, vmax, s_hist[warp], lane);
        uint32_t cnt_le = 0, next_gt = 0xffffffffu;
#pragma unroll 4
        for (int j = lane; j < nq; j += 32) {
            const uint32_t v = dq[j];
            cnt_le += v <= v_lo ? 1u : 0u;
            if (v > v_lo) next_gt = min(next_gt, v);
        }
        cnt_le = __reduce_add_sync(0xffffffffu, cnt_le);
        next_gt = __reduce_min_sync(0xffffffffu, next_gt);
        const uint32_t v_hi = (lo + 1 >= static_cast<uint32_t>(nq) || cnt_le >= lo + 2) ? v_lo : next_gt;
        if (lane == 0) {
            prm.max_sq[unit] = vmax;
            prm.p95_sq[unit * 2 + 0] = v_lo;
            prm.p95_sq[unit * 2 + 1] = v_hi;
            prm.sum_dist[unit] = dsum;
        }
    }
}


// ------------------------------------------------------------------------------ layered pairs: fused distances
// When BOTH contours of an (item, class) pair were verified as height functions (trace_layered_kernel<false>),
// neither vertex list ever reaches HBM.  One WARP per pair keeps, for each map, two int16 tables over the doubled
// lattice columns c in [0, 2 W - 1) in shared memory -- the contour's vertices of column c are the lattice rows
// lo[c], lo[c] + 2, .., hi[c]:
//     c = 2 x      one vertex          lo = hi = 2 h(x) - 1
//     c = 2 x + 1  |h(x+1) - h(x)|     lo = 2 min(h(x), h(x+1)), hi = 2 max(..) - 2;  none: lo = 32767, hi = -32768
// (built from the label pass's boundary rows: 4 KB per map for W = 512).  The squared distance from a query vertex
// (qy, qx) to the vertices of source column c is  max(lo - qy, qy - hi, |c - qx| & 1)^2 + (c - qx)^2  (same column
// parity = same row parity, so an in-range query hits a vertex exactly; the other parity misses by one row), i.e.
// every source COLUMN costs a constant, whatever its vertex count.  A query scans the columns qx, qx -+ 1, qx -+ 2,
// .. until d^2 can no longer beat its best: exact, and ~4 x (contour distance in pixels) columns per query.
// Queries are the vertices of the other map's tables: the W even-column vertices straight from the lanes (lane =
// column), the odd-column runs compacted through a per-warp ring so that every round works on 32 real queries.
// The minima are counted like in the vertex-list search (count_minima / stats_from_counters).  Pairs this kernel
// cannot finish -- a side that has to be walked, a distance the counters cannot hold -- get their verified sides'
// vertices emitted from the tables (column order) and both units marked kNeedsSearch for distance_column_kernel.
constexpr int kLdWarps = 4;           // one CTA = one (item, class) pair: the tables are shared, the query blocks dealt out
constexpr int kLdPad = 32;            // empty columns on both sides of a table: the scan needs no clamping up to d = 32
constexpr int kLdRing = 512;          // odd-column query ring per warp (entries)
constexpr int kLdShort = 64;          // vertex lists up to this length are searched by brute force in the fused kernel
#ifndef OCTM_LD_MINB
#define OCTM_LD_MINB 8
#endif

struct LayeredDistParams {
    const uint32_t* first_pos;   // [n][2][K]
    const int* bnd_t;            // [n][K-1][W]
    const int* bnd_p;
    const uint32_t* unsorted;    // [n] bit m: map m of the item has a column that is not in class order (label pass)
    long long n_pairs;           // n * K
    int H, W, K, max_pts;
    int tab;                     // int16 entries per table: 2 W - 1 + 2 kLdPad, rounded up to a multiple of 8
    uint32_t* verts;             // [n][K][2][max_pts]  written only for pairs left to the vertex-list search
    uint32_t* n_pts;             // [n][K][2]  out: vertex count, 0 (no contour) or kTraceTodo (left to the fallback kernels)
    uint32_t* max_sq;            // [n][K][2]  out; kNeedsSearch for the units left to the vertex-list search
    uint32_t* p95_sq;            // [n][K][2][2]
    double* sum_dist;            // [n][K][2]
    uint32_t* todo_count;        // PASS 1: += 1 per pair handed on (zeroed by the host)
    uint32_t* search_count;      // += 1 per pair left to the vertex-list search (zeroed by the host)
    int vals_cap;                // PASS 2: entries of the per-CTA value list behind the rings (0: none, recompute instead)
};

__device__ __forceinline__ int max3i(int a, int b, int c) { return max(max(a, b), c); }
__device__ __forceinline__ int min3i(int a, int b, int c) { return min(min(a, b), c); }

// Nearest vertices of the tabulated contour (lo / hi point at column 0 of the tables) to TWO queries per lane: the
// two scans share one loop (twice the loads in flight per pass, half the loop overhead per query).  A query whose
// scan is over keeps folding in real distances, which is harmless.
__device__ __forceinline__ void layered_nearest2p(const short* lo0, const short* hi0, const short* lo1, const short* hi1, int qy0,
                                                  int qx0, int qy1, int qx1, int ncol, int& best0, int& best1) {
    const short* l0 = lo0 + qx0;
    const short* h0 = hi0 + qx0;
    const short* l1 = lo1 + qx1;
    const short* h1 = hi1 + qx1;
    const int dy0 = max3i(l0[0] - qy0, qy0 - h0[0], 0), dy1 = max3i(l1[0] - qy1, qy1 - h1[0], 0);
    best0 = dy0 * dy0;
    best1 = dy1 * dy1;
    int d = 1;
    // columns qx -+ d, two distances per pass (odd d: the other row parity, a miss by at least one row)
    while (d * d < max(best0, best1) && d < kLdPad) {
        const int d2 = d * d, e2 = (d + 1) * (d + 1);
        {
            const int a0 = max3i(l0[-d] - qy0, qy0 - h0[-d], 1), a1 = max3i(l0[d] - qy0, qy0 - h0[d], 1);
            const int b0 = max3i(l0[-d - 1] - qy0, qy0 - h0[-d - 1], 0), b1 = max3i(l0[d + 1] - qy0, qy0 - h0[d + 1], 0);
            best0 = min3i(best0, a0 * a0 + d2, a1 * a1 + d2);
            best0 = min3i(best0, b0 * b0 + e2, b1 * b1 + e2);
        }
        {
            const int a0 = max3i(l1[-d] - qy1, qy1 - h1[-d], 1), a1 = max3i(l1[d] - qy1, qy1 - h1[d], 1);
            const int b0 = max3i(l1[-d - 1] - qy1, qy1 - h1[-d - 1], 0), b1 = max3i(l1[d + 1] - qy1, qy1 - h1[d + 1], 0);
            best1 = min3i(best1, a0 * a0 + d2, a1 * a1 + d2);
            best1 = min3i(best1, b0 * b0 + e2, b1 * b1 + e2);
        }
        d += 2;
    }
    // contours further apart than the pad (rare): the same scan with clamped columns (the pads are empty columns)
    for (int e = d; e * e < best0; ++e) {
        const int cl = max(qx0 - e, -1), cr = min(qx0 + e, ncol), par = e & 1;
        const int a0 = max3i(lo0[cl] - qy0, qy0 - hi0[cl], par), a1 = max3i(lo0[cr] - qy0, qy0 - hi0[cr], par);
        best0 = min3i(best0, a0 * a0 + e * e, a1 * a1 + e * e);
    }
    for (int e = d; e * e < best1; ++e) {
        const int cl = max(qx1 - e, -1), cr = min(qx1 + e, ncol), par = e & 1;
        const int a0 = max3i(lo1[cl] - qy1, qy1 - hi1[cl], par), a1 = max3i(lo1[cr] - qy1, qy1 - hi1[cr], par);
        best1 = min3i(best1, a0 * a0 + e * e, a1 * a1 + e * e);
    }
}
__device__ __forceinline__ void layered_nearest2(const short* lo, const short* hi, int qy0, int qx0, int qy1, int qx1,
                                                 int ncol, int& best0, int& best1) {
    layered_nearest2p(lo, hi, lo, hi, qy0, qx0, qy1, qx1, ncol, best0, best1);
}

// Nearest vertices of a SHORT vertex list (shared memory, {y, x} pairs) to two queries: brute force.
__device__ __forceinline__ void list_nearest2(const int2* pts, int n, int qy0, int qx0, int qy1, int qx1, int& best0, int& best1) {
    best0 = best1 = 0x7fffffff;
#pragma unroll 2
    for (int i = 0; i < n; ++i) {
        const int2 v = pts[i];                                   // one address for the whole warp: a broadcast
        const int dy0 = v.x - qy0, dx0 = v.y - qx0, dy1 = v.x - qy1, dx1 = v.y - qx1;
        best0 = min(best0, dy0 * dy0 + dx0 * dx0);
        best1 = min(best1, dy1 * dy1 + dx1 * dx1);
    }
}

// One side of a pair as the fused kernel sees it: a height-function table or a short vertex list.
struct LdSide {
    const short* lo;      // table (column 0), or null
    const short* hi;
    const int2* pts;      // short list, or null
    const int* best;      // list side: squared distance of every vertex to the other side (ld_list_minima)
    int n;                // vertices of the side
};

// Squared distance of every vertex of a short list to the other side, by the whole CTA: a far blob's search radius is
// its distance (hundreds of columns), so each thread takes every 128th source column (or vertex) of every query and
// the minima meet in shared memory.  best[] must hold 0x7fffffff on entry.
template <bool SRC_TABLE>
__device__ __forceinline__ void ld_list_minima(const int2* qpts, int nq, const LdSide& src, int ncol, int tid, int* best) {
    for (int q = 0; q < nq; ++q) {
        const int2 v = qpts[q];
        int b = 0x7fffffff;
        if (SRC_TABLE) {
            for (int c = tid; c < ncol; c += kLdWarps * 32) {
                const int dx = c - v.y;
                const int dy = max3i(src.lo[c] - v.x, v.x - src.hi[c], dx & 1);
                b = min(b, dy * dy + dx * dx);
            }
        } else {
            for (int i = tid; i < src.n; i += kLdWarps * 32) {
                const int dy = src.pts[i].x - v.x, dx = src.pts[i].y - v.y;
                b = min(b, dy * dy + dx * dx);
            }
        }
        b = __reduce_min_sync(0xffffffffu, b);
        if ((tid & 31) == 0) atomicMin(&best[q], b);
    }
}

template <bool SRC_TABLE>
__device__ __forceinline__ void ld_nearest2(const LdSide& src, int qy0, int qx0, int qy1, int qx1, int ncol, int& b0, int& b1) {
    if (SRC_TABLE) layered_nearest2(src.lo, src.hi, qy0, qx0, qy1, qx1, ncol, b0, b1);
    else list_nearest2(src.pts, src.n, qy0, qx0, qy1, qx1, b0, b1);
}

// All vertices of `qry` against `src`; the minima are counted in `bins`.  Table queries: the column blocks are dealt
// out to the CTA's warps (lane = column for the even-column vertices, the odd-column runs compacted through the
// warp's ring); list queries: warp 0 takes them, two per lane.
// What happens to a query's squared distance.  FineCounter: the usual counting (values < 2048 in 16-bit counters).
// WideCounter adds the two roles that serve directions whose distances do not fit (a blob far from its layer, layers
// far apart): a radix select over RECOMPUTED distances -- "coarse" counts value >> shift (and sums the square roots,
// one per query), "window" counts, at full resolution, the values of the one coarse bin that holds the percentile's
// order statistics and keeps the smallest value above that bin.
struct FineCounter {
    uint32_t* bins;
    uint32_t run_max = 0;
    bool run_bad = false, overflow = false;
    __device__ __forceinline__ void add(int best, bool valid, int lane) { count_minima(best, valid, lane, bins, run_max, run_bad); }
};
struct WideCounter {          // PASS 2 only: one code path for the three roles, chosen at run time (CTA-uniform)
    uint32_t* bins;
    int mode;                 // 0 fine, 1 coarse, 2 window
    int shift;
    uint32_t target;
    uint32_t run_max = 0, next_min = 0xffffffffu;
    double sum = 0.0;
    bool run_bad = false, overflow = false;
    uint32_t* vals = nullptr;     // coarse role: every value is also appended here (shared memory; any order), so that the
    uint32_t* vals_n = nullptr;   // window role can run over the stored values instead of recomputing them
    __device__ __forceinline__ void add(int best, bool valid, int lane) {
        if (mode == 0) {
            count_minima(best, valid, lane, bins, run_max, run_bad);
            return;
        }
        if (mode == 1 && vals != nullptr) {
            const uint32_t m = __ballot_sync(0xffffffffu, valid);
            uint32_t base = 0;
            if (lane == 0 && m) base = atomicAdd(vals_n, static_cast<uint32_t>(__popc(m)));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (valid) vals[base + __popc(m & lanemask_lt())] = static_cast<uint32_t>(best);
        }
        const uint32_t v = static_cast<uint32_t>(best), key = v >> shift;
        if (valid) run_max = max(run_max, v);
        const bool in = valid && (mode == 1 || key == target);
        // coarse: value >> shift (< kCountBins by the choice of shift); window: the even values of the target bin
        const uint32_t idx = mode == 1 ? key : (v & ((1u << shift) - 1u)) >> 1;
        const uint32_t peers = __match_any_sync(0xffffffffu, in ? idx : static_cast<uint32_t>(kCountBins) + lane);
        if (in && lane == __ffs(peers) - 1) atomicAdd(&bins[idx >> 1], static_cast<uint32_t>(__popc(peers)) << ((idx & 1u) * 16));
        if (mode == 1) {
            if (valid) sum += sqrt(static_cast<double>(best) / 4.0);
        } else if (valid && key > target) {
            next_min = min(next_min, v);
        }
    }
};

template <bool QRY_TABLE, bool SRC_TABLE, class Counter>
__device__ __forceinline__ void ld_direction(const LdSide& qry, const LdSide& src, int W, int ncol, int warp, int lane,
                                             uint32_t* ring, Counter& ctr) {
    if (!QRY_TABLE) {
        // list queries: their minima were found once by ld_list_minima (all threads of the CTA); warp 0 counts them
        if (warp == 0) {
            const bool v0 = lane < qry.n, v1 = lane + 32 < qry.n;
            ctr.add(qry.best[v0 ? lane : 0], v0, lane);
            ctr.add(qry.best[v1 ? lane + 32 : 0], v1, lane);
        }
        return;
    }
    const short* qlo = qry.lo;
    const short* qhi = qry.hi;
    const int nblk = (W + 63) >> 6;
    uint32_t head = 0, tail = 0;
    // a query that costs (next to) nothing for idle slots: a vertex of the source itself
    const int idle_y = SRC_TABLE ? src.lo[0] : src.pts[0].x, idle_x = SRC_TABLE ? 0 : src.pts[0].y;
    for (int blk = warp; blk < nblk; blk += kLdWarps) {          // 64 columns: lane takes x and x + 32
        const int xa = blk * 64 + lane, xb = xa + 32;
        const bool va = xa < W, vb = xb < W;
        // the even-column vertices of the two columns
        int ba, bb;
        ld_nearest2<SRC_TABLE>(src, va ? qlo[2 * xa] : idle_y, va ? 2 * xa : idle_x, vb ? qlo[2 * xb] : idle_y, vb ? 2 * xb : idle_x, ncol, ba, bb);
        ctr.add(ba, va, lane);
        ctr.add(bb, vb, lane);
        // the runs between columns x and x + 1 join the ring
        const bool ra = xa + 1 < W, rb = xb + 1 < W;
        const int la = ra ? qlo[2 * xa + 1] : 32767, ha = ra ? qhi[2 * xa + 1] : -32768;
        const int lb = rb ? qlo[2 * xb + 1] : 32767, hb = rb ? qhi[2 * xb + 1] : -32768;
        const uint32_t lena = ha >= la ? static_cast<uint32_t>((ha - la) >> 1) + 1u : 0u;
        const uint32_t lenb = hb >= lb ? static_cast<uint32_t>((hb - lb) >> 1) + 1u : 0u;
        const uint32_t len = lena + lenb;
        uint32_t incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total > static_cast<uint32_t>(kLdRing - 64)) { ctr.overflow = true; break; }   // absurdly steep: general path
        uint32_t at = tail + incl - len;
        for (uint32_t t = 0; t < lena; ++t, ++at)
            ring[at & (kLdRing - 1)] = (static_cast<uint32_t>(la + 2 * static_cast<int>(t)) << 16) | static_cast<uint32_t>(2 * xa + 1);
        for (uint32_t t = 0; t < lenb; ++t, ++at)
            ring[at & (kLdRing - 1)] = (static_cast<uint32_t>(lb + 2 * static_cast<int>(t)) << 16) | static_cast<uint32_t>(2 * xb + 1);
        tail += total;
        __syncwarp();
        while (tail - head >= 64u) {
            const uint32_t q0 = ring[(head + lane) & (kLdRing - 1)], q1 = ring[(head + 32 + lane) & (kLdRing - 1)];
            head += 64u;
            int b0, b1;
            ld_nearest2<SRC_TABLE>(src, static_cast<int>(q0 >> 16), static_cast<int>(q0 & 0xffffu), static_cast<int>(q1 >> 16),
                                   static_cast<int>(q1 & 0xffffu), ncol, b0, b1);
            ctr.add(b0, true, lane);
            ctr.add(b1, true, lane);
        }
        __syncwarp();
    }
    if (tail != head) {
        const uint32_t left = tail - head;
        const bool v0 = static_cast<uint32_t>(lane) < left, v1 = static_cast<uint32_t>(lane) + 32u < left;
        const uint32_t q0 = ring[(head + (v0 ? lane : 0)) & (kLdRing - 1)];
        const uint32_t q1 = v1 ? ring[(head + 32 + lane) & (kLdRing - 1)] : q0;
        int b0, b1;
        ld_nearest2<SRC_TABLE>(src, static_cast<int>(q0 >> 16), static_cast<int>(q0 & 0xffffu), static_cast<int>(q1 >> 16),
                               static_cast<int>(q1 & 0xffffu), ncol, b0, b1);
        ctr.add(b0, v0, lane);
        ctr.add(b1, v1, lane);
    }
}

// Statistics of the (at most 64) minima of a short list, by one warp, without counters: maximum, the two order
// statistics of numpy's linear 95th percentile (by rank counting) and the sum of sqrt(D2 / 4) in a fixed order.
__device__ __forceinline__ void ld_list_stats(const int* best, int n, int lane, uint32_t* max_sq, uint32_t* p95_sq, double* sum_dist) {
    const bool v0 = lane < n, v1 = lane + 32 < n;
    const uint32_t a = v0 ? static_cast<uint32_t>(best[lane]) : 0u, b = v1 ? static_cast<uint32_t>(best[lane + 32]) : 0u;
    uint32_t ra = 0, rb = 0;                  // rank = values below + equal values with a smaller index
    for (int i = 0; i < n; ++i) {
        const uint32_t w = static_cast<uint32_t>(best[i]);
        ra += (w < a || (w == a && i < lane)) ? 1u : 0u;
        rb += (w < b || (w == b && i < lane + 32)) ? 1u : 0u;
    }
    const double pos = __dmul_rn(static_cast<double>(n - 1), 0.95);
    const uint32_t lo = static_cast<uint32_t>(floor(pos)), hi = lo + 1 < static_cast<uint32_t>(n) ? lo + 1 : lo;
    uint32_t v_lo = 0, v_hi = 0;
    uint32_t m = __ballot_sync(0xffffffffu, v0 && ra == lo);
    if (m) v_lo = __shfl_sync(0xffffffffu, a, __ffs(m) - 1);
    m = __ballot_sync(0xffffffffu, v1 && rb == lo);
    if (m) v_lo = __shfl_sync(0xffffffffu, b, __ffs(m) - 1);
    m = __ballot_sync(0xffffffffu, v0 && ra == hi);
    if (m) v_hi = __shfl_sync(0xffffffffu, a, __ffs(m) - 1);
    m = __ballot_sync(0xffffffffu, v1 && rb == hi);
    if (m) v_hi = __shfl_sync(0xffffffffu, b, __ffs(m) - 1);
    const uint32_t vmax = __reduce_max_sync(0xffffffffu, max(a, b));
    double ds = __dadd_rn(v0 ? sqrt(static_cast<double>(a) / 4.0) : 0.0, v1 ? sqrt(static_cast<double>(b) / 4.0) : 0.0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ds = __dadd_rn(ds, __shfl_xor_sync(0xffffffffu, ds, o));
    if (lane == 0) {
        *max_sq = vmax;
        p95_sq[0] = v_lo;
        p95_sq[1] = v_hi;
        *sum_dist = ds;
    }
}

// PASS 2, one table and one short list: the direction whose QUERIES are the table's vertices, by the whole CTA without
// rings or counters.  Thread t takes columns t, t + 128, ..: the even-column vertex and the run to the next column, every
// query by brute force over the (at most 64) list vertices; the squared distances go to `vals` in shared memory (any
// order), the sum of square roots and the maximum are kept per thread.  Order statistics by an 8-bit radix select over
// the stored values (block_select, as in the single-kernel check modes).
__device__ __forceinline__ void ld_table_vs_list(const short* qlo, const short* qhi, const int2* pts, int n, int W, int nq, int tid,
                                                 uint32_t* vals, uint32_t* s_nvals, uint32_t* s_scr /*>= 8*/, double* s_dscr /*>= 8*/,
                                                 uint32_t* s_hist /*256*/, uint32_t* s_bc /*2*/, uint32_t* max_sq, uint32_t* p95_sq,
                                                 double* sum_dist) {
    uint32_t vmax = 0;
    double dsum = 0.0;
    auto emit = [&](int qy, int qx) {
        int best = 0x7fffffff;
        for (int i = 0; i < n; ++i) {
            const int2 v = pts[i];
            const int dy = v.x - qy, dx = v.y - qx;
            best = min(best, dy * dy + dx * dx);
        }
        vals[atomicAdd(s_nvals, 1u)] = static_cast<uint32_t>(best);
        vmax = max(vmax, static_cast<uint32_t>(best));
        dsum = __dadd_rn(dsum, sqrt(static_cast<double>(best) / 4.0));
    };
    for (int x = tid; x < W; x += kLdWarps * 32) {
        emit(qlo[2 * x], 2 * x);
        if (x + 1 < W)
            for (int y = qlo[2 * x + 1], yh = qhi[2 * x + 1]; y <= yh; y += 2) emit(y, 2 * x + 1);
    }
    vmax = block_reduce_max<kLdWarps * 32>(vmax, s_scr);                  // (barriers inside: vals complete afterwards)
    dsum = block_reduce_add<kLdWarps * 32>(dsum, s_dscr);
    const double pos = __dmul_rn(static_cast<double>(nq - 1), 0.95);      // numpy linear percentile
    const uint32_t lo = static_cast<uint32_t>(floor(pos));
    const uint32_t v_lo = block_select<kLdWarps * 32>(vals, nq, lo, vmax, s_hist, s_bc);
    uint32_t cnt_le = 0, next_gt = 0xffffffffu;
    for (int j = tid; j < nq; j += kLdWarps * 32) {
        const uint32_t v = vals[j];
        cnt_le += v <= v_lo ? 1u : 0u;
        if (v > v_lo) next_gt = min(next_gt, v);
    }
    cnt_le = block_reduce_add<kLdWarps * 32>(cnt_le, s_scr);
    next_gt = block_reduce_min<kLdWarps * 32>(next_gt, s_scr);
    const uint32_t v_hi = (lo + 1 >= static_cast<uint32_t>(nq) || cnt_le >= lo + 2) ? v_lo : next_gt;
    if (tid == 0) {
        *max_sq = vmax;
        p95_sq[0] = v_lo;
        p95_sq[1] = v_hi;
        *sum_dist = dsum;
    }
}

// PASS 1, both sides tables: BOTH directions in one go.  The even-column vertices of a direction are taken lane = column as
// in ld_direction; the odd-column runs of both directions share the warp's ring (bit 31 of an entry = direction), so a
// warp ends with ONE partially filled round instead of one per direction (the half-empty rounds were 10 % of the kernel).
// side 0 = y_true's tables at tabs, side 1 = y_pred's at tabs + 2 tab; direction d: queries = side 1 - d, sources = side d.
__device__ __forceinline__ void ld_both_tt(const short* tabs, int tab, int W, int ncol, int warp, int lane, uint32_t* ring,
                                           uint32_t* bins2, uint32_t* s_vmax, uint32_t* s_bad) {
    uint32_t rm0 = 0, rm1 = 0, head = 0, tail = 0;
    bool bad0 = false, bad1 = false, overflow = false;
    const int nblk = (W + 63) >> 6;
    // count a squared distance of direction dirq (per lane) like count_minima, the two directions' counters side by side
    auto count = [&](int bestd, bool valid, uint32_t dirq) {
        const uint32_t dv = static_cast<uint32_t>(bestd), h = dv >> 1;
        const bool ok = valid && (dv & 1u) == 0 && h < static_cast<uint32_t>(kCountBins);
        const uint32_t peers = __match_any_sync(0xffffffffu, ok ? (h | (dirq << 16)) : 0x100000u + lane);
        if (ok && lane == __ffs(peers) - 1)
            atomicAdd(&bins2[dirq * (kCountBins / 2) + (h >> 1)], static_cast<uint32_t>(__popc(peers)) << ((h & 1u) * 16));
        if (valid) {
            if (dirq) rm1 = max(rm1, dv); else rm0 = max(rm0, dv);
            if (!ok) { if (dirq) bad1 = true; else bad0 = true; }
        }
    };
    auto ring_round = [&](uint32_t left) {
        const bool v0 = static_cast<uint32_t>(lane) < left, v1 = static_cast<uint32_t>(lane) + 32u < left;
        // idle slots repeat a source vertex of direction 0 (distance 0 at d = 0)
        const uint32_t idle = static_cast<uint32_t>(static_cast<unsigned short>(tabs[kLdPad])) << 16;
        const uint32_t q0 = v0 ? ring[(head + lane) & (kLdRing - 1)] : idle, q1 = v1 ? ring[(head + 32 + lane) & (kLdRing - 1)] : idle;
        head += left < 64u ? left : 64u;
        const uint32_t d0 = q0 >> 31, d1 = q1 >> 31;
        const short* lo0 = tabs + kLdPad + d0 * (2 * tab);
        const short* lo1 = tabs + kLdPad + d1 * (2 * tab);
        int b0, b1;
        layered_nearest2p(lo0, lo0 + tab, lo1, lo1 + tab, static_cast<int>((q0 >> 16) & 0x7fffu), static_cast<int>(q0 & 0xffffu),
                          static_cast<int>((q1 >> 16) & 0x7fffu), static_cast<int>(q1 & 0xffffu), ncol, b0, b1);
        count(b0, v0, d0);
        count(b1, v1, d1);
    };
#pragma unroll 1
    for (int dir = 0; dir < 2 && !overflow; ++dir) {
        const short* slo = tabs + kLdPad + dir * (2 * tab);
        const short* shi = slo + tab;
        const short* qlo = tabs + kLdPad + (1 - dir) * (2 * tab);
        const short* qhi = qlo + tab;
        const uint32_t tag = static_cast<uint32_t>(dir) << 31;
        const int idle_y = slo[0];
        for (int blk = warp; blk < nblk; blk += kLdWarps) {          // 64 columns: lane takes x and x + 32
            const int xa = blk * 64 + lane, xb = xa + 32;
            const bool va = xa < W, vb = xb < W;
            int ba, bb;
            layered_nearest2(slo, shi, va ? qlo[2 * xa] : idle_y, va ? 2 * xa : 0, vb ? qlo[2 * xb] : idle_y, vb ? 2 * xb : 0, ncol, ba, bb);
            count(ba, va, static_cast<uint32_t>(dir));
            count(bb, vb, static_cast<uint32_t>(dir));
            const bool ra = xa + 1 < W, rb = xb + 1 < W;
            const int la = ra ? qlo[2 * xa + 1] : 32767, ha = ra ? qhi[2 * xa + 1] : -32768;
            const int lb = rb ? qlo[2 * xb + 1] : 32767, hb = rb ? qhi[2 * xb + 1] : -32768;
            const uint32_t lena = ha >= la ? static_cast<uint32_t>((ha - la) >> 1) + 1u : 0u;
            const uint32_t lenb = hb >= lb ? static_cast<uint32_t>((hb - lb) >> 1) + 1u : 0u;
            const uint32_t len = lena + lenb;
            uint32_t incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (total > static_cast<uint32_t>(kLdRing - 64)) { overflow = true; break; }   // absurdly steep: general path
            uint32_t at = tail + incl - len;
            for (uint32_t t = 0; t < lena; ++t, ++at)
                ring[at & (kLdRing - 1)] = tag | (static_cast<uint32_t>(la + 2 * static_cast<int>(t)) << 16) | static_cast<uint32_t>(2 * xa + 1);
            for (uint32_t t = 0; t < lenb; ++t, ++at)
                ring[at & (kLdRing - 1)] = tag | (static_cast<uint32_t>(lb + 2 * static_cast<int>(t)) << 16) | static_cast<uint32_t>(2 * xb + 1);
            tail += total;
            __syncwarp();
            while (tail - head >= 64u) ring_round(64u);
            __syncwarp();
        }
    }
    if (!overflow && tail != head) ring_round(tail - head);
    const uint32_t w0 = __reduce_max_sync(0xffffffffu, rm0), w1 = __reduce_max_sync(0xffffffffu, rm1);
    const bool b0 = __any_sync(0xffffffffu, bad0), b1 = __any_sync(0xffffffffu, bad1), ov = __any_sync(0xffffffffu, overflow);
    if (lane == 0) {
        atomicMax(&s_vmax[0], w0);
        atomicMax(&s_vmax[1], w1);
        if (b0 || b1 || ov) atomicOr(s_bad, (b0 ? 1u : 0u) | (b1 ? 2u : 0u) | (ov ? 4u : 0u));
    }
}

// The vertices of a table in column order -> verts (what trace_layered_kernel<true> would have written).
__device__ __forceinline__ void ld_emit(const short* lo, const short* hi, int ncol, int lane, uint32_t* out, uint32_t cap) {
    uint32_t base = 0;
    for (int c0 = 0; c0 < ncol; c0 += 32) {
        const int c = c0 + lane;
        const int l = c < ncol ? lo[c] : 32767, h = c < ncol ? hi[c] : -32768;
        const uint32_t len = h >= l ? static_cast<uint32_t>((h - l) >> 1) + 1u : 0u;
        uint32_t incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        uint32_t at = base + incl - len;
        for (uint32_t t = 0; t < len; ++t, ++at)
            if (at < cap) out[at] = (static_cast<uint32_t>(l + 2 * static_cast<int>(t)) << 16) | static_cast<uint32_t>(c);
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// PASS 1: every pair; verification from the certificate + boundary rows; pairs with two verified sides are measured,
//         the others are handed on (n_pts: verified side = count | kLayeredBit, other side kTraceTodo; units marked).
// PASS 2: after trace_layered_kernel<false> (verification against label pixels) and the walk have settled the
//         kTraceTodo sides: the marked pairs again.  A side is now a table (kLayeredBit) or a vertex list; table x
//         table, table x short list and short x short list are measured here, the rest gets its tables emitted as
//         vertex lists and stays marked for distance_column_kernel.
template <int PASS>
__global__ void __launch_bounds__(kLdWarps * 32, PASS != 2 ? OCTM_LD_MINB : 8) layered_distance_kernel(const LayeredDistParams prm) {
    extern __shared__ __align__(16) uint8_t dsm[];
    __shared__ uint32_t s_vmax[2], s_bad, s_ok[2], s_minkey[2], s_cnt[2], s_seed[2];
    __shared__ int2 s_list[2][kLdShort];
    __shared__ int s_lbest[2][kLdShort];
    __shared__ double s_dsum[kLdWarps];
    __shared__ double s_dsum8[8];
    __shared__ uint32_t s_next, s_tgt[3], s_amax, s_nvals;
    // PASS 3 = PASS 1 for a stream of predictions the certificate mostly rejects: their contours went through the pixel
    // verification BEFORE this pass, and what it verified (a re-centred row in the contour's vertex slot, kRowBit) is a
    // table side like a certified one -- lightly noisy pairs are then measured here instead of by the slower PASS 2.
    // PASS 1 proper stays free of all that (the clean path).
    constexpr bool P1 = PASS == 1 || PASS == 3;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = prm.W, K = prm.K, tab = prm.tab, ncol = 2 * W - 1;
    if (PASS == 2 && *prm.todo_count == 0) return;
    short* const tabs = reinterpret_cast<short*>(dsm);                                   // [map][lo | hi][tab]
    uint32_t* const bins2 = reinterpret_cast<uint32_t*>(dsm + static_cast<size_t>(tab) * 8);   // [dir][kCountBins / 2]
    uint32_t* const ring = bins2 + kCountBins + warp * kLdRing;                          // this warp's ring
    for (int i = tid; i < kCountBins; i += kLdWarps * 32) bins2[i] = 0;
    for (int i = tid; i < 2 * tab; i += kLdWarps * 32) {                                 // the pads never change
        const int m = i >= tab ? 1 : 0, j = i - m * tab;
        if (j < kLdPad || j >= kLdPad + ncol) {
            tabs[(2 * m) * tab + j] = 32767;
            tabs[(2 * m + 1) * tab + j] = -32768;
        }
    }

    for (long long pair0 = blockIdx.x; pair0 < prm.n_pairs; pair0 += gridDim.x) {
      // PASS 1: the class of pixel (0, 0) and the class after it are bounded by the SAME boundary row (#{label <= c00}
      // from below, #{label < c00 + 1} from above).  When both maps start with the same class, the CTA of that class
      // also owns the next class's pair: if its (different) verification holds as well, the tables and therefore all
      // distances are identical and the results are copied; otherwise the pair is measured in a second turn.
      bool own_next = false;             // CTA-uniform
      for (int rep = 0; rep < 2; ++rep) {
        if (rep == 1 && !own_next) break;
        const long long pair = pair0 + rep;
        if (PASS == 2 && prm.max_sq[pair * 2] != kNeedsSearch) continue;                 // CTA-uniform
        __syncthreads();                     // the previous pair is finished with the tables, the counters and s_*
        const long long item = pair / K;
        const int cls = static_cast<int>(pair - item * K);
        if (P1 && rep == 0) {
            const uint32_t ft = lane < K ? prm.first_pos[(item * 2 + 0) * K + lane] : OCTM_NO_SEED;
            const uint32_t fq = lane < K ? prm.first_pos[(item * 2 + 1) * K + lane] : OCTM_NO_SEED;
            const int c00t = __ffs(__ballot_sync(0xffffffffu, ft == 0u)) - 1, c00p = __ffs(__ballot_sync(0xffffffffu, fq == 0u)) - 1;
            if (c00t >= 0 && c00t == c00p && (PASS != 3 || (prm.unsorted[item] & 3u) == 0u)) {
                if (cls == c00t + 1) break;                  // owned by the CTA of pair - 1
                own_next = cls == c00t && cls + 1 < K;
            }
        }
        if (tid < 2) { s_vmax[tid] = 0; s_ok[tid] = 1; s_minkey[tid] = 0xffffffffu; s_cnt[tid] = 0; }
        if (tid == 2) s_bad = 0;
        const uint32_t unsorted = P1 ? prm.unsorted[item] : 0u;
        uint32_t n0 = 0, n1 = 0;
        if (PASS == 2 || (PASS == 3 && (unsorted & 1u))) n0 = prm.n_pts[pair * 2];
        if (PASS == 2 || (PASS == 3 && (unsorted & 2u))) n1 = prm.n_pts[pair * 2 + 1];
        const bool row0 = PASS == 3 && (unsorted & 1u) && n0 != kTraceTodo && (n0 & kRowBit);
        const bool row1 = PASS == 3 && (unsorted & 2u) && n1 != kTraceTodo && (n1 & kRowBit);
        __syncthreads();
        // ---- verification + tables, warps 0-1: map 0, warps 2-3: map 1.  On a map whose columns are all in class
        // order (the label pass's certificate) contour [0] of a class mask is the height function h(x) = #{label <
        // k} of one boundary row iff  (a) 1 <= h <= H - 1,  (b) wherever the polyline steps between neighbouring
        // columns the class really lies on its side of the step -- a condition on the NEXT boundary row (band not
        // thinner than the step) or, for the class of pixel (0, 0), on the previous one --, and (c) the raster-first
        // pixel of the path is the seed the label pass found.  All of it is arithmetic on the boundary rows: what
        // trace_layered_kernel establishes by reading label pixels, without touching the label maps.
        // (PASS 2: the sides flagged kLayeredBit were verified before; only their tables are built.)
        {
            const int m = warp >> 1;
            const uint32_t nm = m ? n1 : n0;
            const bool rowside = PASS == 3 && ((unsorted >> m) & 1u) && nm != kTraceTodo && (nm & kRowBit);
            const bool want = P1 || (nm != kTraceTodo && (nm & kLayeredBit));
            const uint32_t* fp = prm.first_pos + (item * 2 + m) * K;
            const uint32_t myfp = lane < K ? fp[lane] : OCTM_NO_SEED;
            const int c00 = __ffs(__ballot_sync(0xffffffffu, myfp == 0u)) - 1;
            const bool inv = cls == c00;                                   // the class is the region above the path
            const uint32_t others = __reduce_min_sync(0xffffffffu, lane == c00 ? OCTM_NO_SEED : myfp);
            const uint32_t seed = inv ? others : __shfl_sync(0xffffffffu, myfp, cls);
            if (lane == 0 && (warp & 1) == 0) s_seed[m] = seed;
            const int brow = inv ? cls : cls - 1;
            if (want && seed != OCTM_NO_SEED && (!((unsorted >> m) & 1u) || rowside) && brow >= 0 && brow < K - 1) {
                const int* rows = (m ? prm.bnd_p : prm.bnd_t) + item * (K - 1) * static_cast<long long>(W);
                const int* hrow = rows + brow * static_cast<long long>(W);
                if ((PASS == 2 && (nm & kRowBit)) || rowside)      // verified against the label pixels with a re-centred row
                    hrow = reinterpret_cast<const int*>(prm.verts + (pair * 2 + m) * static_cast<long long>(prm.max_pts));
                // the row that bounds the class on the far side of the path: the next boundary (band thickness) for a
                // class below the path, the previous one for the class of pixel (0, 0); null = nothing to check
                const int* orow = inv ? (cls > 0 ? rows + (cls - 1) * static_cast<long long>(W) : nullptr)
                                      : (cls < K - 1 ? rows + cls * static_cast<long long>(W) : nullptr);
                if (PASS == 2 || rowside) orow = nullptr;           // (nothing left to check: the defaults below pass)
                uint32_t* lo32 = reinterpret_cast<uint32_t*>(tabs + (2 * m) * tab + kLdPad);
                uint32_t* hi32 = lo32 + (tab >> 1);
                bool ok = true;
                uint32_t minkey = 0xffffffffu, steps = 0;
                const int H = prm.H;
                const bool vec4 = (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(hrow) | reinterpret_cast<uintptr_t>(orow)) & 15) == 0;
                // four adjacent columns per lane (16-byte loads of the rows, one 16-byte store per table): a warp takes 128
                // columns per turn; the columns left and right of a lane's four come from its neighbours
                for (int x0 = (warp & 1) * 128; vec4 && x0 < W; x0 += 256) {
                    const int x = x0 + 4 * lane;
                    const int xc = min(x, W - 4);
                    const int4 hv = *reinterpret_cast<const int4*>(hrow + xc);
                    const int4 ov = orow != nullptr ? *reinterpret_cast<const int4*>(orow + xc) : (inv ? make_int4(0, 0, 0, 0) : make_int4(H, H, H, H));
                    int hl = __shfl_up_sync(0xffffffffu, hv.w, 1), hn = __shfl_down_sync(0xffffffffu, hv.x, 1);
                    if (lane == 0) hl = hrow[max(xc - 1, 0)];
                    if (lane == 31) hn = hrow[min(xc + 4, W - 1)];
                    if (x + 4 >= W) hn = hv.w;
                    if (x < W) {
                        const int hh[6] = {hl, hv.x, hv.y, hv.z, hv.w, hn};
                        const int oo[4] = {ov.x, ov.y, ov.z, ov.w};
                        uint32_t wl[4], wh[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int h = hh[i + 1], hp = hh[i], hx = hh[i + 2];
                            if (P1) {
                                const int lo_w = min(hp, min(h, hx)), hi_w = max(hp, max(h, hx));
                                ok = ok && h >= 1 && h <= H - 1 && (inv ? oo[i] <= lo_w - 1 : oo[i] >= hi_w + 1);
                                minkey = min(minkey, static_cast<uint32_t>(h) * static_cast<uint32_t>(W) + static_cast<uint32_t>(x + i));
                                steps += static_cast<uint32_t>(abs(hx - h));
                            }
                            const uint32_t e = static_cast<uint32_t>(2 * h - 1) & 0xffffu;
                            const bool run = hx != h;
                            wl[i] = e | ((run ? static_cast<uint32_t>(2 * min(h, hx)) & 0xffffu : 32767u) << 16);
                            wh[i] = e | ((run ? static_cast<uint32_t>(2 * max(h, hx) - 2) & 0xffffu : 0x8000u) << 16);
                        }
                        // (the right neighbour of the last column is the column itself: no run, as in the scalar loop)
                        *reinterpret_cast<uint4*>(lo32 + x) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
                        *reinterpret_cast<uint4*>(hi32 + x) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
                    }
                }
                for (int x0 = (warp & 1) * 32; !vec4 && x0 < W; x0 += 64) {
                    const int x = x0 + lane;
                    const int xc = min(x, W - 1);
                    const int h = hrow[xc];
                    int hl = __shfl_up_sync(0xffffffffu, h, 1), hn = __shfl_down_sync(0xffffffffu, h, 1);
                    if (lane == 0) hl = hrow[max(xc - 1, 0)];
                    if (lane == 31) hn = hrow[min(xc + 1, W - 1)];
                    if (x >= W - 1) hn = h;
                    const int o = orow != nullptr ? orow[xc] : (inv ? 0 : H);
                    if (x < W) {
                        if (P1) {
                            const int lo_w = min(hl, min(h, hn)), hi_w = max(hl, max(h, hn));
                            ok = ok && h >= 1 && h <= H - 1 && (inv ? o <= lo_w - 1 : o >= hi_w + 1);
                            minkey = min(minkey, static_cast<uint32_t>(h) * static_cast<uint32_t>(W) + static_cast<uint32_t>(x));
                            steps += static_cast<uint32_t>(abs(hn - h));
                        }
                        // columns 2 x and 2 x + 1 as one 32-bit store per table (column 2 W - 1 is a pad)
                        const uint32_t e = static_cast<uint32_t>(2 * h - 1) & 0xffffu;
                        const bool run = hn != h;
                        const uint32_t ol = run ? static_cast<uint32_t>(2 * min(h, hn)) & 0xffffu : 32767u;
                        const uint32_t oh = run ? static_cast<uint32_t>(2 * max(h, hn) - 2) & 0xffffu : 0x8000u;
                        lo32[x] = e | (ol << 16);
                        hi32[x] = e | (oh << 16);
                    }
                }
                if (P1) {
                    ok = __all_sync(0xffffffffu, ok);
                    minkey = __reduce_min_sync(0xffffffffu, minkey);
                    steps = __reduce_add_sync(0xffffffffu, steps);
                    if (lane == 0) {
                        if (!ok) s_ok[m] = 0;
                        atomicMin(&s_minkey[m], minkey);
                        atomicAdd(&s_cnt[m], steps);
                    }
                }
            } else if (P1 && lane == 0) {
                s_ok[m] = 0;
            }
            if (PASS == 2 && !want && nm != kTraceTodo && nm != 0 && nm <= static_cast<uint32_t>(kLdShort)) {
                // a short vertex list (a walked blob): into shared memory as {y, x}
                const uint32_t* v = prm.verts + (pair * 2 + m) * static_cast<long long>(prm.max_pts);
                for (int i = (warp & 1) * 32 + lane; i < static_cast<int>(nm); i += 64) {
                    const uint32_t w = v[i];
                    s_list[m][i] = make_int2(static_cast<int>(w >> 16), static_cast<int>(w & 0xffffu));
                    s_lbest[m][i] = 0x7fffffff;
                }
            }
        }
        __syncthreads();
        if (P1) {
            // per map: 0 = no contour, count = verified height function, kTraceTodo = left to the fallback kernels
            const uint32_t c0 = static_cast<uint32_t>(W) + s_cnt[0], c1 = static_cast<uint32_t>(W) + s_cnt[1];
            const bool v0 = s_ok[0] && s_minkey[0] == s_seed[0], v1 = s_ok[1] && s_minkey[1] == s_seed[1];
            // max_pts bounds what is STORED (vertex lists, rows): a pair with two verified sides is measured from the tables
            // and stores nothing, so its contours may be longer (wide images: 2 W + sum |dh| vertices); 16-bit counters
            const bool both = v0 && v1 && s_seed[0] != OCTM_NO_SEED && s_seed[1] != OCTM_NO_SEED && c0 <= 0xffffu && c1 <= 0xffffu;
            n0 = s_seed[0] == OCTM_NO_SEED ? 0u
                 : (v0 && (both || c0 <= static_cast<uint32_t>(prm.max_pts)) ? (c0 | kLayeredBit | (row0 ? kRowBit : 0u)) : kTraceTodo);
            n1 = s_seed[1] == OCTM_NO_SEED ? 0u
                 : (v1 && (both || c1 <= static_cast<uint32_t>(prm.max_pts)) ? (c1 | kLayeredBit | (row1 ? kRowBit : 0u)) : kTraceTodo);
        }
        const bool todo0 = n0 == kTraceTodo, todo1 = n1 == kTraceTodo;        // PASS 2: cannot happen (the walk settles them)
        const bool lay0 = !todo0 && (n0 & kLayeredBit), lay1 = !todo1 && (n1 & kLayeredBit);
        const uint32_t cnt0 = todo0 ? 0u : n0 & ~(kLayeredBit | kRowBit), cnt1 = todo1 ? 0u : n1 & ~(kLayeredBit | kRowBit);
        if (n0 == 0 || n1 == 0) {            // a mask without a contour: nothing to measure (reference: IndexError)
            const bool todo = todo0 || todo1;                                 // the other side is still walked for its n_pts
            if (tid < 2) {
                prm.n_pts[pair * 2 + tid] = tid ? (todo1 ? n1 : cnt1) : (todo0 ? n0 : cnt0);
                prm.max_sq[pair * 2 + tid] = todo ? kNeedsSearch : 0u;
                prm.p95_sq[pair * 4 + tid * 2] = prm.p95_sq[pair * 4 + tid * 2 + 1] = 0;
                prm.sum_dist[pair * 2 + tid] = 0.0;
            }
            if (P1 && todo && tid == 0) atomicAdd(prm.todo_count, 1u);
            continue;
        }
        if (P1 && (todo0 || todo1)) {      // handed on: the verified side keeps its flag, nothing is emitted yet
            if (tid < 2) {
                prm.n_pts[pair * 2 + tid] = tid ? n1 : n0;
                prm.max_sq[pair * 2 + tid] = kNeedsSearch;
            }
            if (tid == 2) atomicAdd(prm.todo_count, 1u);
            continue;
        }
        const bool short0 = !lay0 && cnt0 <= static_cast<uint32_t>(kLdShort), short1 = !lay1 && cnt1 <= static_cast<uint32_t>(kLdShort);
        bool done = (lay0 || short0) && (lay1 || short1);            // CTA-uniform
        if (done) {
            LdSide sd0, sd1;
            bool far_hint = false;           // CTA-uniform
            sd0.lo = lay0 ? tabs + kLdPad : nullptr;
            sd0.hi = lay0 ? tabs + tab + kLdPad : nullptr;
            sd0.pts = s_list[0];
            sd0.best = s_lbest[0];
            sd0.n = static_cast<int>(cnt0);
            sd1.lo = lay1 ? tabs + 2 * tab + kLdPad : nullptr;
            sd1.hi = lay1 ? tabs + 3 * tab + kLdPad : nullptr;
            sd1.pts = s_list[1];
            sd1.best = s_lbest[1];
            sd1.n = static_cast<int>(cnt1);
            if constexpr (PASS == 2) if (!lay0 || !lay1) {         // the short lists' own minima, once
                if (!lay0) { if (lay1) ld_list_minima<true>(s_list[0], sd0.n, sd1, ncol, tid, s_lbest[0]); else ld_list_minima<false>(s_list[0], sd0.n, sd1, ncol, tid, s_lbest[0]); }
                if (!lay1) { if (lay0) ld_list_minima<true>(s_list[1], sd1.n, sd0, ncol, tid, s_lbest[1]); else ld_list_minima<false>(s_list[1], sd1.n, sd0, ncol, tid, s_lbest[1]); }
                __syncthreads();
                // a blob far from the other side (the usual fate of a stray pixel): its distances will not fit the fine
                // counters in either direction, so the fine attempt is skipped
                bool f = false;
                if (!lay0 && tid < sd0.n) f = s_lbest[0][tid] >= 2 * kCountBins;
                if (!lay1 && tid < sd1.n) f = f || s_lbest[1][tid] >= 2 * kCountBins;
                far_hint = __syncthreads_or(f) != 0;
            }
            // direction d (0: queries = pred vertices, sources = true vertices; 1: swapped); the kinds are CTA-uniform
            auto run_dir = [&](int d, auto& ctr) {
                LdSide q, sd;                        // field-wise selects: no indexed struct array (it would live in local memory)
                q.lo = d ? sd0.lo : sd1.lo;  q.hi = d ? sd0.hi : sd1.hi;  q.pts = d ? sd0.pts : sd1.pts;
                q.best = d ? sd0.best : sd1.best;  q.n = d ? sd0.n : sd1.n;
                sd.lo = d ? sd1.lo : sd0.lo;  sd.hi = d ? sd1.hi : sd0.hi;  sd.pts = d ? sd1.pts : sd0.pts;
                sd.best = d ? sd1.best : sd0.best;  sd.n = d ? sd1.n : sd0.n;
                const bool qt = q.lo != nullptr, st = sd.lo != nullptr;
                if constexpr (P1) ld_direction<true, true>(q, sd, W, ncol, warp, lane, ring, ctr);
                else if (qt && st) ld_direction<true, true>(q, sd, W, ncol, warp, lane, ring, ctr);
                else if (qt) ld_direction<true, false>(q, sd, W, ncol, warp, lane, ring, ctr);
                else if (st) ld_direction<false, true>(q, sd, W, ncol, warp, lane, ring, ctr);
                else ld_direction<false, false>(q, sd, W, ncol, warp, lane, ring, ctr);
            };
            using Counter = typename std::conditional<P1, FineCounter, WideCounter>::type;
            // PASS 2: a direction whose QUERIES are a short list is settled at once from the list's minima (warp 0)
            const bool direct0 = PASS == 2 && !lay1, direct1 = PASS == 2 && !lay0;      // direction d: queries = side 1 - d
            if (PASS == 2 && warp == 0) {
                if (direct0) ld_list_stats(s_lbest[1], sd1.n, lane, prm.max_sq + pair * 2, prm.p95_sq + pair * 4, prm.sum_dist + pair * 2);
                if (direct1) ld_list_stats(s_lbest[0], sd0.n, lane, prm.max_sq + pair * 2 + 1, prm.p95_sq + pair * 4 + 2, prm.sum_dist + pair * 2 + 1);
            }
            if constexpr (PASS == 2) {
                // one table, one short list: the table's vertices against the list without rings or counters
                const int nq_tab = static_cast<int>(lay0 ? cnt0 : cnt1);
                if (lay0 != lay1 && nq_tab <= prm.vals_cap) {
                    const int d = lay0 ? 1 : 0;          // direction whose queries are the table
                    uint32_t* const vals = reinterpret_cast<uint32_t*>(dsm + static_cast<size_t>(tab) * 8 + kCountBins * 4 + kLdWarps * kLdRing * 4);
                    if (tid == 0) s_nvals = 0;
                    __syncthreads();
                    ld_table_vs_list(lay0 ? sd0.lo : sd1.lo, lay0 ? sd0.hi : sd1.hi, lay0 ? s_list[1] : s_list[0],
                                     static_cast<int>(lay0 ? cnt1 : cnt0), W, nq_tab, tid, vals, &s_nvals, bins2 + kCountBins /* scratch: warp 0's ring */,
                                     s_dsum8, bins2, s_tgt, prm.max_sq + pair * 2 + d, prm.p95_sq + (pair * 2 + d) * 2,
                                     prm.sum_dist + pair * 2 + d);
                    __syncthreads();
                    for (int i = tid; i < 256; i += kLdWarps * 32) bins2[i] = 0;       // block_select's histogram lives in the counters
                    if (tid >= 64 && tid < 66) prm.n_pts[pair * 2 + (tid - 64)] = tid == 64 ? cnt0 : cnt1;
                    continue;
                }
            }
            if constexpr (P1) ld_both_tt(tabs, tab, W, ncol, warp, lane, ring, bins2, s_vmax, &s_bad);
#pragma unroll 1
            for (int d = P1 ? 2 : 0; d < 2; ++d) {
                if (d ? direct1 : direct0) continue;
                Counter fc;
                fc.bins = bins2 + d * (kCountBins / 2);
                if constexpr (PASS == 2) {
                    fc.mode = 0;
                    if (far_hint) {                  // straight to the wide counting, with the image diagonal as the bound
                        if (tid == 0) { atomicOr(&s_bad, 1u << d); s_vmax[d] = static_cast<uint32_t>(4 * (prm.H * prm.H + W * W)); }
                        continue;
                    }
                }
                run_dir(d, fc);
                const uint32_t wmax = __reduce_max_sync(0xffffffffu, fc.run_max);
                const bool bad = __any_sync(0xffffffffu, fc.run_bad), ovf = __any_sync(0xffffffffu, fc.overflow);
                if (lane == 0) {
                    atomicMax(&s_vmax[d], wmax);
                    if (bad || ovf) atomicOr(&s_bad, (bad ? 1u << d : 0u) | (ovf ? 4u : 0u));
                }
            }
            __syncthreads();                 // both directions counted
            const uint32_t badbits = s_bad;
            done = badbits == 0;
            if (done) {
                if (warp < 2 && !(warp ? direct1 : direct0))
                    stats_from_counters(bins2 + warp * (kCountBins / 2), s_vmax[warp], static_cast<int>(warp ? cnt0 : cnt1), lane,
                                        prm.max_sq + pair * 2 + warp, prm.p95_sq + (pair * 2 + warp) * 2, prm.sum_dist + pair * 2 + warp);
                if (tid >= 64 && tid < 66) prm.n_pts[pair * 2 + (tid - 64)] = tid == 64 ? cnt0 : cnt1;
                if (P1 && rep == 0 && own_next) {
                    // the next class: same rows, so (a) holds and the path is the same; (b) its band must be thicker
                    // than every step, (c) its first pixel must be the path's raster-first pixel
                    __syncthreads();
                    if (tid < 2) s_ok[tid] = 1;
                    __syncthreads();
                    {
                        const int m = warp >> 1, c1 = cls + 1;
                        const int* rows = (m ? prm.bnd_p : prm.bnd_t) + item * (K - 1) * static_cast<long long>(W);
                        const int* orow = c1 < K - 1 ? rows + c1 * static_cast<long long>(W) : nullptr;
                        const short* lo = tabs + (2 * m) * tab + kLdPad;
                        bool ok = prm.first_pos[(item * 2 + m) * K + c1] == s_minkey[m];
                        for (int x0 = (warp & 1) * 32; x0 < W; x0 += 64) {
                            const int x = x0 + lane;
                            if (x < W) {
                                const int h = (lo[2 * x] + 1) >> 1;
                                const int hl = (lo[2 * max(x - 1, 0)] + 1) >> 1, hn = (lo[2 * min(x + 1, W - 1)] + 1) >> 1;
                                const int o = orow != nullptr ? orow[x] : prm.H;
                                ok = ok && o >= max(hl, max(h, hn)) + 1;
                            }
                        }
                        if (!__all_sync(0xffffffffu, ok) && lane == 0) s_ok[m] = 0;
                    }
                    __syncthreads();
                    if (s_ok[0] && s_ok[1]) {
                        if (tid < 2) {
                            prm.n_pts[(pair + 1) * 2 + tid] = tid ? cnt1 : cnt0;
                            prm.max_sq[(pair + 1) * 2 + tid] = prm.max_sq[pair * 2 + tid];
                            prm.p95_sq[(pair + 1) * 4 + tid * 2] = prm.p95_sq[pair * 4 + tid * 2];
                            prm.p95_sq[(pair + 1) * 4 + tid * 2 + 1] = prm.p95_sq[pair * 4 + tid * 2 + 1];
                            prm.sum_dist[(pair + 1) * 2 + tid] = prm.sum_dist[pair * 2 + tid];
                        }
                        own_next = false;
                    }
                }
                continue;
            }
            // (a side longer than max_pts was accepted because nothing had to be stored; now something may: it is handed
            // on as unverified, and the walk reports the overflow)
            const bool over0 = cnt0 > static_cast<uint32_t>(prm.max_pts), over1 = cnt1 > static_cast<uint32_t>(prm.max_pts);
            if constexpr (P1) if (!(badbits & 4u) || over0 || over1) {
                // distances the counters cannot hold: PASS 2 measures the pair again with the wide counting below
                for (int i = tid; i < kCountBins; i += kLdWarps * 32) bins2[i] = 0;
                if (tid < 2) {
                    prm.n_pts[pair * 2 + tid] = tid ? (over1 ? kTraceTodo : n1) : (over0 ? kTraceTodo : n0);
                    prm.max_sq[pair * 2 + tid] = kNeedsSearch;
                }
                if (tid == 2) atomicAdd(prm.todo_count, 1u);
                continue;
            }
            if constexpr (PASS == 2) if (!(badbits & 4u) && max(direct0 ? 0u : s_vmax[0], direct1 ? 0u : s_vmax[1]) < (static_cast<uint32_t>(kCountBins) << 11)) {
                // ---- wide counting: per direction either the usual statistics, or a two-level radix select over
                // recomputed distances (coarse bins of 2^shift values; then the bin of the percentile at full resolution)
#pragma unroll 1
                for (int d = 0; d < 2; ++d) {
                    uint32_t* bins = bins2 + d * (kCountBins / 2);
                    const int nq = static_cast<int>(d ? cnt0 : cnt1);
                    if (d ? direct1 : direct0) continue;
                    if (!((badbits >> d) & 1u)) {
                        if (warp == 0)
                            stats_from_counters(bins, s_vmax[d], nq, lane, prm.max_sq + pair * 2 + d, prm.p95_sq + (pair * 2 + d) * 2,
                                                prm.sum_dist + pair * 2 + d);
                        continue;                        // CTA-uniform
                    }
                    const uint32_t vmax = s_vmax[d];
                    int shift = 1;
                    while ((vmax >> shift) >= static_cast<uint32_t>(kCountBins)) ++shift;        // <= 11 by the guard above
                    __syncthreads();
                    for (int i = tid; i < kCountBins / 2; i += kLdWarps * 32) bins[i] = 0;
                    if (tid < kLdWarps) s_dsum[tid] = 0.0;
                    if (tid == 0) { s_next = 0xffffffffu; s_amax = 0; }
                    __syncthreads();
                    uint32_t* const vals = reinterpret_cast<uint32_t*>(dsm + static_cast<size_t>(tab) * 8 + kCountBins * 4 + kLdWarps * kLdRing * 4);
                    const bool stored = nq <= prm.vals_cap;              // CTA-uniform
                    if (tid == 0) s_nvals = 0;
                    __syncthreads();
                    WideCounter cc;
                    cc.bins = bins;
                    cc.mode = 1;
                    cc.shift = shift;
                    cc.target = 0;
                    if (stored) { cc.vals = vals; cc.vals_n = &s_nvals; }
                    run_dir(d, cc);
                    {
                        const uint32_t wm = __reduce_max_sync(0xffffffffu, cc.run_max);
                        if (lane == 0) atomicMax(&s_amax, wm);
                    }
                    double ws = cc.sum;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) ws = __dadd_rn(ws, __shfl_xor_sync(0xffffffffu, ws, o));
                    if (lane == 0) s_dsum[warp] = ws;
                    __syncthreads();
                    // warp 0: the coarse bin of rank lo (numpy's linear percentile: lo = floor(0.95 (nq - 1)), hi = lo + 1)
                    const double pos = __dmul_rn(static_cast<double>(nq - 1), 0.95);
                    const uint32_t lo_rank = static_cast<uint32_t>(floor(pos));
                    const uint32_t hi_rank = lo_rank + 1 < static_cast<uint32_t>(nq) ? lo_rank + 1 : lo_rank;
                    if (warp == 0) {
                        uint32_t base = 0, tgt = 0, before = 0, inbin = 0;
                        for (uint32_t h0 = 0; h0 <= (s_amax >> shift); h0 += 32) {
                            const uint32_t h = h0 + lane;
                            const uint32_t c = (bins[h >> 1] >> ((h & 1u) * 16)) & 0xffffu;
                            uint32_t incl = c;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                                if (lane >= o) incl += up;
                            }
                            const uint32_t first = base + incl - c;
                            const uint32_t hit = __ballot_sync(0xffffffffu, c != 0 && first <= lo_rank && lo_rank < first + c);
                            if (hit) {
                                const int src_lane = __ffs(hit) - 1;
                                tgt = h0 + src_lane;
                                before = __shfl_sync(0xffffffffu, first, src_lane);
                                inbin = __shfl_sync(0xffffffffu, c, src_lane);
                            }
                            base += __shfl_sync(0xffffffffu, incl, 31);
                        }
                        if (lane == 0) { s_tgt[0] = tgt; s_tgt[1] = before; s_tgt[2] = inbin; }
                    }
                    __syncthreads();
                    for (int i = tid; i < kCountBins / 2; i += kLdWarps * 32) bins[i] = 0;
                    __syncthreads();
                    WideCounter wc;
                    wc.bins = bins;
                    wc.mode = 2;
                    wc.shift = shift;
                    wc.target = s_tgt[0];
                    if (stored) {                        // over the stored values: no second enumeration
                        for (int i0 = 0; i0 < nq; i0 += kLdWarps * 32) {
                            const int i = i0 + tid;
                            wc.add(i < nq ? static_cast<int>(vals[i]) : 0, i < nq, lane);
                        }
                    } else {
                        run_dir(d, wc);
                    }
                    const uint32_t wn = __reduce_min_sync(0xffffffffu, wc.next_min);
                    if (lane == 0) atomicMin(&s_next, wn);
                    __syncthreads();
                    if (warp == 0) {
                        const uint32_t r_lo = lo_rank - s_tgt[1], r_hi = hi_rank - s_tgt[1], inbin = s_tgt[2];
                        uint32_t v_lo = 0, v_hi = 0, base = 0;
                        const uint32_t nfine = 1u << (shift - 1);
                        for (uint32_t h0 = 0; h0 < nfine; h0 += 32) {
                            const uint32_t h = h0 + lane;
                            const uint32_t c = (bins[h >> 1] >> ((h & 1u) * 16)) & 0xffffu;
                            uint32_t incl = c;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                                if (lane >= o) incl += up;
                            }
                            const uint32_t first = base + incl - c;
                            const uint32_t m_lo = __ballot_sync(0xffffffffu, c != 0 && first <= r_lo && r_lo < first + c);
                            const uint32_t m_hi = __ballot_sync(0xffffffffu, c != 0 && first <= r_hi && r_hi < first + c);
                            if (m_lo) v_lo = (s_tgt[0] << shift) + 2 * (h0 + __ffs(m_lo) - 1);
                            if (m_hi) v_hi = (s_tgt[0] << shift) + 2 * (h0 + __ffs(m_hi) - 1);
                            base += __shfl_sync(0xffffffffu, incl, 31);
                        }
                        if (r_hi >= inbin) v_hi = s_next;            // the next order statistic lies in a higher coarse bin
                        if (lane == 0) {
                            prm.max_sq[pair * 2 + d] = s_amax;
                            prm.p95_sq[(pair * 2 + d) * 2] = v_lo;
                            prm.p95_sq[(pair * 2 + d) * 2 + 1] = v_hi;
                            prm.sum_dist[pair * 2 + d] = __dadd_rn(__dadd_rn(s_dsum[0], s_dsum[1]), __dadd_rn(s_dsum[2], s_dsum[3]));
                        }
                    }
                    __syncthreads();
                    for (int i = tid; i < kCountBins / 2; i += kLdWarps * 32) bins[i] = 0;
                }
                if (tid >= 64 && tid < 66) prm.n_pts[pair * 2 + (tid - 64)] = tid == 64 ? cnt0 : cnt1;
                continue;
            }
            for (int i = tid; i < kCountBins; i += kLdWarps * 32) bins2[i] = 0;      // left to the vertex-list search
        }
        // ---- left to the vertex-list search: the tables become vertex lists in column order (warp m: map m)
        if (warp < 2 && (warp ? lay1 : lay0))
            ld_emit(tabs + (2 * warp) * tab + kLdPad, tabs + (2 * warp + 1) * tab + kLdPad, ncol, lane,
                    prm.verts + (pair * 2 + warp) * static_cast<long long>(prm.max_pts), static_cast<uint32_t>(prm.max_pts));
        if (tid >= 64 && tid < 66) {
            prm.n_pts[pair * 2 + (tid - 64)] = tid == 64 ? cnt0 : cnt1;
            prm.max_sq[pair * 2 + (tid - 64)] = kNeedsSearch;
        }
        if (tid == 66) atomicAdd(prm.search_count, 1u);
      }
    }
}

}  // namespace octm

// ----------------------------------------------------------------------------------- C ABI
static int check_shape(int64_t n, int H, int W, int K, int max_pts) {
    if (n < 0) return octm::fail(OCTM_ERR_INVALID, "n_items < 0");
    if (K < 2 || K > OCTM_MAX_CLASSES) return octm::fail(OCTM_ERR_INVALID, "num_classes %d outside [2, 16]", K);
    if (H < 1 || W < 1) return octm::fail(OCTM_ERR_INVALID, "H, W must be >= 1");
    if (H > 8192 || W > 8192)
        return octm::fail(OCTM_ERR_UNSUPPORTED, "image side > 8192: squared lattice distances leave the int32 kernel range");
    if (static_cast<long long>(H) * W >= (1ll << 32) - 1) return octm::fail(OCTM_ERR_UNSUPPORTED, "H*W >= 2^32");
    if (max_pts < 8 || max_pts % 4 != 0) return octm::fail(OCTM_ERR_INVALID, "max_pts must be >= 8 and a multiple of 4");
    return OCTM_OK;
}

extern "C" int octm_first_pos_u8(const uint8_t* labels, int64_t n_items, int64_t item_elems, int num_classes,
                                 uint32_t* first_pos, void* stream) {
    if (n_items < 0 || item_elems < 1 || item_elems >= (1ll << 32) - 1 || num_classes < 2 || num_classes > 16)
        return octm::fail(OCTM_ERR_INVALID, "bad shape");
    if (n_items == 0) return OCTM_OK;
    if (!labels || !first_pos) return octm::fail(OCTM_ERR_INVALID, "null pointer");
    const long long grid = n_items < 148 * 8 ? n_items : 148 * 8;
    OCTM_TIMED("first_pos_kernel", static_cast<cudaStream_t>(stream)) octm::first_pos_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        labels, n_items, item_elems, num_classes, first_pos, num_classes);
    return octm::check_launch("first_pos_kernel");
}

static bool layered_ok(const void* y_true, const void* y_pred, const void* bnd_true, const void* bnd_pred, int H, int W, int K) {
    // layered fast path (needs the label pass's boundary rows): OCTM_TRACE_LAYERED=0 turns it off
    static const bool env_layered = [] { const char* e = getenv("OCTM_TRACE_LAYERED"); return !(e && e[0] == '0'); }();
    return env_layered && bnd_true != nullptr && bnd_pred != nullptr && H >= 2 && W % 2 == 0 && K >= 2 && K <= 32 &&
           reinterpret_cast<uintptr_t>(y_true) % 2 == 0 && reinterpret_cast<uintptr_t>(y_pred) % 2 == 0 &&
           reinterpret_cast<uintptr_t>(bnd_true) % 8 == 0 && reinterpret_cast<uintptr_t>(bnd_pred) % 8 == 0;
}

template <bool EMIT>
static int launch_layered_trace(const octm::TraceParams& p, cudaStream_t s) {
    int fit = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, octm::trace_layered_kernel<EMIT>, 128, 0) != cudaSuccess || fit < 1) {
        cudaGetLastError();
        fit = 8;
    }
    long long grid = (p.n_items * p.K * 2 + 3) / 4;
    const long long cap = static_cast<long long>(octm::sm_count()) * fit;
    if (grid > cap) grid = cap;
    OCTM_TIMED("trace_layered_kernel", s) octm::trace_layered_kernel<EMIT><<<static_cast<unsigned>(grid), 128, 0, s>>>(p);
    return octm::check_launch("trace_layered_kernel");
}

#ifdef OCTM_WALK_STATS
extern "C" __attribute__((visibility("default"))) int octm_debug_walk_stats(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, octm::g_walk_stats, sizeof(unsigned long long) * 8);
    if (reset) {
        unsigned long long z[8] = {};
        cudaMemcpyToSymbol(octm::g_walk_stats, z, sizeof(z));
    }
    return 0;
}
#endif

static int launch_walk(const octm::TraceParams& p_in, cudaStream_t s, int force_words = -1) {
    octm::TraceParams p = p_in;
    bool words = p.W % 16 == 0 && reinterpret_cast<uintptr_t>(p.yt) % 16 == 0 && reinterpret_cast<uintptr_t>(p.yp) % 16 == 0;
    if (p.walk_list != nullptr && words && force_words < 0) {
        // A compacted list is walked by the register-cached kernel when it is long (throughput: fewer load instructions)
        // and by the byte-load kernel when it is short (latency: the step's dependent chain is shorter; measured 12-19 %
        // faster below ~32 k walks, 20 % slower at 131 k).  The length is a device word: both are launched, one returns.
        p.walk_when = 1;
        if (int e = launch_walk(p, s, 1)) return e;
        p.walk_when = 2;
        return launch_walk(p, s, 0);
    }
    if (force_words >= 0) words = words && force_words == 1;
    const long long threads = p.n_items * p.K * 2;
    static const int env_ctas = [] { const char* e = getenv("OCTM_TRACE_CTAS"); return e ? atoi(e) : 0; }();
    int fit = 0;       // persistent grid: every CTA that can be resident (the walk is latency-bound: occupancy hides it)
    if ((words ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, octm::trace_kernel<true>, 128, 0)
               : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, octm::trace_kernel<false>, 128, 0)) != cudaSuccess || fit < 1) {
        cudaGetLastError();
        fit = 8;
    }
    const int ctas_per_sm = env_ctas > 0 ? env_ctas : fit;
    long long grid = (threads + 127) / 128;
    const long long cap = static_cast<long long>(octm::sm_count()) * ctas_per_sm;
    if (grid > cap) grid = cap;
    if (words) OCTM_TIMED("trace_kernel", s) octm::trace_kernel<true><<<static_cast<unsigned>(grid), 128, 0, s>>>(p);
    else OCTM_TIMED("trace_kernel", s) octm::trace_kernel<false><<<static_cast<unsigned>(grid), 128, 0, s>>>(p);
    return octm::check_launch("trace_kernel");
}

extern "C" int octm_contour2d_trace_u8(const uint8_t* y_true, const uint8_t* y_pred, int64_t n_items, int H, int W,
                                       int num_classes, const uint32_t* first_pos, const int32_t* bnd_true,
                                       const int32_t* bnd_pred, int max_pts, uint32_t* verts, uint32_t* n_pts,
                                       uint32_t* flags, void* stream) {
    if (int e = check_shape(n_items, H, W, num_classes, max_pts)) return e;
    if (n_items == 0) return OCTM_OK;
    if (!y_true || !y_pred || !first_pos || !verts || !n_pts || !flags) return octm::fail(OCTM_ERR_INVALID, "null pointer");
    if ((bnd_true == nullptr) != (bnd_pred == nullptr)) return octm::fail(OCTM_ERR_INVALID, "bnd_true/bnd_pred: both or neither");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(flags, 0, sizeof(uint32_t) * n_items * num_classes, s) != cudaSuccess)
        return octm::fail(OCTM_ERR_LAUNCH, "memset(flags) failed");
    octm::TraceParams p{y_true, y_pred, n_items, H, W, num_classes, max_pts, first_pos, verts, n_pts, flags,
                        bnd_true, bnd_pred, false};
    if (layered_ok(y_true, y_pred, bnd_true, bnd_pred, H, W, num_classes)) {
        if (int e = launch_layered_trace<true>(p, s)) return e;
        p.only_todo = true;
    }
    return launch_walk(p, s);
}

static size_t dist_smem(int max_pts) {
    const size_t capb = (max_pts + octm::kBox - 1) / octm::kBox, caps = (capb + octm::kSuper - 1) / octm::kSuper;
    return static_cast<size_t>(max_pts) * 20 + 2 * (capb + caps) * 16;
}

static int run_distance(const uint32_t* verts, const uint32_t* n_pts, int64_t n_items, int num_classes, int max_pts, int H,
                        int W, uint32_t* max_sq, uint32_t* p95_sq, double* sum_dist, uint32_t* d2, int keep_d2,
                        bool only_marked, const uint32_t* todo_count, void* stream) {
    if (n_items < 0 || num_classes < 1 || max_pts < 8) return octm::fail(OCTM_ERR_INVALID, "bad shape");
    if (H < 1 || W < 1 || H > 8192 || W > 8192) return octm::fail(OCTM_ERR_INVALID, "H, W outside [1, 8192]");
    if (n_items == 0) return OCTM_OK;
    if (!verts || !n_pts || !max_sq || !p95_sq || !sum_dist) return octm::fail(OCTM_ERR_INVALID, "null pointer");
    // OCTM_DISTANCE_MODE = column (default: column-sorted cooperative search, falls back to tiled when the sorted
    // contour does not fit shared memory) | tiled (cooperative box search) | lane (per-lane pruned search) |
    // brute (no pruning); all produce identical integers (tests run each).
    static const int env_mode = [] {
        const char* e = getenv("OCTM_DISTANCE_MODE");
        if (e && !strcmp(e, "brute")) return 2;
        if (e && !strcmp(e, "lane")) return 1;
        if (e && !strcmp(e, "tiled")) return 0;
        return 3;
    }();
    int mode = env_mode;
    if (only_marked && mode != 3) mode = 0;         // the single-kernel check modes search every unit
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n_pairs = n_items * num_classes;
    auto run_select = [&]() -> int {
        octm::SelectParams kp{n_pts, d2, n_pairs * 2, max_pts, max_sq, p95_sq, sum_dist, todo_count};
        long long kgrid = (n_pairs * 2 + 3) / 4;
        const long long kcap = static_cast<long long>(octm::sm_count()) * 16;
        if (kgrid > kcap) kgrid = kcap;
        OCTM_TIMED("distance_select_kernel", st) octm::distance_select_kernel<<<static_cast<unsigned>(kgrid), 128, 0, st>>>(kp);
        return octm::check_launch("distance_select_kernel");
    };
    if (mode == 3) {
        if (d2 == nullptr) return octm::fail(OCTM_ERR_INVALID, "d2 scratch [n][K][2][max_pts] is required");
        const int ncol = 2 * W;
        const size_t smem = (static_cast<size_t>(max_pts) + 4) * 16 + (static_cast<size_t>(ncol) + 1) * 4;
        if (smem + 4096 > static_cast<size_t>(octm::max_optin_smem())) {
            mode = 0;                                   // long contours: tile the source instead
        } else {
            if (cudaFuncSetAttribute(octm::distance_column_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)) != cudaSuccess)
                return octm::fail(OCTM_ERR_LAUNCH, "cudaFuncSetAttribute(distance_column_kernel) failed");
            octm::ColumnParams cp{verts, n_pts, n_pairs * 2, max_pts, ncol, d2, max_sq, p95_sq, sum_dist, keep_d2 != 0, only_marked, todo_count};
            long long grid = n_pairs * 2;
            const long long cap = static_cast<long long>(octm::sm_count()) * 32;
            if (grid > cap) grid = cap;
            OCTM_TIMED("distance_column_kernel", st) octm::distance_column_kernel<<<static_cast<unsigned>(grid), octm::kColThreads, smem, st>>>(cp);
            if (int e = octm::check_launch("distance_column_kernel")) return e;
            return run_select();
        }
    }
    if (mode == 0) {
        if (d2 == nullptr) return octm::fail(OCTM_ERR_INVALID, "d2 scratch [n][K][2][max_pts] is required");
        static const int env_tile = [] { const char* e = getenv("OCTM_DIST_TILE"); return e ? atoi(e) : 0; }();
        int tile = env_tile >= octm::kBox ? (env_tile / octm::kBox) * octm::kBox : 2048;
        if (tile > ((max_pts + octm::kBox - 1) / octm::kBox) * octm::kBox) tile = ((max_pts + octm::kBox - 1) / octm::kBox) * octm::kBox;
        const size_t smem = static_cast<size_t>(tile) * 16 + static_cast<size_t>(tile / octm::kBox) * 16;
        if (smem > static_cast<size_t>(octm::max_optin_smem()))
            return octm::fail(OCTM_ERR_UNSUPPORTED, "tile %d needs %zu B of shared memory", tile, smem);
        if (cudaFuncSetAttribute(octm::distance_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)) != cudaSuccess)
            return octm::fail(OCTM_ERR_LAUNCH, "cudaFuncSetAttribute(distance_search_kernel) failed");
        octm::SearchParams sp{verts, n_pts, n_pairs * 2, max_pts, tile, d2, max_sq, p95_sq, sum_dist, keep_d2 != 0, only_marked, todo_count};
        long long grid = n_pairs * 2;
        const long long cap = static_cast<long long>(octm::sm_count()) * 32;
        if (grid > cap) grid = cap;
        OCTM_TIMED("distance_search_kernel", st) octm::distance_search_kernel<<<static_cast<unsigned>(grid), octm::kSearchThreads, smem, st>>>(sp);
        if (int e = octm::check_launch("distance_search_kernel")) return e;
        return run_select();
    }
    const size_t smem = dist_smem(max_pts);
    if (smem > static_cast<size_t>(octm::max_optin_smem()) - 4096)
        return octm::fail(OCTM_ERR_UNSUPPORTED, "max_pts %d needs %zu B of shared memory in this mode", max_pts, smem);
    octm::DistParams p{verts, n_pts, n_pairs, max_pts, max_sq, p95_sq, sum_dist, d2};
    long long grid = n_pairs;
    const long long cap = static_cast<long long>(octm::sm_count()) * 16;
    if (grid > cap) grid = cap;
    auto kern = mode == 2 ? octm::distance_kernel<false> : octm::distance_kernel<true>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, octm::max_optin_smem() - 4096) != cudaSuccess)
        return octm::fail(OCTM_ERR_LAUNCH, "cudaFuncSetAttribute(distance_kernel) failed");
    OCTM_TIMED("distance_kernel", st) kern<<<static_cast<unsigned>(grid), octm::kDistThreads, smem, st>>>(p);
    return octm::check_launch("distance_kernel");
}

extern "C" int octm_contour2d_distance(const uint32_t* verts, const uint32_t* n_pts, int64_t n_items, int num_classes,
                                       int max_pts, int H, int W, uint32_t* max_sq, uint32_t* p95_sq, double* sum_dist,
                                       uint32_t* d2, int keep_d2, void* stream) {
    return run_distance(verts, n_pts, n_items, num_classes, max_pts, H, W, max_sq, p95_sq, sum_dist, d2, keep_d2, false, nullptr, stream);
}

// How many of the last call's contours ended in the walk list (host-mapped word, written by a one-thread kernel, read by
// the next call without synchronisation): verifying the rejected maps' contours BEFORE the boundary-row pass pays when
// most of them pass (stray pixels here and there) and costs a little when most are walked anyway (ragged boundaries,
// heavy noise).  Speed heuristics only.
namespace octm {
__global__ void walk_report_kernel(const uint32_t* walk_count, uint32_t contours, uint32_t* out) {
    out[1] = contours;
    out[0] = *walk_count;
}
struct WalkReport {
    uint32_t* host = nullptr;
    uint32_t* dev = nullptr;
};
static WalkReport g_walk_report[64];
static std::mutex g_walk_mu;
static WalkReport* walk_report() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return nullptr; }
    std::lock_guard<std::mutex> lk(g_walk_mu);
    WalkReport& r = g_walk_report[dev];
    if (r.host == nullptr) {
        void* h = nullptr;
        void* d = nullptr;
        if (cudaHostAlloc(&h, 2 * sizeof(uint32_t), cudaHostAllocMapped) != cudaSuccess || cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        r.host = static_cast<uint32_t*>(h);
        r.dev = static_cast<uint32_t*>(d);
        r.host[0] = r.host[1] = 0;
    }
    return &r;
}
}  // namespace octm

extern "C" int octm_contour2d_metrics_u8(const uint8_t* y_true, const uint8_t* y_pred, int64_t n_items, int H, int W,
                                         int num_classes, const uint32_t* first_pos, const int32_t* bnd_true,
                                         const int32_t* bnd_pred, const uint32_t* unsorted, int max_pts, uint32_t* n_pts,
                                         uint32_t* flags, uint32_t* max_sq, uint32_t* p95_sq, double* sum_dist,
                                         void* workspace, size_t workspace_bytes, void* stream) {
    if (int e = check_shape(n_items, H, W, num_classes, max_pts)) return e;
    if (n_items == 0) return OCTM_OK;
    if (!y_true || !y_pred || !first_pos || !n_pts || !flags || !max_sq || !p95_sq || !sum_dist)
        return octm::fail(OCTM_ERR_INVALID, "null pointer");
    if ((bnd_true == nullptr) != (bnd_pred == nullptr)) return octm::fail(OCTM_ERR_INVALID, "bnd_true/bnd_pred: both or neither");
    const size_t verts_b = (static_cast<size_t>(n_items) * num_classes * 2 * max_pts * sizeof(uint32_t) + 255) & ~static_cast<size_t>(255);
    if (workspace == nullptr || workspace_bytes < 2 * verts_b + 256)
        return octm::fail(OCTM_ERR_WORKSPACE, "workspace too small: need %zu B", 2 * verts_b + 256);
    uint32_t* verts = static_cast<uint32_t*>(workspace);
    uint32_t* d2 = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(workspace) + verts_b);
    uint32_t* todo = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(workspace) + 2 * verts_b);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // OCTM_LAYERED_FUSED=0: always go through vertex lists (the round-1 path; tests compare the two)
    static const bool env_fused = [] { const char* e = getenv("OCTM_LAYERED_FUSED"); return !(e && e[0] == '0'); }();
    const int tab = (2 * W - 1 + 2 * octm::kLdPad + 7) & ~7;
    const size_t smem = static_cast<size_t>(tab) * 8 + octm::kCountBins * 4 + octm::kLdWarps * octm::kLdRing * 4;
    const bool fused = env_fused && unsorted != nullptr && layered_ok(y_true, y_pred, bnd_true, bnd_pred, H, W, num_classes) &&
                       H <= 4095 && max_pts <= 0xffff && smem <= static_cast<size_t>(octm::max_optin_smem());
    if (!fused) {
        if (int e = octm_contour2d_trace_u8(y_true, y_pred, n_items, H, W, num_classes, first_pos, bnd_true, bnd_pred, max_pts,
                                            verts, n_pts, flags, stream))
            return e;
        return run_distance(verts, n_pts, n_items, num_classes, max_pts, H, W, max_sq, p95_sq, sum_dist, d2, 0, false, nullptr, stream);
    }
    if (cudaMemsetAsync(flags, 0, sizeof(uint32_t) * n_items * num_classes, s) != cudaSuccess ||
        cudaMemsetAsync(todo, 0, 3 * sizeof(uint32_t), s) != cudaSuccess)
        return octm::fail(OCTM_ERR_LAUNCH, "memset(flags) failed");
    uint32_t* search = todo + 1;
    const long long n_pairs = n_items * num_classes;
    octm::LayeredDistParams lp{first_pos, bnd_true, bnd_pred, unsorted, n_pairs, H, W, num_classes, max_pts, tab, verts,
                               n_pts, max_sq, p95_sq, sum_dist, todo, search, 0};
    auto launch_fused = [&](int pass) -> int {
        auto kern = pass == 1 ? octm::layered_distance_kernel<1> : (pass == 3 ? octm::layered_distance_kernel<3> : octm::layered_distance_kernel<2>);
        size_t smem_pass = smem;
        lp.vals_cap = 0;
        if (pass == 2 && max_pts <= 4096) {          // the second pass keeps a unit's distances in shared memory (wide counting)
            smem_pass += static_cast<size_t>(max_pts) * 4;
            lp.vals_cap = max_pts;
        }
        const size_t smem = smem_pass;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
            return octm::fail(OCTM_ERR_LAUNCH, "cudaFuncSetAttribute(layered_distance_kernel) failed");
        int fit = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, kern, octm::kLdWarps * 32, smem) != cudaSuccess || fit < 1) {
            cudaGetLastError();
            fit = 1;
        }
        static const int env_ctas = [] { const char* e = getenv("OCTM_LD_CTAS"); return e ? atoi(e) : 0; }();
        if (env_ctas > 0 && env_ctas < fit) fit = env_ctas;             // co-scheduling experiments
        long long grid = n_pairs;
        const long long cap = static_cast<long long>(octm::sm_count()) * fit;
        if (grid > cap) grid = cap;
        // a CTA strides over the pairs by the grid size: keep the stride coprime to the class count, or every CTA would
        // meet one class only (and the classes differ: the class after pixel (0, 0)'s is usually a copy, not a search)
        while (grid > 1 && std::gcd(grid, static_cast<long long>(num_classes)) != 1) --grid;
        OCTM_TIMED(pass == 1 ? "layered_distance_kernel" : (pass == 3 ? "layered_distance_kernel_rows" : "layered_distance_kernel_pass2"), s)
            kern<<<static_cast<unsigned>(grid), octm::kLdWarps * 32, smem, s>>>(lp);
        return octm::check_launch("layered_distance_kernel");
    };
    octm::TraceParams p{y_true, y_pred, n_items, H, W, num_classes, max_pts, first_pos, verts, n_pts, flags,
                        bnd_true, bnd_pred, true, todo, d2, todo + 2, 0, unsorted, 0};     // the walk list borrows the distance scratch
    // The order of the first two steps follows the data (the label pass's report of how many maps its certificate
    // rejected, octm_label_pass_seed_policy): same results either way.
    octm::WalkReport* wr = octm::stream_mostly_rejected() ? octm::walk_report() : nullptr;
    // (pinned "noisy": always rows first, so that tests reach this order whatever ran before them)
    const bool few_walks = wr != nullptr && (octm_label_pass_seed_policy(-1) == 2 ||
                                             static_cast<unsigned long long>(*static_cast<volatile uint32_t*>(wr->host)) * 4ull <=
                                                 *static_cast<volatile uint32_t*>(wr->host + 1));  // (also before the first report)
    if (wr != nullptr && few_walks) {
        // 1'. the contours of the REJECTED maps (predictions with stray pixels) are verified against the label pixels
        // first; what passes becomes a re-centred row, i.e. a table side like a certified one ...
        p.take = 1;
        p.todo_count = nullptr;
        if (int e = launch_layered_trace<false>(p, s)) return e;
        p.todo_count = todo;
        // ... so that the pairs of a lightly noisy batch are measured by the fast pass, not by the second one
        if (int e = launch_fused(3)) return e;
        p.take = 2;              // 2'. what is left of the CERTIFIED maps
    } else {
        // 1. every pair: verification from the certificate and the boundary rows; pairs with two verified sides are measured
        if (int e = launch_fused(1)) return e;
    }
    // What is handed on (nothing on clean layered data: the kernels below then return at once):
    // 2. verification of the remaining contours against the label pixels (-> tables) ...
    if (int e = launch_layered_trace<false>(p, s)) return e;
    p.take = 0;
    if (wr != nullptr) {
        OCTM_TIMED("walk_report_kernel", s) octm::walk_report_kernel<<<1, 1, 0, s>>>(
            todo + 2, static_cast<uint32_t>(std::min<long long>(n_items * num_classes * 2, 0xffffffffll)), wr->dev);
        if (int e = octm::check_launch("walk_report_kernel")) return e;
    }
    // 3. ... the walk for what is not a height function (-> vertex lists) ...
    if (int e = launch_walk(p, s)) return e;
    // 4. ... the handed-on pairs again: tables and short vertex lists are measured in shared memory ...
    if (int e = launch_fused(2)) return e;
    // 5. ... and the vertex-list search for the rest (long lists on both sides, or against a long list)
    return run_distance(verts, n_pts, n_items, num_classes, max_pts, H, W, max_sq, p95_sq, sum_dist, d2, 0, true, search, stream);
}

extern "C" size_t octm_contour2d_workspace_bytes(int64_t n_items, int H, int W, int num_classes, int max_pts) {
    (void)H; (void)W;
    if (n_items <= 0 || num_classes < 1 || max_pts < 1) return 0;
    const size_t verts = static_cast<size_t>(n_items) * num_classes * 2 * max_pts * sizeof(uint32_t);
    const size_t first = static_cast<size_t>(n_items) * 2 * num_classes * sizeof(uint32_t);
    // vertices, squared-distance scratch (same shape), first occurrences
    return 2 * ((verts + 255) & ~static_cast<size_t>(255)) + ((first + 255) & ~static_cast<size_t>(255));
}

extern "C" int octm_contour2d_u8(const uint8_t* y_true, const uint8_t* y_pred, int64_t n_items, int H, int W,
                                 int num_classes, const uint32_t* first_pos, int max_pts, uint32_t* n_pts,
                                 uint32_t* flags, uint32_t* max_sq, uint32_t* p95_sq, double* sum_dist, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    if (int e = check_shape(n_items, H, W, num_classes, max_pts)) return e;
    if (n_items == 0) return OCTM_OK;
    if (workspace == nullptr || workspace_bytes < octm_contour2d_workspace_bytes(n_items, H, W, num_classes, max_pts))
        return octm::fail(OCTM_ERR_WORKSPACE, "workspace too small: need %zu B",
                          octm_contour2d_workspace_bytes(n_items, H, W, num_classes, max_pts));
    uint32_t* verts = static_cast<uint32_t*>(workspace);
    const size_t verts_b = (static_cast<size_t>(n_items) * num_classes * 2 * max_pts * sizeof(uint32_t) + 255) & ~static_cast<size_t>(255);
    uint32_t* d2_ws = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(workspace) + verts_b);
    uint32_t* fp_ws = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(workspace) + 2 * verts_b);
    if (first_pos == nullptr) {
        // no label pass ran: find the first occurrences of both maps here, interleaved [n][2][K]
        const long long grid = n_items < 148 * 8 ? n_items : 148 * 8;
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        OCTM_TIMED("first_pos_kernel", s) octm::first_pos_kernel<<<static_cast<unsigned>(grid), 256, 0, s>>>(y_true, n_items, static_cast<long long>(H) * W,
                                                                         num_classes, fp_ws, 2 * num_classes);
        if (int e = octm::check_launch("first_pos_kernel")) return e;
        OCTM_TIMED("first_pos_kernel", s) octm::first_pos_kernel<<<static_cast<unsigned>(grid), 256, 0, s>>>(y_pred, n_items, static_cast<long long>(H) * W,
                                                                         num_classes, fp_ws + num_classes, 2 * num_classes);
        if (int e = octm::check_launch("first_pos_kernel")) return e;
        first_pos = fp_ws;
    }
    if (int e = octm_contour2d_trace_u8(y_true, y_pred, n_items, H, W, num_classes, first_pos, nullptr, nullptr, max_pts, verts, n_pts,
                                        flags, stream))
        return e;
    return octm_contour2d_distance(verts, n_pts, n_items, num_classes, max_pts, H, W, max_sq, p95_sq, sum_dist, d2_ws, 0, stream);
}
