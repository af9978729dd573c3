// Core of the contour-[0] walk, shared by the CUDA trace kernel and by the host-compiled unit
// test harness under tests/host/ (which exists only to check these tables against the oracle on
// a machine without a GPU; the shipped library contains no host execution path).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define OCTM_HD __host__ __device__ __forceinline__
#else
#define OCTM_HD inline
#endif

namespace octm {

// edges of a 2x2 square: T=0, B=1, L=2, R=3 (opposite = e ^ 1)
// segment table: for each marching-squares case the (from, to) edge pairs in emission order.
OCTM_HD int seg_to(int kase, int from) {
    // 2 bits per (case, from) -> `to` edge.  Built from the table in the header comment of
    // oracle/contours_oracle.py: 1 T>L, 2 R>T, 3 R>L, 4 L>B, 5 T>B, 6 R>T L>B, 7 R>B, 8 B>R,
    // 9 T>L B>R, 10 B>T, 11 B>L, 12 L>R, 13 T>R, 14 L>T.
    constexpr unsigned long long lo = []() {
        unsigned long long v = 0;
        const int tbl[8][4] = {{0, 0, 0, 0}, {2, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 2},
                               {0, 0, 1, 0}, {1, 0, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
        for (int c = 0; c < 8; ++c)
            for (int e = 0; e < 4; ++e) v |= static_cast<unsigned long long>(tbl[c][e]) << (2 * (c * 4 + e));
        return v;
    }();
    constexpr unsigned long long hi = []() {
        unsigned long long v = 0;
        const int tbl[8][4] = {{0, 3, 0, 0}, {2, 3, 0, 0}, {0, 0, 0, 0}, {0, 2, 0, 0},
                               {0, 0, 3, 0}, {3, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
        for (int c = 0; c < 8; ++c)
            for (int e = 0; e < 4; ++e) v |= static_cast<unsigned long long>(tbl[c][e]) << (2 * (c * 4 + e));
        return v;
    }();
    const int idx = (kase & 7) * 4 + from;
    const unsigned long long w = (kase & 8) ? hi : lo;
    return static_cast<int>((w >> (2 * idx)) & 3ull);
}

// inverse: the `from` edge of the segment of this case that ends at edge `to`
OCTM_HD int seg_from(int kase, int to) {
    constexpr unsigned long long lo = []() {
        unsigned long long v = 0;
        // case 1 T>L: from[L]=T | 2 R>T: from[T]=R | 3 R>L: from[L]=R | 4 L>B: from[B]=L | 5 T>B: from[B]=T
        // 6 R>T, L>B: from[T]=R, from[B]=L | 7 R>B: from[B]=R
        const int tbl[8][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {3, 0, 0, 0}, {0, 0, 3, 0},
                               {0, 2, 0, 0}, {0, 0, 0, 0}, {3, 2, 0, 0}, {0, 3, 0, 0}};
        for (int c = 0; c < 8; ++c)
            for (int e = 0; e < 4; ++e) v |= static_cast<unsigned long long>(tbl[c][e]) << (2 * (c * 4 + e));
        return v;
    }();
    constexpr unsigned long long hi = []() {
        unsigned long long v = 0;
        // 8 B>R: from[R]=B | 9 T>L, B>R: from[L]=T, from[R]=B | 10 B>T: from[T]=B | 11 B>L: from[L]=B
        // 12 L>R: from[R]=L | 13 T>R: from[R]=T | 14 L>T: from[T]=L
        const int tbl[8][4] = {{0, 0, 0, 1}, {0, 0, 0, 1}, {1, 0, 0, 0}, {0, 0, 1, 0},
                               {0, 0, 0, 2}, {0, 0, 0, 0}, {2, 0, 0, 0}, {0, 0, 0, 0}};
        for (int c = 0; c < 8; ++c)
            for (int e = 0; e < 4; ++e) v |= static_cast<unsigned long long>(tbl[c][e]) << (2 * (c * 4 + e));
        return v;
    }();
    const int idx = (kase & 7) * 4 + to;
    const unsigned long long w = (kase & 8) ? hi : lo;
    return static_cast<int>((w >> (2 * idx)) & 3ull);
}

// emission order of the segment starting at `from` inside its square (1 only for the second
// segment of a saddle: case 6 L>B, case 9 B>R)
OCTM_HD int seg_order(int kase, int from) {
    return ((kase == 6 && from == 2) || (kase == 9 && from == 1)) ? 1 : 0;
}

// first emitted segment of a mixed square: its `from` edge
OCTM_HD int first_from(int kase) {
    // 1 T, 2 R, 3 R, 4 L, 5 T, 6 R, 7 R, 8 B, 9 T, 10 B, 11 B, 12 L, 13 T, 14 L
    constexpr unsigned int tbl = (0u << 2) | (3u << 4) | (3u << 6) | (2u << 8) | (0u << 10) | (3u << 12) | (3u << 14) |
                                 (1u << 16) | (0u << 18) | (1u << 20) | (1u << 22) | (2u << 24) | (0u << 26) | (2u << 28);
    return static_cast<int>((tbl >> (2 * kase)) & 3u);
}

OCTM_HD uint32_t edge_vertex(int r0, int c0, int e) {
    // doubled lattice: T (2r0, 2c0+1)  B (2r0+2, 2c0+1)  L (2r0+1, 2c0)  R (2r0+1, 2c0+2)
    const int y = 2 * r0 + (e == 1 ? 2 : (e >= 2 ? 1 : 0));
    const int x = 2 * c0 + (e == 3 ? 2 : (e <= 1 ? 1 : 0));
    return (static_cast<uint32_t>(y) << 16) | static_cast<uint32_t>(x);
}

struct TraceResult {
    uint32_t npts;
    bool closed;
};

// Walks contour [0] of the binary image `mask(r, c) -> 0/1` of size H x W (both >= 2).
// `seed` = flat index of the raster-first pixel whose mask value differs from pixel (0, 0).
// Calls emit(i, packed_vertex) for i = 0..npts-1 (forward run from the first segment's `to` end;
// then, if the polyline closed, the repeated vertex; else the first segment's `from` end and the
// backward run).
template <class Mask, class Emit>
OCTM_HD TraceResult trace_first_contour(int H, int W, uint32_t seed, Mask mask, Emit emit) {
    auto kase_at = [&](int r0, int c0) -> int {
        return mask(r0, c0) | (mask(r0, c0 + 1) << 1) | (mask(r0 + 1, c0) << 2) | (mask(r0 + 1, c0 + 1) << 3);
    };
    TraceResult res{0, false};
    const int s0 = mask(0, 0);
    const int rs = static_cast<int>(seed / static_cast<uint32_t>(W)), cs = static_cast<int>(seed % static_cast<uint32_t>(W));
    int r0, c0;
    if (rs >= 1) {
        r0 = rs - 1;
        c0 = cs - 1 > 0 ? cs - 1 : 0;
    } else {
        // row 0 is uniform up to column cs: the first mixed square of square-row 0 may already be
        // caused by row 1
        int q = cs;
        for (int x = 0; x < cs; ++x)
            if (mask(1, x) != s0) { q = x; break; }
        r0 = 0;
        c0 = q - 1 > 0 ? q - 1 : 0;
    }
    const int k0 = kase_at(r0, c0);
    const int from0 = first_from(k0);
    const int to0 = seg_to(k0, from0);
    const int sr = r0, sc = c0;
    unsigned long long last_key = (static_cast<unsigned long long>(r0) * W + c0) * 2 + 0;
    uint32_t last_to = edge_vertex(r0, c0, to0);
    emit(res.npts++, last_to);
    int e = to0;
    bool open_fwd = false;
    for (;;) {
        const int nr = r0 + (e == 1) - (e == 0), nc = c0 + (e == 3) - (e == 2);
        if (nr < 0 || nc < 0 || nr > H - 2 || nc > W - 2) { open_fwd = true; break; }
        const int entry = e ^ 1;
        if (nr == sr && nc == sc && entry == from0) { res.closed = true; break; }
        const int kk = kase_at(nr, nc);
        const int to = seg_to(kk, entry);
        const uint32_t v = edge_vertex(nr, nc, to);
        emit(res.npts++, v);
        const unsigned long long key = (static_cast<unsigned long long>(nr) * W + nc) * 2 + seg_order(kk, entry);
        if (key > last_key) { last_key = key; last_to = v; }
        r0 = nr; c0 = nc; e = to;
    }
    if (res.closed) {
        emit(res.npts++, last_to);   // the vertex find_contours repeats when the polyline closes
    } else if (open_fwd) {
        r0 = sr; c0 = sc; e = from0;
        emit(res.npts++, edge_vertex(r0, c0, e));
        for (;;) {
            const int nr = r0 + (e == 1) - (e == 0), nc = c0 + (e == 3) - (e == 2);
            if (nr < 0 || nc < 0 || nr > H - 2 || nc > W - 2) break;
            const int exit_edge = e ^ 1;                 // that square's segment ends on the shared edge
            const int kk = kase_at(nr, nc);
            const int from = seg_from(kk, exit_edge);
            emit(res.npts++, edge_vertex(nr, nc, from));
            r0 = nr; c0 = nc; e = from;
        }
    }
    return res;
}

}  // namespace octm
