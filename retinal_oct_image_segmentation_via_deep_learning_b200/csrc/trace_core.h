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

// the other end of the segment of case `kase` that touches edge `e` (every edge of a mixed square
// belongs to exactly one of its segments, so this serves the forward and the backward walk alike)
OCTM_HD int other_end(int kase, int e) {
    // 2 bits per (case, edge): seg_to for `from` edges, seg_from for `to` edges
    constexpr unsigned long long lo = []() {
        unsigned long long v = 0;
        // 1 T-L | 2 R-T | 3 R-L | 4 L-B | 5 T-B | 6 R-T, L-B | 7 R-B       (edges T=0 B=1 L=2 R=3)
        const int tbl[8][4] = {{0, 0, 0, 0}, {2, 0, 0, 0}, {3, 0, 0, 0}, {0, 0, 3, 2},
                               {0, 2, 1, 0}, {1, 0, 0, 0}, {3, 2, 1, 0}, {0, 3, 0, 1}};
        for (int c = 0; c < 8; ++c)
            for (int e = 0; e < 4; ++e) v |= static_cast<unsigned long long>(tbl[c][e]) << (2 * (c * 4 + e));
        return v;
    }();
    constexpr unsigned long long hi = []() {
        unsigned long long v = 0;
        // 8 B-R | 9 T-L, B-R | 10 B-T | 11 B-L | 12 L-R | 13 T-R | 14 L-T
        const int tbl[8][4] = {{0, 3, 0, 1}, {2, 3, 0, 1}, {1, 0, 0, 0}, {0, 2, 1, 0},
                               {0, 0, 3, 2}, {3, 0, 0, 0}, {2, 0, 0, 0}, {0, 0, 0, 0}};
        for (int c = 0; c < 8; ++c)
            for (int e = 0; e < 4; ++e) v |= static_cast<unsigned long long>(tbl[c][e]) << (2 * (c * 4 + e));
        return v;
    }();
    const unsigned long long w = (kase & 8) ? hi : lo;
    return static_cast<int>((w >> (2 * ((kase & 7) * 4 + e))) & 3ull);
}

// packed doubled-lattice offset of edge e's midpoint inside its square: (dy << 16) | dx with
// T (0, 1)  B (2, 1)  L (1, 0)  R (1, 2)
OCTM_HD uint32_t edge_offset(int e) {
    constexpr unsigned long long tbl = 0x0001ull | (0x0201ull << 16) | (0x0100ull << 32) | (0x0102ull << 48);
    const uint32_t f = static_cast<uint32_t>(tbl >> (16 * e)) & 0xffffu;     // dy << 8 | dx
    return ((f & 0xff00u) << 8) | (f & 0xffu);
}

// Stepping from a square of case `kase` across its edge e, two pixels of the next square are already
// known.  Returns those carried bits in the next square's case, plus where its two NEW pixels go:
//   e = T (up)    : new ul (bit 0), ur (bit 1);  ll, lr carried from ul, ur
//   e = B (down)  : new ll (bit 2), lr (bit 3);  ul, ur carried from ll, lr
//   e = L (left)  : new ul (bit 0), ll (bit 2);  ur, lr carried from ul, ll
//   e = R (right) : new ur (bit 1), lr (bit 3);  ul, ll carried from ur, lr
OCTM_HD int carried_bits(int kase, int e) {
    return e == 0 ? (kase & 3) << 2 : e == 1 ? kase >> 2 : e == 2 ? (kase & 5) << 1 : (kase >> 1) & 5;
}
OCTM_HD int new_bit_a(int e) { return e == 1 ? 2 : e == 3 ? 1 : 0; }
OCTM_HD int new_bit_b(int e) { return e == 0 ? 1 : e == 2 ? 2 : 3; }

// Everything one step of the walk needs to know about "square of case `kase` entered through edge
// `entry`", packed in one word (the CUDA kernel keeps the 64 words in shared memory):
//   bits 0-1, 16-17 : edge_offset(exit edge)        bits 8-9 : exit edge x
//   bit 12          : 1 for the second-emitted segment of a saddle (case 6 from L, case 9 from B)
//   bits 20-23      : carried_bits(kase, x)          bits 24-25 / 26-27 : new_bit_a(x) / new_bit_b(x)
OCTM_HD uint32_t step_word(int idx6 /* kase * 4 + entry */) {
    const int kase = idx6 >> 2;
    const int x = other_end(kase, idx6 & 3);
    return edge_offset(x) | (static_cast<uint32_t>(x) << 8) | ((idx6 == 26 || idx6 == 37) ? (1u << 12) : 0u) |
           (static_cast<uint32_t>(carried_bits(kase, x)) << 20) | (static_cast<uint32_t>(new_bit_a(x)) << 24) |
           (static_cast<uint32_t>(new_bit_b(x)) << 26);
}

struct TraceResult {
    uint32_t npts;
    bool closed;
};

// Walks contour [0] of a binary image of size H x W (both >= 2).
//   kase(r0, c0)   -> the marching-squares case of the 2x2 square whose top-left pixel is (r0, c0):
//                     ul | ur << 1 | ll << 2 | lr << 3 with 1 = "pixel is in the mask"
//   pix2(r0, c0, e)-> the two pixels of square (r0, c0) that are new after stepping into it across edge e of
//                     the previous square (see carried_bits), as a | b << 1
//   mask(r, c)     -> 0/1, used only to locate the first mixed square when the seed lies in row 0
//   step(idx6)     -> step_word(idx6) (a table look-up on the device)
//   seed           =  flat index of the raster-first pixel whose mask value differs from pixel (0, 0)
// Calls emit(i, packed_vertex) for i = 0..npts-1: the forward run from the first segment's `to` end;
// then, if the polyline closed, the repeated vertex; else the first segment's `from` end and the
// backward run.  Forward and backward steps run through ONE loop body (a lane of a warp that is
// already walking backward does not diverge from its neighbours that still walk forward), and a step
// reads only the two pixels it has not seen yet.
template <class Kase, class Pix2, class Mask, class Step, class Emit>
OCTM_HD TraceResult trace_first_contour(int H, int W, uint32_t seed, Kase kase_at, Pix2 pix2, Mask mask, Step step, Emit emit) {
    TraceResult res{0, false};
    const int rs = static_cast<int>(seed / static_cast<uint32_t>(W)), cs = static_cast<int>(seed % static_cast<uint32_t>(W));
    int r0, c0;
    if (rs >= 1) {
        r0 = rs - 1;
        c0 = cs - 1 > 0 ? cs - 1 : 0;
    } else {
        // row 0 is uniform up to column cs: the first mixed square of square-row 0 may already be
        // caused by row 1
        const int s0 = mask(0, 0);
        int q = cs;
        for (int x = 0; x < cs; ++x)
            if (mask(1, x) != s0) { q = x; break; }
        r0 = 0;
        c0 = q - 1 > 0 ? q - 1 : 0;
    }
    const int k0 = kase_at(r0, c0);
    const int from0 = first_from(k0);
    const int sr = r0, sc = c0;
    uint32_t t = step(k0 * 4 + from0);         // leaves the first square through the first segment's `to` edge
    const int to0 = static_cast<int>((t >> 8) & 3u);
    const uint32_t sq0 = (static_cast<uint32_t>(r0) << 16) | static_cast<uint32_t>(c0);     // r, c < 8192
    // raster-last segment seen so far (closed polylines repeat its `to` vertex): key = square, then order;
    // (r << 16 | c) orders squares like r * W + c
    uint32_t last_key = sq0 << 1;
    uint32_t last_to = (sq0 << 1) + (t & 0x00030003u);
    uint32_t n = 0;
    emit(n++, last_to);
    bool backward = false;
    for (;;) {
        const int e = static_cast<int>((t >> 8) & 3u);
        const int s = (e & 1) ? 1 : -1;
        const int nr = r0 + ((e & 2) ? 0 : s), nc = c0 + ((e & 2) ? s : 0);
        if (static_cast<unsigned>(nr) > static_cast<unsigned>(H - 2) || static_cast<unsigned>(nc) > static_cast<unsigned>(W - 2)) {
            if (backward) break;
            // the forward run left the image: restart from the first segment's `from` end, backwards
            backward = true;
            r0 = sr; c0 = sc;
            t = step(k0 * 4 + to0);            // leaves the first square through `from0`
            emit(n++, (sq0 << 1) + (t & 0x00030003u));
            continue;
        }
        const int entry = e ^ 1;
        if (!backward && nr == sr && nc == sc && entry == from0) { res.closed = true; break; }
        const uint32_t ab = static_cast<uint32_t>(pix2(nr, nc, e));
        const uint32_t kk = ((t >> 20) & 15u) | ((ab & 1u) << ((t >> 24) & 3u)) | ((ab >> 1) << ((t >> 26) & 3u));
        t = step(static_cast<int>(kk * 4u) + entry);
        const uint32_t sq = (static_cast<uint32_t>(nr) << 16) | static_cast<uint32_t>(nc);
        const uint32_t v = (sq << 1) + (t & 0x00030003u);
        emit(n++, v);
        const uint32_t key = (sq << 1) | ((t >> 12) & 1u);
        if (!backward && key > last_key) { last_key = key; last_to = v; }
        r0 = nr; c0 = nc;
    }
    if (res.closed) emit(n++, last_to);   // the vertex find_contours repeats when the polyline closes
    res.npts = n;
    return res;
}

}  // namespace octm
