// Host-compiled harness around csrc/trace_core.h -- TEST ONLY (never linked into liboctm.so).
// Lets the CPU test tier check the marching-squares tables and the contour-[0] walk that the CUDA
// trace kernel instantiates, against oracle/contours_oracle.py, without a GPU.
#include <stdint.h>

#include "../../retinal_oct_image_segmentation_via_deep_learning_b200/csrc/trace_core.h"

extern "C" int trace_check(const uint8_t* mask, int H, int W, uint32_t* out, int cap, int* closed) {
    if (H < 2 || W < 2) return 0;
    const uint8_t s0 = mask[0] ? 1 : 0;
    uint32_t seed = 0xFFFFFFFFu;
    for (long long i = 0; i < static_cast<long long>(H) * W; ++i)
        if ((mask[i] ? 1 : 0) != s0) { seed = static_cast<uint32_t>(i); break; }
    if (seed == 0xFFFFFFFFu) return 0;
    auto px = [&](int rr, int cc) -> int { return mask[static_cast<long long>(rr) * W + cc] ? 1 : 0; };
    const octm::TraceResult r = octm::trace_first_contour(
        H, W, seed,
        [&](int r0, int c0) -> int { return px(r0, c0) | (px(r0, c0 + 1) << 1) | (px(r0 + 1, c0) << 2) | (px(r0 + 1, c0 + 1) << 3); },
        [&](int r0, int c0, int e) -> int {
            const int ar = r0 + (e == 1), ac = c0 + (e == 3);
            const int br = ar + (e >= 2), bc = ac + (e < 2);
            return px(ar, ac) | (px(br, bc) << 1);
        },
        px, [](int idx6) -> uint32_t { return octm::step_word(idx6); },
        [&](uint32_t i, uint32_t v) { if (i < static_cast<uint32_t>(cap)) out[i] = v; });
    *closed = r.closed ? 1 : 0;
    return static_cast<int>(r.npts);
}
