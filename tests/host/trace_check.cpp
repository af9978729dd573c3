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
    const octm::TraceResult r = octm::trace_first_contour(
        H, W, seed, [&](int rr, int cc) -> int { return mask[static_cast<long long>(rr) * W + cc] ? 1 : 0; },
        [&](uint32_t i, uint32_t v) { if (i < static_cast<uint32_t>(cap)) out[i] = v; });
    *closed = r.closed ? 1 : 0;
    return static_cast<int>(r.npts);
}
