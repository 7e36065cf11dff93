// Plane-sweep sampling position: the arithmetic of reference models/module.py:89-114 followed by
// ATen's grid_sampler un-normalisation (align_corners=True), for one (pixel, depth, view).
#pragma once
#include "common.cuh"

namespace imvs {

struct Tap {
    int x0, y0;        // top-left tap in source-feature pixels (may be -1 .. W1)
    float fx, fy;      // bilinear fractions
    unsigned mask;     // bit0 (y0,x0) bit1 (y0,x0+1) bit2 (y0+1,x0) bit3 (y0+1,x0+1): tap inside the map
};

// P: 12 floats, rot row-major then trans.  (X, Y) = reference pixel scaled to the source-feature
// resolution (module.py:95-96).  (Wd, Hd) = size of the DEPTH map: the reference substitutes
// (Wd, Hd, 1) for points with z <= 1e-2 (module.py:105-108) -- in depth-map units, a quirk kept.
__device__ __forceinline__ Tap project_tap(const float* __restrict__ P, float X, float Y, float depth,
                                           float Wd, float Hd, int W1, int H1) {
    float rx = fmaf(P[0], X, fmaf(P[1], Y, P[2]));
    float ry = fmaf(P[3], X, fmaf(P[4], Y, P[5]));
    float rz = fmaf(P[6], X, fmaf(P[7], Y, P[8]));
    float px = fmaf(rx, depth, P[9]);
    float py = fmaf(ry, depth, P[10]);
    float pz = fmaf(rz, depth, P[11]);
    if (!(pz > 1e-2f)) { px = Wd; py = Hd; pz = 1.0f; }
    float u = px / pz, v = py / pz;
    // module.py:112-113 then grid_sampler_unnormalize(align_corners=True): same two roundings
    float hx = (float)(W1 - 1) * 0.5f, hy = (float)(H1 - 1) * 0.5f;
    float gx = u / hx - 1.0f, gy = v / hy - 1.0f;
    float ix = ((gx + 1.0f) * 0.5f) * (float)(W1 - 1);
    float iy = ((gy + 1.0f) * 0.5f) * (float)(H1 - 1);
    // clamp far-out / NaN positions so the float->int conversion is defined; every tap of a
    // clamped position is outside the map, so the clamp never changes a result
    ix = fminf(fmaxf(ix, -2.0f), (float)W1 + 1.0f);
    iy = fminf(fmaxf(iy, -2.0f), (float)H1 + 1.0f);
    float x0f = floorf(ix), y0f = floorf(iy);
    Tap t;
    t.x0 = (int)x0f;
    t.y0 = (int)y0f;
    t.fx = ix - x0f;
    t.fy = iy - y0f;
    bool xa = t.x0 >= 0 && t.x0 <= W1 - 1, xb = t.x0 + 1 >= 0 && t.x0 + 1 <= W1 - 1;
    bool ya = t.y0 >= 0 && t.y0 <= H1 - 1, yb = t.y0 + 1 >= 0 && t.y0 + 1 <= H1 - 1;
    t.mask = (unsigned)(xa && ya) | ((unsigned)(xb && ya) << 1) | ((unsigned)(xa && yb) << 2) | ((unsigned)(xb && yb) << 3);
    return t;
}

}  // namespace imvs
