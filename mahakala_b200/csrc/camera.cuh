// Camera-ray initialisation as device functions (shared by the stand-alone camera kernels and the
// fused render kernel, which builds each ray's initial state in registers from its pixel index).
//
// Restates /root/reference/mahakala/geodesics.py:
//   _Image_to_BH :204-209, _perpendicular :212-216, k = cross(origin - p, perp - p) :174-179 / :127-132,
//   _quadratic :58-66, _Nullify.nullify :69-85 (p = 1), initial_condition :219-230.
// The image-plane arithmetic uses explicit non-contracted IEEE operations (__dmul_rn/__dadd_rn) in the
// reference's order, so initial positions and directions are bit-identical to the NumPy host math of
// the reference; initial-state errors are amplified ~1e5x along near-critical rays.
#pragma once
#include "ks_metric.cuh"
#include "camera_nullify.cuh"

namespace mk {

struct CameraGeom {
    double ci, si, d;      // cos(incl), sin(incl), distance
};

__device__ __forceinline__ void image_to_bh(const CameraGeom& c, double x, double y, double z, double out[3])
{
    // x_BH = -y cos i + z sin i + d sin i ;  y_BH = x ;  z_BH = y sin i + z cos i + d cos i
    out[0] = __dadd_rn(__dadd_rn(__dmul_rn(-y, c.ci), __dmul_rn(z, c.si)), __dmul_rn(c.d, c.si));
    out[1] = x;
    out[2] = __dadd_rn(__dadd_rn(__dmul_rn(y, c.si), __dmul_rn(z, c.ci)), __dmul_rn(c.d, c.ci));
}

// position (t = 0) and un-normalised direction (k^t = 1) of the ray through image point (xi, yi)
__device__ __forceinline__ void camera_point(const CameraGeom& c, double xi, double yi, double x[4], double v[4])
{
    double o[3], p[3], q[3];
    image_to_bh(c, 0.0, 0.0, 0.0, o);
    image_to_bh(c, xi, yi, 0.0, p);
    image_to_bh(c, __dadd_rn(xi, yi), __dsub_rn(yi, xi), 0.0, q);      // _perpendicular
    double a0 = __dadd_rn(-p[0], o[0]), a1 = __dadd_rn(-p[1], o[1]), a2 = __dadd_rn(-p[2], o[2]);
    double b0 = __dsub_rn(q[0], p[0]), b1 = __dsub_rn(q[1], p[1]), b2 = __dsub_rn(q[2], p[2]);
    x[0] = 0.0; x[1] = p[0]; x[2] = p[1]; x[3] = p[2];
    v[0] = 1.0;
    v[1] = __dsub_rn(__dmul_rn(a1, b2), __dmul_rn(a2, b1));
    v[2] = __dsub_rn(__dmul_rn(a2, b0), __dmul_rn(a0, b2));
    v[3] = __dsub_rn(__dmul_rn(a0, b1), __dmul_rn(a1, b0));
}

// np.linspace(lo, hi, 2n+1)[1::2][j]  ==  (2j+1) * step + lo  with step = (hi - lo) / (2n)
__device__ __forceinline__ double pixel_centre(double lo, double step, long j)
{
    return __dadd_rn(__dmul_rn((double)(2 * j + 1), step), lo);
}

// the same with the literal Kerr-Schild metric of geodesics.py:95-104 (IEEE division / sqrt; once per ray)
__device__ __forceinline__ void nullify_state(const KerrSchild& g, const double x[4], const double v[4], double s[8])
{
    double f, l[4];
    l[0] = 1.0;
    {
        double zz = x[3] * x[3];
        double kk = 0.5 * (x[1] * x[1] + x[2] * x[2] + zz - g.aa);
        double rr = sqrt(kk * kk + g.aa * zz) + kk;
        double r = sqrt(rr);
        f = (2.0 * rr * r) / (rr * rr + g.aa * zz);
        l[1] = (r * x[1] + g.a * x[2]) / (rr + g.aa);
        l[2] = (r * x[2] - g.a * x[1]) / (rr + g.aa);
        l[3] = x[3] / r;
    }
    double gm[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) gm[i][j] = ((i == j) ? (i == 0 ? -1.0 : 1.0) : 0.0) + f * (l[i] * l[j]);
    nullify_with_metric(gm, x, v, s);
}

}  // namespace mk
