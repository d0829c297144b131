// geom.cuh -- the reference's image/distance arithmetic, shared by K1 and K2/K4 so
// that every kernel sees bit-identical neighbour coordinates and distances.
#pragma once

namespace gapcu {

// xyz = pos(j,:) + n1*lat(1,:) + n2*lat(2,:) + n3*lat(3,:), dr = pos(i,:) - xyz,
// dis = dsqrt(dr(1)**2 + dr(2)**2 + dr(3)**2)        (gap_calc.f90:98-100)
// evaluated left to right with IEEE round-to-nearest and no FMA contraction.
__device__ __forceinline__ double image_distance_xyz(double xj, double yj, double zj, const double *lat,
                                                     int n1, int n2, int n3, double xi, double yi, double zi,
                                                     double &ox, double &oy, double &oz) {
    const double d1 = (double)n1, d2 = (double)n2, d3 = (double)n3;
    ox = __dadd_rn(__dadd_rn(__dadd_rn(xj, __dmul_rn(d1, lat[0])), __dmul_rn(d2, lat[3])), __dmul_rn(d3, lat[6]));
    oy = __dadd_rn(__dadd_rn(__dadd_rn(yj, __dmul_rn(d1, lat[1])), __dmul_rn(d2, lat[4])), __dmul_rn(d3, lat[7]));
    oz = __dadd_rn(__dadd_rn(__dadd_rn(zj, __dmul_rn(d1, lat[2])), __dmul_rn(d2, lat[5])), __dmul_rn(d3, lat[8]));
    const double dx = __dsub_rn(xi, ox), dy = __dsub_rn(yi, oy), dz = __dsub_rn(zi, oz);
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}
// The same image and difference, stopped before the square root: dis > c  <=>  d2 > T(c) with
// T(c) = max{x : sqrt_rn(x) <= c} (sqrt_rn is monotone; potential.cpp:sqrt_threshold), so the
// reference's test "dis .gt. rcut" can be taken on d2 without evaluating the root.
__device__ __forceinline__ double image_dist2_xyz(double xj, double yj, double zj, const double *lat,
                                                  int n1, int n2, int n3, double xi, double yi, double zi) {
    const double d1 = (double)n1, d2 = (double)n2, d3 = (double)n3;
    const double ox = __dadd_rn(__dadd_rn(__dadd_rn(xj, __dmul_rn(d1, lat[0])), __dmul_rn(d2, lat[3])), __dmul_rn(d3, lat[6]));
    const double oy = __dadd_rn(__dadd_rn(__dadd_rn(yj, __dmul_rn(d1, lat[1])), __dmul_rn(d2, lat[4])), __dmul_rn(d3, lat[7]));
    const double oz = __dadd_rn(__dadd_rn(__dadd_rn(zj, __dmul_rn(d1, lat[2])), __dmul_rn(d2, lat[5])), __dmul_rn(d3, lat[8]));
    const double dx = __dsub_rn(xi, ox), dy = __dsub_rn(yi, oy), dz = __dsub_rn(zi, oz);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
__device__ __forceinline__ double image_distance(const double *pos, int ntot, int j, const double *lat,
                                                 int n1, int n2, int n3, double xi, double yi, double zi,
                                                 double &ox, double &oy, double &oz) {
    return image_distance_xyz(pos[j], pos[ntot + j], pos[2 * ntot + j], lat, n1, n2, n3, xi, yi, zi, ox, oy, oz);
}

// rjk**2 before the square root (wacsf.f90:241), from absolute image coordinates
__device__ __forceinline__ double pair_dist2(double xa, double ya, double za, double xb, double yb, double zb) {
    const double dx = __dsub_rn(xa, xb), dy = __dsub_rn(ya, yb), dz = __dsub_rn(za, zb);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

}  // namespace gapcu
