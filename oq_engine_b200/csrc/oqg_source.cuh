// Rupture generation of point sources on the device (SURVEY.md section 8f
// row 1): the planar ruptures of PointSource.get_planar, one thread per
// rupture, written in the reference's planar_array_dt record layout so that
// the rest of the path (prep_rupture -> fused kernel) is unchanged.
//
// Reference: openquake/hazardlib/source/point.py:192-289 (get_planin,
// get_planar), geo/surface/planar.py:68-95 (get_rupdims), :107-185
// (_build_corners), :209-251 (build_planar), :261-320 (build_planar_array),
// geo/geodetic.py:483-517 (fast_point_at), geo/utils.py
// (spherical_to_cartesian), scalerel/wc1994.py (get_median_area),
// scalerel/point.py:40-52 (PointMSR).
//
// Ruptures of a source are ordered (magnitude, nodal plane, hypocentral
// depth) as in the reference's (M, N, D, 3) planar array.
#pragma once
#include <math.h>
#include "../../include/oqg.h"
#include "oqg_geo.cuh"

namespace oqg {

// geodetic.fast_point_at, geodetic.py:483-517 (degrees in, degrees out)
__device__ __forceinline__ void point_at(double lon, double lat, double azimuth,
                                         double distance, double &olon,
                                         double &olat) {
  const double D2R = 0.017453292519943295;
  lon *= D2R;
  lat *= D2R;
  const double tc = -(azimuth * D2R);
  double sd, cd, sl, cl, st, ct;
  sincos(distance / OQG_EARTH_RADIUS, &sd, &cd);
  sincos(lat, &sl, &cl);
  sincos(tc, &st, &ct);
  const double sin_lats = sl * cd + cl * sd * ct;
  const double dlon = atan2(st * sd * cl, cd - sl * sin_lats);
  const double PI = 3.141592653589793;
  // numpy.mod(lon - dlon + pi, 2 pi) - pi: result has the sign of 2 pi
  double x = lon - dlon + PI;
  double m = fmod(x, 2. * PI);
  if (m < 0.) m += 2. * PI;
  olon = (m - PI) * OQG_R2D;
  olat = asin(sin_lats) * OQG_R2D;
}

__device__ __forceinline__ double median_area(int msr, double par, double mag,
                                              double rake) {
  if (msr == OQG_MSR_POINT) return 1e-4;
  if (msr == OQG_MSR_PEER) return pow(10.0, mag - 4.0);
  if (msr == OQG_MSR_CSCALING) return pow(10.0, mag - par);
  // WC1994: strike slip |rake| <= 45 or >= 135; reverse rake > 0; normal
  if (fabs(rake) <= 45. || fabs(rake) >= 135.)
    return pow(10.0, -3.42 + 0.90 * mag);
  if (rake > 0.) return pow(10.0, -3.99 + 0.98 * mag);
  return pow(10.0, -2.87 + 0.82 * mag);
}

// get_rupdims planar.py:68-95: rupture length and horizontal projection of
// the width, km
__device__ __forceinline__ void rupture_dims(double usd, double lsd, double rar,
                                             double area, double dip,
                                             double &length, double &wproj) {
  double sdip, cdip;
  sincos(dip * 0.017453292519943295, &sdip, &cdip);
  const double max_width = (lsd - usd) / sdip;
  length = sqrt(area * rar);
  double width = area / length;
  if (width > max_width) {
    width = max_width;
    length = area / width;
  }
  wproj = width * cdip;
}

// one planar rupture from its parameters; pl = 42 doubles [k][14]
__device__ inline void build_planar_rupture(
    double usd, double lsd, double rar, double area, double strike, double dip,
    double rake, double clon, double clat, double dep, double dip_frac,
    double rate_weight, int shift_hypo, double *pl) {
  const double D2R = 0.017453292519943295;
  // get_rupdims planar.py:68-95
  double sdip, cdip;
  sincos(dip * D2R, &sdip, &cdip);
  const double max_width = (lsd - usd) / sdip;
  double length = sqrt(area * rar);
  double width = area / length;
  if (width > max_width) {
    width = max_width;
    length = area / width;
  }
  const double hl = length / 2., hw = (width * cdip) / 2.,
               hh = (width * sdip) / 2.;
  // _build_corners planar.py:107-185
  const double theta = atan(hw / hl) * OQG_R2D;
  const double hor = sqrt(hl * hl + hw * hw);
  const double shift = (1. - 2. * dip_frac) * hh;
  const double top = dep + shift - hh, bot = dep + shift + hh;
  double v = usd - top;
  if (v < 0.) {
    v = lsd - bot;
    if (v > 0.) v = 0.;
  }
  v = shift + v;
  double lon = clon, lat = clat, cdep = dep;
  if (v != 0.) {
    const double azi = v < 0. ? strike + 90. + 90. + 90. : strike + 90.;
    point_at(clon, clat, azi, fabs(v / hh * hw), lon, lat);
    cdep = dep + v;
  }
  double clo[4], cla[4], cde[4];
  point_at(lon, lat, strike + 180. + theta, hor, clo[0], cla[0]);
  point_at(lon, lat, strike - theta, hor, clo[1], cla[1]);
  point_at(lon, lat, strike + 180. - theta, hor, clo[2], cla[2]);
  point_at(lon, lat, strike + theta, hor, clo[3], cla[3]);
  cde[0] = cde[1] = cdep - hh;
  cde[2] = cde[3] = cdep + hh;
  // build_planar_array planar.py:261-320
  double xyz[4][3];
  for (int c = 0; c < 4; c++) {
    double sl, cl, sp, cp;
    sincos(clo[c] * D2R, &sl, &cl);
    sincos(cla[c] * D2R, &sp, &cp);
    const double rr = OQG_EARTH_RADIUS - cde[c];
    const double ctr = rr * cp;
    xyz[c][0] = ctr * cl;
    xyz[c][1] = ctr * sl;
    xyz[c][2] = rr * sp;
  }
  double a[3], b[3], n[3], uv1[3], uv2[3];
  for (int k = 0; k < 3; k++) {
    a[k] = xyz[0][k] - xyz[1][k];   // tl - tr
    b[k] = xyz[0][k] - xyz[2][k];   // tl - bl
  }
  n[0] = a[1] * b[2] - a[2] * b[1];
  n[1] = a[2] * b[0] - a[0] * b[2];
  n[2] = a[0] * b[1] - a[1] * b[0];
  double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  for (int k = 0; k < 3; k++) n[k] /= nn;
  for (int k = 0; k < 3; k++) uv1[k] = -a[k];   // tr - tl
  double un = sqrt(uv1[0] * uv1[0] + uv1[1] * uv1[1] + uv1[2] * uv1[2]);
  for (int k = 0; k < 3; k++) uv1[k] /= un;
  uv2[0] = n[1] * uv1[2] - n[2] * uv1[1];
  uv2[1] = n[2] * uv1[0] - n[0] * uv1[2];
  uv2[2] = n[0] * uv1[1] - n[1] * uv1[0];
  double xx[4], yy[4];
  xx[0] = yy[0] = 0.;
  for (int c = 1; c < 4; c++) {
    double d0 = xyz[c][0] - xyz[0][0], d1 = xyz[c][1] - xyz[0][1],
           d2 = xyz[c][2] - xyz[0][2];
    xx[c] = d0 * uv1[0] + d1 * uv1[1] + d2 * uv1[2];
    yy[c] = d0 * uv2[0] + d1 * uv2[1] + d2 * uv2[2];
  }
  const double plen = ((xx[1] - xx[0]) + (xx[3] - xx[2])) / 2.0;
  const double pwid = ((yy[2] - yy[0]) + (yy[3] - yy[1])) / 2.0;
  // record layout [k][14]: corners[4] xyz[4] normal uv1 uv2 wlr sdr hypo
  for (int c = 0; c < 4; c++) {
    pl[0 * 14 + c] = clo[c];
    pl[1 * 14 + c] = cla[c];
    pl[2 * 14 + c] = cde[c];
    for (int k = 0; k < 3; k++) pl[k * 14 + 4 + c] = xyz[c][k];
  }
  for (int k = 0; k < 3; k++) {
    pl[k * 14 + 8] = n[k];
    pl[k * 14 + 9] = uv1[k];
    pl[k * 14 + 10] = uv2[k];
  }
  pl[0 * 14 + 11] = pwid; pl[1 * 14 + 11] = plen; pl[2 * 14 + 11] = rate_weight;
  pl[0 * 14 + 12] = strike; pl[1 * 14 + 12] = dip; pl[2 * 14 + 12] = rake;
  pl[0 * 14 + 13] = clon; pl[1 * 14 + 13] = clat;
  pl[2 * 14 + 13] = shift_hypo ? cdep : dep;
}

}  // namespace oqg
