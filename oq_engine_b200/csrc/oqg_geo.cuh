// Rupture-to-site distance metrics for planar ruptures as device functions.
//
// Reference: openquake/hazardlib/geo/surface/planar.py (project :324-403,
// get_rjb :460-529, get_rx :533-545, get_ry0 :549-568, get_rhypo :572-586,
// get_repi :589-602, get_azimuth :605-621), geo/geodetic.py (distances
// :38-48, distances_to_arc :51-66) and geo/utils.py (min_distance :710-720).
//
// The reference evaluates the cross-track distance to a great-circle arc as
//    (pi/2 - acos(sin(azimuth_to_target - azimuth) * sin(dist / R))) * R
// with ~20 transcendental calls.  Here the same quantity is computed from
// unit vectors: with p the arc origin, q the site and w = e_p cos(azi) -
// n_p sin(azi) (e_p, n_p the local east / north unit vectors at p) one has
//    sin(azimuth_to_target - azimuth) * sin(dist / R) == (q - p) . w
// exactly, so the distance is R * asin((q - p) . w): one dot product and one
// asin, better conditioned than the reference formula (which loses ~1e-12
// relative to cancellation in lon - lons).  The great-circle distance comes
// from the chord: sin(dist / 2R) = |q - p| / 2.
//
// A rupture is prepared once (prep_rupture) into NF_RUP doubles stored
// field-major in HBM (rup[field][R]) so that rupture tiles load coalesced.
#pragma once
#include <math.h>
#include "../../include/oqg.h"

namespace oqg {

// min of two doubles as one DSETP and two selects: fmin() on sm_100 is a
// DSETP.MIN plus NaN handling and moves, 8 instructions (there is no DMNMX).
// A NaN second argument is ignored like in fmin(); the first one must not be
// NaN (a running minimum).
__device__ __forceinline__ double dmin(double a, double b) {
  return b < a ? b : a;
}

#define OQG_EARTH_RADIUS 6371.0  // geodetic.py:32
#define OQG_R2D 57.29577951308232

// fields of a prepared rupture
enum {
  RF_TL = 0,       // xyz of the top-left corner, km (3)
  RF_NRM = 3,      // unit normal (3)
  RF_UV1 = 6,      // unit vector along strike (3)
  RF_UV2 = 9,      // unit vector down dip (3)
  RF_WIDTH = 12,
  RF_LENGTH = 13,
  RF_W0 = 14,      // arc normals: BL/strike, TL/strike, TL/downdip, TR/downdip
  RF_P0 = 26,      // arc origins as unit vectors BL, TL, TL, TR -> only 3
                   // distinct: BL (26), TL (29), TR (32)
  RF_CS = 35,      // surface projections of the corners TL TR BL BR, km (12)
  RF_MAG = 47, RF_RAKE = 48, RF_DIP = 49, RF_ZTOR = 50, RF_HYPO_DEPTH = 51,
  RF_RATE = 52, RF_MAGDIST = 53,
  // the fields only repi, rhypo and the azimuth read come last: the fused
  // kernel stages the first NF_RUP_NOHYPO fields for the other models
  NF_RUP_NOHYPO = 54,
  RF_HYP = 54,     // hypocentre unit vector (3)
  RF_HDEP = 57,    // hypocentre depth
  RF_HE = 58,      // east unit vector at the hypocentre (3)
  RF_HN = 61,      // north unit vector at the hypocentre (3)
  RF_STRIKE = 64,
  NF_RUP = 65
};

// fields of a prepared site
enum {
  SF_XYZ = 0,   // km (3)
  SF_Q = 3,     // unit vector (3)
  SF_DEPTH = 6, // EARTH_RADIUS - |xyz|
  NF_SITE = 7
};

struct Dist {
  double rrup, rjb, rx, ry0, repi, rhypo, azimuth, clon, clat;
};

__device__ __forceinline__ void unit_from_lonlat(double lon_deg, double lat_deg,
                                                 double *p, double *e,
                                                 double *n) {
  double sl, cl, sp, cp;
  sincos(lon_deg * 0.017453292519943295, &sl, &cl);
  sincos(lat_deg * 0.017453292519943295, &sp, &cp);
  p[0] = cp * cl; p[1] = cp * sl; p[2] = sp;
  if (e) { e[0] = -sl; e[1] = cl; e[2] = 0.; }
  if (n) { n[0] = -sp * cl; n[1] = -sp * sl; n[2] = cp; }
}

// planar: 42 doubles, the memory of one (3,) record array of
// planar_array_dt (planar.py:45-53);
// writes field f of the prepared rupture to out[f * stride]
__device__ inline void prep_rupture(const double *pl, double mag,
                                    double magdist, double *out,
                                    int64_t stride) {
  // memory layout of a (3,) record array of planar_array_dt: [k][14] with
  // k = lon/lat/dep (x/y/z) and 14 = corners[4] xyz[4] normal uv1 uv2 wlr
  // sdr hypo
#define CORNER(k, c) pl[(k) * 14 + (c)]
#define XYZ(k, c) pl[(k) * 14 + 4 + (c)]
  const double wlr[3] = {pl[11], pl[14 + 11], pl[28 + 11]};
  const double sdr[3] = {pl[12], pl[14 + 12], pl[28 + 12]};
  const double hypo[3] = {pl[13], pl[14 + 13], pl[28 + 13]};
#define PUT(f, v) out[(int64_t)(f) * stride] = (v)
  for (int k = 0; k < 3; k++) {
    PUT(RF_TL + k, XYZ(k, 0));
    PUT(RF_NRM + k, pl[k * 14 + 8]);
    PUT(RF_UV1 + k, pl[k * 14 + 9]);
    PUT(RF_UV2 + k, pl[k * 14 + 10]);
  }
  PUT(RF_WIDTH, wlr[0]);
  PUT(RF_LENGTH, wlr[1]);
  const double strike = sdr[0];
  const double downdip = fmod(strike + 90., 360.);
  double ss, cs, sd, cd;
  sincos(strike * 0.017453292519943295, &ss, &cs);
  sincos(downdip * 0.017453292519943295, &sd, &cd);
  // corner order TL TR BL BR; arcs: (BL, strike) (TL, strike) (TL, downdip)
  // (TR, downdip), planar.py:481-498
  const int arc_corner[4] = {2, 0, 0, 1};
  for (int a = 0; a < 4; a++) {
    int c = arc_corner[a];
    double p[3], e[3], n[3];
    unit_from_lonlat(CORNER(0, c), CORNER(1, c), p, e, n);
    double sa = a < 2 ? ss : sd, ca = a < 2 ? cs : cd;
    for (int k = 0; k < 3; k++) PUT(RF_W0 + 3 * a + k, e[k] * ca - n[k] * sa);
  }
  const int origin_corner[3] = {2, 0, 1};  // BL, TL, TR
  for (int j = 0; j < 3; j++) {
    double p[3];
    int c = origin_corner[j];
    unit_from_lonlat(CORNER(0, c), CORNER(1, c), p, 0, 0);
    for (int k = 0; k < 3; k++) PUT(RF_P0 + 3 * j + k, p[k]);
  }
  for (int c = 0; c < 4; c++) {
    double p[3];
    unit_from_lonlat(CORNER(0, c), CORNER(1, c), p, 0, 0);
    for (int k = 0; k < 3; k++)
      PUT(RF_CS + 3 * c + k, p[k] * OQG_EARTH_RADIUS);
  }
  {
    double p[3], e[3], n[3];
    unit_from_lonlat(hypo[0], hypo[1], p, e, n);
    for (int k = 0; k < 3; k++) {
      PUT(RF_HYP + k, p[k]);
      PUT(RF_HE + k, e[k]);
      PUT(RF_HN + k, n[k]);
    }
  }
  PUT(RF_HDEP, hypo[2]);
  PUT(RF_STRIKE, strike);
  PUT(RF_MAG, mag);
  PUT(RF_RAKE, sdr[2]);
  PUT(RF_DIP, sdr[1]);
  PUT(RF_ZTOR, CORNER(2, 0));
  PUT(RF_HYPO_DEPTH, hypo[2]);
  PUT(RF_RATE, wlr[2]);
  PUT(RF_MAGDIST, magdist);
#undef PUT
#undef CORNER
#undef XYZ
}

// site: lon/lat in degrees, depth km -> xyz (geodetic.py:445-480), the unit
// vector and the depth recovered from xyz (geo/utils.py:688-707)
__device__ inline void prep_site(double lon, double lat, double depth,
                                 double *out, int64_t stride) {
  double p[3];
  unit_from_lonlat(lon, lat, p, 0, 0);
  double rr = OQG_EARTH_RADIUS - depth;
  double x = rr * p[0], y = rr * p[1], z = rr * p[2];
  double nrm = sqrt(x * x + y * y + z * z);
  out[(SF_XYZ + 0) * stride] = x;
  out[(SF_XYZ + 1) * stride] = y;
  out[(SF_XYZ + 2) * stride] = z;
  out[(SF_Q + 0) * stride] = x / nrm;
  out[(SF_Q + 1) * stride] = y / nrm;
  out[(SF_Q + 2) * stride] = z / nrm;
  out[SF_DEPTH * stride] = OQG_EARTH_RADIUS - nrm;
}

// Distances of one rupture-site pair.  `rup` points to field 0 of the
// rupture, fields are `rs` doubles apart (shared or global memory); the site
// is given by its xyz (km), unit vector q and depth.  `mask` selects what to
// compute (rrup always).
template <typename RupPtr>
__device__ __forceinline__ void pair_distances(RupPtr rup, int rs,
                                               const double *sx,
                                               const double *sq, double sdep,
                                               int mask, Dist &d) {
#define RUP(f) rup[(f) * rs]
  // --- rrup, planar.py:324-403
  {
    double m0 = sx[0] - RUP(RF_TL), m1 = sx[1] - RUP(RF_TL + 1),
           m2 = sx[2] - RUP(RF_TL + 2);
    double dn = m0 * RUP(RF_NRM) + m1 * RUP(RF_NRM + 1) + m2 * RUP(RF_NRM + 2);
    double xx = m0 * RUP(RF_UV1) + m1 * RUP(RF_UV1 + 1) + m2 * RUP(RF_UV1 + 2);
    double yy = m0 * RUP(RF_UV2) + m1 * RUP(RF_UV2 + 1) + m2 * RUP(RF_UV2 + 2);
    double len = RUP(RF_LENGTH), wid = RUP(RF_WIDTH);
    double mxx = xx < 0. ? xx : (xx > len ? xx - len : 0.);
    double myy = yy < 0. ? yy : (yy > wid ? yy - wid : 0.);
    d.rrup = sqrt(dn * dn + mxx * mxx + myy * myy);
    if (mask & ((1 << OQG_CLON) | (1 << OQG_CLAT))) {
      // project_back planar.py:407-424: the closest point of the rupture
      // plane, back to spherical coordinates (geo/utils.py:689-707)
      const double cx = fmin(fmax(xx, 0.), len), cy = fmin(fmax(yy, 0.), wid);
      const double p0 = RUP(RF_TL) + RUP(RF_UV1) * cx + RUP(RF_UV2) * cy,
                   p1 = RUP(RF_TL + 1) + RUP(RF_UV1 + 1) * cx + RUP(RF_UV2 + 1) * cy,
                   p2 = RUP(RF_TL + 2) + RUP(RF_UV1 + 2) * cx + RUP(RF_UV2 + 2) * cy;
      const double rr = sqrt(p0 * p0 + p1 * p1 + p2 * p2);
      d.clon = atan2(p1, p0) * OQG_R2D;
      d.clat = asin(fmin(fmax(p2 / rr, -1.), 1.)) * OQG_R2D;
    }
  }
  if (mask & ((1 << OQG_RJB) | (1 << OQG_RX) | (1 << OQG_RY0))) {
    // signed sines of the four cross-track angles
    double s[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const int po = a == 0 ? 0 : (a == 3 ? 2 : 1);  // BL TL TL TR
      double e0 = sq[0] - RUP(RF_P0 + 3 * po), e1 = sq[1] - RUP(RF_P0 + 3 * po + 1),
             e2 = sq[2] - RUP(RF_P0 + 3 * po + 2);
      s[a] = e0 * RUP(RF_W0 + 3 * a) + e1 * RUP(RF_W0 + 3 * a + 1) +
             e2 * RUP(RF_W0 + 3 * a + 2);
    }
    if (mask & (1 << OQG_RX)) d.rx = asin(s[1]) * OQG_EARTH_RADIUS;
    bool same23 = (s[2] > 0.) == (s[3] > 0.) && (s[2] < 0.) == (s[3] < 0.);
    if (mask & (1 << OQG_RY0)) {
      d.ry0 = same23 ? asin(dmin(fabs(s[2]), fabs(s[3]))) * OQG_EARTH_RADIUS
                     : 0.;
    }
    if (mask & (1 << OQG_RJB)) {
      bool same01 = (s[0] > 0.) == (s[1] > 0.) && (s[0] < 0.) == (s[1] < 0.);
      if (same01 && same23) {
        // min distance to the surface projections of the corners
        double best = 1e300;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          double e0 = sx[0] - RUP(RF_CS + 3 * c), e1 = sx[1] - RUP(RF_CS + 3 * c + 1),
                 e2 = sx[2] - RUP(RF_CS + 3 * c + 2);
          best = dmin(best, e0 * e0 + e1 * e1 + e2 * e2);
        }
        d.rjb = sqrt(best);
      } else if (same01) {
        d.rjb = asin(dmin(fabs(s[0]), fabs(s[1]))) * OQG_EARTH_RADIUS;
      } else if (same23) {
        d.rjb = asin(dmin(fabs(s[2]), fabs(s[3]))) * OQG_EARTH_RADIUS;
      } else {
        d.rjb = 0.;
      }
    }
  }
  if (mask & ((1 << OQG_REPI) | (1 << OQG_RHYPO) | (1 << OQG_AZIMUTH))) {
    double e0 = sq[0] - RUP(RF_HYP), e1 = sq[1] - RUP(RF_HYP + 1),
           e2 = sq[2] - RUP(RF_HYP + 2);
    double half = 0.5 * sqrt(e0 * e0 + e1 * e1 + e2 * e2);
    d.repi = asin(fmin(half, 1.)) * 2. * OQG_EARTH_RADIUS;
    double vd = RUP(RF_HDEP) - sdep;
    d.rhypo = sqrt(d.repi * d.repi + vd * vd);
    if (mask & (1 << OQG_AZIMUTH)) {
      double E = e0 * RUP(RF_HE) + e1 * RUP(RF_HE + 1) + e2 * RUP(RF_HE + 2);
      double N = e0 * RUP(RF_HN) + e1 * RUP(RF_HN + 1) + e2 * RUP(RF_HN + 2);
      // fast_azimuth (geodetic.py:117-130) then (azimuth - strike) % 360
      double az = atan2(E, N) * OQG_R2D - RUP(RF_STRIKE);
      az = fmod(az, 360.);
      if (az < 0.) az += 360.;
      d.azimuth = az;
    }
  }
#undef RUP
}

}  // namespace oqg

// ---------------------------------------------------------------------------
// Non-planar surfaces, first step: the rupture
// distance of mesh surfaces, Mesh.get_min_distance
// (openquake/hazardlib/geo/mesh.py:365-383: cdist(self.xyz, mesh.xyz).min(0))
// for S meshes at once, and its minimum over the sections of a multi-fault
// rupture (contexts.py:969-971: numpy.min(_get(surfaces, 'rrup', dparam), 0)).
namespace oqg {

// points: [3][P] km (x, y, z planes); one CTA = (512 sites, mesh s), four
// sites per thread so that one shared-memory read of a point feeds four
// evaluations; the mesh points are staged through shared memory in chunks
// of 256.  The squared distance is expanded around the first point c of the
// mesh, |q - p|^2 = |q~|^2 + (|p~|^2 - 2 q~ . p~) with p~ = p - c, q~ = q - c:
// three DFMA per pair instead of three subtractions and three multiply-adds
// (the kernel is bound by the FP64 pipe).  With coordinates relative to the
// mesh (|p~|, |q~| <= 1e3 km instead of 6371) the cancellation costs
// <= 1e-10 km^2, i.e. <= 5e-9 km at 10 m and 5e-11 km at 1 km.
#define MESH_SPT 4
__global__ void __launch_bounds__(128)
k_mesh_min_distance(const double *__restrict__ pts, int64_t P,
                    const int64_t *__restrict__ off, int S,
                    const double *__restrict__ site, int N,
                    double *__restrict__ out, int y0) {
  __shared__ double sx[256], sy[256], sz[256], sp[256];
  const int s = blockIdx.y + y0;   // y0: the host launches grid.y in chunks
  const int n0 = blockIdx.x * (blockDim.x * MESH_SPT) + threadIdx.x;
  const int64_t pa = off[s], pb = off[s + 1];
  const double cx = pts[pa], cy = pts[P + pa], cz = pts[2 * P + pa];
  double qx[MESH_SPT], qy[MESH_SPT], qz[MESH_SPT], best[MESH_SPT];
#pragma unroll
  for (int k = 0; k < MESH_SPT; k++) {
    const int n = n0 + k * blockDim.x;
    const bool ok = n < N;
    qx[k] = (ok ? site[(int64_t)(SF_XYZ + 0) * N + n] : cx) - cx;
    qy[k] = (ok ? site[(int64_t)(SF_XYZ + 1) * N + n] : cy) - cy;
    qz[k] = (ok ? site[(int64_t)(SF_XYZ + 2) * N + n] : cz) - cz;
    best[k] = INFINITY;
  }
  for (int64_t p0 = pa; p0 < pb; p0 += 256) {
    const int np = (int)min((int64_t)256, pb - p0);
    __syncthreads();
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
      const double x = pts[p0 + i] - cx, y = pts[P + p0 + i] - cy,
                   z = pts[2 * P + p0 + i] - cz;
      sx[i] = -2. * x;
      sy[i] = -2. * y;
      sz[i] = -2. * z;
      sp[i] = fma(x, x, fma(y, y, z * z));
    }
    __syncthreads();
#pragma unroll 2
    for (int i = 0; i < np; i++) {
      const double px = sx[i], py = sy[i], pz = sz[i], pp = sp[i];
#pragma unroll
      for (int k = 0; k < MESH_SPT; k++)
        best[k] = dmin(best[k], fma(qx[k], px, fma(qy[k], py, fma(qz[k], pz, pp))));
    }
  }
#pragma unroll
  for (int k = 0; k < MESH_SPT; k++) {
    const int n = n0 + k * blockDim.x;
    if (n < N) {
      const double qq = fma(qx[k], qx[k], fma(qy[k], qy[k], qz[k] * qz[k]));
      out[(int64_t)s * N + n] = sqrt(fmax(best[k] + qq, 0.));
    }
  }
}

// lon / lat / depth -> [3][P] cartesian km (geo/utils.py spherical_to_cartesian)
__global__ void k_mesh_xyz(const double *__restrict__ lon,
                           const double *__restrict__ lat,
                           const double *__restrict__ dep, int64_t P,
                           double *__restrict__ pts) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= P) return;
  double sl, cl, sp, cp;
  sincos(lon[i] * 0.017453292519943295, &sl, &cl);
  sincos(lat[i] * 0.017453292519943295, &sp, &cp);
  const double rr = OQG_EARTH_RADIUS - (dep ? dep[i] : 0.);
  pts[i] = rr * cp * cl;
  pts[P + i] = rr * cp * sl;
  pts[2 * P + i] = rr * sp;
}

// out[u][n] = min over the sections of rupture u of sec[s][n]
__global__ void k_section_min(const double *__restrict__ sec, int N,
                              const int64_t *__restrict__ roff,
                              const int32_t *__restrict__ rsec, int U,
                              double *__restrict__ out, int y0) {
  const int u = blockIdx.y + y0;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double best = INFINITY;
  for (int64_t k = roff[u]; k < roff[u + 1]; k++)
    best = fmin(best, sec[(int64_t)rsec[k] * N + n]);
  out[(int64_t)u * N + n] = best;
}

// ---- Joyner-Boore distance of rectangular mesh surfaces ---------------------
// RectangularMesh.get_joyner_boore_distance (geo/mesh.py:453-540): the
// reference takes the minimum CHORD distance between the surface projections
// of the mesh points and the site (min_geodetic_distance, geodetic.py:331-343)
// and, for the sites closer than 40 km, replaces it by the distance to the
// polygon enclosing the mesh in the orthographic projection centred on it
// (_get_proj_enclosing_polygon :541-597: the union of the stripes between
// consecutive mesh rows, each grown by DIST_TOLERANCE = 10 m; zero inside).
//
// Here both come out of one loop over the mesh.  The mesh points are stored
// in the frame of the projection: x = R p.e, y = R p.n (the projected
// coordinates, e / n the east / north unit vectors at the projection
// centre c) and w = R p.c, so that the 3D chord is |(x, y, w)_q - (x, y, w)_p|
// and the 2D geometry reuses x and y.  The distance to the union of the
// stripes is the minimum over the grid segments (row and column
// neighbours), and zero if the site is inside any cell (even-odd test on its
// four edges).  The reference grows every stripe by 10 m and simplifies the
// union by 10 m: the net effect lies between 0 and 10 m and depends on the
// shapely / libgeos version (its test_version accepts two answers, both 2-3 m
// inside the exact outline); the exact union is taken here (MESH_JB_TOL 0),
// within the 0.02 km to which the reference's own tests pin this distance.
#define MESH_JB_TOL 0.0
#define MESH_JB_THRESHOLD 40.0

__global__ void k_mesh_frame(const double *__restrict__ lon,
                             const double *__restrict__ lat, int64_t P,
                             const int32_t *__restrict__ owner,
                             const double *__restrict__ frame,
                             double *__restrict__ pts) {
  // pts [3][P] = (x, y, w) of every mesh point in the frame [S][9] = e, n, c
  // of its mesh
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= P) return;
  double sl, cl, sp, cp;
  sincos(lon[i] * 0.017453292519943295, &sl, &cl);
  sincos(lat[i] * 0.017453292519943295, &sp, &cp);
  const double px = cp * cl, py = cp * sl, pz = sp;
  const double *f = frame + (int64_t)owner[i] * 9;
  pts[i] = OQG_EARTH_RADIUS * (px * f[0] + py * f[1] + pz * f[2]);
  pts[P + i] = OQG_EARTH_RADIUS * (px * f[3] + py * f[4] + pz * f[5]);
  pts[2 * P + i] = OQG_EARTH_RADIUS * (px * f[6] + py * f[7] + pz * f[8]);
}

__device__ __forceinline__ double seg_dist2(double x, double y, double ax,
                                            double ay, double bx, double by) {
  const double dx = bx - ax, dy = by - ay;
  const double l2 = dx * dx + dy * dy;
  double t = l2 > 0. ? ((x - ax) * dx + (y - ay) * dy) / l2 : 0.;
  t = fmin(fmax(t, 0.), 1.);
  const double ex = x - (ax + t * dx), ey = y - (ay + t * dy);
  return ex * ex + ey * ey;
}

__device__ __forceinline__ int crosses(double x, double y, double ax,
                                       double ay, double bx, double by) {
  return ((ay > y) != (by > y)) &&
         (x < (bx - ax) * (y - ay) / (by - ay) + ax);
}

__global__ void __launch_bounds__(128)
k_mesh_rjb(const double *__restrict__ pts, int64_t P,
           const int64_t *__restrict__ off, const int32_t *__restrict__ shape,
           const double *__restrict__ frame, const double *__restrict__ site,
           int N, double *__restrict__ out, int y0) {
  const int s = blockIdx.y + y0;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const double *f = frame + (int64_t)s * 9;
  const double q0 = site[(int64_t)(SF_Q + 0) * N + n],
               q1 = site[(int64_t)(SF_Q + 1) * N + n],
               q2 = site[(int64_t)(SF_Q + 2) * N + n];
  const double x = OQG_EARTH_RADIUS * (q0 * f[0] + q1 * f[1] + q2 * f[2]);
  const double y = OQG_EARTH_RADIUS * (q0 * f[3] + q1 * f[4] + q2 * f[5]);
  const double w = OQG_EARTH_RADIUS * (q0 * f[6] + q1 * f[7] + q2 * f[8]);
  const int R = shape[2 * s], C = shape[2 * s + 1];
  const double *X = pts + off[s], *Y = pts + P + off[s],
               *W = pts + 2 * P + off[s];
  double chord2 = INFINITY, seg2 = INFINITY;
  int inside = 0;
  for (int r = 0; r < R; r++)
    for (int c = 0; c < C; c++) {
      const int i = r * C + c;
      const double ax = X[i], ay = Y[i], aw = W[i];
      const double dx = x - ax, dy = y - ay, dw = w - aw;
      const double d2 = dx * dx + dy * dy;
      chord2 = fmin(chord2, d2 + dw * dw);
      seg2 = fmin(seg2, d2);
      const bool right = c + 1 < C, down = r + 1 < R;
      if (right) seg2 = fmin(seg2, seg_dist2(x, y, ax, ay, X[i + 1], Y[i + 1]));
      if (down) seg2 = fmin(seg2, seg_dist2(x, y, ax, ay, X[i + C], Y[i + C]));
      if (right && down) {
        const double bx = X[i + 1], by = Y[i + 1];
        const double cx = X[i + C + 1], cy = Y[i + C + 1];
        const double ex = X[i + C], ey = Y[i + C];
        inside |= (crosses(x, y, ax, ay, bx, by) + crosses(x, y, bx, by, cx, cy) +
                   crosses(x, y, cx, cy, ex, ey) + crosses(x, y, ex, ey, ax, ay)) & 1;
      }
    }
  const double chord = sqrt(chord2);
  const double poly = inside ? 0. : fmax(sqrt(seg2) - MESH_JB_TOL, 0.);
  out[(int64_t)s * N + n] = chord < MESH_JB_THRESHOLD ? poly : chord;
}

// ---- Rx and Ry0 of mesh surfaces ------------------------------------------
// BaseSurface.get_rx_distance (geo/surface/base.py:319-384) and
// get_ry0_distance (base.py:193-239) for S top edges at once.  The reference
// solves spherical triangles with azimuths and arc lengths (geodetic.py
// distance_to_arc :612-647, distance_to_semi_arc :556-609,
// min_distance_to_segment :182-257); here the same quantities are dot
// products with unit vectors: at a vertex p with east / north vectors e, n, a
// direction of azimuth a is t = sin a e + cos a n and the direction to its
// right is r = t x p, and for a site q at arc length d and azimuth z from p
//     sin(d) cos(z - a) = (q - p) . t      (in front of / behind p)
//     sin(d) sin(z - a) = (q - p) . r      (distance_to_arc = R asin of it)
//     geodetic distance = 2 R asin(|q - p| / 2), chord = R |q - p|
// so the branches of the reference are sign tests on two dot products and
// each distance is one asin or one sqrt.  One thread per site, one CTA row
// per surface; the per-segment vectors are built by the CTA in shared memory.
struct EdgeSeg {
  double pa[3], ta[3], ra[3], pb[3], tb[3];
};
#define RX_CHUNK 32

__device__ __forceinline__ double dot3(const double *a, const double *b) {
  return fma(a[0], b[0], fma(a[1], b[1], a[2] * b[2]));
}

// segment from vertex a to vertex b: tangent ta at a towards b, its right
// hand side ra, and the direction tb at b with the same azimuth as ta (the
// reference compares the azimuth of the segment, measured at its first
// point, with azimuths measured at its second point, geodetic.py:224-226)
__device__ inline void edge_seg_setup(double lon_a, double lat_a, double lon_b,
                                      double lat_b, EdgeSeg &g) {
  double ea[3], na[3], eb[3], nb[3], d[3];
  unit_from_lonlat(lon_a, lat_a, g.pa, ea, na);
  unit_from_lonlat(lon_b, lat_b, g.pb, eb, nb);
  for (int k = 0; k < 3; k++) d[k] = g.pb[k] - g.pa[k];
  const double along = dot3(d, g.pa);
  for (int k = 0; k < 3; k++) g.ta[k] = fma(-along, g.pa[k], d[k]);
  const double nrm = sqrt(dot3(g.ta, g.ta));
  // coincident vertices: azimuth(p, p) = 0 in the reference (atan2(0, 0))
  for (int k = 0; k < 3; k++) g.ta[k] = nrm > 0. ? g.ta[k] / nrm : na[k];
  const double sa = dot3(g.ta, ea), ca = dot3(g.ta, na);
  for (int k = 0; k < 3; k++) g.tb[k] = fma(sa, eb[k], ca * nb[k]);
  g.ra[0] = g.ta[1] * g.pa[2] - g.ta[2] * g.pa[1];
  g.ra[1] = g.ta[2] * g.pa[0] - g.ta[0] * g.pa[2];
  g.ra[2] = g.ta[0] * g.pa[1] - g.ta[1] * g.pa[0];
}

__device__ __forceinline__ double arc_of(double x) {
  return OQG_EARTH_RADIUS * asin(fmin(fmax(x, -1.), 1.));
}

__global__ void __launch_bounds__(128)
k_mesh_rx_ry0(const double *__restrict__ tlon, const double *__restrict__ tlat,
              const int64_t *__restrict__ off, const double *__restrict__ strike,
              const double *__restrict__ site, int N, double *__restrict__ rx,
              double *__restrict__ ry0, int y0) {
  __shared__ EdgeSeg s_seg[RX_CHUNK];
  __shared__ double s_end[2][6];   // first / last vertex: p, r of strike + 90
  const int s = blockIdx.y + y0;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t pa = off[s];
  const int K = (int)(off[s + 1] - pa);
  const bool ok = n < N;
  double q[3] = {0., 0., 1.};
  if (ok)
    for (int k = 0; k < 3; k++) q[k] = site[(int64_t)(SF_Q + k) * N + n];
  if (K < 2) {          // no segment: not a surface
    if (ok && rx) rx[(int64_t)s * N + n] = NAN;
    if (ok && ry0) ry0[(int64_t)s * N + n] = NAN;
    return;
  }
  if (ry0) {
    if (threadIdx.x < 2) {
      const int64_t v = threadIdx.x == 0 ? pa : pa + K - 1;
      double p[3], e[3], nn[3], sa, ca;
      unit_from_lonlat(tlon[v], tlat[v], p, e, nn);
      sincos((strike[s] + 90.) * 0.017453292519943295, &sa, &ca);
      for (int k = 0; k < 3; k++) {
        s_end[threadIdx.x][k] = p[k];
        s_end[threadIdx.x][3 + k] = fma(ca, e[k], -sa * nn[k]);
      }
    }
    __syncthreads();
    if (ok) {
      double d[2];
      for (int v = 0; v < 2; v++) {
        const double dx[3] = {q[0] - s_end[v][0], q[1] - s_end[v][1],
                              q[2] - s_end[v][2]};
        d[v] = arc_of(dot3(dx, &s_end[v][3]));
      }
      const int s0 = (d[0] > 0.) - (d[0] < 0.), s1 = (d[1] > 0.) - (d[1] < 0.);
      ry0[(int64_t)s * N + n] = s0 == s1 ? fmin(fabs(d[0]), fabs(d[1])) : 0.;
    }
  }
  if (!rx) return;
  double best = 0., babs = INFINITY;
  for (int i0 = 0; i0 < K - 1; i0 += RX_CHUNK) {
    const int ns = min(RX_CHUNK, K - 1 - i0);
    __syncthreads();
    if (threadIdx.x < ns) {
      const int i = i0 + threadIdx.x;
      // the first segment of a polyline runs backwards (base.py:360-361)
      const int64_t a = (K > 2 && i == 0) ? pa + 1 : pa + i;
      const int64_t b = (K > 2 && i == 0) ? pa : pa + i + 1;
      edge_seg_setup(tlon[a], tlat[a], tlon[b], tlat[b], s_seg[threadIdx.x]);
    }
    __syncthreads();
    if (!ok) continue;
    for (int j = 0; j < ns; j++) {
      const EdgeSeg &g = s_seg[j];
      const int i = i0 + j;
      const double dx[3] = {q[0] - g.pa[0], q[1] - g.pa[1], q[2] - g.pa[2]};
      const double c = dot3(dx, g.ta), x = dot3(dx, g.ra);
      double d;
      if (K == 2) {
        d = arc_of(x);                              // distance_to_arc
      } else if (i == 0 || i == K - 2) {            // distance_to_semi_arc
        if (c > 0.) {
          d = arc_of(x);
        } else {
          d = 2. * arc_of(0.5 * sqrt(dot3(dx, dx)));
          if (x < 0.) d = -d;
        }
        if (i == 0) d = -d;
      } else {                                      // min_distance_to_segment
        const double ex[3] = {q[0] - g.pb[0], q[1] - g.pb[1], q[2] - g.pb[2]};
        if (c >= 0. && dot3(ex, g.tb) <= 0.)
          d = fabs(arc_of(x));
        else
          d = OQG_EARTH_RADIUS * sqrt(fmin(dot3(dx, dx), dot3(ex, ex)));
        if (x < 0.) d = -d;
      }
      if (fabs(d) < babs) {     // numpy argmin: the first of equal minima
        babs = fabs(d);
        best = d;
      }
    }
  }
  if (ok) rx[(int64_t)s * N + n] = best;
}

// ---- mean dip, strike and width of rectangular meshes ---------------------
// RectangularMesh.get_mean_inclination_and_azimuth (geo/mesh.py:728-783 with
// calc_inclination :101-133, calc_azimuth :136-196) and get_mean_width
// (:861-885, get_cell_dimensions :785-827): the rupture parameters dip,
// strike (needed by Ry0) and width of a non-planar surface.  Every mesh cell
// is split in a top-left and a bottom-right triangle; inclination and
// azimuth are area-weighted circular means over the triangles, the width is
// the mean over the columns weighted by the mean cell length of the column.
// One CTA per mesh, one thread per column of cells walking down the rows.
__device__ __forceinline__ void mesh_point(const double *lon, const double *lat,
                                           const double *dep, int64_t i,
                                           double *p) {
  double sl, cl, sp, cp;
  sincos(lon[i] * 0.017453292519943295, &sl, &cl);
  sincos(lat[i] * 0.017453292519943295, &sp, &cp);
  const double rr = OQG_EARTH_RADIUS - (dep ? dep[i] : 0.);
  p[0] = rr * cp * cl; p[1] = rr * cp * sl; p[2] = rr * sp;
}

__device__ __forceinline__ void cross3(const double *a, const double *b,
                                       double *c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

__device__ __forceinline__ double heron(double a, double b, double c) {
  const double s = (a + b + c) / 2.0;
  return sqrt(s * (s - a) * (s - b) * (s - c));
}

// one triangle with azimuthal edge e1 and up-dip edge e2 at the corner p:
// adds its share to the inclination sums (ix, iy) and azimuth sums (ax, ay)
__device__ inline void mesh_triangle(const double *p, const double *e1,
                                     const double *e2, double area, double *acc) {
  double nrm[3], west[3], north[3], pz[3] = {p[0], p[1], p[2] + 1.0};
  cross3(e1, e2, nrm);
  const double ic = fmin(fmax(dot3(p, nrm) * rsqrt(dot3(p, p)) /
                                  sqrt(dot3(nrm, nrm)), -1.), 1.);
  acc[0] += area * ic;
  acc[1] += area * sqrt(fmax(1. - ic * ic, 0.));
  cross3(pz, p, west);
  cross3(p, west, north);
  const double l1 = sqrt(dot3(e1, e1));
  const double aw = dot3(e1, west) / (l1 * sqrt(dot3(west, west)));
  const double ac = dot3(e1, north) / (l1 * sqrt(dot3(north, north)));
  acc[2] += area * ac;
  const double y = area * sqrt(fmax(1. - ac * ac, 0.));
  acc[3] += aw < 0. ? -y : y;    // sign(sign(.) + 0.1): zero counts as +1
}

#define MESH_DIP_TPB 64
__global__ void __launch_bounds__(MESH_DIP_TPB)
k_mesh_dip_strike_width(const double *__restrict__ lon,
                        const double *__restrict__ lat,
                        const double *__restrict__ dep,
                        const int64_t *__restrict__ off,
                        const int32_t *__restrict__ shape,
                        double *__restrict__ out) {
  __shared__ double s_red[MESH_DIP_TPB / 32][8];
  const int s = blockIdx.x;
  const int R = shape[2 * s], C = shape[2 * s + 1];
  const int64_t base = off[s];
  // acc: ix iy ax ay, sum widths * mcl, sum mcl, any depth != 0
  double acc[7] = {0., 0., 0., 0., 0., 0., 0.};
  for (int c = threadIdx.x; c < C - 1; c += blockDim.x) {
    double width = 0., col_area = 0., len_area = 0.;
    double p00[3], p01[3], p10[3], p11[3];
    mesh_point(lon, lat, dep, base + c, p00);
    mesh_point(lon, lat, dep, base + c + 1, p01);
    if (dep && (dep[base + c] != 0. || dep[base + c + 1] != 0.)) acc[6] = 1.;
    for (int r = 0; r < R - 1; r++) {
      const int64_t i10 = base + (int64_t)(r + 1) * C + c;
      mesh_point(lon, lat, dep, i10, p10);
      mesh_point(lon, lat, dep, i10 + 1, p11);
      if (dep && (dep[i10] != 0. || dep[i10 + 1] != 0.)) acc[6] = 1.;
      double top[3], left[3], bottom[3], right[3], diag[3];
      for (int k = 0; k < 3; k++) {
        top[k] = p01[k] - p00[k];        // along_azimuth[r][c]
        left[k] = p00[k] - p10[k];       // updip[r][c]
        bottom[k] = p11[k] - p10[k];     // along_azimuth[r + 1][c]
        right[k] = p01[k] - p11[k];      // updip[r][c + 1]
        diag[k] = p01[k] - p10[k];
      }
      const double lt = sqrt(dot3(top, top)), ll = sqrt(dot3(left, left)),
                   lb = sqrt(dot3(bottom, bottom)),
                   lr = sqrt(dot3(right, right)), ld = sqrt(dot3(diag, diag));
      const double tl_area = heron(lt, ll, ld), br_area = heron(lb, lr, ld);
      mesh_triangle(p00, top, left, tl_area, acc);
      mesh_triangle(p11, bottom, right, br_area, acc);
      const double area = tl_area + br_area;
      width += (ll * tl_area + lr * br_area) / area;
      len_area += lt * tl_area + lb * br_area;   // cell_length * cell_area
      col_area += area;
      for (int k = 0; k < 3; k++) {
        p00[k] = p10[k];
        p01[k] = p11[k];
      }
    }
    const double mcl = len_area / col_area;
    acc[4] += width * mcl;
    acc[5] += mcl;
  }
  for (int k = 0; k < 7; k++) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < 7; k++) {
      acc[k] = 0.;
      for (int w = 0; w < MESH_DIP_TPB / 32; w++) acc[k] += s_red[w][k];
    }
    if (R < 2 || C < 2) {     // an assertion in the reference
      out[3 * s] = out[3 * s + 1] = out[3 * s + 2] = NAN;
      return;
    }
    double incl = acc[6] != 0. ? atan2(acc[1], acc[0]) * OQG_R2D : 0.;
    double azim = atan2(acc[3], acc[2]) * OQG_R2D;
    if (azim < 0.) azim += 360.;
    if (incl > 90.) {
      incl = 180. - incl;
      azim = fmod(azim + 180., 360.);
    }
    out[3 * s] = incl;
    out[3 * s + 1] = azim;
    out[3 * s + 2] = acc[4] / acc[5];
  }
}


// ---- GC2 coordinates of multi-fault ruptures --------------------------------
// Line.get_tuw (geo/line.py:178-266: get_ui_ti, get_ti_weights, line_get_tu)
// for S2 oriented lines (every section's top-of-rupture line and its flip)
// and N sites: out [S2][N][3] float32 = t, u and the sum of the segment
// weights, as get_dparam(sec, sites, 'tuw') stores them (calc/filters.py:
// 54-59).  seg [T][7] = projected start (x, y) of each segment, its length,
// uhat (2), that (2); loff [S2 + 1] the segments of each line; proj [S2][2] =
// lam0, phi0 of the line's orthographic projection.  One thread per (line,
// site); the arithmetic is float64 like the reference's, the result float32.
#define GC2_TOL 0.1    // geo/line.py:26
__global__ void __launch_bounds__(128)
k_line_tuw(const double *__restrict__ seg, const int64_t *__restrict__ loff,
           const double *__restrict__ proj, int S2,
           const double *__restrict__ slon, const double *__restrict__ slat,
           int N, float *__restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = blockIdx.y;
  if (n >= N || s >= S2) return;
  const double lam0 = proj[2 * s], phi0 = proj[2 * s + 1];
  // project_direct (geo/utils.py:599-607)
  const double lam = slon[n] * 0.017453292519943295 - lam0;
  const double phi = slat[n] * 0.017453292519943295;
  double sl, cl, sp, cp, sp0, cp0;
  sincos(lam, &sl, &cl);
  sincos(phi, &sp, &cp);
  sincos(phi0, &sp0, &cp0);
  const double sx = cp * sl * OQG_EARTH_RADIUS;
  const double sy = (cp0 * sp - sp0 * cp * cl) * OQG_EARTH_RADIUS;
  double wsum = 0., tacc = 0., uacc = 0., delta = 0.;
  bool on_trace = false;
  // two passes as the reference: the weights first (t = sum(ti w / wsum))
  for (int pass = 0; pass < 2; pass++) {
    delta = 0.;
    for (int64_t k = loff[s]; k < loff[s + 1]; k++) {
      const double *g = seg + k * 7;
      const double dx = sx - g[0], dy = sy - g[1], len = g[2];
      const double ui = dx * g[3] + dy * g[4];
      const double ti = dx * g[5] + dy * g[6];
      double w;
      if (fabs(ti) >= GC2_TOL) {
        w = (atan((len - ui) / ti) - atan(-ui / ti)) / ti;
      } else if (ui < 0. - GC2_TOL || ui > len + GC2_TOL) {
        w = 1. / (ui - len) - 1. / ui;       // on the extension of the segment
      } else {
        w = 1 / (-0.01 - len) + 1 / 0.01;    // on the segment
        on_trace = true;
      }
      if (pass == 0) {
        wsum += w;
      } else {
        tacc += ti * w / wsum;
        uacc += (ui + delta) * w;
      }
      delta += len;
    }
  }
  float *o = out + ((int64_t)s * N + n) * 3;
  o[0] = on_trace ? 0.f : (float)tacc;
  o[1] = (float)(uacc / wsum);
  o[2] = (float)wsum;
}

// genctxs with the dparam cache (contexts.py:959-1029, _get_tu :85-98,
// set_distances :113-135, multiline.get_tu geo/multiline.py:246-261): the
// distances of U ruptures made of sections from the section planes.
//   rrup, rjb [U][N] = minimum over the sections of the rupture
//   rx = T, ry0 from U and u_max, T and U the weighted means of the
//   sections' (t, u + shift) in the order of the rupture's MultiLine, taken
//   from the orientation (flip) the MultiLine gives each section; the
//   accumulation is float32 like the reference's
// roff [U + 1], rsec / rflip / rshift per (rupture, section) in MultiLine
// order; tuw [S][2][N][3] (orientation 0 = as given, 1 = flipped).
__global__ void __launch_bounds__(128)
k_multi_compose(const double *__restrict__ sec_rrup,
                const double *__restrict__ sec_rjb,
                const float *__restrict__ tuw, int N,
                const int64_t *__restrict__ roff,
                const int32_t *__restrict__ rsec,
                const uint8_t *__restrict__ rflip,
                const float *__restrict__ rshift,
                const float *__restrict__ umax, int U,
                double *__restrict__ rrup, double *__restrict__ rjb,
                double *__restrict__ rx, double *__restrict__ ry0, int y0) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int u = blockIdx.y + y0;
  if (n >= N || u >= U) return;
  double best = INFINITY, bestj = INFINITY;
  float ts = 0.f, us = 0.f, ws = 0.f;
  for (int64_t k = roff[u]; k < roff[u + 1]; k++) {
    const int64_t s = rsec[k];
    if (sec_rrup) best = fmin(best, sec_rrup[s * N + n]);
    if (sec_rjb) bestj = fmin(bestj, sec_rjb[s * N + n]);
    if (tuw) {
      const float *t = tuw + ((s * 2 + (rflip[k] ? 1 : 0)) * N + n) * 3;
      const float w = t[2];
      ts += t[0] * w;
      us += (t[1] + rshift[k]) * w;
      ws += w;
    }
  }
  const int64_t at = (int64_t)u * N + n;
  if (rrup) rrup[at] = best;
  if (rjb) rjb[at] = bestj;
  if (tuw) {
    const float tt = ts / ws, uu = us / ws;
    if (rx) rx[at] = (double)tt;
    if (ry0) {
      const float um = umax[u];
      ry0[at] = uu < 0.f ? (double)fabsf(uu) : (uu > um ? (double)(uu - um) : 0.);
    }
  }
}

}  // namespace oqg
