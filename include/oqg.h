/*
 * oqg.h -- C ABI of the B200-native classical-PSHA hot path.
 *
 * The reference (gem/oq-engine, openquake/hazardlib) has no FFI layer: its
 * boundary for this path is a Python operator surface.
 * Each entry point below names the reference operator it replaces
 * (paths relative to the reference checkout):
 *
 *   oqg_distances      planar.project / get_rjb / get_rx / get_ry0 / get_repi /
 *                      get_rhypo / get_azimuth
 *                      (hazardlib/geo/surface/planar.py:324-403, 460-621),
 *                      the distance part of ContextMaker.get_ctxs
 *                      (hazardlib/contexts.py:374-440, 1051-1099)
 *   oqg_mean_stds      ContextMaker.get_mean_stds / get_4MN -> GMPE.compute
 *                      (hazardlib/contexts.py:1250-1303, gsim/base.py:389-401)
 *   oqg_poes           _set_poes / set_poes / truncnorm_sf / ndtr
 *                      (hazardlib/contexts.py:482-489, 1415-1475,
 *                      hazardlib/stats.py:27-73)
 *   oqg_update         MapArray.update_indep -> update_pmap_i / update_pmap_r
 *                      with tom.get_pnes (hazardlib/map_array.py:238-260,
 *                      443-455, hazardlib/tom.py:185-217)
 *   oqg_hazard_planar  the whole chain ContextMaker.get_ctxs -> update for
 *                      planar ruptures, fused, with no context rows
 *                      materialised (hazardlib/contexts.py:1232-1247 driven by
 *                      RmapMaker._make_src_indep :1548-1579)
 *   oqg_hazard_planar_host  the same call taking HOST buffers (copies inside)
 *
 * Conventions: every function returns 0 on success, non-zero on error with a
 * message available from oqg_last_error(); all `dev` pointers are device
 * pointers on the handle's device (any allocator: cudaMalloc, a torch CUDA
 * tensor's data_ptr()); all work is queued on the handle's stream and the
 * functions do NOT synchronise unless stated; a handle is not thread-safe.
 * There is no CPU fallback: without a CUDA device oqg_create fails.
 */
#ifndef OQG_H
#define OQG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oqg_handle oqg_handle;

/* GMPE families with a device implementation (reference module) */
enum {
  OQG_BSSA14 = 0,  /* gsim/boore_2014.py BooreEtAl2014 (+HighQ/LowQ tables) */
  OQG_CY14 = 1,    /* gsim/chiou_youngs_2014.py ChiouYoungs2014 */
  OQG_ASK14 = 2,   /* gsim/abrahamson_2014.py AbrahamsonEtAl2014 */
  OQG_CB14 = 3,    /* gsim/campbell_bozorgnia_2014.py CampbellBozorgnia2014 */
  OQG_BA08 = 4,    /* gsim/boore_atkinson_2008.py BooreAtkinson2008 */
  OQG_AB06 = 5,    /* gsim/atkinson_boore_2006.py AtkinsonBoore2006 */
  OQG_AKKAR14 = 6, /* gsim/akkar_2014.py AkkarEtAl{Rjb,Repi,Rhyp}2014 */
  OQG_CY08 = 7,    /* gsim/chiou_youngs_2008.py ChiouYoungs2008 */
  OQG_TORO02 = 8,  /* gsim/toro_2002.py ToroEtAl2002 */
  OQG_C03 = 9,     /* gsim/campbell_2003.py Campbell2003 */
  OQG_SADIGH97 = 10, /* gsim/sadigh_1997.py SadighEtAl1997 */
  OQG_NFAMILIES = 11
};

/* distance metrics: bit positions for masks and order of output planes */
enum {
  OQG_RRUP = 0, OQG_RJB = 1, OQG_RX = 2, OQG_RY0 = 3, OQG_REPI = 4,
  OQG_RHYPO = 5, OQG_AZIMUTH = 6, OQG_NDIST = 7,
  /* outputs of oqg_distances only: longitude and latitude (degrees) of the
   * point of the rupture closest to the site, project_back
   * (geo/surface/planar.py:407-424), the clon / clat columns of the context
   * rows when there are few sites (contexts.py:385-386, 425-427) */
  OQG_CLON = 7, OQG_CLAT = 8, OQG_NOUT = 9
};

#define OQG_KMAX 56       /* doubles per coefficient row */
#define OQG_PLANAR_LEN 42 /* doubles per planar rupture: the memory of one
                             (3,) record array of planar_array_dt
                             (planar.py:45-53), i.e. [k][14] with k = lon/lat/
                             dep (or x/y/z) and 14 = corners[4] xyz[4] normal
                             uv1 uv2 wlr sdr hypo; a reference planar array
                             .view(float) uploads as it is */
#define OQG_MAX_OCC 16    /* max length of probs_occur */

/* per-model configuration; `flags` and `par` meaning per family:
 *  BSSA14 : flags bit0-1 region (0 nobasin, 1 CAL, 2 JPN), bit4 sof;
 *           par[0] sigma_mu_epsilon
 *  CY14   : flags bit0-1 region (0 CAL, 1 JPN, 2 ITA, 3 WEN), bit4 use_hw,
 *           bit5 peer; par[0] sigma_mu_epsilon, par[1] alpha_nm
 *  ASK14  : flags bit0-1 region (0 none/CAL, 1 JPN, 2 TWN, 3 CHN);
 *           par[0] sigma_mu_epsilon
 *  CB14   : flags bit0 SJ; par[0] sigma_mu_epsilon
 *  BA08   : -
 *  AB06   : flags bit0-1 mag_eq (0 NA, 1 Mblg87, 2 Mblg96); par[0] scale_fac
 *  AKKAR14: flags bit0-1 distance (0 rjb, 1 repi, 2 rhypo);
 *           par[0] log(adjustment_factor)
 *  CY08: - ; TORO02, C03: flags bit0 the SHARE variant (style of faulting and
 *           rock adjustment, its factors in the coefficient row)
 *  SADIGH97: - (the model of the PEER verification tests)
 */
typedef struct {
  int32_t family;
  int32_t flags;
  double par[4];
} oqg_gsim;

/* context rows as a struct of device arrays of length P (the columns of the
 * reference's ctx recarray, contexts.py:660-678); pointers of columns that
 * no configured model reads may be NULL */
typedef struct {
  const double *mag, *rake, *dip, *ztor, *width, *hypo_depth;
  const double *rrup, *rjb, *rx, *ry0, *repi, *rhypo;
  const double *vs30, *z1pt0, *z2pt5;
  const uint8_t *vs30measured;
} oqg_ctx;

/* site collection as a struct of arrays of length N (site.py:249-324) */
typedef struct {
  const double *lon, *lat, *depth;       /* degrees, degrees, km */
  const double *vs30, *z1pt0, *z2pt5;    /* may be NULL if unused */
  const uint8_t *vs30measured;           /* may be NULL if unused */
} oqg_sites;

int oqg_create(oqg_handle **out, int device);
void oqg_destroy(oqg_handle *h);
/* message of the last error (h may be NULL for errors of oqg_create) */
const char *oqg_last_error(oqg_handle *h);
/* queue the work on an existing cudaStream_t (e.g. torch's current stream;
 * NULL is the legacy default stream) / back on the handle's own stream */
int oqg_set_stream(oqg_handle *h, void *cuda_stream);
int oqg_use_own_stream(oqg_handle *h);
int oqg_synchronize(oqg_handle *h);

/* comma separated names of the coefficient row of a family, in the order the
 * device expects them; names of the form table.column; NULL if unknown */
const char *oqg_family_fields(int family);
/* bit mask (1 << OQG_RRUP ...) of the distances / a string of context
 * columns a configured model reads */
int oqg_required_distances(const oqg_gsim *g);

/* models: coeffs[G][M+1][OQG_KMAX] (host), row M = the PGA row of the model.
 * An unknown family or flag is an error (there is no fallback). */
int oqg_set_gsims(oqg_handle *h, int G, const oqg_gsim *gsims, int M,
                  const double *coeffs);
/* ContextMaker.horiz_comp_to_geom_mean (hazardlib/contexts.py:761-776, the
 * conversion tables of set_imts_conv :726-746): conv [G][M][3] = conv_median,
 * conv_sigma, rstd per (gsim, IMT), all zero where there is no conversion;
 * oqg_mean_stds then returns mean - log(conv_median) and
 * sqrt((sig^2 - conv_sigma^2) / rstd^2) (tau and phi unchanged, the zero
 * sigma check comes first, as in get_4MN :1297-1302).  NULL switches it off;
 * oqg_set_gsims resets it.  The fused pass refuses to run while it is set. */
int oqg_set_conversion(oqg_handle *h, const double *conv);

/* mask[G][M] (host, may be NULL = none): 1 where the PoEs of (gsim, IMT) are
 * zeroed -- the ImtWeight == 0 branch of set_poes (contexts.py:1469-1475);
 * reset by oqg_set_gsims */
int oqg_set_imt_mask(oqg_handle *h, const unsigned char *mask);
/* levels: loglevels[M][L1] (host), as ContextMaker.loglevels
 * (contexts.py:718-724); phi_b = ndtr(truncation_level) (contexts.py:628) */
int oqg_set_levels(oqg_handle *h, int M, int L1, const double *loglevels,
                   double truncation_level);

/* distances of U planar ruptures x N sites: out[k][U][N] (dev) for the k-th
 * set bit of `mask` in the order of the OQG_R* enum; planar is dev [U][42] */
int oqg_distances(oqg_handle *h, const double *planar_dev, int U,
                  const oqg_sites *sites_dev, int N, int mask,
                  double *out_dev);

/* out[4][G][M][P] (dev) = mean, sig, tau, phi */
int oqg_mean_stds(oqg_handle *h, const oqg_ctx *ctx_dev, int64_t P,
                  double *out_dev);

/* verification entry point: out[2][G][M][P] (dev) = mean, sig as evaluated
 * by the staged model code of the fused kernel (per-site / per-rupture /
 * per-pair / per-IMT stages) on the same context rows */
int oqg_mean_stds_staged(oqg_handle *h, const oqg_ctx *ctx_dev, int64_t P,
                         double *out_dev);

/* poes[P][L][G] float32 (dev) from mean_std[4][G][M][P] (dev), L = M*L1 */
int oqg_poes(oqg_handle *h, const double *mean_std_dev, int64_t P,
             float *poes_dev);

/* fold rows into the map: map[N][L][G] float32 (dev), poes[P][L][G] float32
 * (dev); the rows of site row s are order[seg[s] .. seg[s+1]) (dev; a stable
 * grouping of the context rows by site row, seg has N+1 entries);
 * rate[P]; probs_occur[P][K] or NULL with K = 0 (Poisson);
 * rates_mode 0: map *= pne, 1: map += -log(pne) (update_pmap_i / _r) */
int oqg_update(oqg_handle *h, float *map_dev, int N, const float *poes_dev,
               int64_t P, const int64_t *order_dev, const int64_t *seg_dev,
               const double *rate_dev, const double *probs_occur_dev, int K,
               double itime, int rates_mode);
/* mutually exclusive ruptures: map += (1 - pne) * weight[row]
 * (MapArray.update_mutex -> update_pmap_m, hazardlib/map_array.py:262-269,
 * 455-466); same row grouping as oqg_update; weight_dev[P] */
int oqg_update_mutex(oqg_handle *h, float *map_dev, int N,
                     const float *poes_dev, int64_t P,
                     const int64_t *order_dev, const int64_t *seg_dev,
                     const double *rate_dev, const double *probs_occur_dev,
                     int K, double itime, const double *weight_dev);


/* the fused path.  planar dev [U][42] (wlr[2] is the occurrence rate),
 * mag[U], magdist[U] (rows with rrup >= magdist are dropped,
 * contexts.py:472), probs_occur [U][K] or NULL, sites dev arrays, sidx[N] the
 * map row of each site (NULL: identity); map dev [Nmap][L][G] float32 is
 * updated in place: rates_mode 0: *= prod pne, 1: += sum rates.
 * counters (dev, may be NULL): [0] += number of rupture-site pairs kept */
int oqg_hazard_planar(oqg_handle *h, const double *planar_dev,
                      const double *mag_dev, const double *magdist_dev,
                      const double *probs_occur_dev, int K, int U,
                      const oqg_sites *sites_dev, const uint32_t *sidx_dev,
                      int N, double minimum_distance, double itime,
                      int rates_mode, float *map_dev,
                      unsigned long long *counters_dev);

/* the name SURVEY.md section 8b gives to the fused entry point: the same
 * call for Poissonian ruptures (no probs_occur, identity site index) */
int oqg_fused_planar(oqg_handle *h, const double *planar_dev,
                     const double *mag_dev, int R, const oqg_sites *sites_dev,
                     int N, const double *magdist_dev, double minimum_distance,
                     double itime, int rates_mode, float *map_dev,
                     unsigned long long *counters_dev);

/* oqg_hazard_planar with the per-rupture visibility of the
 * pointsource_distance approximation: aux_dev [U][OQG_AUX_LEN] from
 * oqg_point_ruptures, or NULL (every site sees every rupture) */
int oqg_hazard_planar_aux(oqg_handle *h, const double *planar_dev,
                          const double *mag_dev, const double *magdist_dev,
                          const double *aux_dev, const double *probs_occur_dev,
                          int K, int U, const oqg_sites *sites_dev,
                          const uint32_t *sidx_dev, int N,
                          double minimum_distance, double itime,
                          int rates_mode, float *map_dev,
                          unsigned long long *counters_dev);

/* The same fused pass (GMPE -> PoEs -> map) for ruptures that are NOT planar
 * (the ruptures of fault sources: ContextMaker.genctxs, contexts.py:959-1029,
 * get_distances, calc/filters.py:75-116): their distances come from the mesh
 * kernels (oqg_mesh_min_distance, oqg_mesh_rjb, oqg_mesh_rx_ry0) as planes
 * [U][N] float64 in HBM.  rup_params_dev [U][8] = mag, rake, dip, ztor,
 * width, hypo_depth, occurrence rate, integration distance of each rupture
 * (get_rparams, contexts.py:905-957); planes (HOST array of 6 device
 * pointers, order OQG_RRUP .. OQG_RHYPO; rrup is required, the others may be
 * NULL when no configured model reads them).  A pair is kept when rrup <=
 * the integration distance (contexts.py:974; planar ruptures: <).  The other
 * arguments are those of oqg_hazard_planar. */
int oqg_hazard_dist(oqg_handle *h, const double *rup_params_dev,
                    const double *const *planes, const double *probs_occur_dev,
                    int K, int U, const oqg_sites *sites_dev,
                    const uint32_t *sidx_dev, int N, double minimum_distance,
                    double itime, int rates_mode, float *map_dev,
                    unsigned long long *counters_dev);

/* same with HOST buffers: copies the ruptures and sites in, the map in and
 * out, and synchronises; map is host [N][L][G] float32 (sidx = identity) */
int oqg_hazard_planar_host(oqg_handle *h, const double *planar,
                           const double *mag, const double *magdist,
                           const double *probs_occur, int K, int U,
                           const oqg_sites *sites, int N,
                           double minimum_distance, double itime,
                           int rates_mode, float *map,
                           unsigned long long *pairs_kept);

/* A stream of sources from host memory into one map (the reference's loop
 * `for src in sources: cmaker.update(pmap, ctxs)`, hazardlib/contexts.py
 * :1548-1579): oqg_host_begin uploads the sites and the map once (map NULL:
 * the identity, 1 for a map of probabilities of no exceedance, 0 for rates),
 * oqg_host_add folds the ruptures of one source -- the arguments of
 * oqg_hazard_planar_host -- with its upload overlapping the kernels of the
 * previous source (two device slots, a copy stream); it returns when the
 * caller's buffers can be reused, not when the kernels are done.
 * oqg_host_end downloads the map (and the number of pairs kept, may be NULL)
 * and synchronises.  One session per handle at a time. */
int oqg_host_begin(oqg_handle *h, const oqg_sites *sites, int N, int rates_mode,
                   const float *map);
int oqg_host_add(oqg_handle *h, const double *planar, const double *mag,
                 const double *magdist, const double *probs, int K, int U,
                 double mindist, double itime);
int oqg_host_end(oqg_handle *h, float *map, unsigned long long *pairs_kept);

/* ---- rupture generation of point sources on the device --
 * replaces PointSource.get_planar / get_planin (source/point.py:192-289) and
 * build_planar / _build_corners (geo/surface/planar.py:107-251) for a batch
 * of S point sources given as a struct of DEVICE arrays.  The magnitude
 * frequency distribution, nodal plane distribution and hypocentral depth
 * distribution of source s are the slices [beg, beg + cnt) of the shared
 * lists below (area sources: all the points share one slice).  The ruptures
 * of a source are ordered (magnitude, nodal plane, depth) as in the
 * reference; rup_off[s] is the index of the first rupture of source s,
 * rup_off[S] = U = sum of mag_cnt * np_cnt * hd_cnt. */
enum { OQG_MSR_WC1994 = 0,   /* scalerel/wc1994.py get_median_area */
       OQG_MSR_POINT = 1,    /* scalerel/point.py PointMSR */
       OQG_MSR_PEER = 2,     /* scalerel/peer.py PeerMSR: 10 ** (mag - 4) */
       OQG_MSR_CSCALING = 3 }; /* scalerel/cscaling.py CScalingMSR:
                                  10 ** (mag - msr_par) */
typedef struct {
  const double *lon, *lat;            /* [S] location, degrees */
  const double *usd, *lsd, *rar;      /* [S] upper / lower seismogenic depth
                                         (km), rupture aspect ratio */
  const int32_t *mag_beg, *mag_cnt;   /* [S] */
  const int32_t *np_beg, *np_cnt;     /* [S] */
  const int32_t *hd_beg, *hd_cnt;     /* [S] */
  const int64_t *rup_off;             /* [S + 1] */
  const double *mag, *mag_rate;       /* magnitudes and annual rates
                                         (get_annual_occurrence_rates) */
  const double *mag_dist;             /* integration distance of each
                                         magnitude, km; may be NULL */
  const double *np_strike, *np_dip, *np_rake, *np_prob;
  const double *hd_depth, *hd_prob;
  const double *hd_frac;              /* hypo_dip_fracs; NULL = 0.5 */
  int32_t msr;                        /* OQG_MSR_* */
  int32_t shift_hypo;                 /* ContextMaker.shift_hypo */
  /* the pointsource_distance approximation of _quintets
   * (hazardlib/contexts.py:321-370; engine default 100 km): 0 = off.  When
   * positive, a source with np_cnt * hd_cnt > 1 has, per magnitude,
   * np_cnt * hd_cnt + 1 ruptures: all the planes and depths for the sites
   * within psdist(mag) = min(radius(mag), mag_dist / 2) +
   * pointsource_distance of the source location (chord distance,
   * SiteCollection.get_cdist), and the average rupture of
   * PointSource.iruptures (source/point.py:311-318) for the other sites;
   * rup_off must count them. */
  double pointsource_distance;
  /* optional [U]: overrides the visibility of each rupture (0 all sites,
   * 1 near, 2 far, 3 none) -- the all-or-nothing rule of the reference for
   * few sites (contexts.py:358-362) is decided by the caller */
  const uint8_t *mode;
  double msr_par;                     /* C of OQG_MSR_CSCALING */
} oqg_point_sources;

#define OQG_AUX_LEN 5 /* per rupture: x, y, z of the source location (km),
                         psdist (km), mode */

/* planar_dev[U][42] in the planar_array_dt record layout, except that wlr[2]
 * holds the occurrence rate mag_rate * (np_prob * hd_prob) (the convention
 * of oqg_hazard_planar; the reference keeps the weight there and multiplies
 * by the magnitude rate in _get_ctx_planar, contexts.py:416);
 * mag_dev[U]; magdist_dev[U] (NULL if mag_dist is NULL); aux_dev
 * [U][OQG_AUX_LEN] (may be NULL when pointsource_distance is 0 and mode is
 * NULL), to be passed to oqg_hazard_planar_aux.  An unknown msr is an
 * error. */
int oqg_point_ruptures(oqg_handle *h, const oqg_point_sources *src, int S,
                       int64_t U, double *planar_dev, double *mag_dev,
                       double *magdist_dev, double *aux_dev);

/* ---- non-planar surfaces, first step: rupture distance of
 * mesh surfaces.  oqg_mesh_min_distance is Mesh.get_min_distance
 * (hazardlib/geo/mesh.py:365-383, cdist(...).min(0)) for S meshes at once:
 * mesh points lon/lat/depth_dev[P] (depth may be NULL = 0), the points of
 * mesh s are off_dev[s] .. off_dev[s + 1]; out_dev[S][N] km.
 * oqg_section_min is the multi-fault composition of contexts.py:969-971:
 * out_dev[U][N] = min over the sections rup_sec_dev[rup_off_dev[u] ..
 * rup_off_dev[u + 1]) of sec_dev[section][N]. */
int oqg_mesh_min_distance(oqg_handle *h, const double *lon_dev,
                          const double *lat_dev, const double *depth_dev,
                          int64_t P, const int64_t *off_dev, int S,
                          const oqg_sites *sites_dev, int N, double *out_dev);
/* Rx and Ry0 of S mesh surfaces at once: BaseSurface.get_rx_distance
 * (hazardlib/geo/surface/base.py:319-384) and get_ry0_distance
 * (base.py:193-239).  top_lon/top_lat_dev: the top edges (first mesh row) in
 * degrees, the vertices of surface s are off_dev[s] .. off_dev[s + 1] (at
 * least two; a shorter edge gives NaN); strike_dev[S]: the mean strike of
 * each surface in degrees (surface.get_strike(); only read for ry0).
 * rx_dev[S][N], ry0_dev[S][N] km; either may be NULL. */
int oqg_mesh_rx_ry0(oqg_handle *h, const double *top_lon_dev,
                    const double *top_lat_dev, const int64_t *off_dev,
                    const double *strike_dev, int S,
                    const oqg_sites *sites_dev, int N, double *rx_dev,
                    double *ry0_dev);
/* RectangularMesh.get_joyner_boore_distance (hazardlib/geo/mesh.py:453-540,
 * _get_proj_enclosing_polygon :541-597) for S rectangular meshes at once:
 * lon_dev / lat_dev [P] the points of all the meshes (row-major), off_dev
 * [S + 1] their offsets, shape_dev [S][2] rows and columns, owner_dev [P] the
 * mesh of every point, frame_dev [S][9] the east, north and radial unit
 * vectors at the centre of the orthographic projection of each mesh (the
 * middle point of its spherical bounding box, geo/utils.py:652-658) ->
 * out_dev [S][N] km: the minimum chord distance of the surface projections
 * beyond 40 km, the distance to the union of the mesh cells (zero inside)
 * below.  Agrees with the reference to the 0.02 km of the reference's own
 * tests (shapely grows, simplifies and rounds the polygon by up to 10 m). */
int oqg_mesh_rjb(oqg_handle *h, const double *lon_dev, const double *lat_dev,
                 int64_t P, const int64_t *off_dev, const int32_t *shape_dev,
                 const int32_t *owner_dev, const double *frame_dev, int S,
                 const oqg_sites *sites, int N, double *out_dev);

/* Multi-fault ruptures: the `dparam` cache of ContextMaker
 * (hazardlib/contexts.py:1039-1049, calc/filters.py:51-71).
 *
 * oqg_line_tuw: Line.get_tuw (hazardlib/geo/line.py:178-266), the GC2
 * coordinates of N sites for S2 oriented top-of-rupture lines: seg_dev
 * [T][7] = projected start (x, y), length, uhat (2), that (2) of every
 * segment (Line.sut_hat :543-565), loff_dev [S2 + 1] the segments of each
 * line, proj_dev [S2][2] = lam0, phi0 of its orthographic projection ->
 * out_dev [S2][N][3] float32 = t, u, sum of the weights.
 *
 * oqg_multi_compose: genctxs with the cache (contexts.py:959-1029, _get_tu
 * :85-98, set_distances :113-135, multiline.get_tu geo/multiline.py:246-261):
 * from the section planes sec_rrup_dev / sec_rjb_dev [S][N] and tuw_dev
 * [S][2][N][3] (orientation 0 as given, 1 flipped) the planes [U][N] of U
 * ruptures: rrup / rjb = minimum over the sections, rx = T, ry0 = |U| below
 * 0, U - u_max above u_max, else 0.  roff_dev [U + 1]; rsec_dev, rflip_dev,
 * rshift_dev per (rupture, section) in the order of the rupture's MultiLine;
 * umax_dev [U].  Output pointers may be NULL. */
int oqg_line_tuw(oqg_handle *h, const double *seg_dev, const int64_t *loff_dev,
                 const double *proj_dev, int S2, const oqg_sites *sites_dev,
                 int N, float *out_dev);
int oqg_multi_compose(oqg_handle *h, const double *sec_rrup_dev,
                      const double *sec_rjb_dev, const float *tuw_dev, int S,
                      int N, const int64_t *roff_dev, const int32_t *rsec_dev,
                      const uint8_t *rflip_dev, const float *rshift_dev,
                      const float *umax_dev, int U, double *rrup_dev,
                      double *rjb_dev, double *rx_dev, double *ry0_dev);

/* Mean dip, strike and width of S rectangular meshes:
 * RectangularMesh.get_mean_inclination_and_azimuth (hazardlib/geo/mesh.py:
 * 728-783) and get_mean_width (mesh.py:861-885), the surface.get_dip() /
 * get_strike() / get_width() of fault surfaces.  Mesh s has shape_dev[2 s]
 * rows x shape_dev[2 s + 1] columns (row-major from off_dev[s]) in
 * lon/lat/depth_dev (depth may be NULL = 0, which makes the dip 0 as in the
 * reference); out_dev[S][3] = inclination, azimuth (degrees), width (km);
 * NaN for a mesh with less than two rows or columns. */
int oqg_mesh_dip_strike_width(oqg_handle *h, const double *lon_dev,
                              const double *lat_dev, const double *depth_dev,
                              const int64_t *off_dev, const int32_t *shape_dev,
                              int S, double *out_dev);
int oqg_section_min(oqg_handle *h, const double *sec_dev, int S, int N,
                    const int64_t *rup_off_dev, const int32_t *rup_sec_dev,
                    int U, double *out_dev);

/* ---- disaggregation: _disaggregate
 * (hazardlib/calc/disagg.py:188-243, Poissonian lane) with _disagg_eps
 * (:246-256), get_eps4 (:176-186) and _build_disagg_matrix (:262-296) for the
 * U context rows of one site and one magnitude bin and one gsim.
 *   rrup/clon/clat/rate_dev [U]; mea_dev, std_dev [M][U]
 *   iml2 [M][P] log intensities (host; -inf = zero hazard, skipped)
 *   dist/lon/lat edges (host, D + 1, Lo + 1, La + 1 values), eps_edges (host,
 *   E + 1 <= 33 values); epsstar as in the reference; gp = group probability
 *   out_dev [D][Lo][La][E][M][P] float64 = 1 - prod pne */
int oqg_disaggregate(oqg_handle *h, const double *rrup_dev,
                     const double *clon_dev, const double *clat_dev,
                     const double *rate_dev, int64_t U, const double *mea_dev,
                     const double *std_dev, int M, int P, const double *iml2,
                     const double *dist_edges, int D, const double *lon_edges,
                     int Lo, const double *lat_edges, int La,
                     const double *eps_edges, int E, double truncation_level,
                     double itime, int epsstar, double gp, double *out_dev);

/* conditional spectrum: _cs_out (hazardlib/calc/cond_spectra.py:35-85) for the
 * U context rows of one site and one gsim.  mu_dev, sig_dev [M][U] (mean and
 * total standard deviation of log ground motion); probs_dev [U] probability
 * of at least one occurrence; host arrays rho[M] (correlation with the
 * conditioning IMT of index imti), imls[P], cs_poes[P], prev[M][P] (the
 * previously computed c[m, 1, p], or NULL); out (HOST) [M][3][P] =
 * sum of weights, weighted mean term, weighted variance term (0 without
 * prev).  Synchronises. */
int oqg_cs_out(oqg_handle *h, const double *mu_dev, const double *sig_dev,
               const double *probs_dev, int64_t U, int M, const double *rho,
               int imti, const double *imls, const double *cs_poes, int P,
               double truncation_level, double invtime, const double *prev,
               double *out);

/* ground motion fields: GmfComputer._compute (hazardlib/calc/gmf.py:476-525)
 * given the epsilons, float32 as in the reference.  mean_stds_dev [4][M][N]
 * (mean, sig, tau, phi), within_eps_dev [M][N][E], between_eps_dev [E][M]
 * (may be NULL for modes 0 and 1), not_mmi (host) [M] (0 for MMI: not
 * exponentiated); mode 0: exp(mean); 1: exp(mean + sig * within);
 * 2: exp(mean + phi * within + tau * between); 3: exp(mean + within + tau *
 * between) with `within` the spatially correlated residuals; out_dev
 * [M][N][E] */
int oqg_gmf(oqg_handle *h, const float *mean_stds_dev,
            const float *within_eps_dev, const float *between_eps_dev,
            const unsigned char *not_mmi, int M, int N, int E, int mode,
            float *out_dev);

/* ---- sampling side of the ground motion fields: what
 * GmfComputer.compute_all (hazardlib/calc/gmf.py:394-474) does around
 * _compute ----------------------------------------------------------------
 * oqg_truncnorm_rvs: n samples of scipy.stats.truncnorm(-tl, tl).rvs(size,
 * rng) for rng = numpy.random.default_rng(seed) (PCG64), float32 as the
 * reference stores its epsilons (gmf.py:443-446; NoCrossCorrelation.
 * get_inter_eps, correlation_models/cross_imt/no_cross_correlation.py:33-36).
 * state / inc (host, [hi, lo] 64-bit words) = numpy.random.PCG64(seed)
 * .state['state']; `offset` = how many uniforms of the stream were consumed
 * before (the draws of one rng follow each other).  The uniforms are numpy's
 * bit for bit; the inverse CDF is evaluated in float64 by another routine
 * than scipy's, so a sample may differ by one float32 rounding. */
int oqg_truncnorm_rvs(oqg_handle *h, const uint64_t *state, const uint64_t *inc,
                      uint64_t offset, int64_t n, double truncation_level,
                      float *out_dev);
/* correlation matrix exp(coeff * d_ij) over the geodetic distance matrix
 * (Mesh.get_distance_matrix, geo/mesh.py:401-429) of N sites: JB2009 with
 * coeff = -3 / decay_range (correlation_models/spatial/jayaram_baker_2009.py
 * :33-44).  lon / lat in degrees, out_dev [N][N] float64. */
int oqg_corr_matrix(oqg_handle *h, const double *lon_dev, const double *lat_dev,
                    int N, double coeff, double *out_dev);
/* set_max_min (gmf.py:71-91) in place on gmf_dev [M][N][E] float32: values
 * above max_iml[m] become the median exp(mean) (mean_dev [M][N] float32, MMI
 * not exponentiated), then the (site, event) columns whose values are below
 * min_iml[m] for every IMT are zeroed.  max_iml, min_iml, not_mmi: host. */
int oqg_gmf_max_min(oqg_handle *h, float *gmf_dev, const float *mean_dev,
                    const double *max_iml, const double *min_iml,
                    const unsigned char *not_mmi, int M, int N, int E);

/* ---- post-processing of the rates map -------------------
 * oqg_postclassical replaces, for all the sites at once, the loop of
 * postclassical (calculators/classical.py:258-325): MapGetter.get_hcurve
 * (calculators/getters.py:351-364) and build_stat_curve (getters.py:47-76)
 * with mean_curve / std_curve / quantile_curve (hazardlib/stats.py:156-214)
 * and, if use_rates, to_rates / to_probs (calc/mean_rates.py:28-50).
 *   rates_dev   [N][L][Gt] float32, the array of a rates MapArray
 *   g_off_dev   [Gt + 1], g_rlz_dev: the realizations (trt_rlzs % 2^24) of
 *               each gsim column g are g_rlz[g_off[g] .. g_off[g + 1])
 *   weights_dev [W][R] float64, W = 1 (one weight per realization) or M
 *               (IMT-dependent weights, one row per IMT); an IMT whose
 *               weights sum to zero gives zero curves (getters.py:61-63)
 *   kinds, qs   S statistics (host arrays): OQG_STAT_MEAN, OQG_STAT_STD,
 *               OQG_STAT_MAX (stats.max_curve, the weights are ignored),
 *               OQG_STAT_QUANTILE with its quantile in qs[s]
 *   rlz_out_dev [R][N][L] float32 or NULL ('hcurves-rlzs')
 *   stat_out_dev[S][N][L] float32 or NULL ('hcurves-stats')
 * S <= 16, W <= 64.  Realizations using the same set of gsim columns have
 * the same curve and are handled as one group (exactly: see oqg_post.cuh);
 * quantiles need at most 256 distinct groups or R <= 4096; mean, std and max
 * have no limit on R (an error otherwise). */
enum { OQG_STAT_MEAN = 0, OQG_STAT_STD = 1, OQG_STAT_QUANTILE = 2,
       OQG_STAT_MAX = 3 };
int oqg_postclassical(oqg_handle *h, const float *rates_dev, int N, int L,
                      int Gt, int M, const int32_t *g_off_dev,
                      const int32_t *g_rlz_dev, int R,
                      const double *weights_dev, int W,
                      const double *weight_sums, int S, const int32_t *kinds,
                      const double *qs, int use_rates, float *rlz_out_dev,
                      float *stat_out_dev);
/* grouping of the realizations with equal gsim columns in oqg_postclassical
 * (default on; off = one group per realization, for tests and measurements) */
int oqg_set_post_grouping(oqg_handle *h, int on);
/* the kernel path of the last oqg_postclassical call: 0 groups in registers /
 * shared memory, 1 streaming (no quantile), 2 one CTA per element, 3 generic;
 * *groups (may be NULL) = the number of groups it worked on */
int oqg_last_post_path(oqg_handle *h, int *groups);

/* calc_mean_rates (hazardlib/calc/mean_rates.py:62-74): the mean rates
 * out_dev[N][L] (float64) = rates_dev[N][L][Gt] (float32) times the weights
 * of the gsim columns weights_dev[W][Gt], W = 1 (one weight per column, the
 * "fast mean") or W = M (IMT-dependent weights, one row per IMT). */
int oqg_mean_rates(oqg_handle *h, const float *rates_dev, int64_t N, int L,
                   int Gt, int M, const double *weights_dev, int W,
                   double *out_dev);

/* compute_hazard_maps (hazardlib/map_array.py:94-141) for K curves:
 * curves_dev [K][M][L1] float32 PoEs, imls_dev [M][L1], poes_dev [P]
 * -> out_dev [K][M][P] float32 intensity levels (0 where the PoE is above
 * the curve) */
int oqg_hazard_maps(oqg_handle *h, const float *curves_dev, int64_t K, int M,
                    int L1, const double *imls_dev, int P,
                    const double *poes_dev, float *out_dev);

/* The fused kernel evaluates the PoEs from a float32 polynomial table of
 * the truncated-normal survival function (argument in float64) when the
 * table built for the truncation level reproduces the exact formula to 2e-6
 * relative for PoE >= 1e-9; otherwise, or after oqg_set_exact_poe(h, 1), it
 * uses the reference formula with FP64 erf.  oqg_poe_table_info reports
 * whether the table is in use and its measured maximum relative error. */
int oqg_set_exact_poe(oqg_handle *h, int on);
int oqg_poe_table_info(oqg_handle *h, int *in_use, double *max_rel_err);

/* Shape of the fused kernel's CTAs (tuning / A-B measurements; the default
 * 0, 0 is automatic).  subs = 1..4 caps the number of (site tile x rupture
 * chunk) work items that share one CTA and its 8-fold replicated,
 * bank-conflict-free PoE table; plain_table != 0 forces the one-work-item
 * CTAs with the unreplicated table.  oqg_last_fused_subs returns what the
 * last launch used. */
int oqg_set_fused_config(oqg_handle *h, int subs, int plain_table);
int oqg_last_fused_subs(oqg_handle *h);

/* MapArray.to_array + the chunk grouping of _store (hazardlib/map_array.py:
 * 396-405, 520-540, 49-64; calculators/classical.py:58-101): the rates
 * map_dev [N][L][G] float32 as records of rates_dt -- {uint32 sid, uint16 lid,
 * uint16 gid, float32 rate}, 12 bytes -- for the elements with rate > 0.
 * rows_dev [R] (may be NULL = all the N * G rows in natural order) lists the
 * curves g * N + s in output order (the reference: by chunk = sid %
 * num_chunks, then gsim column, then site); inside a curve the levels are in
 * order.  sids_dev [N] (NULL = 0..N-1), gid_dev [G] the gsim IDs, level0 the
 * offset of the level index.  out_dev has room for `capacity` records (NULL:
 * count only); *total gets the number of records.  Synchronises. */
int oqg_rates_to_records(oqg_handle *h, const float *rates_dev, int N, int L,
                         int G, const uint32_t *sids_dev,
                         const int32_t *gid_dev, int level0,
                         const int64_t *rows_dev, int64_t R, void *out_dev,
                         int64_t capacity, int64_t *total);

/* Site grouping of the fused pass (default on, cells of 0.16 degrees): the
 * sites are sorted by (cell, vs30) on the device before the pass so that the
 * 32 sites of a warp are alike and the exact-zero level skipping fires more
 * often; the map rows follow through the permuted site index, the results
 * are bit-identical.  on = 0 keeps the caller's site order. */
int oqg_set_site_grouping(oqg_handle *h, int on, double cell_deg);

/* Statistics pass of the fused kernel: while on, a call with `counters`
 * also adds to counters[1] the PoE evaluations the kernel EXECUTES (the
 * levels its level skipping does not cut off: skipped levels have a PoE of
 * exactly 0), so that a benchmark can state executed next to nominal
 * evaluations.  It costs time; leave it off in timed runs. */
int oqg_set_count_evals(oqg_handle *h, int on);

/* number of kernels launched by this handle so far (for bench accounting) */
unsigned long long oqg_launch_count(oqg_handle *h);
/* average device milliseconds of the last fused kernel launches measured
 * with CUDA events on the handle's stream: fills ms[G], returns G or <0 */
int oqg_last_fused_ms(oqg_handle *h, float *ms, int cap);
/* FP64 FMA microbenchmark: measured DFMA rate of the device in TFLOP/s
 * (2 flop per FMA); synchronises */
int oqg_measure_fp64_peak(oqg_handle *h, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* OQG_H */
