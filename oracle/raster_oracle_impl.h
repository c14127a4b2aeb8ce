/*
 * raster_oracle_impl.h -- CPU restatement of the reference 3DGS rasterizer.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md): never linked into or called
 * from the product library.  Included twice by raster_oracle.c:
 *     REAL=float  SUF=32   fp32, following the reference's operation order; where
 *                          nvcc (-fmad=true) contracts a*b+c the fused operation is
 *                          written explicitly (FMA) -- see `dot3` below;
 *     REAL=double SUF=64   same algorithm in fp64 (arbiter for tolerance disputes,
 *                          finite-difference checks).
 * Every function cites the reference lines it follows
 * (dgr = dgmesh/submodules/diff-gaussian-rasterization).
 */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

#if IS_FLOAT
#define FMA(a, b, c) fmaf((a), (b), (c))
#define SQRT(x) sqrtf(x)
#define EXP(x) expf(x)
#define CEIL(x) ceilf(x)
#define FMAX(a, b) fmaxf(a, b)
#define FMIN(a, b) fminf(a, b)
#define R(x) ((float)(x))
#else
#define FMA(a, b, c) ((a) * (b) + (c))
#define SQRT(x) sqrt(x)
#define EXP(x) exp(x)
#define CEIL(x) ceil(x)
#define FMAX(a, b) fmax(a, b)
#define FMIN(a, b) fmin(a, b)
#define R(x) ((double)(x))
#endif

/* a0*b0 + a1*b1 + a2*b2 evaluated left to right with the contraction nvcc applies:
 * the first sum fuses its LEFT product (right product rounded), the second sum fuses
 * its right product:  fma(a2,b2, fma(a0,b0, rn(a1*b1))). */
static inline REAL FN(dot3_)(REAL a0, REAL b0, REAL a1, REAL b1, REAL a2, REAL b2) {
  return FMA(a2, b2, FMA(a0, b0, a1 * b1));
}
#define DOT3 FN(dot3_)

typedef struct { REAL c[3][3]; } FN(M3_);
#define M3 FN(M3_)

/* glm mat3 * mat3, column-major (third_party/glm/glm/detail/type_mat3x3.inl:486-520) */
static inline M3 FN(m3mul_)(const M3* a, const M3* b) {
  M3 r;
  for (int col = 0; col < 3; ++col)
    for (int row = 0; row < 3; ++row)
      r.c[col][row] = DOT3(a->c[0][row], b->c[col][0], a->c[1][row], b->c[col][1], a->c[2][row], b->c[col][2]);
  return r;
}
static inline M3 FN(m3T_)(const M3* a) {
  M3 r;
  for (int col = 0; col < 3; ++col)
    for (int row = 0; row < 3; ++row) r.c[col][row] = a->c[row][col];
  return r;
}
#define M3MUL FN(m3mul_)
#define M3T FN(m3T_)

/* auxiliary.h:58-77: m0*x + m4*y + m8*z + m12 */
static inline REAL FN(aff_)(REAL m0, REAL m4, REAL m8, REAL m12, REAL x, REAL y, REAL z) {
  return DOT3(m0, x, m4, y, m8, z) + m12;
}
#define AFF FN(aff_)

/* dgr/cuda_rasterizer/forward.cu:118-152 (computeCov3D): Sigma = (S R)^T (S R) */
static void FN(cov3d_)(const REAL* scale, REAL mod, const REAL* rot, REAL* cov3D) {
  M3 S = {{{0}}};
  S.c[0][0] = mod * scale[0];
  S.c[1][1] = mod * scale[1];
  S.c[2][2] = mod * scale[2];
  const REAL r = rot[0], x = rot[1], y = rot[2], z = rot[3];
  M3 Rm;
  Rm.c[0][0] = FMA(R(-2.0), FMA(y, y, z * z), R(1.0));
  Rm.c[0][1] = R(2.0) * FMA(x, y, -(r * z));
  Rm.c[0][2] = R(2.0) * FMA(x, z, r * y);
  Rm.c[1][0] = R(2.0) * FMA(x, y, r * z);
  Rm.c[1][1] = FMA(R(-2.0), FMA(x, x, z * z), R(1.0));
  Rm.c[1][2] = R(2.0) * FMA(y, z, -(r * x));
  Rm.c[2][0] = R(2.0) * FMA(x, z, -(r * y));
  Rm.c[2][1] = R(2.0) * FMA(y, z, r * x);
  Rm.c[2][2] = FMA(R(-2.0), FMA(x, x, y * y), R(1.0));
  M3 M = M3MUL(&S, &Rm);
  M3 Mt = M3T(&M);
  M3 Sg = M3MUL(&Mt, &M);
  cov3D[0] = Sg.c[0][0];
  cov3D[1] = Sg.c[0][1];
  cov3D[2] = Sg.c[0][2];
  cov3D[3] = Sg.c[1][1];
  cov3D[4] = Sg.c[1][2];
  cov3D[5] = Sg.c[2][2];
}

typedef struct {
  REAL t[3], txtz, tytz;
  M3 W, T, Vrk;
} FN(Ewa_);
#define EWA FN(Ewa_)

/* forward.cu:74-104 / backward.cu:163-199: clamped view position, J, W, T = W*J, Vrk */
static void FN(ewa_frame_)(const REAL* mean, REAL fx, REAL fy, REAL tan_fovx, REAL tan_fovy, const REAL* cov3D,
                           const REAL* view, EWA* f) {
  REAL t0 = AFF(view[0], view[4], view[8], view[12], mean[0], mean[1], mean[2]);
  REAL t1 = AFF(view[1], view[5], view[9], view[13], mean[0], mean[1], mean[2]);
  REAL t2 = AFF(view[2], view[6], view[10], view[14], mean[0], mean[1], mean[2]);
  const REAL limx = R(1.3) * tan_fovx, limy = R(1.3) * tan_fovy;
  f->txtz = t0 / t2;
  f->tytz = t1 / t2;
  t0 = FMIN(limx, FMAX(-limx, f->txtz)) * t2;
  t1 = FMIN(limy, FMAX(-limy, f->tytz)) * t2;
  f->t[0] = t0; f->t[1] = t1; f->t[2] = t2;
  M3 J = {{{0}}};
  J.c[0][0] = fx / t2;
  J.c[0][2] = -(fx * t0) / (t2 * t2);
  J.c[1][1] = fy / t2;
  J.c[1][2] = -(fy * t1) / (t2 * t2);
  f->W.c[0][0] = view[0]; f->W.c[0][1] = view[4]; f->W.c[0][2] = view[8];
  f->W.c[1][0] = view[1]; f->W.c[1][1] = view[5]; f->W.c[1][2] = view[9];
  f->W.c[2][0] = view[2]; f->W.c[2][1] = view[6]; f->W.c[2][2] = view[10];
  f->T = M3MUL(&f->W, &J);
  f->Vrk.c[0][0] = cov3D[0]; f->Vrk.c[0][1] = cov3D[1]; f->Vrk.c[0][2] = cov3D[2];
  f->Vrk.c[1][0] = cov3D[1]; f->Vrk.c[1][1] = cov3D[3]; f->Vrk.c[1][2] = cov3D[4];
  f->Vrk.c[2][0] = cov3D[2]; f->Vrk.c[2][1] = cov3D[4]; f->Vrk.c[2][2] = cov3D[5];
}
/* forward.cu:106-112: cov = T^T Vrk^T T, +0.3 on the diagonal */
static void FN(ewa_cov2d_)(const EWA* f, REAL* cov) {
  M3 Tt = M3T(&f->T), Vt = M3T(&f->Vrk);
  M3 A = M3MUL(&Tt, &Vt);
  M3 C = M3MUL(&A, &f->T);
  cov[0] = C.c[0][0] + R(0.3);
  cov[1] = C.c[0][1];
  cov[2] = C.c[1][1] + R(0.3);
}

static const double FN(kC1_) = 0.4886025119029199;
static const double FN(kC0_) = 0.28209479177387814;
static const double FN(kC2_)[5] = {1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
                                   -1.0925484305920792, 0.5462742152960396};
static const double FN(kC3_)[7] = {-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
                                   -0.4570457994644658, 1.445305721320277, -0.5900435899266435};
#define C0 R((float)FN(kC0_))
#define C1 R((float)FN(kC1_))
#define C2(i) R((float)FN(kC2_)[i])
#define C3(i) R((float)FN(kC3_)[i])

/* SH basis weights for direction (x,y,z), degree deg: colour = sum_k w[k]*sh[k] (forward.cu:20-71) */
static void FN(sh_weights_)(int deg, REAL x, REAL y, REAL z, REAL* w) {
  for (int k = 0; k < 16; ++k) w[k] = 0;
  w[0] = C0;
  if (deg > 0) {
    w[1] = -C1 * y; w[2] = C1 * z; w[3] = -C1 * x;
    if (deg > 1) {
      const REAL xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      w[4] = C2(0) * xy; w[5] = C2(1) * yz; w[6] = C2(2) * (R(2.0) * zz - xx - yy);
      w[7] = C2(3) * xz; w[8] = C2(4) * (xx - yy);
      if (deg > 2) {
        w[9] = C3(0) * y * (R(3.0) * xx - yy);
        w[10] = C3(1) * xy * z;
        w[11] = C3(2) * y * (R(4.0) * zz - xx - yy);
        w[12] = C3(3) * z * (R(2.0) * zz - R(3.0) * xx - R(3.0) * yy);
        w[13] = C3(4) * x * (R(4.0) * zz - xx - yy);
        w[14] = C3(5) * z * (xx - yy);
        w[15] = C3(6) * x * (xx - R(3.0) * yy);
      }
    }
  }
}

/* auxiliary.h:46-56 (getRect) */
static void FN(rect_)(REAL px, REAL py, int rad, int gx, int gy, int* r) {
  int v;
  v = (int)((px - rad) / 16); r[0] = v < 0 ? 0 : (v > gx ? gx : v);
  v = (int)((py - rad) / 16); r[1] = v < 0 ? 0 : (v > gy ? gy : v);
  v = (int)((px + rad + 16 - 1) / 16); r[2] = v < 0 ? 0 : (v > gx ? gx : v);
  v = (int)((py + rad + 16 - 1) / 16); r[3] = v < 0 ? 0 : (v > gy ? gy : v);
}

/* ---------------------------------------------------------------------------
 * preprocess (forward.cu:155-256).  All arrays are caller-allocated; outputs are
 * zero-filled first, as the reference's torch::full / cudaMemset'd buffers are. */
void FN(orc_preprocess_)(int P, int D, int M, const REAL* means3D, const REAL* scales, REAL scale_modifier,
                         const REAL* rotations, const REAL* opacities, const REAL* shs, const REAL* cov3D_precomp,
                         const REAL* colors_precomp, const REAL* view, const REAL* proj, const REAL* campos, int W,
                         int H, REAL tan_fovx, REAL tan_fovy, int* radii, REAL* means2D, REAL* depths, REAL* cov3Ds,
                         REAL* rgb, REAL* conic_opacity, uint32_t* tiles_touched, uint8_t* clamped) {
  const REAL focal_y = H / (R(2.0) * tan_fovy), focal_x = W / (R(2.0) * tan_fovx);
  const int gx = (W + 15) / 16, gy = (H + 15) / 16;
  for (int i = 0; i < P; ++i) {
    radii[i] = 0; tiles_touched[i] = 0;
    means2D[2 * i] = means2D[2 * i + 1] = 0; depths[i] = 0;
    for (int k = 0; k < 6; ++k) cov3Ds[6 * i + k] = 0;
    for (int k = 0; k < 3; ++k) { rgb[3 * i + k] = 0; clamped[3 * i + k] = 0; }
    for (int k = 0; k < 4; ++k) conic_opacity[4 * i + k] = 0;
    const REAL* p = means3D + 3 * i;
    /* auxiliary.h:139-164 in_frustum */
    const REAL hx = AFF(proj[0], proj[4], proj[8], proj[12], p[0], p[1], p[2]);
    const REAL hy = AFF(proj[1], proj[5], proj[9], proj[13], p[0], p[1], p[2]);
    const REAL hw = AFF(proj[3], proj[7], proj[11], proj[15], p[0], p[1], p[2]);
    const REAL p_w = R(1.0) / (hw + R(0.0000001));
    const REAL projx = hx * p_w, projy = hy * p_w;
    const REAL vz = AFF(view[2], view[6], view[10], view[14], p[0], p[1], p[2]);
    if (vz <= R(0.2)) continue;
    const REAL* cov3D;
    if (cov3D_precomp) cov3D = cov3D_precomp + 6 * i;
    else { FN(cov3d_)(scales + 3 * i, scale_modifier, rotations + 4 * i, cov3Ds + 6 * i); cov3D = cov3Ds + 6 * i; }
    EWA f;
    FN(ewa_frame_)(p, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, view, &f);
    REAL cov[3];
    FN(ewa_cov2d_)(&f, cov);
    const REAL det = FMA(cov[0], cov[2], -(cov[1] * cov[1]));
    if (det == 0) continue;
    const REAL det_inv = R(1.0) / det;
    const REAL con[3] = {cov[2] * det_inv, -cov[1] * det_inv, cov[0] * det_inv};
    const REAL mid = R(0.5) * (cov[0] + cov[2]);
    const REAL disc = SQRT(FMAX(R(0.1), FMA(mid, mid, -det)));
    const REAL l1 = mid + disc, l2 = mid - disc;
    const REAL my_radius = CEIL(R(3.0) * SQRT(FMAX(l1, l2)));
    /* ndc2Pix in double (auxiliary.h:41-44) */
    const REAL pix = (REAL)((((double)projx + 1.0) * W - 1.0) * 0.5);
    const REAL piy = (REAL)((((double)projy + 1.0) * H - 1.0) * 0.5);
    int rc[4];
    FN(rect_)(pix, piy, (int)my_radius, gx, gy, rc);
    if ((rc[2] - rc[0]) * (rc[3] - rc[1]) == 0) continue;
    if (!colors_precomp) {
      REAL d[3] = {p[0] - campos[0], p[1] - campos[1], p[2] - campos[2]};
      const REAL len = SQRT(FMA(d[0], d[0], d[1] * d[1]) + d[2] * d[2]);
      d[0] = d[0] / len; d[1] = d[1] / len; d[2] = d[2] / len;
      REAL w[16];
      FN(sh_weights_)(D, d[0], d[1], d[2], w);
      const int nc = (D + 1) * (D + 1);
      for (int ch = 0; ch < 3; ++ch) {
        REAL r = 0;
        for (int k = 0; k < nc; ++k) r = FMA(w[k], shs[((size_t)i * M + k) * 3 + ch], r);
        r += R(0.5);
        clamped[3 * i + ch] = r < 0;
        rgb[3 * i + ch] = r < 0 ? 0 : r;
      }
    }
    depths[i] = vz;
    radii[i] = (int)my_radius;
    means2D[2 * i] = pix; means2D[2 * i + 1] = piy;
    conic_opacity[4 * i + 0] = con[0]; conic_opacity[4 * i + 1] = con[1];
    conic_opacity[4 * i + 2] = con[2]; conic_opacity[4 * i + 3] = opacities[i];
    tiles_touched[i] = (uint32_t)((rc[3] - rc[1]) * (rc[2] - rc[0]));
  }
}

/* ---------------------------------------------------------------------------
 * binning: duplicateWithKeys + stable sort on (tile | depth bits) + identifyTileRanges
 * (rasterizer_impl.cu:70-138, 277-318).  depth bits are always taken from the fp32
 * value (the key the reference sorts on).  keys/point_list sized R = sum(tiles_touched). */
typedef struct { uint64_t key; uint32_t val; uint32_t seq; } FN(KV_);
static int FN(kvcmp_)(const void* a, const void* b) {
  const FN(KV_)* x = (const FN(KV_)*)a; const FN(KV_)* y = (const FN(KV_)*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->seq < y->seq ? -1 : (x->seq > y->seq);
}
int64_t FN(orc_bin_)(int P, int W, int H, const int* radii, const REAL* means2D, const REAL* depths, uint64_t* keys,
                     uint32_t* point_list, uint32_t* ranges, int64_t R_cap) {
  const int gx = (W + 15) / 16, gy = (H + 15) / 16;
  int64_t n = 0;
  FN(KV_)* kv = (FN(KV_)*)malloc(sizeof(FN(KV_)) * (size_t)(R_cap > 0 ? R_cap : 1));
  for (int i = 0; i < P; ++i) {
    if (radii[i] <= 0) continue;
    int rc[4];
    FN(rect_)(means2D[2 * i], means2D[2 * i + 1], radii[i], gx, gy, rc);
    float df = (float)depths[i];
    uint32_t db;
    memcpy(&db, &df, 4);
    for (int y = rc[1]; y < rc[3]; ++y)
      for (int x = rc[0]; x < rc[2]; ++x) {
        if (n < R_cap) {
          kv[n].key = ((uint64_t)(uint32_t)(y * gx + x) << 32) | db;
          kv[n].val = (uint32_t)i;
          kv[n].seq = (uint32_t)n;
        }
        ++n;
      }
  }
  if (n > R_cap) { free(kv); return -n; }
  qsort(kv, (size_t)n, sizeof(FN(KV_)), FN(kvcmp_));
  for (int t = 0; t < gx * gy; ++t) ranges[2 * t] = ranges[2 * t + 1] = 0;
  for (int64_t i = 0; i < n; ++i) {
    keys[i] = kv[i].key; point_list[i] = kv[i].val;
    const uint32_t cur = (uint32_t)(kv[i].key >> 32);
    if (i == 0) ranges[2 * cur] = 0;
    else {
      const uint32_t prev = (uint32_t)(kv[i - 1].key >> 32);
      if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * cur] = (uint32_t)i; }
    }
    if (i == n - 1) ranges[2 * cur + 1] = (uint32_t)n;
  }
  free(kv);
  return n;
}

/* ---------------------------------------------------------------------------
 * forward blend (forward.cu:261-374), one pixel at a time */
void FN(orc_render_fwd_)(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const REAL* means2D,
                         const REAL* features, const REAL* conic_opacity, const REAL* bg, REAL* final_T,
                         uint32_t* n_contrib, REAL* out_color) {
  const int gx = (W + 15) / 16, gy = (H + 15) / 16;
#pragma omp parallel for schedule(dynamic, 4)
  for (int tile = 0; tile < gx * gy; ++tile) {
    const int tx = tile % gx, ty = tile / gx;
    const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    for (int ly = 0; ly < 16; ++ly)
      for (int lx = 0; lx < 16; ++lx) {
        const int px = tx * 16 + lx, py = ty * 16 + ly;
        if (px >= W || py >= H) continue;
        REAL T = 1, C[3] = {0, 0, 0};
        uint32_t contributor = 0, last = 0;
        for (uint32_t k = r0; k < r1; ++k) {
          contributor++;
          const uint32_t id = point_list[k];
          const REAL dx = means2D[2 * id] - (REAL)px, dy = means2D[2 * id + 1] - (REAL)py;
          const REAL* co = conic_opacity + 4 * id;
          /* -0.5f*(a*dx*dx + c*dy*dy) - b*dx*dy */
          const REAL q = FMA(co[0] * dx, dx, (co[2] * dy) * dy);
          const REAL power = FMA(R(-0.5), q, -((co[1] * dx) * dy));
          if (power > 0) continue;
          const REAL alpha = FMIN(R(0.99), co[3] * EXP(power));
          if (alpha < R(1.0) / R(255.0)) continue;
          const REAL test_T = T * (1 - alpha);
          if (test_T < R(0.0001)) break; /* done = true */
          for (int ch = 0; ch < 3; ++ch) C[ch] = FMA(features[3 * id + ch] * alpha, T, C[ch]);
          T = test_T;
          last = contributor;
        }
        const size_t pid = (size_t)W * py + px;
        final_T[pid] = T;
        n_contrib[pid] = last;
        for (int ch = 0; ch < 3; ++ch) out_color[(size_t)ch * H * W + pid] = FMA(T, bg[ch], C[ch]);
      }
  }
}

/* ---------------------------------------------------------------------------
 * backward blend (backward.cu:399-557).  Gradients are accumulated in double (the
 * reference's atomic order is arbitrary) and written as REAL.
 * dL_dmean2D[P,3], dL_dconic[P,4] (.x .y .w used), dL_dopacity[P], dL_dcolor[P,3] */
void FN(orc_render_bwd_)(int P, int W, int H, const uint32_t* ranges, const uint32_t* point_list, const REAL* bg,
                         const REAL* means2D, const REAL* conic_opacity, const REAL* colors, const REAL* final_Ts,
                         const uint32_t* n_contrib, const REAL* dL_dpix, REAL* dL_dmean2D, REAL* dL_dconic,
                         REAL* dL_dopacity, REAL* dL_dcolor) {
  const int gx = (W + 15) / 16, gy = (H + 15) / 16;
  double* acc = (double*)calloc((size_t)P * 9, sizeof(double));
  const REAL ddelx_dx = (REAL)(0.5 * W), ddely_dy = (REAL)(0.5 * H);
  for (int tile = 0; tile < gx * gy; ++tile) {
    const int tx = tile % gx, ty = tile / gx;
    const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    for (int ly = 0; ly < 16; ++ly)
      for (int lx = 0; lx < 16; ++lx) {
        const int px = tx * 16 + lx, py = ty * 16 + ly;
        if (px >= W || py >= H) continue;
        const size_t pid = (size_t)W * py + px;
        const REAL T_final = final_Ts[pid];
        REAL T = T_final;
        const uint32_t last_contributor = n_contrib[pid];
        REAL dLp[3], accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0;
        for (int ch = 0; ch < 3; ++ch) dLp[ch] = dL_dpix[(size_t)ch * H * W + pid];
        REAL bg_dot = 0;
        for (int ch = 0; ch < 3; ++ch) bg_dot = FMA(bg[ch], dLp[ch], bg_dot);
        for (uint32_t kk = last_contributor; kk-- > 0;) {
          const uint32_t id = point_list[r0 + kk];
          (void)r1;
          const REAL dx = means2D[2 * id] - (REAL)px, dy = means2D[2 * id + 1] - (REAL)py;
          const REAL* co = conic_opacity + 4 * id;
          const REAL q = FMA(co[0] * dx, dx, (co[2] * dy) * dy);
          const REAL power = FMA(R(-0.5), q, -((co[1] * dx) * dy));
          if (power > 0) continue;
          const REAL G = EXP(power);
          const REAL alpha = FMIN(R(0.99), co[3] * G);
          if (alpha < R(1.0) / R(255.0)) continue;
          T = T / (R(1.0) - alpha);
          const REAL dchannel_dcolor = alpha * T;
          REAL dL_dalpha = 0;
          double* a = acc + (size_t)id * 9;
          for (int ch = 0; ch < 3; ++ch) {
            const REAL c = colors[3 * id + ch];
            accum_rec[ch] = FMA(last_alpha, last_color[ch], (R(1.0) - last_alpha) * accum_rec[ch]);
            last_color[ch] = c;
            dL_dalpha = FMA(c - accum_rec[ch], dLp[ch], dL_dalpha);
            a[6 + ch] += (double)(dchannel_dcolor * dLp[ch]);
          }
          dL_dalpha *= T;
          last_alpha = alpha;
          dL_dalpha = FMA(-T_final / (R(1.0) - alpha), bg_dot, dL_dalpha);
          const REAL dL_dG = co[3] * dL_dalpha;
          const REAL gdx = G * dx, gdy = G * dy;
          const REAL dG_ddelx = FMA(-gdx, co[0], -(gdy * co[1]));
          const REAL dG_ddely = FMA(-gdy, co[2], -(gdx * co[1]));
          a[0] += (double)(dL_dG * dG_ddelx * ddelx_dx);
          a[1] += (double)(dL_dG * dG_ddely * ddely_dy);
          a[2] += (double)(R(-0.5) * gdx * dx * dL_dG);
          a[3] += (double)(R(-0.5) * gdx * dy * dL_dG);
          a[4] += (double)(R(-0.5) * gdy * dy * dL_dG);
          a[5] += (double)(G * dL_dalpha);
        }
      }
  }
  for (int i = 0; i < P; ++i) {
    const double* a = acc + (size_t)i * 9;
    dL_dmean2D[3 * i] = (REAL)a[0]; dL_dmean2D[3 * i + 1] = (REAL)a[1]; dL_dmean2D[3 * i + 2] = 0;
    dL_dconic[4 * i] = (REAL)a[2]; dL_dconic[4 * i + 1] = (REAL)a[3]; dL_dconic[4 * i + 2] = 0;
    dL_dconic[4 * i + 3] = (REAL)a[4];
    dL_dopacity[i] = (REAL)a[5];
    for (int ch = 0; ch < 3; ++ch) dL_dcolor[3 * i + ch] = (REAL)a[6 + ch];
  }
  free(acc);
}

/* ---------------------------------------------------------------------------
 * preprocess backward: computeCov2DCUDA + preprocessCUDA (backward.cu:144-396).
 * Outputs are zero for radii == 0 (the reference hands in zero-filled tensors). */
void FN(orc_preprocess_bwd_)(int P, int D, int M, const REAL* means3D, const int* radii, const REAL* shs,
                             const uint8_t* clamped, const REAL* scales, const REAL* rotations, REAL scale_modifier,
                             const REAL* cov3Ds, const REAL* view, const REAL* proj, int W, int H, REAL tan_fovx,
                             REAL tan_fovy, const REAL* campos, const REAL* dL_dmean2D, const REAL* dL_dconics,
                             const REAL* dL_dcolor, REAL* dL_dmeans, REAL* dL_dcov, REAL* dL_dsh, REAL* dL_dscale,
                             REAL* dL_drot) {
  const REAL h_y = H / (R(2.0) * tan_fovy), h_x = W / (R(2.0) * tan_fovx);
  for (int i = 0; i < P; ++i) {
    for (int k = 0; k < 3; ++k) dL_dmeans[3 * i + k] = 0;
    for (int k = 0; k < 6; ++k) dL_dcov[6 * i + k] = 0;
    if (dL_dsh) for (int k = 0; k < 3 * M; ++k) dL_dsh[(size_t)i * M * 3 + k] = 0;
    for (int k = 0; k < 3; ++k) dL_dscale[3 * i + k] = 0;
    for (int k = 0; k < 4; ++k) dL_drot[4 * i + k] = 0;
    if (!(radii[i] > 0)) continue;
    const REAL* mean = means3D + 3 * i;
    const REAL* cov3D = cov3Ds + 6 * i;
    const REAL dcon[3] = {dL_dconics[4 * i], dL_dconics[4 * i + 1], dL_dconics[4 * i + 3]};
    EWA f;
    FN(ewa_frame_)(mean, h_x, h_y, tan_fovx, tan_fovy, cov3D, view, &f);
    const REAL limx = R(1.3) * tan_fovx, limy = R(1.3) * tan_fovy;
    const REAL xg = (f.txtz < -limx || f.txtz > limx) ? 0 : 1, yg = (f.tytz < -limy || f.tytz > limy) ? 0 : 1;
    REAL c2[3];
    FN(ewa_cov2d_)(&f, c2);
    const REAL a = c2[0], b = c2[1], c = c2[2];
    const REAL denom = a * c - b * b;
    REAL dL_da = 0, dL_db = 0, dL_dc = 0;
    const REAL denom2inv = R(1.0) / ((denom * denom) + R(0.0000001));
#define T_(i_, j_) f.T.c[i_][j_]
#define V_(i_, j_) f.Vrk.c[i_][j_]
#define W_(i_, j_) f.W.c[i_][j_]
    if (denom2inv != 0) {
      dL_da = denom2inv * (-c * c * dcon[0] + 2 * b * c * dcon[1] + (denom - a * c) * dcon[2]);
      dL_dc = denom2inv * (-a * a * dcon[2] + 2 * a * b * dcon[1] + (denom - a * c) * dcon[0]);
      dL_db = denom2inv * 2 * (b * c * dcon[0] - (denom + 2 * b * b) * dcon[1] + a * b * dcon[2]);
      dL_dcov[6 * i + 0] = (T_(0, 0) * T_(0, 0) * dL_da + T_(0, 0) * T_(1, 0) * dL_db + T_(1, 0) * T_(1, 0) * dL_dc);
      dL_dcov[6 * i + 3] = (T_(0, 1) * T_(0, 1) * dL_da + T_(0, 1) * T_(1, 1) * dL_db + T_(1, 1) * T_(1, 1) * dL_dc);
      dL_dcov[6 * i + 5] = (T_(0, 2) * T_(0, 2) * dL_da + T_(0, 2) * T_(1, 2) * dL_db + T_(1, 2) * T_(1, 2) * dL_dc);
      dL_dcov[6 * i + 1] = 2 * T_(0, 0) * T_(0, 1) * dL_da + (T_(0, 0) * T_(1, 1) + T_(0, 1) * T_(1, 0)) * dL_db +
                           2 * T_(1, 0) * T_(1, 1) * dL_dc;
      dL_dcov[6 * i + 2] = 2 * T_(0, 0) * T_(0, 2) * dL_da + (T_(0, 0) * T_(1, 2) + T_(0, 2) * T_(1, 0)) * dL_db +
                           2 * T_(1, 0) * T_(1, 2) * dL_dc;
      dL_dcov[6 * i + 4] = 2 * T_(0, 2) * T_(0, 1) * dL_da + (T_(0, 1) * T_(1, 2) + T_(0, 2) * T_(1, 1)) * dL_db +
                           2 * T_(1, 1) * T_(1, 2) * dL_dc;
    }
    REAL dT[2][3];
    for (int k = 0; k < 3; ++k) {
      const REAL r0 = T_(0, 0) * V_(k, 0) + T_(0, 1) * V_(k, 1) + T_(0, 2) * V_(k, 2);
      const REAL r1 = T_(1, 0) * V_(k, 0) + T_(1, 1) * V_(k, 1) + T_(1, 2) * V_(k, 2);
      dT[0][k] = 2 * r0 * dL_da + r1 * dL_db;
      dT[1][k] = 2 * r1 * dL_dc + r0 * dL_db;
    }
    const REAL dJ00 = W_(0, 0) * dT[0][0] + W_(0, 1) * dT[0][1] + W_(0, 2) * dT[0][2];
    const REAL dJ02 = W_(2, 0) * dT[0][0] + W_(2, 1) * dT[0][1] + W_(2, 2) * dT[0][2];
    const REAL dJ11 = W_(1, 0) * dT[1][0] + W_(1, 1) * dT[1][1] + W_(1, 2) * dT[1][2];
    const REAL dJ12 = W_(2, 0) * dT[1][0] + W_(2, 1) * dT[1][1] + W_(2, 2) * dT[1][2];
    const REAL tz = R(1.0) / f.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    const REAL dtx = xg * -h_x * tz2 * dJ02, dty = yg * -h_y * tz2 * dJ12;
    const REAL dtz = -h_x * tz2 * dJ00 - h_y * tz2 * dJ11 + (2 * h_x * f.t[0]) * tz3 * dJ02 +
                     (2 * h_y * f.t[1]) * tz3 * dJ12;
    REAL dmean[3];
    dmean[0] = view[0] * dtx + view[1] * dty + view[2] * dtz;
    dmean[1] = view[4] * dtx + view[5] * dty + view[6] * dtz;
    dmean[2] = view[8] * dtx + view[9] * dty + view[10] * dtz;
    /* backward.cu:366-381 */
    const REAL mhw = proj[3] * mean[0] + proj[7] * mean[1] + proj[11] * mean[2] + proj[15];
    const REAL m_w = R(1.0) / (mhw + R(0.0000001));
    const REAL mul1 = (proj[0] * mean[0] + proj[4] * mean[1] + proj[8] * mean[2] + proj[12]) * m_w * m_w;
    const REAL mul2 = (proj[1] * mean[0] + proj[5] * mean[1] + proj[9] * mean[2] + proj[13]) * m_w * m_w;
    const REAL g2x = dL_dmean2D[3 * i], g2y = dL_dmean2D[3 * i + 1];
    dmean[0] += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
    dmean[1] += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
    dmean[2] += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;
    /* SH (backward.cu:20-139) */
    if (shs) {
      REAL dorig[3] = {mean[0] - campos[0], mean[1] - campos[1], mean[2] - campos[2]};
      const REAL len = SQRT(dorig[0] * dorig[0] + dorig[1] * dorig[1] + dorig[2] * dorig[2]);
      const REAL x = dorig[0] / len, y = dorig[1] / len, z = dorig[2] / len;
      REAL dRGB[3];
      for (int ch = 0; ch < 3; ++ch) dRGB[ch] = clamped[3 * i + ch] ? 0 : dL_dcolor[3 * i + ch];
      REAL w[16];
      FN(sh_weights_)(D, x, y, z, w);
      const int nc = (D + 1) * (D + 1);
      const REAL* sh = shs + (size_t)i * M * 3;
      for (int k = 0; k < nc; ++k)
        for (int ch = 0; ch < 3; ++ch) dL_dsh[((size_t)i * M + k) * 3 + ch] = w[k] * dRGB[ch];
      REAL ddir[3] = {0, 0, 0};
      const REAL xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      for (int ch = 0; ch < 3; ++ch) {
#define S_(k_) sh[(k_)*3 + ch]
        REAL dx = 0, dy = 0, dz = 0;
        if (D > 0) {
          dx = -C1 * S_(3); dy = -C1 * S_(1); dz = C1 * S_(2);
          if (D > 1) {
            dx += C2(0) * y * S_(4) + C2(2) * R(2.0) * -x * S_(6) + C2(3) * z * S_(7) + C2(4) * R(2.0) * x * S_(8);
            dy += C2(0) * x * S_(4) + C2(1) * z * S_(5) + C2(2) * R(2.0) * -y * S_(6) + C2(4) * R(2.0) * -y * S_(8);
            dz += C2(1) * y * S_(5) + C2(2) * R(2.0) * R(2.0) * z * S_(6) + C2(3) * x * S_(7);
            if (D > 2) {
              dx += (C3(0) * S_(9) * R(3.0) * R(2.0) * xy + C3(1) * S_(10) * yz + C3(2) * S_(11) * R(-2.0) * xy +
                     C3(3) * S_(12) * R(-3.0) * R(2.0) * xz + C3(4) * S_(13) * (R(-3.0) * xx + R(4.0) * zz - yy) +
                     C3(5) * S_(14) * R(2.0) * xz + C3(6) * S_(15) * R(3.0) * (xx - yy));
              dy += (C3(0) * S_(9) * R(3.0) * (xx - yy) + C3(1) * S_(10) * xz +
                     C3(2) * S_(11) * (R(-3.0) * yy + R(4.0) * zz - xx) + C3(3) * S_(12) * R(-3.0) * R(2.0) * yz +
                     C3(4) * S_(13) * R(-2.0) * xy + C3(5) * S_(14) * R(-2.0) * yz +
                     C3(6) * S_(15) * R(-3.0) * R(2.0) * xy);
              dz += (C3(1) * S_(10) * xy + C3(2) * S_(11) * R(4.0) * R(2.0) * yz +
                     C3(3) * S_(12) * R(3.0) * (R(2.0) * zz - xx - yy) + C3(4) * S_(13) * R(4.0) * R(2.0) * xz +
                     C3(5) * S_(14) * (xx - yy));
            }
          }
        }
#undef S_
        ddir[0] += dx * dRGB[ch]; ddir[1] += dy * dRGB[ch]; ddir[2] += dz * dRGB[ch];
      }
      /* auxiliary.h:112-122 dnormvdv */
      const REAL sum2 = dorig[0] * dorig[0] + dorig[1] * dorig[1] + dorig[2] * dorig[2];
      const REAL inv32 = R(1.0) / SQRT(sum2 * sum2 * sum2);
      dmean[0] += ((+sum2 - dorig[0] * dorig[0]) * ddir[0] - dorig[1] * dorig[0] * ddir[1] -
                   dorig[2] * dorig[0] * ddir[2]) * inv32;
      dmean[1] += (-dorig[0] * dorig[1] * ddir[0] + (sum2 - dorig[1] * dorig[1]) * ddir[1] -
                   dorig[2] * dorig[1] * ddir[2]) * inv32;
      dmean[2] += (-dorig[0] * dorig[2] * ddir[0] - dorig[1] * dorig[2] * ddir[1] +
                   (sum2 - dorig[2] * dorig[2]) * ddir[2]) * inv32;
    }
    for (int k = 0; k < 3; ++k) dL_dmeans[3 * i + k] = dmean[k];
    /* scale / rotation (backward.cu:279-341) */
    if (scales) {
      const REAL* q = rotations + 4 * i;
      const REAL r = q[0], x = q[1], y = q[2], z = q[3];
      M3 Rm;
      Rm.c[0][0] = R(1.0) - R(2.0) * (y * y + z * z); Rm.c[0][1] = R(2.0) * (x * y - r * z);
      Rm.c[0][2] = R(2.0) * (x * z + r * y);          Rm.c[1][0] = R(2.0) * (x * y + r * z);
      Rm.c[1][1] = R(1.0) - R(2.0) * (x * x + z * z); Rm.c[1][2] = R(2.0) * (y * z - r * x);
      Rm.c[2][0] = R(2.0) * (x * z - r * y);          Rm.c[2][1] = R(2.0) * (y * z + r * x);
      Rm.c[2][2] = R(1.0) - R(2.0) * (x * x + y * y);
      const REAL s[3] = {scale_modifier * scales[3 * i], scale_modifier * scales[3 * i + 1],
                         scale_modifier * scales[3 * i + 2]};
      M3 Mm;
      for (int cc = 0; cc < 3; ++cc) for (int rr = 0; rr < 3; ++rr) Mm.c[cc][rr] = s[rr] * Rm.c[cc][rr];
      const REAL* dc = dL_dcov + 6 * i;
      M3 dS;
      dS.c[0][0] = dc[0]; dS.c[0][1] = R(0.5) * dc[1]; dS.c[0][2] = R(0.5) * dc[2];
      dS.c[1][0] = R(0.5) * dc[1]; dS.c[1][1] = dc[3]; dS.c[1][2] = R(0.5) * dc[4];
      dS.c[2][0] = R(0.5) * dc[2]; dS.c[2][1] = R(0.5) * dc[4]; dS.c[2][2] = dc[5];
      M3 M2;
      for (int cc = 0; cc < 3; ++cc) for (int rr = 0; rr < 3; ++rr) M2.c[cc][rr] = R(2.0) * Mm.c[cc][rr];
      M3 dM = M3MUL(&M2, &dS);
      M3 Rt = M3T(&Rm), dMt = M3T(&dM);
      for (int k = 0; k < 3; ++k)
        dL_dscale[3 * i + k] = Rt.c[k][0] * dMt.c[k][0] + Rt.c[k][1] * dMt.c[k][1] + Rt.c[k][2] * dMt.c[k][2];
      for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) dMt.c[k][j] *= s[k];
#define D_(i_, j_) dMt.c[i_][j_]
      dL_drot[4 * i + 0] = 2 * z * (D_(0, 1) - D_(1, 0)) + 2 * y * (D_(2, 0) - D_(0, 2)) + 2 * x * (D_(1, 2) - D_(2, 1));
      dL_drot[4 * i + 1] = 2 * y * (D_(1, 0) + D_(0, 1)) + 2 * z * (D_(2, 0) + D_(0, 2)) + 2 * r * (D_(1, 2) - D_(2, 1)) -
                           4 * x * (D_(2, 2) + D_(1, 1));
      dL_drot[4 * i + 2] = 2 * x * (D_(1, 0) + D_(0, 1)) + 2 * r * (D_(2, 0) - D_(0, 2)) + 2 * z * (D_(1, 2) + D_(2, 1)) -
                           4 * y * (D_(2, 2) + D_(0, 0));
      dL_drot[4 * i + 3] = 2 * r * (D_(0, 1) - D_(1, 0)) + 2 * x * (D_(2, 0) + D_(0, 2)) + 2 * y * (D_(1, 2) + D_(2, 1)) -
                           4 * z * (D_(1, 1) + D_(0, 0));
#undef D_
    }
#undef T_
#undef V_
#undef W_
  }
}

#undef FMA
#undef SQRT
#undef EXP
#undef CEIL
#undef FMAX
#undef FMIN
#undef R
#undef DOT3
#undef M3
#undef M3MUL
#undef M3T
#undef AFF
#undef EWA
#undef C0
#undef C1
#undef C2
#undef C3
#undef FN
#undef CAT
#undef CAT_
