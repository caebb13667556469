/*
 * TEST INFRASTRUCTURE — "Oracle B": scalar CPU restatement of GS-LIVM's rasterizer hot path.
 * See gsr_oracle.h for scope, pinning status and the reference lines each function follows.
 * Build: oracle/Makefile target `cpu` (gcc -O2 -ffp-contract=off: the compiler must not fuse or
 * reorder float operations; every fused multiply-add below is an explicit fmaf()).
 */
#include "gsr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define TILE 16

/* reference include/gs/cuda_rasterizer/auxiliary.h:22-33 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                               0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f,  -0.5900435899266435f};

/* CUDA float->int conversions saturate and map NaN to 0 */
static int f2i_trunc(float x) {
  if (x != x) return 0;
  if (x >= 2147483648.0f) return 2147483647;
  if (x <= -2147483648.0f) return (-2147483647 - 1);
  return (int)x;
}
static int f2i_ceil(float x) { return f2i_trunc(ceilf(x)); }
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* a*b + c*d + e*f the way nvcc contracts the reference's sums: FMUL(c*d), FFMA(a,b,.), FFMA(e,f,.) */
static float dot3_ref(float a, float b, float c, float d, float e, float f) { return fmaf(e, f, fmaf(a, b, c * d)); }

/* transformPoint4x3 / 4x4 row r, auxiliary.h:48-64 */
static float xform_row(const float* m, int r, float x, float y, float z) {
  return dot3_ref(x, m[r], y, m[4 + r], z, m[8 + r]) + m[12 + r];
}

static void rect_of(float px, float py, int radius, int gx, int gy, int* rminx, int* rminy, int* rmaxx, int* rmaxy) {
  /* getRect, auxiliary.h:39-46 */
  const float rf = (float)radius;
  *rminx = imin(gx, imax(0, f2i_trunc((px - rf) * 0.0625f)));
  *rminy = imin(gy, imax(0, f2i_trunc((py - rf) * 0.0625f)));
  *rmaxx = imin(gx, imax(0, f2i_trunc((((px + rf) + 16.0f) - 1.0f) * 0.0625f)));
  *rmaxy = imin(gy, imax(0, f2i_trunc((((py + rf) + 16.0f) - 1.0f) * 0.0625f)));
}

static void rotation_rows(const float* q4, float R[9]) {
  /* forward.cu:146-162 with the quaternion used as given; R[3*i+j] = row i of (glm R as a math matrix) */
  const float r = q4[0], x = q4[1], y = q4[2], z = q4[3];
  const float xz = x * z, rx = r * x, rz = r * z, yy = y * y, zz = z * z;
  const float xz_p_ry = fmaf(r, y, xz), yz_p_rx = fmaf(y, z, rx), xz_m_ry = fmaf(-r, y, xz);
  const float xy_m_rz = fmaf(x, y, -rz), xy_p_rz = fmaf(x, y, rz), yz_m_rx = fmaf(y, z, -rx);
  const float xx_p_yy = fmaf(x, x, yy), yy_p_zz = yy + zz, xx_p_zz = fmaf(x, x, zz);
  R[0] = -(yy_p_zz + yy_p_zz) + 1.0f, R[1] = xy_p_rz + xy_p_rz, R[2] = xz_m_ry + xz_m_ry;
  R[3] = xy_m_rz + xy_m_rz, R[4] = -(xx_p_zz + xx_p_zz) + 1.0f, R[5] = yz_p_rx + yz_p_rx;
  R[6] = xz_p_ry + xz_p_ry, R[7] = yz_m_rx + yz_m_rx, R[8] = -(xx_p_yy + xx_p_yy) + 1.0f;
}

static void cov3d_of(const float s[3], const float R[9], float c[6]) {
  /* forward.cu:164-175: M = S*R, Sigma = M^T M */
  float M[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M[3 * i + j] = s[i] * R[3 * i + j];
  c[0] = dot3_ref(M[0], M[0], M[3], M[3], M[6], M[6]);
  c[1] = dot3_ref(M[0], M[1], M[3], M[4], M[6], M[7]);
  c[2] = dot3_ref(M[0], M[2], M[3], M[5], M[6], M[8]);
  c[3] = dot3_ref(M[1], M[1], M[4], M[4], M[7], M[7]);
  c[4] = dot3_ref(M[1], M[2], M[4], M[5], M[7], M[8]);
  c[5] = dot3_ref(M[2], M[2], M[5], M[5], M[8], M[8]);
}

typedef struct {
  float T0[3], T1[3]; /* rows of the 2x3 part of T = W*J */
  float u[3], v[3];   /* u_k = sum_j T0j Vrk[j][k], v_k likewise */
  float a, b, c;      /* cov2D incl. low-pass */
  float tx, ty, tz, txtz, tytz, limx, limy;
} cov2d_t;

static void cov2d_of(const float* V, float px, float py, float pz, float fx, float fy, float tan_fovx, float tan_fovy,
                     const float c[6], cov2d_t* o) {
  /* computeCov2D, forward.cu:79-133 */
  const float tz = xform_row(V, 2, px, py, pz);
  const float tx0 = xform_row(V, 0, px, py, pz), ty0 = xform_row(V, 1, px, py, pz);
  o->limx = tan_fovx * 1.3f, o->limy = tan_fovy * 1.3f;
  o->txtz = tx0 / tz, o->tytz = ty0 / tz;
  const float txc = fminf(fmaxf(o->txtz, -o->limx), o->limx);
  const float tyc = fminf(fmaxf(o->tytz, -o->limy), o->limy);
  const float J00 = fx / tz, J11 = fy / tz, tz2 = tz * tz;
  const float J02 = ((tz * -txc) * fx) / tz2, J12 = ((tz * -tyc) * fy) / tz2;
  o->tx = txc * tz, o->ty = tyc * tz, o->tz = tz;
  o->T0[0] = fmaf(J02, V[2], V[0] * J00), o->T0[1] = fmaf(J02, V[6], V[4] * J00), o->T0[2] = fmaf(J02, V[10], V[8] * J00);
  o->T1[0] = fmaf(J12, V[2], V[1] * J11), o->T1[1] = fmaf(J12, V[6], V[5] * J11), o->T1[2] = fmaf(J12, V[10], V[9] * J11);
  const float* T0 = o->T0;
  const float* T1 = o->T1;
  o->u[0] = dot3_ref(T0[0], c[0], T0[1], c[1], T0[2], c[2]);
  o->u[1] = dot3_ref(T0[0], c[1], T0[1], c[3], T0[2], c[4]);
  o->u[2] = dot3_ref(T0[0], c[2], T0[1], c[4], T0[2], c[5]);
  o->v[0] = dot3_ref(T1[0], c[0], T1[1], c[1], T1[2], c[2]);
  o->v[1] = dot3_ref(T1[0], c[1], T1[1], c[3], T1[2], c[4]);
  o->v[2] = dot3_ref(T1[0], c[2], T1[1], c[4], T1[2], c[5]);
  o->a = dot3_ref(T0[0], o->u[0], T0[1], o->u[1], T0[2], o->u[2]) + 0.3f;
  o->b = dot3_ref(T0[0], o->v[0], T0[1], o->v[1], T0[2], o->v[2]);
  o->c = dot3_ref(T1[0], o->v[0], T1[1], o->v[1], T1[2], o->v[2]) + 0.3f;
}

static void sh_to_rgb(int deg, const float* sh, float x, float y, float z, float out[3]) {
  /* computeColorFromSH, forward.cu:45-68 (sh is [M][3]) */
  for (int c = 0; c < 3; ++c) {
    float r = SH_C0 * sh[c];
    if (deg > 0) {
      r = r - SH_C1 * y * sh[3 + c] + SH_C1 * z * sh[6 + c] - SH_C1 * x * sh[9 + c];
      if (deg > 1) {
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        r = r + SH_C2[0] * xy * sh[12 + c] + SH_C2[1] * yz * sh[15 + c] + SH_C2[2] * (2.0f * zz - xx - yy) * sh[18 + c] +
            SH_C2[3] * xz * sh[21 + c] + SH_C2[4] * (xx - yy) * sh[24 + c];
        if (deg > 2) {
          r = r + SH_C3[0] * y * (3.0f * xx - yy) * sh[27 + c] + SH_C3[1] * xy * z * sh[30 + c] +
              SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[33 + c] +
              SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + c] +
              SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[39 + c] + SH_C3[5] * z * (xx - yy) * sh[42 + c] +
              SH_C3[6] * x * (xx - 3.0f * yy) * sh[45 + c];
        }
      }
    }
    out[c] = r + 0.5f;
  }
}

int gsro_preprocess(const gsro_inputs* in, int32_t* radii, float* means2D, float* depths, float* cov3D,
                    float* conic_opacity, float* rgb, uint8_t* clamped, uint32_t* tiles_touched) {
  const int gx = (in->W + TILE - 1) / TILE, gy = (in->H + TILE - 1) / TILE;
  const float focal_y = in->H / (2.0f * in->tan_fovy); /* rasterizer_impl.cu:210-211 */
  const float focal_x = in->W / (2.0f * in->tan_fovx);
  const float* V = in->view;
  const float* PM = in->proj;
  for (int idx = 0; idx < in->P; ++idx) {
    radii[idx] = 0; /* forward.cu:217-218 */
    tiles_touched[idx] = 0;
    const float px = in->means3D[3 * idx], py = in->means3D[3 * idx + 1], pz = in->means3D[3 * idx + 2];
    const float depth = xform_row(V, 2, px, py, pz);
    if (depth <= 0.2f) continue; /* forward.cu:223 */
    float s[3] = {0.f, 0.f, 0.f};
    if (in->scales) { /* scale_abnormal, forward.cu:19-25,227 */
      for (int k = 0; k < 3; ++k) s[k] = in->scales[3 * idx + k] * in->scale_modifier;
      if (s[0] > 0.3f || s[1] > 0.3f || s[2] > 0.3f) continue;
    }
    const float hom_x = xform_row(PM, 0, px, py, pz), hom_y = xform_row(PM, 1, px, py, pz);
    const float hom_w = xform_row(PM, 3, px, py, pz) + 0.0000001f;
    const float p_w = 1.0f / hom_w;
    const float proj_x = hom_x * p_w, proj_y = hom_y * p_w;

    float c[6];
    if (in->cov3D_precomp) {
      memcpy(c, in->cov3D_precomp + 6 * (size_t)idx, sizeof(c));
    } else {
      float R[9];
      rotation_rows(in->rotations + 4 * (size_t)idx, R);
      cov3d_of(s, R, c);
      memcpy(cov3D + 6 * (size_t)idx, c, sizeof(c));
    }
    cov2d_t q;
    cov2d_of(V, px, py, pz, focal_x, focal_y, in->tan_fovx, in->tan_fovy, c, &q);
    const float det = fmaf(q.a, q.c, -(q.b * q.b)); /* forward.cu:250 */
    if (det == 0.0f) continue;
    const float det_inv = 1.0f / det;
    const float con_x = q.c * det_inv, con_y = q.b * -det_inv, con_z = q.a * det_inv;
    const float mid = (q.a + q.c) * 0.5f;
    const float disc = sqrtf(fmaxf(fmaf(mid, mid, -det), 0.1f));
    const float lam = fmaxf(mid + disc, mid - disc);
    const int radius = f2i_ceil(sqrtf(lam) * 3.0f); /* forward.cu:263 */
    /* ndc2Pix in double, auxiliary.h:35-37 */
    const float pix_x = (float)(fma((double)proj_x + 1.0, (double)in->W, -1.0) * 0.5);
    const float pix_y = (float)(fma((double)proj_y + 1.0, (double)in->H, -1.0) * 0.5);
    int x0, y0, x1, y1;
    rect_of(pix_x, pix_y, radius, gx, gy, &x0, &y0, &x1, &y1);
    const uint32_t tiles = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);
    if (tiles == 0) continue;

    if (!in->colors_precomp) { /* forward.cu:272-277, :40-42, :68-75 */
      float dx = px - in->campos[0], dy = py - in->campos[1], dz = pz - in->campos[2];
      const float len = sqrtf(fmaf(dz, dz, fmaf(dx, dx, dy * dy)));
      dx = dx / len, dy = dy / len, dz = dz / len;
      float col[3];
      sh_to_rgb(in->D, in->shs + (size_t)idx * in->M * 3, dx, dy, dz, col);
      for (int k = 0; k < 3; ++k) {
        clamped[3 * (size_t)idx + k] = col[k] < 0.f;
        rgb[3 * (size_t)idx + k] = fmaxf(col[k], 0.0f);
      }
    }
    depths[idx] = depth;
    radii[idx] = radius;
    means2D[2 * (size_t)idx] = pix_x, means2D[2 * (size_t)idx + 1] = pix_y;
    conic_opacity[4 * (size_t)idx] = con_x, conic_opacity[4 * (size_t)idx + 1] = con_y;
    conic_opacity[4 * (size_t)idx + 2] = con_z, conic_opacity[4 * (size_t)idx + 3] = in->opacities[idx];
    tiles_touched[idx] = tiles;
  }
  return 0;
}

int64_t gsro_scan(int P, const uint32_t* tiles_touched, uint32_t* offsets) {
  uint32_t acc = 0;
  for (int i = 0; i < P; ++i) {
    acc += tiles_touched[i];
    offsets[i] = acc;
  }
  return P > 0 ? (int64_t)offsets[P - 1] : 0;
}

static uint32_t higher_msb(uint32_t n) { /* getHigherMsb, rasterizer_impl.cu:35-48 */
  uint32_t msb = sizeof(n) * 4, step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb) msb += step; else msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

int gsro_bin(int P, int64_t R, int W, int H, const int32_t* radii, const float* means2D, const float* depths,
             const uint32_t* offsets, uint64_t* keys_unsorted, uint32_t* vals_unsorted, uint64_t* keys,
             uint32_t* vals, uint32_t* ranges) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  for (int idx = 0; idx < P; ++idx) { /* duplicateWithKeys */
    if (radii[idx] <= 0) continue;
    uint32_t off = idx == 0 ? 0u : offsets[idx - 1];
    int x0, y0, x1, y1;
    rect_of(means2D[2 * (size_t)idx], means2D[2 * (size_t)idx + 1], radii[idx], gx, gy, &x0, &y0, &x1, &y1);
    uint32_t dbits;
    memcpy(&dbits, &depths[idx], 4);
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        keys_unsorted[off] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
        vals_unsorted[off] = (uint32_t)idx;
        ++off;
      }
  }
  /* stable LSD radix sort on bits [0, 32+bit), 8 bits per pass == what cub::DeviceRadixSort guarantees */
  const int end_bit = 32 + (int)higher_msb((uint32_t)(gx * gy));
  uint64_t* kt = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(R > 0 ? R : 1));
  uint32_t* vt = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(R > 0 ? R : 1));
  if (!kt || !vt) return -1;
  memcpy(keys, keys_unsorted, sizeof(uint64_t) * (size_t)R);
  memcpy(vals, vals_unsorted, sizeof(uint32_t) * (size_t)R);
  uint64_t *ka = keys, *kb = kt;
  uint32_t *va = vals, *vb = vt;
  for (int shift = 0; shift < end_bit; shift += 8) {
    const int bits = end_bit - shift < 8 ? end_bit - shift : 8;
    const uint64_t mask = ((uint64_t)1 << bits) - 1;
    size_t count[257];
    memset(count, 0, sizeof(count));
    for (int64_t i = 0; i < R; ++i) count[((ka[i] >> shift) & mask) + 1]++;
    for (int d = 0; d < 256; ++d) count[d + 1] += count[d];
    for (int64_t i = 0; i < R; ++i) {
      const size_t dst = count[(ka[i] >> shift) & mask]++;
      kb[dst] = ka[i], vb[dst] = va[i];
    }
    uint64_t* tk = ka; ka = kb; kb = tk;
    uint32_t* tv = va; va = vb; vb = tv;
  }
  if (ka != keys) {
    memcpy(keys, ka, sizeof(uint64_t) * (size_t)R);
    memcpy(vals, va, sizeof(uint32_t) * (size_t)R);
  }
  free(kt), free(vt);
  memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)(gx * gy)); /* rasterizer_impl.cu:311 */
  for (int64_t i = 0; i < R; ++i) { /* identifyTileRanges */
    const uint32_t tile = (uint32_t)(keys[i] >> 32);
    if (i == 0) ranges[2 * tile] = 0;
    else {
      const uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
      if (tile != prev) ranges[2 * prev + 1] = (uint32_t)i, ranges[2 * tile] = (uint32_t)i;
    }
    if (i == R - 1) ranges[2 * tile + 1] = (uint32_t)R;
  }
  return 0;
}


/* ---- minimal pthread parallel-for over tiles (this toolchain has no libgomp) ------------------- */
typedef void (*tile_fn)(int tile, void* ctx, int worker);
typedef struct { tile_fn fn; void* ctx; int ntiles; int* next; int worker; } pf_task;
static void* pf_main(void* p) {
  pf_task* t = (pf_task*)p;
  for (;;) {
    const int tile = __atomic_fetch_add(t->next, 1, __ATOMIC_RELAXED);
    if (tile >= t->ntiles) break;
    t->fn(tile, t->ctx, t->worker);
  }
  return 0;
}
#define GSRO_MAX_THREADS 256
static void parallel_tiles(int ntiles, int nthreads, tile_fn fn, void* ctx) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > GSRO_MAX_THREADS) nthreads = GSRO_MAX_THREADS;
  int next = 0;
  pf_task tasks[GSRO_MAX_THREADS];
  pthread_t th[GSRO_MAX_THREADS];
  for (int w = 0; w < nthreads; ++w) {
    tasks[w].fn = fn, tasks[w].ctx = ctx, tasks[w].ntiles = ntiles, tasks[w].next = &next, tasks[w].worker = w;
    if (w > 0) pthread_create(&th[w], 0, pf_main, &tasks[w]);
  }
  pf_main(&tasks[0]);
  for (int w = 1; w < nthreads; ++w) pthread_join(th[w], 0);
}

static int near_rel(float v, float thr, float rel) { return fabsf(v - thr) <= rel * thr; }

typedef struct {
  int W, H, gx;
  const uint32_t *ranges, *point_list;
  const float *means2D, *conic_opacity, *features, *depths, *bg;
  float *out_color, *out_depth, *out_acc, *final_T;
  uint32_t* n_contrib;
  uint8_t* fragile;
  uint64_t (*cnt)[8]; /* per worker: X, Xp, Xa, Xb (+ padding) */
} fwd_ctx;

static void blend_fwd_tile(int tile, void* vctx, int worker) {
  fwd_ctx* c = (fwd_ctx*)vctx;
  const int W = c->W, H = c->H;
  const int tx = tile % c->gx, ty = tile / c->gx;
  const uint32_t r0 = c->ranges[2 * tile], r1 = c->ranges[2 * tile + 1];
  uint64_t X = 0, Xp = 0, Xa = 0, Xb = 0;
  for (int ly = 0; ly < TILE; ++ly)
    for (int lx = 0; lx < TILE; ++lx) {
      const int px = tx * TILE + lx, py = ty * TILE + ly;
      if (px >= W || py >= H) continue;
      const size_t pix = (size_t)W * py + px;
      const float pxf = (float)px, pyf = (float)py;
      float T = 1.0f, A = 0.f, D = 0.f, C[3] = {0.f, 0.f, 0.f};
      uint32_t contributor = 0, last = 0;
      uint8_t frag = 0;
      for (uint32_t i = r0; i < r1; ++i) { /* forward.cu:357-394 */
        contributor++;
        X++;
        const uint32_t id = c->point_list[i];
        const float dx = c->means2D[2 * (size_t)id] - pxf, dy = c->means2D[2 * (size_t)id + 1] - pyf;
        const float* co = c->conic_opacity + 4 * (size_t)id;
        /* SASS order: power = fma(fma(dx, dx*a, dy*(dy*c)), -0.5, -(dy*(dx*b))) */
        const float power = fmaf(fmaf(dx, dx * co[0], dy * (dy * co[2])), -0.5f, -(dy * (dx * co[1])));
        if (fabsf(power) < 1e-6f && power != 0.0f) frag = 1;
        if (power > 0.0f) continue;
        Xp++;
        const float alpha = fminf(0.99f, co[3] * expf(power));
        if (near_rel(alpha, 1.0f / 255.0f, 1e-5f)) frag = 1;
        if (alpha < 1.0f / 255.0f) continue;
        Xa++;
        const float test_T = T * (1.0f - alpha);
        if (near_rel(test_T, 0.0001f, 1e-4f)) frag = 1;
        if (test_T < 0.0001f) break; /* done = true: nothing later can change this pixel */
        Xb++;
        A = fmaf(T, alpha, A);
        D = fmaf(T, alpha * c->depths[id], D);
        for (int ch = 0; ch < 3; ++ch) C[ch] = fmaf(T, alpha * c->features[3 * (size_t)id + ch], C[ch]);
        T = test_T;
        last = contributor;
      }
      c->final_T[pix] = T;
      c->n_contrib[pix] = last;
      for (int ch = 0; ch < 3; ++ch) c->out_color[(size_t)ch * H * W + pix] = fmaf(c->bg[ch], T, C[ch]);
      c->out_depth[pix] = D;
      c->out_acc[pix] = A;
      if (c->fragile) c->fragile[pix] = frag;
    }
  c->cnt[worker][0] += X, c->cnt[worker][1] += Xp, c->cnt[worker][2] += Xa, c->cnt[worker][3] += Xb;
}

int gsro_blend_fwd(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* means2D,
                   const float* conic_opacity, const float* features, const float* depths, const float* bg,
                   float* out_color, float* out_depth, float* out_acc, float* final_T, uint32_t* n_contrib,
                   uint8_t* fragile, uint64_t counters[4], int nthreads) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  static uint64_t cnt[GSRO_MAX_THREADS][8];
  memset(cnt, 0, sizeof(cnt));
  fwd_ctx c = {W, H, gx, ranges, point_list, means2D, conic_opacity, features, depths, bg,
               out_color, out_depth, out_acc, final_T, n_contrib, fragile, cnt};
  parallel_tiles(gx * gy, nthreads, blend_fwd_tile, &c);
  if (counters) {
    for (int k = 0; k < 4; ++k) counters[k] = 0;
    for (int w = 0; w < GSRO_MAX_THREADS; ++w)
      for (int k = 0; k < 4; ++k) counters[k] += cnt[w][k];
  }
  return 0;
}

static void atomic_add(double* p, double v) { /* CAS loop: workers of different tiles share Gaussians */
  uint64_t old_bits, new_bits;
  double old_v, new_v;
  __atomic_load((uint64_t*)p, &old_bits, __ATOMIC_RELAXED);
  do {
    memcpy(&old_v, &old_bits, 8);
    new_v = old_v + v;
    memcpy(&new_bits, &new_v, 8);
  } while (!__atomic_compare_exchange((uint64_t*)p, &old_bits, &new_bits, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}

typedef struct {
  int W, H, gx;
  const uint32_t *ranges, *point_list;
  const float *means2D, *conic_opacity, *features, *bg, *final_T;
  const uint32_t* n_contrib;
  const float *dL_dpix, *dL_dacc;
  double *dL_dmean2D, *dL_dconic, *dL_dopacity, *dL_dcolor;
  uint64_t (*cnt)[8]; /* per worker: X' (visited), X'_p (power <= 0), X_b (blended) */
} bwd_ctx;

static void blend_bwd_tile(int tile, void* vctx, int worker) {
  bwd_ctx* cx = (bwd_ctx*)vctx;
  const int W = cx->W, H = cx->H, gx = cx->gx;
  const uint32_t *ranges = cx->ranges, *point_list = cx->point_list, *n_contrib = cx->n_contrib;
  const float *means2D = cx->means2D, *conic_opacity = cx->conic_opacity, *features = cx->features, *bg = cx->bg;
  const float *final_T = cx->final_T, *dL_dpix = cx->dL_dpix, *dL_dacc = cx->dL_dacc;
  double *dL_dmean2D = cx->dL_dmean2D, *dL_dconic = cx->dL_dconic, *dL_dopacity = cx->dL_dopacity,
         *dL_dcolor = cx->dL_dcolor;
  const double ddelx_dx = 0.5 * W, ddely_dy = 0.5 * H; /* backward.cu:505-506 */
    const int tx = tile % gx, ty = tile / gx;
    const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    for (int ly = 0; ly < TILE; ++ly)
      for (int lx = 0; lx < TILE; ++lx) {
        const int px = tx * TILE + lx, py = ty * TILE + ly;
        if (px >= W || py >= H) continue;
        const size_t pix = (size_t)W * py + px;
        const float pxf = (float)px, pyf = (float)py;
        const double T_final = final_T[pix];
        double T = T_final;
        const uint32_t last_contributor = n_contrib[pix];
        double dpx[3], dacc = dL_dacc[pix], bg_dot = 0.0;
        for (int ch = 0; ch < 3; ++ch) {
          dpx[ch] = dL_dpix[(size_t)ch * H * W + pix];
          bg_dot += (double)bg[ch] * dpx[ch];
        }
        double accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, accum_acc = 0, last_acc = 0, last_alpha = 0;
        for (uint32_t k = r0 + last_contributor; k-- > r0;) { /* back to front; backward.cu:526-601 */
          const uint32_t id = point_list[k];
          cx->cnt[worker][0]++;
          const float dxf = means2D[2 * (size_t)id] - pxf, dyf = means2D[2 * (size_t)id + 1] - pyf;
          const float* co = conic_opacity + 4 * (size_t)id;
          const float power = fmaf(fmaf(dxf, dxf * co[0], dyf * (dyf * co[2])), -0.5f, -(dyf * (dxf * co[1])));
          if (power > 0.0f) continue;
          cx->cnt[worker][1]++;
          const float Gf = expf(power);
          const float alphaf = fminf(0.99f, co[3] * Gf);
          if (alphaf < 1.0f / 255.0f) continue;
          cx->cnt[worker][2]++;
          const double G = Gf, alpha = alphaf, dx = dxf, dy = dyf;
          T = T / (1.0 - alpha);
          const double dchannel_dcolor = alpha * T;
          double dL_dalpha = 0.0;
          for (int ch = 0; ch < 3; ++ch) {
            const double c = features[3 * (size_t)id + ch];
            accum_rec[ch] = last_alpha * last_color[ch] + (1.0 - last_alpha) * accum_rec[ch];
            last_color[ch] = c;
            dL_dalpha += (c - accum_rec[ch]) * dpx[ch];
            atomic_add(&dL_dcolor[3 * (size_t)id + ch], dchannel_dcolor * dpx[ch]);
          }
          accum_acc = last_alpha * last_acc + (1.0 - last_alpha) * accum_acc; /* c_d == 1, backward.cu:567-571 */
          last_acc = 1.0;
          dL_dalpha += (1.0 - accum_acc) * dacc;
          dL_dalpha *= T;
          last_alpha = alpha;
          dL_dalpha += (-T_final / (1.0 - alpha)) * bg_dot;
          const double dL_dG = (double)co[3] * dL_dalpha;
          const double gdx = G * dx, gdy = G * dy;
          const double dG_ddelx = -gdx * co[0] - gdy * co[1];
          const double dG_ddely = -gdy * co[2] - gdx * co[1];
          atomic_add(&dL_dmean2D[3 * (size_t)id], dL_dG * dG_ddelx * ddelx_dx);
          atomic_add(&dL_dmean2D[3 * (size_t)id + 1], dL_dG * dG_ddely * ddely_dy);
          atomic_add(&dL_dconic[4 * (size_t)id], -0.5 * gdx * dx * dL_dG);
          atomic_add(&dL_dconic[4 * (size_t)id + 1], -0.5 * gdx * dy * dL_dG);
          atomic_add(&dL_dconic[4 * (size_t)id + 3], -0.5 * gdy * dy * dL_dG);
          atomic_add(&dL_dopacity[id], G * dL_dalpha);
        }
      }
}

int gsro_blend_bwd(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* means2D,
                   const float* conic_opacity, const float* features, const float* bg, const float* final_T,
                   const uint32_t* n_contrib, const float* dL_dpix, const float* dL_dacc, double* dL_dmean2D,
                   double* dL_dconic, double* dL_dopacity, double* dL_dcolor, uint64_t counters[3], int nthreads) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  static uint64_t cnt[GSRO_MAX_THREADS][8];
  memset(cnt, 0, sizeof(cnt));
  bwd_ctx c = {W, H, gx, ranges, point_list, means2D, conic_opacity, features, bg, final_T, n_contrib,
               dL_dpix, dL_dacc, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, cnt};
  parallel_tiles(gx * gy, nthreads, blend_bwd_tile, &c);
  if (counters) {
    for (int k = 0; k < 3; ++k) counters[k] = 0;
    for (int w = 0; w < GSRO_MAX_THREADS; ++w)
      for (int k = 0; k < 3; ++k) counters[k] += cnt[w][k];
  }
  return 0;
}

int gsro_geom_bwd(const gsro_inputs* in, const int32_t* radii, const float* cov3D, const uint8_t* clamped,
                  const double* dL_dmean2D, const double* dL_dconic, const double* dL_dcolor, double* dL_dmean3D,
                  double* dL_dcov3D, double* dL_dsh, double* dL_dscale, double* dL_drot) {
  const float fy = in->H / (2.0f * in->tan_fovy), fx = in->W / (2.0f * in->tan_fovx);
  const float* V = in->view;
  const float* PM = in->proj;
  for (int idx = 0; idx < in->P; ++idx) {
    if (!(radii[idx] > 0)) continue;
    const float pxf = in->means3D[3 * idx], pyf = in->means3D[3 * idx + 1], pzf = in->means3D[3 * idx + 2];
    const float* cf = in->cov3D_precomp ? in->cov3D_precomp + 6 * (size_t)idx : cov3D + 6 * (size_t)idx;
    cov2d_t q;
    cov2d_of(V, pxf, pyf, pzf, fx, fy, in->tan_fovx, in->tan_fovy, cf, &q);
    /* ---- computeCov2DCUDA, backward.cu:196-274 ---- */
    const double a = q.a, b = q.b, c = q.c;
    const double dcx = dL_dconic[4 * (size_t)idx], dcy = dL_dconic[4 * (size_t)idx + 1], dcw = dL_dconic[4 * (size_t)idx + 3];
    const double denom = a * c - b * b;
    const float denom_f = (float)denom;
    const float denom2inv_f = 1.0f / ((denom_f * denom_f) + 0.0000001f);
    const double denom2inv = 1.0 / (denom * denom + 0.0000001);
    double dL_da = 0, dL_db = 0, dL_dc = 0, dcov[6] = {0, 0, 0, 0, 0, 0};
    const double T00 = q.T0[0], T01 = q.T0[1], T02 = q.T0[2], T10 = q.T1[0], T11 = q.T1[1], T12 = q.T1[2];
    if (denom2inv_f != 0.f) {
      dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcw);
      dL_dc = denom2inv * (-a * a * dcw + 2 * a * b * dcy + (denom - a * c) * dcx);
      dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcw);
      dcov[0] = T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc;
      dcov[3] = T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc;
      dcov[5] = T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc;
      dcov[1] = 2 * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2 * T10 * T11 * dL_dc;
      dcov[2] = 2 * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2 * T10 * T12 * dL_dc;
      dcov[4] = 2 * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2 * T11 * T12 * dL_dc;
    }
    for (int k = 0; k < 6; ++k) dL_dcov3D[6 * (size_t)idx + k] = dcov[k];
    const double Vrk[3][3] = {{cf[0], cf[1], cf[2]}, {cf[1], cf[3], cf[4]}, {cf[2], cf[4], cf[5]}};
    const double T0[3] = {T00, T01, T02}, T1[3] = {T10, T11, T12};
    double dT0[3], dT1[3];
    for (int k = 0; k < 3; ++k) {
      double u = 0, v = 0;
      for (int j = 0; j < 3; ++j) u += T0[j] * Vrk[k][j], v += T1[j] * Vrk[k][j];
      dT0[k] = 2 * u * dL_da + v * dL_db;
      dT1[k] = 2 * v * dL_dc + u * dL_db;
    }
    const double dJ00 = V[0] * dT0[0] + V[4] * dT0[1] + V[8] * dT0[2];
    const double dJ02 = V[2] * dT0[0] + V[6] * dT0[1] + V[10] * dT0[2];
    const double dJ11 = V[1] * dT1[0] + V[5] * dT1[1] + V[9] * dT1[2];
    const double dJ12 = V[2] * dT1[0] + V[6] * dT1[1] + V[10] * dT1[2];
    const double xg = (q.txtz < -q.limx || q.txtz > q.limx) ? 0.0 : 1.0;
    const double yg = (q.tytz < -q.limy || q.tytz > q.limy) ? 0.0 : 1.0;
    const double itz = 1.0 / q.tz, itz2 = itz * itz, itz3 = itz2 * itz;
    const double dtx = xg * -fx * itz2 * dJ02, dty = yg * -fy * itz2 * dJ12;
    const double dtz = -fx * itz2 * dJ00 - fy * itz2 * dJ11 + (2.0 * fx * q.tx) * itz3 * dJ02 + (2.0 * fy * q.ty) * itz3 * dJ12;
    double g[3] = {V[0] * dtx + V[1] * dty + V[2] * dtz, V[4] * dtx + V[5] * dty + V[6] * dtz,
                   V[8] * dtx + V[9] * dty + V[10] * dtz};
    /* ---- projection, backward.cu:399-416 ---- */
    {
      const double mx = pxf, my = pyf, mz = pzf;
      const double hw = (PM[3] * mx + PM[7] * my + PM[11] * mz + PM[15]) + 0.0000001;
      const double m_w = 1.0 / hw;
      const double mul1 = (PM[0] * mx + PM[4] * my + PM[8] * mz + PM[12]) * m_w * m_w;
      const double mul2 = (PM[1] * mx + PM[5] * my + PM[9] * mz + PM[13]) * m_w * m_w;
      const double d2x = dL_dmean2D[3 * (size_t)idx], d2y = dL_dmean2D[3 * (size_t)idx + 1];
      g[0] += (PM[0] * m_w - PM[3] * mul1) * d2x + (PM[1] * m_w - PM[3] * mul2) * d2y;
      g[1] += (PM[4] * m_w - PM[7] * mul1) * d2x + (PM[5] * m_w - PM[7] * mul2) * d2y;
      g[2] += (PM[8] * m_w - PM[11] * mul1) * d2x + (PM[9] * m_w - PM[11] * mul2) * d2y;
    }
    /* ---- SH, backward.cu:20-135 ---- */
    if (in->shs && !in->colors_precomp) {
      const double ox = (double)pxf - in->campos[0], oy = (double)pyf - in->campos[1], oz = (double)pzf - in->campos[2];
      const double len = sqrt(ox * ox + oy * oy + oz * oz);
      const double x = ox / len, y = oy / len, z = oz / len;
      double dRGB[3];
      for (int ch = 0; ch < 3; ++ch) dRGB[ch] = clamped[3 * (size_t)idx + ch] ? 0.0 : dL_dcolor[3 * (size_t)idx + ch];
      double bs[16] = {0}, bx[16] = {0}, by[16] = {0}, bz[16] = {0};
      const double xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      bs[0] = SH_C0;
      if (in->D > 0) {
        bs[1] = -SH_C1 * y, bs[2] = SH_C1 * z, bs[3] = -SH_C1 * x;
        by[1] = -SH_C1, bz[2] = SH_C1, bx[3] = -SH_C1;
      }
      if (in->D > 1) {
        bs[4] = SH_C2[0] * xy, bs[5] = SH_C2[1] * yz, bs[6] = SH_C2[2] * (2 * zz - xx - yy), bs[7] = SH_C2[3] * xz,
        bs[8] = SH_C2[4] * (xx - yy);
        bx[4] = SH_C2[0] * y, by[4] = SH_C2[0] * x, by[5] = SH_C2[1] * z, bz[5] = SH_C2[1] * y;
        bx[6] = SH_C2[2] * -2 * x, by[6] = SH_C2[2] * -2 * y, bz[6] = SH_C2[2] * 4 * z;
        bx[7] = SH_C2[3] * z, bz[7] = SH_C2[3] * x, bx[8] = SH_C2[4] * 2 * x, by[8] = SH_C2[4] * -2 * y;
      }
      if (in->D > 2) {
        bs[9] = SH_C3[0] * y * (3 * xx - yy), bs[10] = SH_C3[1] * xy * z, bs[11] = SH_C3[2] * y * (4 * zz - xx - yy);
        bs[12] = SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy), bs[13] = SH_C3[4] * x * (4 * zz - xx - yy);
        bs[14] = SH_C3[5] * z * (xx - yy), bs[15] = SH_C3[6] * x * (xx - 3 * yy);
        bx[9] = SH_C3[0] * 6 * xy, by[9] = SH_C3[0] * 3 * (xx - yy);
        bx[10] = SH_C3[1] * yz, by[10] = SH_C3[1] * xz, bz[10] = SH_C3[1] * xy;
        bx[11] = SH_C3[2] * -2 * xy, by[11] = SH_C3[2] * (4 * zz - xx - 3 * yy), bz[11] = SH_C3[2] * 8 * yz;
        bx[12] = SH_C3[3] * -6 * xz, by[12] = SH_C3[3] * -6 * yz, bz[12] = SH_C3[3] * 3 * (2 * zz - xx - yy);
        bx[13] = SH_C3[4] * (4 * zz - 3 * xx - yy), by[13] = SH_C3[4] * -2 * xy, bz[13] = SH_C3[4] * 8 * xz;
        bx[14] = SH_C3[5] * 2 * xz, by[14] = SH_C3[5] * -2 * yz, bz[14] = SH_C3[5] * (xx - yy);
        bx[15] = SH_C3[6] * 3 * (xx - yy), by[15] = SH_C3[6] * -6 * xy;
      }
      const int ncoef = (in->D + 1) * (in->D + 1);
      const float* sh = in->shs + (size_t)idx * in->M * 3;
      double ddx = 0, ddy = 0, ddz = 0;
      for (int k = 0; k < ncoef && k < in->M; ++k)
        for (int ch = 0; ch < 3; ++ch) {
          dL_dsh[((size_t)idx * in->M + k) * 3 + ch] = bs[k] * dRGB[ch];
          if (k > 0) {
            const double w = (double)sh[3 * k + ch] * dRGB[ch];
            ddx += bx[k] * w, ddy += by[k] * w, ddz += bz[k] * w;
          }
        }
      const double sum2 = ox * ox + oy * oy + oz * oz, inv32 = 1.0 / sqrt(sum2 * sum2 * sum2); /* dnormvdv */
      g[0] += ((sum2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * inv32;
      g[1] += (-ox * oy * ddx + (sum2 - oy * oy) * ddy - oz * oy * ddz) * inv32;
      g[2] += (-ox * oz * ddx - oy * oz * ddy + (sum2 - oz * oz) * ddz) * inv32;
    }
    for (int k = 0; k < 3; ++k) dL_dmean3D[3 * (size_t)idx + k] = g[k];
    /* ---- scale / rotation, backward.cu:279-366 ---- */
    if (in->scales && !in->cov3D_precomp) {
      float Rf[9];
      rotation_rows(in->rotations + 4 * (size_t)idx, Rf);
      double s[3], M[9], dM[9];
      for (int k = 0; k < 3; ++k) s[k] = (double)(in->scales[3 * idx + k] * in->scale_modifier);
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[3 * i + j] = s[i] * Rf[3 * i + j];
      const double S[3][3] = {{dcov[0], 0.5 * dcov[1], 0.5 * dcov[2]},
                              {0.5 * dcov[1], dcov[3], 0.5 * dcov[4]},
                              {0.5 * dcov[2], 0.5 * dcov[4], dcov[5]}};
      for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) {
          double acc = 0;
          for (int j = 0; j < 3; ++j) acc += M[3 * i + j] * S[j][k];
          dM[3 * i + k] = 2.0 * acc;
        }
      double E[3][3];
      for (int i = 0; i < 3; ++i) {
        double ds = 0;
        for (int j = 0; j < 3; ++j) ds += (double)Rf[3 * i + j] * dM[3 * i + j], E[i][j] = s[i] * dM[3 * i + j];
        dL_dscale[3 * (size_t)idx + i] = ds;
      }
      const double r = in->rotations[4 * idx], x = in->rotations[4 * idx + 1], y = in->rotations[4 * idx + 2],
                   z = in->rotations[4 * idx + 3];
      dL_drot[4 * (size_t)idx] = 2 * z * (E[0][1] - E[1][0]) + 2 * y * (E[2][0] - E[0][2]) + 2 * x * (E[1][2] - E[2][1]);
      dL_drot[4 * (size_t)idx + 1] = 2 * y * (E[1][0] + E[0][1]) + 2 * z * (E[2][0] + E[0][2]) + 2 * r * (E[1][2] - E[2][1]) -
                                     4 * x * (E[2][2] + E[1][1]);
      dL_drot[4 * (size_t)idx + 2] = 2 * x * (E[1][0] + E[0][1]) + 2 * r * (E[2][0] - E[0][2]) + 2 * z * (E[1][2] + E[2][1]) -
                                     4 * y * (E[2][2] + E[0][0]);
      dL_drot[4 * (size_t)idx + 3] = 2 * r * (E[0][1] - E[1][0]) + 2 * x * (E[2][0] + E[0][2]) + 2 * y * (E[1][2] + E[2][1]) -
                                     4 * z * (E[1][1] + E[0][0]);
    }
  }
  return 0;
}

int gsro_mark_visible(int P, const float* means3D, const float* view, uint8_t* present) {
  for (int i = 0; i < P; ++i)
    present[i] = xform_row(view, 2, means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]) <= 0.2f ? 0 : 1;
  return 0;
}
