/*
 * vkgs_oracle.c — CPU restatement of the jaesung-cs/vkgs per-frame hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file, and only as the checker or the timed CPU arm.  The
 * product (vkgs_b200/csrc) has no CPU path and never links or calls this.
 *
 * PARITY PIN STATUS
 *   - camera matrices (a3): pinned against the reference's own camera.cc + glm,
 *     compiled from /root/reference by oracle/build_ref.py (oracle/_ref).
 *   - shader arithmetic (a2 parse_ply.comp, a4 rank.comp, a6 inverse_index.comp,
 *     a7 projection.comp, a8 splat.vert/.frag): pinned against the reference's
 *     own GLSL sources executed on the CPU through vendored glm by the harness
 *     oracle/build_ref.py generates (oracle/_ref/libref_shaders.so); goldens
 *     from that harness are committed under tests/golden/.
 *   - sort (a5): pinned as a property (== std::stable_sort by key, the
 *     reference's own CPU statement bench/cpu_benchmark.cc:29-51, also compiled
 *     into oracle/_ref).
 *   - fixed-function raster + ROP blending (coverage rule, UNORM8 rounding):
 *     PARITY UNPINNED — implementation-defined inside the Vulkan driver, no
 *     reference test, golden image or runnable Vulkan stack exists here.
 *
 * All paths cited below are relative to /root/reference.
 *
 * ARITHMETIC PIN.  IEEE binary32 throughout, operands evaluated in the order GLSL
 * writes them (mat*vec = sum over columns, left to right).  A sum of products is
 * ONE explicit fused chain, s = a0*b0; s = fma(a1,b1,s); s = fma(a2,b2,s) ... -
 * what a GPU shader compiler emits for GLSL's mat*vec / mat*mat / dot - and
 * nothing else is ever contracted (compile with -ffp-contract=off; fmaf() is the
 * only source of FMAs).  The CUDA kernels spell the same chains with fmaf() in a
 * translation unit built with -fmad=false, and IEEE division / square root, so
 * visible count, keys, ids and instance records are bit-exact against this
 * file.  Transcendentals (exp in activation and in the fragment alpha) are not
 * correctly rounded on any device and are tolerance-checked.
 *
 * Matrices are column-major float[16]: m[c*4+r], like glm / GLSL.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define VKO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ helpers */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float h2f(uint16_t h) { _Float16 x; memcpy(&x, &h, 2); return (float)x; }
static inline uint16_t f2h(float f) { _Float16 x = (_Float16)f; uint16_t h; memcpy(&h, &x, 2); return h; }

/* C = A*B for column-major 3x3 stored as m[c*3+r]; element = fma(a2,b2, fma(a1,b1, a0*b0)). */
static void mat3_mul(const float* A, const float* B, float* C) {
  float t[9];
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r)
      t[c * 3 + r] = fmaf(A[2 * 3 + r], B[c * 3 + 2], fmaf(A[1 * 3 + r], B[c * 3 + 1], A[0 * 3 + r] * B[c * 3 + 0]));
  memcpy(C, t, sizeof t);
}
static void mat3_transpose(const float* A, float* T) {
  float t[9];
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) t[c * 3 + r] = A[r * 3 + c];
  memcpy(T, t, sizeof t);
}
/* C = A*B for column-major 4x4. */
static void mat4_mul(const float* A, const float* B, float* C) {
  float t[16];
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r)
      t[c * 4 + r] = fmaf(A[3 * 4 + r], B[c * 4 + 3],
                          fmaf(A[2 * 4 + r], B[c * 4 + 2], fmaf(A[1 * 4 + r], B[c * 4 + 1], A[0 * 4 + r] * B[c * 4 + 0])));
  memcpy(C, t, sizeof t);
}
/* r = M*v */
static void mat4_vec(const float* M, const float* v, float* r) {
  float t[4];
  for (int i = 0; i < 4; ++i)
    t[i] = fmaf(M[3 * 4 + i], v[3], fmaf(M[2 * 4 + i], v[2], fmaf(M[1 * 4 + i], v[1], M[0 * 4 + i] * v[0])));
  memcpy(r, t, sizeof t);
}
static void mat3_of_mat4(const float* M, float* m3) {
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) m3[c * 3 + r] = M[c * 4 + r];
}

/* 4x4 inverse by cofactor expansion (adjugate / determinant), the same scheme
 * glm::inverse uses (third_party/glm/glm/detail/func_matrix.inl compute_inverse<4,4>). */
static void mat4_inverse(const float* m, float* out) {
#define M_(c, r) m[(c) * 4 + (r)]
  float c00 = M_(2, 2) * M_(3, 3) - M_(3, 2) * M_(2, 3);
  float c02 = M_(1, 2) * M_(3, 3) - M_(3, 2) * M_(1, 3);
  float c03 = M_(1, 2) * M_(2, 3) - M_(2, 2) * M_(1, 3);
  float c04 = M_(2, 1) * M_(3, 3) - M_(3, 1) * M_(2, 3);
  float c06 = M_(1, 1) * M_(3, 3) - M_(3, 1) * M_(1, 3);
  float c07 = M_(1, 1) * M_(2, 3) - M_(2, 1) * M_(1, 3);
  float c08 = M_(2, 1) * M_(3, 2) - M_(3, 1) * M_(2, 2);
  float c10 = M_(1, 1) * M_(3, 2) - M_(3, 1) * M_(1, 2);
  float c11 = M_(1, 1) * M_(2, 2) - M_(2, 1) * M_(1, 2);
  float c12 = M_(2, 0) * M_(3, 3) - M_(3, 0) * M_(2, 3);
  float c14 = M_(1, 0) * M_(3, 3) - M_(3, 0) * M_(1, 3);
  float c15 = M_(1, 0) * M_(2, 3) - M_(2, 0) * M_(1, 3);
  float c16 = M_(2, 0) * M_(3, 2) - M_(3, 0) * M_(2, 2);
  float c18 = M_(1, 0) * M_(3, 2) - M_(3, 0) * M_(1, 2);
  float c19 = M_(1, 0) * M_(2, 2) - M_(2, 0) * M_(1, 2);
  float c20 = M_(2, 0) * M_(3, 1) - M_(3, 0) * M_(2, 1);
  float c22 = M_(1, 0) * M_(3, 1) - M_(3, 0) * M_(1, 1);
  float c23 = M_(1, 0) * M_(2, 1) - M_(2, 0) * M_(1, 1);
  float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
  float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
  float v0[4] = {M_(1, 0), M_(0, 0), M_(0, 0), M_(0, 0)}, v1[4] = {M_(1, 1), M_(0, 1), M_(0, 1), M_(0, 1)};
  float v2[4] = {M_(1, 2), M_(0, 2), M_(0, 2), M_(0, 2)}, v3[4] = {M_(1, 3), M_(0, 3), M_(0, 3), M_(0, 3)};
  static const float sa[4] = {+1, -1, +1, -1}, sb[4] = {-1, +1, -1, +1};
  float inv[16];
  for (int i = 0; i < 4; ++i) {
    inv[0 * 4 + i] = ((v1[i] * f0[i] - v2[i] * f1[i]) + v3[i] * f2[i]) * sa[i];
    inv[1 * 4 + i] = ((v0[i] * f0[i] - v2[i] * f3[i]) + v3[i] * f4[i]) * sb[i];
    inv[2 * 4 + i] = ((v0[i] * f1[i] - v1[i] * f3[i]) + v3[i] * f5[i]) * sa[i];
    inv[3 * 4 + i] = ((v0[i] * f2[i] - v1[i] * f4[i]) + v2[i] * f5[i]) * sb[i];
  }
  float det = ((M_(0, 0) * inv[0] + M_(0, 1) * inv[4]) + (M_(0, 2) * inv[8] + M_(0, 3) * inv[12]));
  float rdet = 1.0f / det;
  for (int i = 0; i < 16; ++i) out[i] = inv[i] * rdet;
#undef M_
}

/* ------------------------------------------------------------ a2: activation */

/* parse_ply.comp:43-97 with the 60-entry float-offset table of
 * splat_load_thread.cc:114-135.  rows = PLY body as floats, stride = offsets[59]. */
VKO_API void vko_activate(uint32_t n, const float* rows, const uint32_t* off, float* pos, float* cov,
                          float* opacity, uint16_t* sh) {
  const uint64_t base = off[59];
#pragma omp parallel for schedule(static)
  for (int64_t id = 0; id < (int64_t)n; ++id) {
    const float* p = rows + base * (uint64_t)id;
    float s0 = expf(p[off[3]]), s1 = expf(p[off[4]]), s2 = expf(p[off[5]]); /* parse_ply.comp:46-48 */
    float qx = p[off[6]], qy = p[off[7]], qz = p[off[8]], qw = p[off[9]];
    float len = sqrtf(((qx * qx + qy * qy) + qz * qz) + qw * qw); /* length(q), :52 */
    qx = qx / len; qy = qy / len; qz = qz / len; qw = qw / len;
    float xx = qx * qx, yy = qy * qy, zz = qz * qz, xy = qx * qy, xz = qx * qz, yz = qy * qz;
    float wx = qw * qx, wy = qw * qy, wz = qw * qz;
    float rot[9]; /* rot[c*3+r], parse_ply.comp:64-72 */
    rot[0] = 1.f - 2.f * (yy + zz); rot[1] = 2.f * (xy + wz);       rot[2] = 2.f * (xz - wy);
    rot[3] = 2.f * (xy - wz);       rot[4] = 1.f - 2.f * (xx + zz); rot[5] = 2.f * (yz + wx);
    rot[6] = 2.f * (xz + wy);       rot[7] = 2.f * (yz - wx);       rot[8] = 1.f - 2.f * (xx + yy);
    float ss[9] = {s0 * s0, 0, 0, 0, s1 * s1, 0, 0, 0, s2 * s2};
    float rt[9], c3[9];
    mat3_mul(rot, ss, c3); /* rot * ss * transpose(rot), :78 */
    mat3_transpose(rot, rt);
    mat3_mul(c3, rt, c3);
    cov[6 * id + 0] = c3[0 * 3 + 0]; cov[6 * id + 1] = c3[1 * 3 + 0]; cov[6 * id + 2] = c3[2 * 3 + 0];
    cov[6 * id + 3] = c3[1 * 3 + 1]; cov[6 * id + 4] = c3[2 * 3 + 1]; cov[6 * id + 5] = c3[2 * 3 + 2];
    pos[3 * id + 0] = p[off[0]]; pos[3 * id + 1] = p[off[1]]; pos[3 * id + 2] = p[off[2]];
    for (int i = 0; i < 48; ++i) sh[48 * id + i] = f2h(p[off[10 + i]]); /* :91-94, f16 RNE */
    opacity[id] = 1.f / (1.f + expf(-p[off[58]]));                       /* sigmoid, :30,96 */
  }
}

/* ------------------------------------------------------- a3: camera block */

/* (projection*view)*model, the left-to-right product rank.comp:32 writes. */
VKO_API void vko_compose_pvm(const float* proj, const float* view, const float* model, float* pvm) {
  float pv[16];
  mat4_mul(proj, view, pv);
  mat4_mul(pv, model, pvm);
}
/* inverse(model) * vec4(eye,1), /w   (projection.comp:85-86), hoisted per frame. */
VKO_API void vko_camera_in_model(const float* model, const float* eye, float* out3) {
  float inv[16], e[4] = {eye[0], eye[1], eye[2], 1.f}, r[4];
  mat4_inverse(model, inv);
  mat4_vec(inv, e, r);
  out3[0] = r[0] / r[3]; out3[1] = r[1] / r[3]; out3[2] = r[2] / r[3];
}

/* ---------------------------------------------------------- a4: cull + key */

static inline int cull_one(const float* pvm, const float* p3, uint32_t* key) {
  float v[4] = {p3[0], p3[1], p3[2], 1.f}, c[4];
  mat4_vec(pvm, v, c);
  float iw = 1.f / c[3]; /* pos / pos.w, rank.comp:33, pinned as one IEEE reciprocal and three products */
  float x = c[0] * iw, y = c[1] * iw, z = c[2] * iw;
  if (fabsf(x) <= 1.f && fabsf(y) <= 1.f && z >= 0.f && z <= 1.f) { /* rank.comp:37 */
    *key = f2u(1.f - z);                                             /* rank.comp:39 */
    return 1;
  }
  return 0;
}

/* rank.comp:27-42 with deterministic (ascending-id) compaction.  Returns V. */
VKO_API uint32_t vko_cull(uint32_t n, const float* pos, const float* pvm, uint32_t* keys, uint32_t* ids) {
  const int64_t CH = 1 << 16;
  int64_t nch = ((int64_t)n + CH - 1) / CH;
  uint32_t* cnt = (uint32_t*)calloc((size_t)nch + 1, 4);
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t c = 0; c < nch; ++c) {
    int64_t e = (c + 1) * CH < (int64_t)n ? (c + 1) * CH : (int64_t)n;
    uint32_t k, m = 0;
    for (int64_t i = c * CH; i < e; ++i) m += cull_one(pvm, pos + 3 * i, &k);
    cnt[c + 1] = m;
  }
  for (int64_t c = 0; c < nch; ++c) cnt[c + 1] += cnt[c];
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t c = 0; c < nch; ++c) {
    int64_t e = (c + 1) * CH < (int64_t)n ? (c + 1) * CH : (int64_t)n;
    uint32_t o = cnt[c], k;
    for (int64_t i = c * CH; i < e; ++i)
      if (cull_one(pvm, pos + 3 * i, &k)) { keys[o] = k; ids[o] = (uint32_t)i; ++o; }
  }
  uint32_t v = cnt[nch];
  free(cnt);
  return v;
}

/* ------------------------------------------------------------------ a5: sort */

/* Ascending stable sort of (key,value): the semantics of vrdxCmdSortKeyValue*
 * (third_party/vulkan_radix_sort/src/vk_radix_sort.cc:262-416), stated on the CPU by
 * the author as std::stable_sort by key (bench/cpu_benchmark.cc:29-51).  4x8-bit LSD. */
VKO_API void vko_sort_pairs(uint32_t n, uint32_t* keys, uint32_t* vals) {
  if (n < 2) return;
  uint32_t* k2 = (uint32_t*)malloc((size_t)n * 4);
  uint32_t* v2 = (uint32_t*)malloc((size_t)n * 4);
  uint32_t *ki = keys, *vi = vals, *ko = k2, *vo = v2;
  for (int pass = 0; pass < 4; ++pass) {
    size_t hist[257] = {0};
    int sh = pass * 8;
    for (uint32_t i = 0; i < n; ++i) hist[((ki[i] >> sh) & 255) + 1]++;
    for (int d = 0; d < 256; ++d) hist[d + 1] += hist[d];
    for (uint32_t i = 0; i < n; ++i) {
      size_t o = hist[(ki[i] >> sh) & 255]++;
      ko[o] = ki[i]; vo[o] = vi[i];
    }
    uint32_t* t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
  }
  /* 4 passes: result is back in keys/vals */
  free(k2); free(v2);
}

/* inverse_index.comp:13-18 after the fill(-1) of engine.cc:1226. */
VKO_API void vko_inverse_index(uint32_t n, uint32_t v, const uint32_t* index, int32_t* inverse) {
  for (uint32_t i = 0; i < n; ++i) inverse[i] = -1;
  for (uint32_t i = 0; i < v; ++i) inverse[index[i]] = (int32_t)i;
}

/* ------------------------------------------------------------ a7: projection */

typedef struct {
  float proj[16], view[16], model[16];
  float eye[3];
  uint32_t width, height;
} vko_camera;

/* Per-frame constants of the projection, hoisted out of the per-splat work exactly as the CUDA path's FrameParams
 * carries them (renderer.cu fill_params): products of the camera matrices and reciprocals of the viewport size. */
typedef struct {
  float vm[16];   /* view * model                                   (projection.comp:98-102 composed) */
  float w3[9];    /* mat3(view) * mat3(model), column-major m[c*3+r]  (:95-101 composed)               */
  float ps[4];    /* mat2(projection), column-major m[c*2+r]          (:112)                           */
  float lpx, lpy; /* 1/W/W, 1/H/H                                     (:116-117)                       */
  float cam_m[3]; /* inverse(model) * eye / w                         (:85-86)                         */
} vko_frame;

static void frame_setup(const vko_camera* cam, vko_frame* f) {
  float m3[9], v3[9];
  mat4_mul(cam->view, cam->model, f->vm);
  mat3_of_mat4(cam->model, m3);
  mat3_of_mat4(cam->view, v3);
  mat3_mul(v3, m3, f->w3);
  f->ps[0] = cam->proj[0]; f->ps[1] = cam->proj[1]; f->ps[2] = cam->proj[4]; f->ps[3] = cam->proj[5];
  float fw = (float)cam->width, fh = (float)cam->height;
  f->lpx = 1.f / fw / fw;
  f->lpy = 1.f / fh / fh;
  vko_camera_in_model(cam->model, cam->eye, f->cam_m);
}

/* One splat of projection.comp:77-179, restated with the frame-constant matrix products hoisted and every division
 * by a shared denominator turned into ONE IEEE reciprocal and multiplications (the GLSL leaves both the association
 * of its matrix products' roundings and `a / b` vs `a * (1/b)` to the driver; this is the order both this file and the
 * CUDA kernel commit to, bit for bit).  With W = mat3(view)*mat3(model), t = (view*model)*pos and the first two rows
 * of J (the third never reaches cov2d), cov2d = K * Sigma * K^T for the 2x3 matrix K = mat2(proj) * J * W.
 * variant 0 = pinned half-angle (sqrt/div only, bit-exact on any IEEE device); variant 1 = the literal atan/cos/sin
 * of projection.comp:130-132 through libm (cross-check of the restatement). */
static void project_one(const vko_camera* cam, const vko_frame* f, const float* pos3, const float* cov6,
                        float opac, const uint16_t* sh48, int variant, float* inst) {
  /* dir = normalize(pos - cam_model)   projection.comp:87 */
  float dx = pos3[0] - f->cam_m[0], dy = pos3[1] - f->cam_m[1], dz = pos3[2] - f->cam_m[2];
  float il = 1.f / sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
  float x = dx * il, y = dy * il, z = dz * il;

  /* t = view * model * pos   :98,102 */
  float p4[4] = {pos3[0], pos3[1], pos3[2], 1.f}, pv[4];
  mat4_vec(f->vm, p4, pv);
  float px = pv[0], py = pv[1], pz = pv[2];

  /* rows 0 and 1 of J (:105-108): (-1/z, 0, x/z/z), (0, -1/z, y/z/z); PJ = mat2(proj) * J, 2x3, PJ[r][c] */
  float iz = 1.f / pz, niz = -iz;
  float j02 = (px * iz) * iz, j12 = (py * iz) * iz;
  float P00 = f->ps[0], P10 = f->ps[1], P01 = f->ps[2], P11 = f->ps[3]; /* P[r][c] = ps[c*2+r] */
  float PJ[2][3] = {{P00 * niz, P01 * niz, fmaf(P01, j12, P00 * j02)}, {P10 * niz, P11 * niz, fmaf(P11, j12, P10 * j02)}};
  /* K = PJ * W,  K[r][c] = sum_k PJ[r][k] * W[k][c],  W[k][c] = w3[c*3+k] */
  float K[2][3], M[2][3];
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 3; ++c)
      K[r][c] = fmaf(PJ[r][2], f->w3[c * 3 + 2], fmaf(PJ[r][1], f->w3[c * 3 + 1], PJ[r][0] * f->w3[c * 3 + 0]));
  /* M = K * Sigma,  Sigma = mat3(v0, v0.y, v1.xy, v0.z, v1.yz) symmetric   :92 */
  float S[3][3] = {{cov6[0], cov6[1], cov6[2]}, {cov6[1], cov6[3], cov6[4]}, {cov6[2], cov6[4], cov6[5]}};
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 3; ++c) M[r][c] = fmaf(K[r][2], S[2][c], fmaf(K[r][1], S[1][c], K[r][0] * S[0][c]));
  /* cov2d = M * K^T + low-pass   :112-117;  a = [0][0], b = [1][1], c = [1][0] (column 1, row 0) */
  float a = fmaf(M[0][2], K[0][2], fmaf(M[0][1], K[0][1], M[0][0] * K[0][0])) + f->lpx;
  float b = fmaf(M[1][2], K[1][2], fmaf(M[1][1], K[1][1], M[1][0] * K[1][0])) + f->lpy;
  float c = fmaf(M[0][2], K[1][2], fmaf(M[0][1], K[1][1], M[0][0] * K[1][0]));

  /* eigendecomposition   :122-134 */
  float D = sqrtf(fmaf(4.f * c, c, (a - b) * (a - b)));
  float s0 = sqrtf(0.5f * ((a + b) + D));
  float s1 = sqrtf(0.5f * ((a + b) - D));
  float iD = 1.f / D;
  float sin2t = (2.f * c) * iD, cos2t = (a - b) * iD;
  float ct, st;
  if (variant == 1) {
    float theta = atan2f(sin2t, cos2t) / 2.f;
    ct = cosf(theta); st = sinf(theta);
  } else {
    /* half-angle identities: h = cos or |sin| of the half angle, whichever is >= 1/sqrt(2) (theta in [-pi/4, pi/4] when
     * cos 2t >= 0, |theta| in (pi/4, pi/2] otherwise); the other one is (sin 2t / 2) * (1/h) - one IEEE reciprocal, as
     * for every other quotient of this function.  The NaN lane (D == 0) stays NaN. */
    float h = sqrtf(0.5f * (1.f + fabsf(cos2t)));
    float ih = 1.f / h;
    float q = (0.5f * sin2t) * ih;
    if (cos2t >= 0.f) { ct = h; st = q; }
    else { ct = fabsf(q); st = copysignf(h, sin2t); }
  }

  /* pos = projection * pos; pos /= pos.w   :136-137 */
  float pc[4];
  mat4_vec(cam->proj, pv, pc);
  float iw = 1.f / pc[3];
  float nx = pc[0] * iw, ny = pc[1] * iw, nz = pc[2] * iw;

  /* SH degree 3   :140-174 */
  const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f, C20 = 1.0925484305920792f,
              C21 = 0.31539156525252005f, C22 = 0.5462742152960396f, C30 = 0.5900435899266435f,
              C31 = 2.890611442640554f, C32 = 0.4570457994644658f, C33 = 0.3731763325901154f,
              C34 = 1.445305721320277f;
  float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  float bs[16];
  bs[0] = C0;                 bs[1] = -C1 * y;
  bs[2] = C1 * z;             bs[3] = -C1 * x;
  bs[4] = C20 * xy;           bs[5] = -C20 * yz;
  bs[6] = C21 * ((2.f * zz - xx) - yy);
  bs[7] = -C20 * xz;
  bs[8] = C22 * (xx - yy);    bs[9] = -C30 * y * (3.f * xx - yy);
  bs[10] = C31 * xy * z;      bs[11] = -C32 * y * ((4.f * zz - xx) - yy);
  bs[12] = C33 * z * ((2.f * zz - 3.f * xx) - 3.f * yy);
  bs[13] = -C32 * x * ((4.f * zz - xx) - yy);
  bs[14] = C34 * z * (xx - yy);
  bs[15] = -C30 * x * (xx - 3.f * yy);
  float col[3];
  for (int ch = 0; ch < 3; ++ch) {
    const uint16_t* s = sh48 + 16 * ch;
    float q[4];
    for (int g = 0; g < 4; ++g)
      q[g] = fmaf(bs[4 * g + 3], h2f(s[4 * g + 3]),
                  fmaf(bs[4 * g + 2], h2f(s[4 * g + 2]), fmaf(bs[4 * g + 1], h2f(s[4 * g + 1]), bs[4 * g + 0] * h2f(s[4 * g + 0]))));
    float cc = ((q[0] + q[1]) + q[2]) + q[3];
    cc = cc + 0.5f;
    col[ch] = cc > 0.f ? cc : 0.f; /* max(color + 0.5, 0); NaN -> 0 like GLSL max(NaN,0) is undefined, pin 0 */
  }
  inst[0] = nx; inst[1] = ny; inst[2] = nz; inst[3] = 0.f;
  inst[4] = s0 * ct; inst[5] = s0 * st; inst[6] = -s1 * st; inst[7] = s1 * ct;
  inst[8] = col[0]; inst[9] = col[1]; inst[10] = col[2]; inst[11] = opac;
}

/* inst[i*12..] for i in [0,v): record of splat ids[i] (i.e. written at its sorted slot, :177-179). */
VKO_API void vko_project(uint32_t v, const uint32_t* ids, const float* pos, const float* cov, const float* opacity,
                         const uint16_t* sh, const vko_camera* cam, int variant, float* inst) {
  vko_frame f;
  frame_setup(cam, &f);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)v; ++i) {
    uint64_t id = ids[i];
    project_one(cam, &f, pos + 3 * id, cov + 6 * id, opacity[id], sh + 48 * id, variant, inst + 12 * i);
  }
}

/* ------------------------------------------------------ a8: raster + blend */

/*
 * splat.vert:10-26 + splat.frag:8-12 + fixed-function state (engine.cc:281-299 blend,
 * engine.cc:1382-1387 clear (0,0,0,1) / depth 1, graphics_pipeline.cc:79-81 depth LESS no write,
 * render_pass.cc:15 B8G8R8A8_UNORM).  Splats are blended in array order (sorted: far -> near).
 *
 * Pinned per-fragment arithmetic (the Vulkan rasteriser's own interpolation is
 * implementation-defined; this is the restatement both sides share):
 *   hw = W/2, hh = H/2;  cpx = fma(ndc.x, hw, hw-0.5), cpy likewise    (pixel i centre <-> ndc (i+.5)*2/W-1)
 *   m = diag(hw,hh)*RS;  det = m00*m11 - m01*m10;  A = m^-1 = [m11,-m01;-m10,m00] * (1/det)
 *   per TILE-aligned origin (tx,ty):  ox = tx-cpx, oy = ty-cpy;  bx = A00*ox + A01*oy;  by = A10*ox + A11*oy
 *   per pixel (lx,ly) in tile:        px = fma(A00,lx,fma(A01,ly,bx));  py = fma(A10,lx,fma(A11,ly,by))
 *   covered <=> |px|<=3 && |py|<=3 && ndc.z<1;   alpha = opacity*exp(-0.5*(px*px+py*py))
 * mode 0 (fp32): C = src*a + C*(1-a), A = a*a + A*(1-a) in fp32, quantise once (RNE) at the end.
 * mode 1 (unorm8 ROP): destination re-quantised after every blend; state q in [0,255]:
 *          q = rint(fma(255*src, a, q*(1-a)))   (RNE).
 * out: RGBA8, row-major, y down.  Optional fout = un-quantised fp32 RGBA (mode 0 only).
 */
typedef struct {
  float cpx, cpy, a00, a01, a10, a11, r, g, b, op, z;
  int x0, x1, y0, y1; /* inclusive pixel bbox, clipped; x0>x1 => nothing */
} raster_splat;

static void raster_setup(const float* inst, uint32_t W, uint32_t H, raster_splat* s) {
  float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
  s->x0 = 1; s->x1 = 0; s->y0 = 1; s->y1 = 0;
  if (!(inst[2] < 1.f)) return; /* depth LESS vs cleared 1.0 */
  s->z = inst[2];
  s->cpx = fmaf(inst[0], hw, hw - 0.5f);
  s->cpy = fmaf(inst[1], hh, hh - 0.5f);
  float m00 = inst[4] * hw, m10 = inst[5] * hh, m01 = inst[6] * hw, m11 = inst[7] * hh; /* m[r][c] */
  float det = m00 * m11 - m01 * m10;
  float idet = 1.f / det;
  s->a00 = m11 * idet; s->a01 = -m01 * idet; s->a10 = -m10 * idet; s->a11 = m00 * idet;
  float cr = inst[8], cg = inst[9], cb = inst[10];
  s->r = cr < 0.f ? 0.f : (cr > 1.f ? 1.f : cr);
  s->g = cg < 0.f ? 0.f : (cg > 1.f ? 1.f : cg);
  s->b = cb < 0.f ? 0.f : (cb > 1.f ? 1.f : cb);
  s->op = inst[11];
  float ex = 3.f * (fabsf(m00) + fabsf(m01)), ey = 3.f * (fabsf(m10) + fabsf(m11));
  if (!(ex == ex) || !(ey == ey) || !(det == det) || isinf(ex) || isinf(ey)) return; /* NaN lane: nothing drawn */
  /* conservative bbox (exact test is per pixel) */
  float fx0 = floorf(s->cpx - ex) - 1.f, fx1 = ceilf(s->cpx + ex) + 1.f;
  float fy0 = floorf(s->cpy - ey) - 1.f, fy1 = ceilf(s->cpy + ey) + 1.f;
  if (fx0 < 0.f) fx0 = 0.f;
  if (fy0 < 0.f) fy0 = 0.f;
  if (fx1 > (float)W - 1.f) fx1 = (float)W - 1.f;
  if (fy1 > (float)H - 1.f) fy1 = (float)H - 1.f;
  if (fx0 > fx1 || fy0 > fy1) return;
  s->x0 = (int)fx0; s->x1 = (int)fx1; s->y0 = (int)fy0; s->y1 = (int)fy1;
}

static inline uint8_t q8(float x) {
  float v = rintf(x * 255.f);
  if (!(v > 0.f)) return 0;
  if (v > 255.f) return 255;
  return (uint8_t)v;
}

/* ---------------------------------------------------------------- opaque line layer
 * The reference draws its axis and grid lines BEFORE the splats with depth test and depth write (engine.cc:1440-1469,
 * pipeline state engine.cc:398-415: LINE_LIST; color.vert: gl_Position = projection * view * model * position;
 * color.frag: premultiplied colour), then the splats with depth test LESS and no depth write (engine.cc:298-299): a
 * splat fragment behind a line pixel is discarded, the rest blend over the line colour.
 * Vulkan leaves non-strict line rasterisation to the implementation (and the reference has no test for it): PARITY
 * UNPINNED.  The rule pinned here and in csrc/lines.cu, operation for operation:
 *   clip-space endpoints by the fma chains of mat4_vec; parametric clip against the six frustum planes (t = d0/(d0-d1));
 *   NDC by one reciprocal of w; screen x = fma(ndc.x, W/2, W/2) (pixel i covers [i, i+1)), likewise y;
 *   walk the major axis (|dx| >= |dy| ? x : y) from the smaller to the larger coordinate: every pixel whose centre
 *   c + 0.5 lies in [a0, a1) gets one fragment at u = ((c + 0.5) - a0) * (1 / (a1 - a0)), minor = floor(fma(u, dm, m0)),
 *   depth = fma(u, dz, z0), colour = fma(u, dc, c0) premultiplied and rounded to UNORM8 (the target's format);
 *   depth test LESS with write: the nearest fragment of a pixel wins, equal depths by the smaller packed colour.
 * layer_depth: W*H floats, 1.0 where no line; layer_rgba: W*H*4 bytes, (0,0,0,255) (the clear colour) where no line. */
static inline uint8_t q8(float x);
VKO_API void vko_raster_lines(uint32_t n_lines, const float* pos /* 2n x 3 */, const float* col /* 2n x 4 */,
                              const float* pvm /* proj*view*model of the lines */, uint32_t W, uint32_t H,
                              float* layer_depth, uint8_t* layer_rgba) {
  for (size_t i = 0; i < (size_t)W * H; ++i) {
    layer_depth[i] = 1.f;
    layer_rgba[4 * i + 0] = 0; layer_rgba[4 * i + 1] = 0; layer_rgba[4 * i + 2] = 0; layer_rgba[4 * i + 3] = 255;
  }
  uint64_t* best = (uint64_t*)malloc((size_t)W * H * sizeof(uint64_t));
  for (size_t i = 0; i < (size_t)W * H; ++i) best[i] = ~0ull;
  const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
  for (uint32_t l = 0; l < n_lines; ++l) {
    float p0[4] = {pos[6 * l + 0], pos[6 * l + 1], pos[6 * l + 2], 1.f}, p1[4] = {pos[6 * l + 3], pos[6 * l + 4], pos[6 * l + 5], 1.f};
    float c0[4], c1[4];
    mat4_vec(pvm, p0, c0);
    mat4_vec(pvm, p1, c1);
    float t0 = 0.f, t1 = 1.f;
    int reject = 0;
    for (int pl = 0; pl < 6 && !reject; ++pl) {
      float d0, d1;
      switch (pl) {
        case 0: d0 = c0[3] + c0[0]; d1 = c1[3] + c1[0]; break;
        case 1: d0 = c0[3] - c0[0]; d1 = c1[3] - c1[0]; break;
        case 2: d0 = c0[3] + c0[1]; d1 = c1[3] + c1[1]; break;
        case 3: d0 = c0[3] - c0[1]; d1 = c1[3] - c1[1]; break;
        case 4: d0 = c0[2]; d1 = c1[2]; break;
        default: d0 = c0[3] - c0[2]; d1 = c1[3] - c1[2]; break;
      }
      if (d0 < 0.f && d1 < 0.f) { reject = 1; break; }
      if (!(d0 == d0) || !(d1 == d1)) { reject = 1; break; }
      if (d0 < 0.f) { float t = d0 / (d0 - d1); if (t > t0) t0 = t; }
      else if (d1 < 0.f) { float t = d0 / (d0 - d1); if (t < t1) t1 = t; }
    }
    if (reject || !(t0 < t1)) continue;
    float e0[4], e1[4];
    for (int k = 0; k < 4; ++k) {
      float d = c1[k] - c0[k];
      e0[k] = fmaf(t0, d, c0[k]);
      e1[k] = fmaf(t1, d, c0[k]);
    }
    /* premultiplied colours at the clipped ends */
    float q0[4], q1[4];
    for (int k = 0; k < 4; ++k) {
      float a = col[8 * l + k], b = col[8 * l + 4 + k], d = b - a;
      q0[k] = fmaf(t0, d, a);
      q1[k] = fmaf(t1, d, a);
    }
    for (int k = 0; k < 3; ++k) { q0[k] = q0[k] * q0[3]; q1[k] = q1[k] * q1[3]; }
    float iw0 = 1.f / e0[3], iw1 = 1.f / e1[3];
    float sx0 = fmaf(e0[0] * iw0, hw, hw), sy0 = fmaf(e0[1] * iw0, hh, hh), z0 = e0[2] * iw0;
    float sx1 = fmaf(e1[0] * iw1, hw, hw), sy1 = fmaf(e1[1] * iw1, hh, hh), z1 = e1[2] * iw1;
    int xmajor = fabsf(sx1 - sx0) >= fabsf(sy1 - sy0);
    float a0 = xmajor ? sx0 : sy0, a1 = xmajor ? sx1 : sy1, m0 = xmajor ? sy0 : sx0, m1 = xmajor ? sy1 : sx1;
    if (a0 > a1) { /* walk from the smaller major coordinate */
      float t;
      t = a0; a0 = a1; a1 = t;
      t = m0; m0 = m1; m1 = t;
      t = z0; z0 = z1; z1 = t;
      for (int k = 0; k < 4; ++k) { t = q0[k]; q0[k] = q1[k]; q1[k] = t; }
    }
    if (!(a1 > a0)) continue; /* degenerate (or NaN) */
    float inv = 1.f / (a1 - a0), dm = m1 - m0, dz = z1 - z0;
    float lim_a = xmajor ? (float)W : (float)H;
    float fa = ceilf(a0 - 0.5f), fb = ceilf(a1 - 0.5f) - 1.f; /* centres c + 0.5 in [a0, a1) */
    if (fa < 0.f) fa = 0.f;
    if (fb > lim_a - 1.f) fb = lim_a - 1.f;
    if (!(fa <= fb)) continue;
    int ca = (int)fa, cb = (int)fb, lim_m = xmajor ? (int)H : (int)W;
    for (int c = ca; c <= cb; ++c) {
      float u = (((float)c + 0.5f) - a0) * inv;
      float mf = floorf(fmaf(u, dm, m0));
      if (!(mf >= 0.f && mf <= (float)(lim_m - 1))) continue;
      int m = (int)mf;
      float z = fmaf(u, dz, z0);
      if (!(z >= 0.f && z <= 1.f)) continue;
      uint32_t rgba = 0;
      for (int k = 0; k < 4; ++k) rgba |= (uint32_t)q8(fmaf(u, q1[k] - q0[k], q0[k])) << (8 * k);
      uint64_t packed = ((uint64_t)f2u(z) << 32) | rgba;
      size_t pix = xmajor ? (size_t)m * W + (size_t)c : (size_t)c * W + (size_t)m;
      if (packed < best[pix]) best[pix] = packed;
    }
  }
  for (size_t i = 0; i < (size_t)W * H; ++i)
    if (best[i] != ~0ull) {
      uint32_t zb = (uint32_t)(best[i] >> 32), rgba = (uint32_t)best[i];
      memcpy(&layer_depth[i], &zb, 4);
      /* premultiplied source over the clear colour (0,0,0,1), ONE / ONE_MINUS_SRC_ALPHA: rgb stays, alpha -> 1 */
      layer_rgba[4 * i + 0] = (uint8_t)rgba; layer_rgba[4 * i + 1] = (uint8_t)(rgba >> 8);
      layer_rgba[4 * i + 2] = (uint8_t)(rgba >> 16); layer_rgba[4 * i + 3] = 255;
    }
  free(best);
}

/* Rows outside [row0,row1) (rounded outwards to whole tile bands) are left untouched in out.
 * layer_depth / layer_rgba (both NULL, or both W*H): the opaque layer under the splats - the accumulators start from its
 * colour and a fragment is kept only if ndc.z < layer_depth (depth test LESS, no write). */
/* One tile band (rows [band * tile, band * tile + tile)) of the frame: every splat, back to front. */
static void raster_band(const raster_splat* S, uint32_t v, uint32_t W, uint32_t H, uint32_t tile, int mode, int band,
                        const float* layer_depth, const uint8_t* layer_rgba, uint8_t* out, float* fout) {
  {
    int by0 = band * (int)tile, by1 = by0 + (int)tile - 1;
    if (by1 > (int)H - 1) by1 = (int)H - 1;
    size_t npx = (size_t)W * (size_t)(by1 - by0 + 1);
    float* acc = (float*)malloc(npx * 4 * sizeof(float)); /* mode 0: [0,1] floats; mode 1: [0,255] integers */
    for (size_t k = 0; k < npx; ++k) {
      acc[4 * k + 0] = 0.f; acc[4 * k + 1] = 0.f; acc[4 * k + 2] = 0.f;
      acc[4 * k + 3] = mode == 1 ? 255.f : 1.f;
      if (layer_rgba) { /* the target holds UNORM8 values when the splats start */
        const uint8_t* lc = layer_rgba + 4 * ((size_t)by0 * W + k);
        for (int ch = 0; ch < 4; ++ch) acc[4 * k + ch] = mode == 1 ? (float)lc[ch] : (float)lc[ch] / 255.f;
      }
    }
    for (uint32_t i = 0; i < v; ++i) {
      const raster_splat* s = &S[i];
      if (s->x0 > s->x1) continue;
      int y0 = s->y0 > by0 ? s->y0 : by0, y1 = s->y1 < by1 ? s->y1 : by1;
      if (y0 > y1) continue;
      int ty = by0;
      float oy = (float)ty - s->cpy;
      float r255 = 255.f * s->r, g255 = 255.f * s->g, b255 = 255.f * s->b;
      for (int txi = s->x0 / (int)tile; txi <= s->x1 / (int)tile; ++txi) {
        int tx = txi * (int)tile;
        float ox = (float)tx - s->cpx;
        float bx = s->a00 * ox + s->a01 * oy, byy = s->a10 * ox + s->a11 * oy;
        int xa = tx > s->x0 ? tx : s->x0, xb = tx + (int)tile - 1 < s->x1 ? tx + (int)tile - 1 : s->x1;
        for (int yy = y0; yy <= y1; ++yy) {
          float ly = (float)(yy - ty);
          float ex = fmaf(s->a01, ly, bx), ey = fmaf(s->a11, ly, byy);
          float* row = acc + 4 * ((size_t)(yy - by0) * W);
          for (int xx = xa; xx <= xb; ++xx) {
            float lx = (float)(xx - tx);
            float px = fmaf(s->a00, lx, ex), py = fmaf(s->a10, lx, ey);
            if (!(fabsf(px) <= 3.f && fabsf(py) <= 3.f)) continue;
            if (layer_depth && !(s->z < layer_depth[(size_t)yy * W + xx])) continue; /* depth test LESS */
            float al = s->op * expf(-0.5f * fmaf(py, py, px * px));
            if (al > 1.f) al = 1.f;
            if (!(al >= 0.f)) al = 0.f;
            float om = 1.f - al;
            float* d = row + 4 * xx;
            if (mode == 1) {
              d[0] = rintf(fmaf(r255, al, d[0] * om));
              d[1] = rintf(fmaf(g255, al, d[1] * om));
              d[2] = rintf(fmaf(b255, al, d[2] * om));
              d[3] = rintf(fmaf(255.f * al, al, d[3] * om));
            } else {
              d[0] = s->r * al + d[0] * om;
              d[1] = s->g * al + d[1] * om;
              d[2] = s->b * al + d[2] * om;
              d[3] = al * al + d[3] * om;
            }
          }
        }
      }
    }
    for (int yy = by0; yy <= by1; ++yy)
      for (uint32_t xx = 0; xx < W; ++xx) {
        const float* d = acc + 4 * ((size_t)(yy - by0) * W + xx);
        uint8_t* o = out + 4 * ((size_t)yy * W + xx);
        if (mode == 1) {
          for (int k = 0; k < 4; ++k) { float q = d[k]; o[k] = (uint8_t)(q < 0.f ? 0.f : (q > 255.f ? 255.f : q)); }
        } else {
          for (int k = 0; k < 4; ++k) o[k] = q8(d[k]);
          if (fout) memcpy(fout + 4 * ((size_t)yy * W + xx), d, 16);
        }
      }
    free(acc);
  }
}

static raster_splat* raster_setup_all(uint32_t v, const float* inst, uint32_t W, uint32_t H) {
  raster_splat* S = (raster_splat*)malloc((size_t)(v ? v : 1) * sizeof(raster_splat));
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)v; ++i) raster_setup(inst + 12 * i, W, H, &S[i]);
  return S;
}

VKO_API void vko_raster_rows_layer(uint32_t v, const float* inst, uint32_t W, uint32_t H, uint32_t tile, int mode,
                                   uint32_t row0, uint32_t row1, const float* layer_depth, const uint8_t* layer_rgba,
                                   uint8_t* out, float* fout) {
  raster_splat* S = raster_setup_all(v, inst, W, H);
  int nband = (int)((H + tile - 1) / tile);
  int band_lo = (int)(row0 / tile), band_hi = (int)((row1 + tile - 1) / tile);
  if (band_hi > nband) band_hi = nband;
#pragma omp parallel for schedule(dynamic, 1)
  for (int band = band_lo; band < band_hi; ++band)
    raster_band(S, v, W, H, tile, mode, band, layer_depth, layer_rgba, out, fout);
  free(S);
}

/* A list of tile bands (indices in units of `tile` rows), one thread each: the stratified CPU sample of bench.py. */
VKO_API void vko_raster_band_list(uint32_t v, const float* inst, uint32_t W, uint32_t H, uint32_t tile, int mode,
                                  uint32_t nbands, const uint32_t* bands, uint8_t* out) {
  raster_splat* S = raster_setup_all(v, inst, W, H);
  int nband = (int)((H + tile - 1) / tile);
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < (int)nbands; ++b)
    if ((int)bands[b] < nband) raster_band(S, v, W, H, tile, mode, (int)bands[b], NULL, NULL, out, NULL);
  free(S);
}

VKO_API void vko_raster_rows(uint32_t v, const float* inst, uint32_t W, uint32_t H, uint32_t tile, int mode,
                             uint32_t row0, uint32_t row1, uint8_t* out, float* fout) {
  vko_raster_rows_layer(v, inst, W, H, tile, mode, row0, row1, NULL, NULL, out, fout);
}

VKO_API void vko_raster(uint32_t v, const float* inst, uint32_t W, uint32_t H, uint32_t tile, int mode,
                        uint8_t* out, float* fout) {
  vko_raster_rows(v, inst, W, H, tile, mode, 0, H, out, fout);
}

/* ------------------------------------------------------------ whole frame */

typedef struct {
  uint32_t visible;
  double ms_cull, ms_sort, ms_project, ms_raster;
} vko_frame_stats;

static double now_ms(void) {
#ifdef _OPENMP
  return omp_get_wtime() * 1e3;
#else
  return 0.0;
#endif
}

/* rank -> sort -> projection -> draw, engine.cc:1164-1290.  keys/ids/inst are caller buffers of
 * capacity n (n*12 floats for inst); any may be inspected afterwards. */
VKO_API void vko_render(uint32_t n, const float* pos, const float* cov, const float* opacity, const uint16_t* sh,
                        const vko_camera* cam, uint32_t tile, int mode, uint32_t* keys, uint32_t* ids, float* inst,
                        uint8_t* out, vko_frame_stats* st) {
  float pvm[16];
  double t0 = now_ms();
  vko_compose_pvm(cam->proj, cam->view, cam->model, pvm);
  uint32_t v = vko_cull(n, pos, pvm, keys, ids);
  double t1 = now_ms();
  vko_sort_pairs(v, keys, ids);
  double t2 = now_ms();
  vko_project(v, ids, pos, cov, opacity, sh, cam, 0, inst);
  double t3 = now_ms();
  vko_raster(v, inst, cam->width, cam->height, tile, mode, out, NULL);
  double t4 = now_ms();
  if (st) { st->visible = v; st->ms_cull = t1 - t0; st->ms_sort = t2 - t1; st->ms_project = t3 - t2; st->ms_raster = t4 - t3; }
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm asks for the host's cores explicitly */
VKO_API void vko_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

VKO_API int vko_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
