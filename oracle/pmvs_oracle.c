/*
 * oracle/pmvs_oracle.c -- TEST INFRASTRUCTURE (see pmvs_oracle.h).
 *
 * Plain-C restatement of the reference's patch-optimisation path.  Each function cites the reference
 * lines it follows (paths relative to /root/reference).  Float semantics are kept deliberately:
 * f32 operations in source order, no FMA (-ffp-contract=off), and f64 exactly where the reference's
 * unqualified sin/cos/asin/acos/log resolve to the double libm entry points.
 */
#include "pmvs_oracle.h"

#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "nm3.h"

#define MAXIMG_LOCAL 256

typedef struct {
  float P[8][3][4]; /* per level */
  float centre[4];
  float oaxis[4];
  float xaxis[3], yaxis[3], zaxis[3];
  float ipscale;
} cam_t;

struct pmvso_ctx {
  int num, tnum, level, csize, wsize, min_image_num, tau, nlevels;
  float ncc_threshold, ncc_threshold_before;
  float angle_threshold0, angle_threshold1, max_angle_threshold;
  double xtol, step;
  int maxeval;
  cam_t* cams;
  unsigned char** pix; /* [index*nlevels + level] */
  int* w;
  int* h;
  int** vis2;
  int* nvis2;
  /* filter-stage state */
  int depth;
  int P;
  float *s_coords, *s_normals, *s_ncc, *s_dscale;
  int *s_img_off, *s_images, *s_grids, *s_vimg_off, *s_vimages, *s_vgrids, *s_timages;
  int** dp;        /* [target image][cell] patch id or -1 */
  int* cell_off;   /* pgrids as CSR over (image, cell): cell_base[image] + cell */
  int* cell_base;
  int* cell_patch;
  int* vcell_off;   /* _vpgrids as CSR (patchOrganizerS.cpp:333-346) */
  int* vcell_patch;
};

/* ---------------------------------------------------------------- small vector helpers (f32, source order) */
static float dot4(const float* u, const float* v) { /* include/numeric/vec4.hpp:199-201 */
  return u[0] * v[0] + u[1] * v[1] + u[2] * v[2] + u[3] * v[3];
}
static float dot3(const float* u, const float* v) { /* include/numeric/vec3.hpp:191-193 */
  return u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
}
static void cross3(const float* u, const float* v, float* o) { /* vec3.hpp:195-202 */
  o[0] = u[1] * v[2] - v[1] * u[2];
  o[1] = -u[0] * v[2] + v[0] * u[2];
  o[2] = u[0] * v[1] - v[0] * u[1];
}
static void unitize3(float* v) { /* vec3.hpp:233-238 */
  const float l = dot3(v, v);
  if (l != 1.0 && l != 0.0) {
    const float s = sqrtf(l);
    v[0] /= s; v[1] /= s; v[2] /= s;
  }
}
static void unitize4(float* v) { /* vec4.hpp:252-257 */
  const float l = dot4(v, v);
  if (l != 1.0 && l != 0.0) {
    const float s = sqrtf(l);
    v[0] /= s; v[1] /= s; v[2] /= s; v[3] /= s;
  }
}
static float norm3(const float* v) { return sqrtf(dot3(v, v)); }
static float norm4(const float* v) { return sqrtf(dot4(v, v)); }
static float fminf_(float a, float b) { return b < a ? b : a; } /* std::min */
static float fmaxf_(float a, float b) { return a < b ? b : a; } /* std::max */

/* ---------------------------------------------------------------- context */
pmvso_ctx* pmvso_create(int num, int tnum, int level, int csize, int wsize, int min_image_num,
                        float threshold, float max_angle_deg) {
  pmvso_ctx* c = (pmvso_ctx*)calloc(1, sizeof(*c));
  c->num = num; c->tnum = tnum; c->level = level; c->csize = csize; c->wsize = wsize;
  c->min_image_num = min_image_num;
  c->tau = min_image_num * 2 < num ? min_image_num * 2 : num; /* source/pmvs/findMatch.cpp:56 */
  c->nlevels = level + 3;                                     /* findMatch.cpp:72 */
  c->ncc_threshold = threshold;
  c->ncc_threshold_before = threshold - 0.3f;                 /* findMatch.cpp:104 */
  c->angle_threshold0 = 60.0f * M_PI / 180.0f;               /* findMatch.cpp:92-93 */
  c->angle_threshold1 = 60.0f * M_PI / 180.0f;
  c->max_angle_threshold = max_angle_deg;                     /* source/pmvs/option.cpp:105-106 */
  c->max_angle_threshold *= M_PI / 180.0f;
  c->xtol = 1.0e-3; c->step = 1.0; c->maxeval = 1000;
  c->cams = (cam_t*)calloc(num, sizeof(cam_t));
  c->pix = (unsigned char**)calloc((size_t)num * c->nlevels, sizeof(unsigned char*));
  c->w = (int*)calloc((size_t)num * c->nlevels, sizeof(int));
  c->h = (int*)calloc((size_t)num * c->nlevels, sizeof(int));
  c->vis2 = (int**)calloc(num, sizeof(int*));
  c->nvis2 = (int*)calloc(num, sizeof(int));
  for (int i = 0; i < num; ++i) { /* option.cpp initVisdata: every other image */
    c->vis2[i] = (int*)malloc(sizeof(int) * (num > 1 ? num - 1 : 1));
    int k = 0;
    for (int j = 0; j < num; ++j)
      if (j != i) c->vis2[i][k++] = j;
    c->nvis2[i] = k;
  }
  return c;
}

void pmvso_destroy(pmvso_ctx* c) {
  if (!c) return;
  for (int i = 0; i < c->num * c->nlevels; ++i) free(c->pix[i]);
  for (int i = 0; i < c->num; ++i) free(c->vis2[i]);
  free(c->pix); free(c->w); free(c->h); free(c->vis2); free(c->nvis2); free(c->cams); free(c);
}

void pmvso_set_thresholds(pmvso_ctx* c, float ncc, float ncc_before) { c->ncc_threshold = ncc; c->ncc_threshold_before = ncc_before; }
void pmvso_set_xtol(pmvso_ctx* c, double xtol, double step, int maxeval) { c->xtol = xtol; c->step = step; c->maxeval = maxeval; }
void pmvso_set_visdata2(pmvso_ctx* c, int index, const int* list, int n) {
  free(c->vis2[index]);
  c->vis2[index] = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
  memcpy(c->vis2[index], list, sizeof(int) * n);
  c->nvis2[index] = n;
}

/* source/image/camera.cpp:56-68 (levels), 109-136 (updateCamera), 138-175 (getOpticalCenter),
 * source/pmvs/optim.cpp:43-64 (setAxesScales) */
void pmvso_set_camera(pmvso_ctx* c, int index, const float* P) {
  cam_t* cam = &c->cams[index];
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 4; ++k) cam->P[0][r][k] = P[4 * r + k];
  for (int l = 1; l < c->nlevels && l < 8; ++l) {
    memcpy(cam->P[l], cam->P[l - 1], sizeof(cam->P[0]));
    for (int k = 0; k < 4; ++k) { cam->P[l][0][k] /= 2.0f; cam->P[l][1][k] /= 2.0f; }
  }
  /* optical axis */
  float oa[4] = {cam->P[0][2][0], cam->P[0][2][1], cam->P[0][2][2], 0.0f};
  const float ftmp = norm4(oa);
  oa[3] = cam->P[0][2][3];
  for (int k = 0; k < 4; ++k) cam->oaxis[k] = oa[k] / ftmp;
  /* optical centre: -A^-1 b in double via adjoint (include/numeric/mat3.hpp:275-292) */
  if (cam->P[0][2][0] == 0.0 && cam->P[0][2][1] == 0.0 && cam->P[0][2][2] == 0.0) {
    float v0[3] = {cam->P[0][0][0], cam->P[0][0][1], cam->P[0][0][2]};
    float v1[3] = {cam->P[0][1][0], cam->P[0][1][1], cam->P[0][1][2]};
    float v2[3];
    cross3(v0, v1, v2);
    unitize3(v2);
    cam->centre[0] = v2[0]; cam->centre[1] = v2[1]; cam->centre[2] = v2[2]; cam->centre[3] = 0.f;
  } else {
    double A[3][3], adj[3][3], b[3], inv[3][3];
    for (int y = 0; y < 3; ++y) {
      for (int x = 0; x < 3; ++x) A[y][x] = cam->P[0][y][x];
      b[y] = -cam->P[0][y][3];
    }
#define CROSSD(u, v, o)                      \
  do {                                       \
    (o)[0] = (u)[1] * (v)[2] - (v)[1] * (u)[2]; \
    (o)[1] = -(u)[0] * (v)[2] + (v)[0] * (u)[2]; \
    (o)[2] = (u)[0] * (v)[1] - (v)[0] * (u)[1]; \
  } while (0)
    CROSSD(A[1], A[2], adj[0]);
    CROSSD(A[2], A[0], adj[1]);
    CROSSD(A[0], A[1], adj[2]);
    const double d = adj[0][0] * A[0][0] + adj[0][1] * A[0][1] + adj[0][2] * A[0][2];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) inv[i][j] = adj[j][i] / d;
    for (int y = 0; y < 3; ++y)
      cam->centre[y] = (float)(inv[y][0] * b[0] + inv[y][1] * b[1] + inv[y][2] * b[2]);
    cam->centre[3] = 1.f;
  }
  /* COptim axes (optim.cpp:47-53) */
  cam->zaxis[0] = cam->oaxis[0]; cam->zaxis[1] = cam->oaxis[1]; cam->zaxis[2] = cam->oaxis[2];
  float xa[3] = {cam->P[0][0][0], cam->P[0][0][1], cam->P[0][0][2]};
  cross3(cam->zaxis, xa, cam->yaxis);
  unitize3(cam->yaxis);
  cross3(cam->yaxis, cam->zaxis, cam->xaxis);
  /* ipscale (optim.cpp:56-63) */
  const float xe[4] = {cam->xaxis[0], cam->xaxis[1], cam->xaxis[2], 0.0f};
  const float ye[4] = {cam->yaxis[0], cam->yaxis[1], cam->yaxis[2], 0.0f};
  const float fx = dot4(xe, cam->P[0][0]);
  const float fy = dot4(ye, cam->P[0][1]);
  cam->ipscale = fx + fy;
}

/* source/image/image.cpp:136-139 (sizes), 228-325 (buildImage, filter 0).
 * Weights are k/64 in double; every partial sum is an exact dyadic rational, so integer arithmetic
 * reproduces (unsigned char)(int)floor(sum/denom + 0.5f) exactly: floor((2S + D) / (2D)). */
void pmvso_set_image(pmvso_ctx* c, int index, int w, int h, const unsigned char* rgb) {
  static const int wt[4] = {1, 3, 3, 1};
  const int base = index * c->nlevels;
  c->w[base] = w; c->h[base] = h;
  free(c->pix[base]);
  c->pix[base] = (unsigned char*)malloc((size_t)w * h * 3);
  memcpy(c->pix[base], rgb, (size_t)w * h * 3);
  for (int l = 1; l < c->nlevels; ++l) {
    const int pw = c->w[base + l - 1], ph = c->h[base + l - 1];
    const int nw = pw / 2, nh = ph / 2;
    c->w[base + l] = nw; c->h[base + l] = nh;
    free(c->pix[base + l]);
    c->pix[base + l] = (unsigned char*)malloc((size_t)(nw > 0 ? nw : 1) * (nh > 0 ? nh : 1) * 3);
    const unsigned char* src = c->pix[base + l - 1];
    unsigned char* dst = c->pix[base + l];
    for (int y = 0; y < nh; ++y)
      for (int x = 0; x < nw; ++x) {
        long S[3] = {0, 0, 0};
        long D = 0;
        for (int j = -1; j < 3; ++j) {
          const int yt = 2 * y + j;
          if (yt < 0 || ph - 1 < yt) continue;
          for (int i = -1; i < 3; ++i) {
            const int xt = 2 * x + i;
            if (xt < 0 || pw - 1 < xt) continue;
            const int k = wt[j + 1] * wt[i + 1];
            const unsigned char* p = src + ((size_t)yt * pw + xt) * 3;
            S[0] += k * p[0]; S[1] += k * p[1]; S[2] += k * p[2];
            D += k;
          }
        }
        for (int k = 0; k < 3; ++k) dst[((size_t)y * nw + x) * 3 + k] = (unsigned char)((2 * S[k] + D) / (2 * D));
      }
  }
}

void pmvso_image_dims(const pmvso_ctx* c, int index, int level, int* w, int* h) {
  *w = c->w[index * c->nlevels + level]; *h = c->h[index * c->nlevels + level];
}
void pmvso_image_bytes(const pmvso_ctx* c, int index, int level, unsigned char* out) {
  const int k = index * c->nlevels + level;
  memcpy(out, c->pix[k], (size_t)c->w[k] * c->h[k] * 3);
}
void pmvso_camera(const pmvso_ctx* c, int index, int level, float* P, float* centre, float* oaxis,
                  float* xaxis, float* yaxis, float* zaxis, float* ipscale) {
  const cam_t* cam = &c->cams[index];
  memcpy(P, cam->P[level], sizeof(float) * 12);
  memcpy(centre, cam->centre, sizeof(float) * 4);
  memcpy(oaxis, cam->oaxis, sizeof(float) * 4);
  memcpy(xaxis, cam->xaxis, sizeof(float) * 3);
  memcpy(yaxis, cam->yaxis, sizeof(float) * 3);
  memcpy(zaxis, cam->zaxis, sizeof(float) * 3);
  *ipscale = cam->ipscale;
}

/* ---------------------------------------------------------------- geometry primitives */
/* include/image/camera.hpp:89-108 */
static void project(const pmvso_ctx* c, int index, const float* X, int level, float* o) {
  const cam_t* cam = &c->cams[index];
  for (int i = 0; i < 3; ++i) o[i] = dot4(cam->P[level][i], X);
  if (o[2] <= 0.0) {
    o[0] = -0xffff; o[1] = -0xffff; o[2] = -1.0f;
    return;
  }
  const float z = o[2];
  o[0] /= z; o[1] /= z; o[2] /= z;
  o[0] = fmaxf_((float)(INT_MIN + 3.0f), fminf_((float)(INT_MAX - 3.0f), o[0]));
  o[1] = fmaxf_((float)(INT_MIN + 3.0f), fminf_((float)(INT_MAX - 3.0f), o[1]));
}
void pmvso_project(const pmvso_ctx* c, int index, const float* coord, int level, float* out3) { project(c, index, coord, level, out3); }

/* source/pmvs/optim.cpp:1116-1124 */
static float get_unit(const pmvso_ctx* c, int index, const float* X) {
  const cam_t* cam = &c->cams[index];
  const float d[4] = {X[0] - cam->centre[0], X[1] - cam->centre[1], X[2] - cam->centre[2], X[3] - cam->centre[3]};
  const float fz = norm4(d);
  const float ftmp = cam->ipscale;
  if (ftmp == 0.0) return 1.0;
  return (float)(2.0 * fz * (0x0001 << c->level) / ftmp);
}
float pmvso_get_unit(const pmvso_ctx* c, int index, const float* coord) { return get_unit(c, index, coord); }

/* include/image/image.hpp:435-476 (bilinear branch) */
static void get_color(const pmvso_ctx* c, int index, float x, float y, int level, float* rgb) {
  const int k = index * c->nlevels + level;
  const int W = c->w[k];
  const unsigned char* im = c->pix[k];
  const int lx = (int)x;
  const int ly = (int)y;
  const int idx = 3 * (ly * W + lx);
  const float dx1 = x - lx; const float dx0 = 1.0f - dx1;
  const float dy1 = y - ly; const float dy0 = 1.0f - dy1;
  const float f00 = dx0 * dy0; const float f01 = dx0 * dy1;
  const float f10 = dx1 * dy0; const float f11 = dx1 * dy1;
  const int idx2 = idx + 3 * W;
  const unsigned char* p0 = im + idx;
  const unsigned char* p1 = im + idx2;
  for (int ch = 0; ch < 3; ++ch) {
    float r = 0.0f;
    r += p0[ch] * f00 + p1[ch] * f01;
    r += p0[ch + 3] * f10 + p1[ch + 3] * f11;
    rgb[ch] = r;
  }
}
void pmvso_get_color(const pmvso_ctx* c, int index, float x, float y, int level, float* rgb) { get_color(c, index, x, y, level, rgb); }

/* optim.cpp:1127-1144 */
static void get_paxes(const pmvso_ctx* c, int index, const float* coord, const float* normal, float* px, float* py) {
  const cam_t* cam = &c->cams[index];
  const float pscale = get_unit(c, index, coord);
  const float n3[3] = {normal[0], normal[1], normal[2]};
  float y3[3], x3[3];
  cross3(n3, cam->xaxis, y3);
  unitize3(y3);
  cross3(y3, n3, x3);
  px[0] = x3[0]; px[1] = x3[1]; px[2] = x3[2]; px[3] = 0.0f;
  py[0] = y3[0]; py[1] = y3[1]; py[2] = y3[2]; py[3] = 0.0f;
  for (int k = 0; k < 4; ++k) { px[k] *= pscale; py[k] *= pscale; }
  float c0[3], c1[3], d[3], t[4];
  project(c, index, coord, c->level, c0);
  for (int k = 0; k < 4; ++k) t[k] = coord[k] + px[k];
  project(c, index, t, c->level, c1);
  for (int k = 0; k < 3; ++k) d[k] = c1[k] - c0[k];
  const float xdis = norm3(d);
  for (int k = 0; k < 4; ++k) t[k] = coord[k] + py[k];
  project(c, index, t, c->level, c1);
  for (int k = 0; k < 3; ++k) d[k] = c1[k] - c0[k];
  const float ydis = norm3(d);
  for (int k = 0; k < 4; ++k) { px[k] /= xdis; py[k] /= ydis; }
}
void pmvso_get_paxes(const pmvso_ctx* c, int index, const float* coord, const float* normal, float* px, float* py) { get_paxes(c, index, coord, normal, px, py); }

/* optim.cpp:808-811 */
static float my_pow2(int x) {
  static const float answers[] = {0.0625, 0.125, 0.25, 0.5, 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024};
  return answers[x + 4];
}

/* optim.cpp:783-805 */
static int grab_safe(const pmvso_ctx* c, int index, int size, const float* center, const float* dx, const float* dy, int level) {
  const int margin = size / 2;
  float tl[2], tr[2], bl[2], br[2];
  for (int k = 0; k < 2; ++k) {
    tl[k] = center[k] - dx[k] * margin - dy[k] * margin;
    tr[k] = center[k] + dx[k] * margin - dy[k] * margin;
    bl[k] = center[k] - dx[k] * margin + dy[k] * margin;
    br[k] = center[k] + dx[k] * margin + dy[k] * margin;
  }
  const float minx = fminf_(tl[0], fminf_(tr[0], fminf_(bl[0], br[0])));
  const float maxx = fmaxf_(tl[0], fmaxf_(tr[0], fmaxf_(bl[0], br[0])));
  const float miny = fminf_(tl[1], fminf_(tr[1], fminf_(bl[1], br[1])));
  const float maxy = fmaxf_(tl[1], fmaxf_(tr[1], fmaxf_(bl[1], br[1])));
  const int margin2 = 3;
  const int k = index * c->nlevels + level;
  if (minx < margin2 || c->w[k] - 1 - margin2 <= maxx || miny < margin2 || c->h[k] - 1 - margin2 <= maxy) return 0;
  return 1;
}

/* optim.cpp:815-863.  Returns 1 = rejected (tex untouched), 0 = ok. */
static int grab_tex(const pmvso_ctx* c, const float* coord, const float* pxaxis, const float* pyaxis,
                    const float* pzaxis, int index, int size, float* tex, int* newlevel_out) {
  const cam_t* cam = &c->cams[index];
  if (newlevel_out) *newlevel_out = -1;
  float ray[4] = {cam->centre[0] - coord[0], cam->centre[1] - coord[1], cam->centre[2] - coord[2], cam->centre[3] - coord[3]};
  unitize4(ray);
  const float weight = fmaxf_(0.0f, dot4(ray, pzaxis));
  if (weight < cos(c->angle_threshold1)) return 1;
  const int margin = size / 2;
  float center[3], dx[3], dy[3], t[4], q[3];
  project(c, index, coord, c->level, center);
  for (int k = 0; k < 4; ++k) t[k] = coord[k] + pxaxis[k];
  project(c, index, t, c->level, q);
  for (int k = 0; k < 3; ++k) dx[k] = q[k] - center[k];
  for (int k = 0; k < 4; ++k) t[k] = coord[k] + pyaxis[k];
  project(c, index, t, c->level, q);
  for (int k = 0; k < 3; ++k) dy[k] = q[k] - center[k];
  const float ratio = (norm3(dx) + norm3(dy)) / 2.0f;
  static const float Log2 = 0.693147180559945309417f; /* static float Log2 = log(2.0f); optim.cpp:813 */
  const double lv = floor(log(ratio) / Log2 + 0.5f);
  int leveldif;
  if (!(lv > -1.0e9)) leveldif = INT_MIN; /* (int) of -inf / NaN: cvttsd2si yields INT_MIN */
  else if (lv > 1.0e9) leveldif = INT_MIN;
  else leveldif = (int)lv;
  leveldif = leveldif < 2 ? leveldif : 2;                    /* std::min(2, leveldif) */
  leveldif = -c->level < leveldif ? leveldif : -c->level;    /* std::max(-level, .) */
  const float scale = my_pow2(leveldif);
  const int newlevel = c->level + leveldif;
  for (int k = 0; k < 3; ++k) { center[k] /= scale; dx[k] /= scale; dy[k] /= scale; }
  if (grab_safe(c, index, size, center, dx, dy, newlevel) == 0) return 1;
  if (newlevel_out) *newlevel_out = newlevel;
  float left[3];
  for (int k = 0; k < 3; ++k) left[k] = center[k] - dx[k] * margin - dy[k] * margin;
  float* tp = tex;
  for (int y = 0; y < size; ++y) {
    float v[3] = {left[0], left[1], left[2]};
    for (int k = 0; k < 3; ++k) left[k] += dy[k];
    for (int x = 0; x < size; ++x) {
      get_color(c, index, v[0], v[1], newlevel, tp);
      tp += 3;
      for (int k = 0; k < 3; ++k) v[k] += dx[k];
    }
  }
  return 0;
}

int pmvso_grab_tex(const pmvso_ctx* c, const float* coord, const float* normal, int ref, int index, float* tex, int* newlevel) {
  float px[4], py[4];
  get_paxes(c, ref, coord, normal, px, py);
  return grab_tex(c, coord, px, py, normal, index, c->wsize, tex, newlevel);
}

/* optim.cpp:1031-1067 */
static void normalize_tex(float* tex, int size) {
  const int size3 = size / 3;
  float ave[3] = {0.f, 0.f, 0.f};
  for (int i = 0; i < size3; ++i) { ave[0] += tex[3 * i]; ave[1] += tex[3 * i + 1]; ave[2] += tex[3 * i + 2]; }
  ave[0] /= (float)size3; ave[1] /= (float)size3; ave[2] /= (float)size3;
  float ave2 = 0.0f;
  for (int i = 0; i < size3; ++i) {
    const float f0 = ave[0] - tex[3 * i];
    const float f1 = ave[1] - tex[3 * i + 1];
    const float f2 = ave[2] - tex[3 * i + 2];
    ave2 += f0 * f0 + f1 * f1 + f2 * f2;
  }
  ave2 = sqrtf(ave2 / size);
  if (ave2 == 0.0f) ave2 = 1.0f;
  for (int i = 0; i < size3; ++i)
    for (int k = 0; k < 3; ++k) { tex[3 * i + k] -= ave[k]; tex[3 * i + k] /= ave2; }
}
void pmvso_normalize(float* tex, int n) { normalize_tex(tex, n); }

/* optim.cpp:1069-1077 */
static float dot_tex(const float* a, const float* b, int size) {
  float ans = 0.0f;
  for (int i = 0; i < size; ++i) ans += a[i] * b[i];
  return ans / size;
}
float pmvso_dot(const float* a, const float* b, int n) { return dot_tex(a, b, n); }

/* include/pmvs/optim.hpp:86-92 */
static float robustincc(const float rhs) { return rhs / (1 + 3 * rhs); }
static float unrobustincc(const float rhs) { return rhs / (1 - 3 * rhs); }

/* ---------------------------------------------------------------- per-refinement context (optim.cpp:584-596) */
typedef struct {
  const pmvso_ctx* c;
  float centre[4], ray[4];
  const int* images;
  int n;
  float dscale, ascale;
  float weights[MAXIMG_LOCAL];
  float* texs; /* tau * tsize floats */
  unsigned char valid[MAXIMG_LOCAL];
} rctx_t;

/* optim.cpp:446-471 + 1146-1152 */
static void set_weights(const pmvso_ctx* c, const float* coord, const float* normal, const int* images, int n, float* w) {
  for (int i = 0; i < n; ++i) {
    const cam_t* cam = &c->cams[images[i]];
    float u = get_unit(c, images[i], coord);
    float ray[4] = {cam->centre[0] - coord[0], cam->centre[1] - coord[1], cam->centre[2] - coord[2], cam->centre[3] - coord[3]};
    unitize4(ray);
    const float denom = dot4(ray, normal);
    if (0.0 < denom) u /= denom; else u = INT_MAX / 2;
    w[i] = u;
  }
  for (int i = 1; i < n; ++i) w[i] = fminf_(1.0f, w[0] / w[i]);
  if (n > 0) w[0] = 1.0f;
}

static void rctx_init(rctx_t* r, const pmvso_ctx* c, const float* coord, const float* normal, const int* images, int n, float dscale) {
  r->c = c;
  memcpy(r->centre, coord, sizeof(float) * 4);
  const cam_t* cam = &c->cams[images[0]];
  for (int k = 0; k < 4; ++k) r->ray[k] = coord[k] - cam->centre[k];
  unitize4(r->ray);
  r->images = images; r->n = n;
  r->dscale = dscale;
  r->ascale = (float)(M_PI / 48.0f);
  set_weights(c, coord, normal, images, n < MAXIMG_LOCAL ? n : MAXIMG_LOCAL, r->weights);
}

/* optim.cpp:690-707 */
static void decode(const rctx_t* r, const double* x, float* coord, float* normal) {
  const cam_t* cam = &r->c->cams[r->images[0]];
  const double s = r->dscale * x[0];
  for (int k = 0; k < 4; ++k) coord[k] = r->centre[k] + (float)(r->ray[k] * s);
  const float angle1 = (float)(x[1] * r->ascale);
  const float angle2 = (float)(x[2] * r->ascale);
  const float fx = (float)(sin(angle1) * cos(angle2));
  const float fy = (float)sin(angle2);
  const float fz = (float)(-cos(angle1) * cos(angle2));
  for (int k = 0; k < 3; ++k) normal[k] = cam->xaxis[k] * fx + cam->yaxis[k] * fy + cam->zaxis[k] * fz;
  normal[3] = 0.0f;
}

/* optim.cpp:660-688 */
static void encode(const rctx_t* r, const float* coord, const float* normal, double* x) {
  const cam_t* cam = &r->c->cams[r->images[0]];
  const float d[4] = {coord[0] - r->centre[0], coord[1] - r->centre[1], coord[2] - r->centre[2], coord[3] - r->centre[3]};
  x[0] = dot4(d, r->ray) / r->dscale;
  float n3[3] = {normal[0], normal[1], normal[2]};
  if (normal[3] != 1.0 && normal[3] != 0.0) { n3[0] /= normal[3]; n3[1] /= normal[3]; n3[2] /= normal[3]; }
  const float fx = dot3(cam->xaxis, n3);
  const float fy = dot3(cam->yaxis, n3);
  const float fz = dot3(cam->zaxis, n3);
  x[2] = asin(fmaxf_(-1.0f, fminf_(1.0f, fy)));
  const float cosb = (float)cos(x[2]);
  if (cosb == 0.0) {
    x[1] = 0.0;
  } else {
    const float sina = fx / cosb;
    const float cosa = -fz / cosb;
    x[1] = acos(fmaxf_(-1.0f, fminf_(1.0f, cosa)));
    if (sina < 0.0) x[1] = -x[1];
  }
  x[1] = x[1] / r->ascale;
  x[2] = x[2] / r->ascale;
}

/* grab + normalise textures of the first `size` images; valid[i] = texture present */
static void grab_all(const pmvso_ctx* c, const float* coord, const float* normal, const float* px, const float* py,
                     const int* images, int size, float* texs, unsigned char* valid) {
  const int tsize = 3 * c->wsize * c->wsize;
  for (int i = 0; i < size; ++i) {
    const int flag = grab_tex(c, coord, px, py, normal, images[i], c->wsize, texs + (size_t)i * tsize, NULL);
    valid[i] = (flag == 0);
    if (flag == 0) normalize_tex(texs + (size_t)i * tsize, tsize);
  }
}

/* optim.cpp:507-578 (pairwise == 0 branch) */
static double my_f(unsigned nn, const double* x, void* data) {
  rctx_t* r = (rctx_t*)data;
  const pmvso_ctx* c = r->c;
  (void)nn;
  float coord[4], normal[4], px[4], py[4];
  decode(r, x, coord, normal);
  get_paxes(c, r->images[0], coord, normal, px, py);
  const int size = c->tau < r->n ? c->tau : r->n;
  const int mininum = c->min_image_num < size ? c->min_image_num : size;
  const int tsize = 3 * c->wsize * c->wsize;
  grab_all(c, coord, normal, px, py, r->images, size, r->texs, r->valid);
  if (!r->valid[0]) return 2.0;
  double ans = 0.0f;
  int denom = 0;
  for (int i = 1; i < size; ++i) {
    if (!r->valid[i]) continue;
    ans += robustincc((float)(1.0 - dot_tex(r->texs, r->texs + (size_t)i * tsize, tsize)));
    denom++;
  }
  if (denom < mininum - 1) return 2.0f;
  return ans / denom + 0.0;
}

/* optim.cpp:865-938 (non-PAIRNCC branch) */
static double compute_incc(rctx_t* r, const float* coord, const float* normal, int robust) {
  const pmvso_ctx* c = r->c;
  if (r->n < 2) return 2.0;
  float px[4], py[4];
  get_paxes(c, r->images[0], coord, normal, px, py);
  const int size = c->tau < r->n ? c->tau : r->n;
  const int tsize = 3 * c->wsize * c->wsize;
  grab_all(c, coord, normal, px, py, r->images, size, r->texs, r->valid);
  if (!r->valid[0]) return 2.0;
  double score = 0.0;
  float totalweight = 0.0;
  for (int i = 1; i < size; ++i) {
    if (!r->valid[i]) continue;
    totalweight += r->weights[i];
    const float d = dot_tex(r->texs, r->texs + (size_t)i * tsize, tsize);
    if (robust) score += robustincc((float)(1.0 - d)) * r->weights[i];
    else score += (1.0 - d) * r->weights[i];
  }
  if (totalweight == 0.0) score = 2.0; else score /= totalweight;
  return score;
}

static float* alloc_texs(const pmvso_ctx* c, int n) {
  return (float*)malloc(sizeof(float) * 3 * c->wsize * c->wsize * (size_t)(n > 1 ? n : 1));
}

void pmvso_encode(const pmvso_ctx* c, const float* coord, const float* normal, const int* images, int n, float dscale, double* x) {
  rctx_t r; rctx_init(&r, c, coord, normal, images, n, dscale);
  encode(&r, coord, normal, x);
}
void pmvso_decode(const pmvso_ctx* c, const float* coord, const float* normal, const int* images, int n, float dscale,
                  const double* x, float* ocoord, float* onormal) {
  rctx_t r; rctx_init(&r, c, coord, normal, images, n, dscale);
  decode(&r, x, ocoord, onormal);
}
double pmvso_my_f(const pmvso_ctx* c, const float* coord, const float* normal, const int* images, int n, float dscale, const double* x) {
  rctx_t r; rctx_init(&r, c, coord, normal, images, n, dscale);
  r.texs = alloc_texs(c, c->tau);
  const double f = my_f(3, x, &r);
  free(r.texs);
  return f;
}
double pmvso_compute_incc(const pmvso_ctx* c, const float* coord, const float* normal, const int* images, int n, int robust) {
  rctx_t r; rctx_init(&r, c, coord, normal, images, n, 1.0f);
  r.texs = alloc_texs(c, c->tau);
  const double f = compute_incc(&r, coord, normal, robust);
  free(r.texs);
  return f;
}

/* optim.cpp:580-658 with nm3 in place of nlopt.  texs scratch supplied by caller. */
static int refine_one(const pmvso_ctx* c, float* coord, float* normal, const int* images, int n, float dscale,
                      float* ncc, int* evals, float* texs) {
  rctx_t r; rctx_init(&r, c, coord, normal, images, n, dscale);
  r.texs = texs;
  double p[3];
  encode(&r, coord, normal, p);
  const double lb[3] = {-HUGE_VAL, -23.99999, -23.99999};
  const double ub[3] = {HUGE_VAL, 23.99999, 23.99999};
  double x[3];
  for (int i = 0; i < 3; ++i) x[i] = fmax(fmin(p[i], ub[i]), lb[i]);
  double minf;
  int nev = 0;
  const int res = nm3_minimize(3, my_f, &r, lb, ub, x, &minf, c->step, c->xtol, c->maxeval, &nev);
  if (evals) *evals = nev;
  if (res != NM3_XTOL_REACHED) return 0;
  decode(&r, x, coord, normal);
  *ncc = (float)(1.0 - unrobustincc((float)compute_incc(&r, coord, normal, 1)));
  return 1;
}

int pmvso_refine(const pmvso_ctx* c, float* coord, float* normal, const int* images, int n, float dscale, float* ncc, int* evals) {
  float* texs = alloc_texs(c, c->tau);
  const int ok = refine_one(c, coord, normal, images, n, dscale, ncc, evals, texs);
  free(texs);
  return ok;
}

typedef struct {
  const pmvso_ctx* c; int P, V; float* coords; float* normals; const int* images; const float* dscales;
  float* nccs; int* evals; unsigned char* ok; int* next; pthread_mutex_t* mu;
} batch_t;

static void* batch_worker(void* arg) {
  batch_t* b = (batch_t*)arg;
  float* texs = alloc_texs(b->c, b->c->tau);
  for (;;) {
    pthread_mutex_lock(b->mu);
    const int s = *b->next; *b->next += 64;
    pthread_mutex_unlock(b->mu);
    if (s >= b->P) break;
    const int e = s + 64 < b->P ? s + 64 : b->P;
    for (int p = s; p < e; ++p) {
      int ev = 0;
      float ncc = -1.0f;
      b->ok[p] = (unsigned char)refine_one(b->c, b->coords + 4 * p, b->normals + 4 * p, b->images + (size_t)b->V * p, b->V,
                                           b->dscales[p], &ncc, &ev, texs);
      b->nccs[p] = ncc; b->evals[p] = ev;
    }
  }
  free(texs);
  return NULL;
}

double pmvso_refine_batch(const pmvso_ctx* c, int P, int V, float* coords, float* normals, const int* images,
                          const float* dscales, float* nccs, int* evals, unsigned char* ok, int threads) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  int next = 0;
  batch_t b = {c, P, V, coords, normals, images, dscales, nccs, evals, ok, &next, &mu};
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  pthread_t th[256];
  for (int i = 1; i < threads; ++i) pthread_create(&th[i], NULL, batch_worker, &b);
  batch_worker(&b);
  for (int i = 1; i < threads; ++i) pthread_join(th[i], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ---------------------------------------------------------------- setINCCs (optim.cpp:709-744, 746-781) */
void pmvso_set_inccs(const pmvso_ctx* c, const float* coord, const float* normal, const int* images, int n, int robust, float* out) {
  float px[4], py[4];
  get_paxes(c, images[0], coord, normal, px, py);
  const int tsize = 3 * c->wsize * c->wsize;
  float* texs = alloc_texs(c, n);
  unsigned char* valid = (unsigned char*)malloc(n > 0 ? n : 1);
  grab_all(c, coord, normal, px, py, images, n, texs, valid);
  if (!valid[0]) {
    for (int i = 0; i < n; ++i) out[i] = 2.0f;
  } else {
    for (int i = 0; i < n; ++i) {
      if (i == 0) out[i] = 0.0f;
      else if (valid[i]) {
        const float d = dot_tex(texs, texs + (size_t)i * tsize, tsize);
        out[i] = robust == 0 ? 1.0f - d : robustincc(1.0f - d);
      } else out[i] = 2.0f;
    }
  }
  free(texs); free(valid);
}

void pmvso_set_inccs_matrix(const pmvso_ctx* c, const float* coord, const float* normal, const int* images, int n, int robust, float* out) {
  float px[4], py[4];
  get_paxes(c, images[0], coord, normal, px, py);
  const int tsize = 3 * c->wsize * c->wsize;
  float* texs = alloc_texs(c, n);
  unsigned char* valid = (unsigned char*)malloc(n > 0 ? n : 1);
  grab_all(c, coord, normal, px, py, images, n, texs, valid);
  for (int i = 0; i < n; ++i) {
    out[i * n + i] = 0.0f;
    for (int j = i + 1; j < n; ++j) {
      float v;
      if (valid[i] && valid[j]) {
        const float d = dot_tex(texs + (size_t)i * tsize, texs + (size_t)j * tsize, tsize);
        v = robust == 0 ? 1.0f - d : robustincc(1.0f - d);
      } else v = 2.0f;
      out[j * n + i] = out[i * n + j] = v;
    }
  }
  free(texs); free(valid);
}

/* ---------------------------------------------------------------- setScales (patchOrganizerS.cpp:663-684) */
static void set_scales(const pmvso_ctx* c, const float* coord, const int* images, int n, float* dscale_io, float* ascale) {
  const cam_t* cam = &c->cams[images[0]];
  const float unit = get_unit(c, images[0], coord);
  const float unit2 = 2.0f * unit;
  float ray[4];
  for (int k = 0; k < 4; ++k) ray[k] = coord[k] - cam->centre[k];
  unitize4(ray);
  const int inum = c->tau < n ? c->tau : n;
  float ds = *dscale_io;
  for (int i = 1; i < inum; ++i) {
    float a[3], b[3], t[4], d[3];
    project(c, images[i], coord, c->level, a);
    for (int k = 0; k < 4; ++k) t[k] = coord[k] - ray[k] * unit2;
    project(c, images[i], t, c->level, b);
    for (int k = 0; k < 3; ++k) d[k] = a[k] - b[k];
    ds += norm3(d);
  }
  ds /= inum - 1;
  ds = unit2 / ds;
  *dscale_io = ds;
  *ascale = (float)atan(ds / (unit * c->wsize / 2.0f));
}
void pmvso_set_scales(const pmvso_ctx* c, const float* coord, const int* images, int n, float* dscale, float* ascale) {
  *dscale = 0.0f; /* CPatch() zero-initialises _dscale (include/pmvs/patch.hpp:24) */
  set_scales(c, coord, images, n, dscale, ascale);
}

/* ---------------------------------------------------------------- image-set selection */
/* optim.cpp:398-444.  No edge images in scope (getEdge == 1 when _edges is empty, photo.hpp:52-53). */
static int add_images(const pmvso_ctx* c, const float* coord, const float* normal, int* images, int n, int cap) {
  unsigned char used[MAXIMG_LOCAL];
  memset(used, 0, sizeof(used));
  for (int i = 0; i < n; ++i) used[images[i]] = 1;
  const int ref = images[0];
  const float athreshold = (float)cos(c->angle_threshold0); /* optim.cpp:416 narrows to float */
  for (int k = 0; k < c->nvis2[ref]; ++k) {
    const int im = c->vis2[ref][k];
    if (used[im]) continue;
    float ic[3];
    project(c, im, coord, c->level, ic);
    const int W = c->w[im * c->nlevels + c->level], H = c->h[im * c->nlevels + c->level];
    if (ic[0] < 0.0f || W - 1 <= ic[0] || ic[1] < 0.0f || H - 1 <= ic[1]) continue;
    const cam_t* cam = &c->cams[im];
    float ray[4] = {cam->centre[0] - coord[0], cam->centre[1] - coord[1], cam->centre[2] - coord[2], cam->centre[3] - coord[3]};
    unitize4(ray);
    const float ftmp = dot4(ray, normal);
    if (athreshold <= ftmp && n < cap) images[n++] = im;
  }
  return n;
}

/* optim.cpp:192-206 */
static int constraint_images(const pmvso_ctx* c, const float* coord, const float* normal, int* images, int n, float ncc_threshold) {
  float* inccs = (float*)malloc(sizeof(float) * (n > 0 ? n : 1));
  pmvso_set_inccs(c, coord, normal, images, n, 0, inccs);
  int m = 1;
  for (int i = 1; i < n; ++i)
    if (inccs[i] < 1.0f - ncc_threshold) images[m++] = images[i];
  free(inccs);
  return m;
}

/* optim.cpp:284-321 (newm == 1) with computeUnits 473-494 */
static int sort_images(const pmvso_ctx* c, const float* coord, const float* normal, int* images, int n) {
  const float threshold = (float)(1.0f - cos(10.0 * M_PI / 180.0));
  int idx[MAXIMG_LOCAL]; float units[MAXIMG_LOCAL]; float rays[MAXIMG_LOCAL][4];
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const cam_t* cam = &c->cams[images[i]];
    float ray[4] = {cam->centre[0] - coord[0], cam->centre[1] - coord[1], cam->centre[2] - coord[2], cam->centre[3] - coord[3]};
    unitize4(ray);
    const float d = dot4(ray, normal);
    if (d <= 0.0f) continue;
    const float scale = get_unit(c, images[i], coord);
    idx[m] = images[i]; units[m] = scale / d; memcpy(rays[m], ray, sizeof(ray)); ++m;
  }
  if (m < 2) return 0;
  units[0] = 0.0f;
  int out = 0;
  while (m > 0) {
    int best = 0;
    for (int j = 1; j < m; ++j)
      if (units[j] < units[best]) best = j;
    images[out++] = idx[best];
    float br[4]; memcpy(br, rays[best], sizeof(br));
    int k = 0;
    for (int j = 0; j < m; ++j) {
      if (j == best) continue;
      const float ftmp = fminf_(threshold, fmaxf_(threshold / 2.0f, 1.0f - dot4(br, rays[j])));
      const float u = units[j] * (threshold / ftmp);
      idx[k] = idx[j]; memcpy(rays[k], rays[j], sizeof(br)); units[k] = u; ++k;
    }
    m = k;
  }
  return out;
}

/* source/image/photoSetS.cpp:164-189 */
static int check_angles(const pmvso_ctx* c, const float* coord, const int* images, int n, float min_angle, float max_angle) {
  float rays[MAXIMG_LOCAL][4];
  for (int i = 0; i < n; ++i) {
    const cam_t* cam = &c->cams[images[i]];
    for (int k = 0; k < 4; ++k) rays[i][k] = cam->centre[k] - coord[k];
    unitize4(rays[i]);
  }
  int count = 0;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      const float d = fmaxf_(-1.0f, fminf_(1.0f, dot4(rays[i], rays[j])));
      const float angle = (float)acos(d);
      if (min_angle < angle && angle < max_angle) ++count;
    }
  return count < 1 ? 1 : 0;
}

/* optim.cpp:95-122 */
int pmvso_pre_process(const pmvso_ctx* c, const float* coord, const float* normal, int* images, int* n, int cap,
                      float* dscale, float* ascale) {
  int m = *n;
  *dscale = 0.0f; *ascale = 0.0f;
  m = add_images(c, coord, normal, images, m, cap);
  m = constraint_images(c, coord, normal, images, m, c->ncc_threshold_before);
  m = sort_images(c, coord, normal, images, m);
  if (m > 0) set_scales(c, coord, images, m, dscale, ascale);
  *n = m;
  if (m < c->min_image_num) return 1;
  if (check_angles(c, coord, images, m, c->max_angle_threshold, c->angle_threshold1)) { *n = 0; return 1; }
  return 0;
}

/* optim.cpp:124-148 */
static int filter_images_by_angle(const pmvso_ctx* c, const float* coord, const float* normal, int* images, int n) {
  int m = 0;
  const double th = cos(c->angle_threshold1);
  for (int i = 0; i < n; ++i) {
    const cam_t* cam = &c->cams[images[i]];
    float ray[4] = {cam->centre[0] - coord[0], cam->centre[1] - coord[1], cam->centre[2] - coord[2], cam->centre[3] - coord[3]};
    unitize4(ray);
    if (dot4(ray, normal) < th) {
      if (i == 0) return 0;
    } else images[m++] = images[i];
  }
  return m;
}

/* patchOrganizerS.cpp:400-414 */
static void set_grids(const pmvso_ctx* c, const float* coord, const int* images, int n, int* grids) {
  for (int i = 0; i < n; ++i) {
    float ic[3];
    project(c, images[i], coord, c->level, ic);
    grids[2 * i] = ((int)floorf(ic[0] + 0.5f)) / c->csize;
    grids[2 * i + 1] = ((int)floorf(ic[1] + 0.5f)) / c->csize;
  }
}

/* optim.cpp:208-254 */
static int set_ref_image(const pmvso_ctx* c, const float* coord, const float* normal, int* images, int n) {
  int idx[MAXIMG_LOCAL]; int m = 0;
  for (int i = 0; i < n; ++i)
    if (images[i] < c->tnum) idx[m++] = images[i];
  if (m == 0) return 0;
  float* inccs = (float*)malloc(sizeof(float) * m * m);
  pmvso_set_inccs_matrix(c, coord, normal, idx, m, 1, inccs);
  int refindex = -1;
  float refncc = INT_MAX / 2;
  for (int i = 0; i < m; ++i) {
    float sum = 0.0f;
    for (int j = 0; j < m; ++j) sum += inccs[i * m + j];
    if (sum < refncc) { refncc = sum; refindex = i; }
  }
  free(inccs);
  const int refIndex = idx[refindex];
  for (int i = 0; i < n; ++i)
    if (images[i] == refIndex) { const int t = images[0]; images[0] = refIndex; images[i] = t; break; }
  return n;
}

/* optim.cpp:150-190 at _depth == 0; no masks / bounding images in scope (getMask == 1) */
int pmvso_post_process(const pmvso_ctx* c, const float* coord, const float* normal, float ncc, int* images, int* n,
                       int cap, int* grids, int* timages, float* tmp) {
  int m = *n;
  *timages = 0; *tmp = 0.0f;
  if (m < c->min_image_num) return 1;
  m = add_images(c, coord, normal, images, m, cap);
  m = constraint_images(c, coord, normal, images, m, c->ncc_threshold);
  m = filter_images_by_angle(c, coord, normal, images, m);
  *n = m;
  if (m < c->min_image_num) return 1;
  set_grids(c, coord, images, m, grids);
  m = set_ref_image(c, coord, normal, images, m);
  *n = m;
  if (m == 0) return 1;
  m = constraint_images(c, coord, normal, images, m, c->ncc_threshold);
  *n = m;
  if (m < c->min_image_num) return 1;
  set_grids(c, coord, images, m, grids);
  int t = 0;
  for (int i = 0; i < m; ++i)
    if (images[i] < c->tnum) ++t;
  *timages = t;
  *tmp = fmaxf_(0.0f, ncc - c->ncc_threshold) * t; /* include/pmvs/patch.hpp:48-50 score2 */
  return 0;
}


/* ================================================================ filter stage */
void pmvso_set_depth(pmvso_ctx* c, int depth) { c->depth = depth; }

void pmvso_grid_dims(const pmvso_ctx* c, int image, int* gw, int* gh) { /* patchOrganizerS.cpp:72-77 */
  const int k = image * c->nlevels + c->level;
  *gw = (c->w[k] + c->csize - 1) / c->csize;
  *gh = (c->h[k] + c->csize - 1) / c->csize;
}

static void* dup_mem(const void* p, size_t n) { void* q = malloc(n ? n : 1); if (n) memcpy(q, p, n); return q; }

void pmvso_store_set(pmvso_ctx* c, int P, const float* coords, const float* normals, const float* ncc, const float* dscale,
                     const int* img_off, const int* images, const int* grids,
                     const int* vimg_off, const int* vimages, const int* vgrids, const int* timages) {
  free(c->s_coords); free(c->s_normals); free(c->s_ncc); free(c->s_dscale); free(c->s_img_off); free(c->s_images); free(c->s_grids);
  free(c->s_vimg_off); free(c->s_vimages); free(c->s_vgrids); free(c->s_timages); free(c->cell_off); free(c->cell_base); free(c->cell_patch);
  free(c->vcell_off); free(c->vcell_patch);
  c->P = P;
  c->s_coords = dup_mem(coords, sizeof(float) * 4 * P); c->s_normals = dup_mem(normals, sizeof(float) * 4 * P);
  c->s_ncc = dup_mem(ncc, sizeof(float) * P); c->s_dscale = dup_mem(dscale, sizeof(float) * P);
  c->s_img_off = dup_mem(img_off, sizeof(int) * (P + 1));
  c->s_images = dup_mem(images, sizeof(int) * img_off[P]); c->s_grids = dup_mem(grids, sizeof(int) * 2 * img_off[P]);
  c->s_vimg_off = dup_mem(vimg_off, sizeof(int) * (P + 1));
  c->s_vimages = dup_mem(vimages, sizeof(int) * vimg_off[P]); c->s_vgrids = dup_mem(vgrids, sizeof(int) * 2 * vimg_off[P]);
  c->s_timages = dup_mem(timages, sizeof(int) * P);
  /* _pgrids (patchOrganizerS.cpp:315-331) as CSR; the order inside a cell does not matter for the maxima taken over it */
  c->cell_base = (int*)malloc(sizeof(int) * (c->tnum + 1));
  int total = 0;
  for (int i = 0; i < c->tnum; ++i) { int gw, gh; pmvso_grid_dims(c, i, &gw, &gh); c->cell_base[i] = total; total += gw * gh; }
  c->cell_base[c->tnum] = total;
  c->cell_off = (int*)calloc(total + 1, sizeof(int));
  for (int p = 0; p < P; ++p)
    for (int e = img_off[p]; e < img_off[p + 1]; ++e) {
      const int im = images[e];
      if (c->tnum <= im) continue;
      int gw, gh; pmvso_grid_dims(c, im, &gw, &gh);
      c->cell_off[c->cell_base[im] + grids[2 * e + 1] * gw + grids[2 * e] + 1]++;
    }
  for (int i = 0; i < total; ++i) c->cell_off[i + 1] += c->cell_off[i];
  c->cell_patch = (int*)malloc(sizeof(int) * (c->cell_off[total] ? c->cell_off[total] : 1));
  int* fill = (int*)calloc(total, sizeof(int));
  for (int p = 0; p < P; ++p)
    for (int e = img_off[p]; e < img_off[p + 1]; ++e) {
      const int im = images[e];
      if (c->tnum <= im) continue;
      int gw, gh; pmvso_grid_dims(c, im, &gw, &gh);
      const int cell = c->cell_base[im] + grids[2 * e + 1] * gw + grids[2 * e];
      c->cell_patch[c->cell_off[cell] + fill[cell]++] = p;
    }
  /* _vpgrids the same way, from _vimages / _vgrids */
  c->vcell_off = (int*)calloc(total + 1, sizeof(int));
  for (int p = 0; p < P; ++p)
    for (int e = vimg_off[p]; e < vimg_off[p + 1]; ++e) {
      const int im = vimages[e];
      int gw, gh; pmvso_grid_dims(c, im, &gw, &gh);
      c->vcell_off[c->cell_base[im] + vgrids[2 * e + 1] * gw + vgrids[2 * e] + 1]++;
    }
  for (int i = 0; i < total; ++i) c->vcell_off[i + 1] += c->vcell_off[i];
  c->vcell_patch = (int*)malloc(sizeof(int) * (c->vcell_off[total] ? c->vcell_off[total] : 1));
  memset(fill, 0, sizeof(int) * total);
  for (int p = 0; p < P; ++p)
    for (int e = vimg_off[p]; e < vimg_off[p + 1]; ++e) {
      const int im = vimages[e];
      int gw, gh; pmvso_grid_dims(c, im, &gw, &gh);
      const int cell = c->cell_base[im] + vgrids[2 * e + 1] * gw + vgrids[2 * e];
      c->vcell_patch[c->vcell_off[cell] + fill[cell]++] = p;
    }
  free(fill);
}

/* CFilter::setDepthMapsThread (filter.cpp:687-732): patches in table order, strictly nearer replaces */
void pmvso_build_depth_maps(pmvso_ctx* c) {
  if (!c->dp) c->dp = (int**)calloc(c->tnum, sizeof(int*));
  for (int index = 0; index < c->tnum; ++index) {
    int gw, gh; pmvso_grid_dims(c, index, &gw, &gh);
    free(c->dp[index]);
    c->dp[index] = (int*)malloc(sizeof(int) * gw * gh);
    for (int i = 0; i < gw * gh; ++i) c->dp[index][i] = -1;
    const cam_t* cam = &c->cams[index];
    for (int p = 0; p < c->P; ++p) {
      const float* X = c->s_coords + 4 * p;
      float ic[3];
      project(c, index, X, c->level, ic);
      const float fx = ic[0] / c->csize;
      const int xs[2] = {(int)floor(fx), (int)ceil(fx)};
      const float fy = ic[1] / c->csize;
      const int ys[2] = {(int)floor(fy), (int)ceil(fy)};
      const float depth = dot4(cam->oaxis, X);
      for (int j = 0; j < 2; ++j)
        for (int i = 0; i < 2; ++i) {
          if (xs[i] < 0 || gw <= xs[i] || ys[j] < 0 || gh <= ys[j]) continue;
          const int cell = ys[j] * gw + xs[i];
          if (c->dp[index][cell] < 0) c->dp[index][cell] = p;
          else {
            const float dtmp = dot4(cam->oaxis, c->s_coords + 4 * c->dp[index][cell]);
            if (depth < dtmp) c->dp[index][cell] = p;
          }
        }
    }
  }
}
void pmvso_get_depth_map(const pmvso_ctx* c, int image, int* out) {
  int gw, gh; pmvso_grid_dims(c, image, &gw, &gh);
  memcpy(out, c->dp[image], sizeof(int) * gw * gh);
}

/* CPatchOrganizerS::isVisible (patchOrganizerS.cpp:487-526) */
int pmvso_is_visible(const pmvso_ctx* c, const float* coord, const float* normal, int image, int ix, int iy, float strict) {
  int gw, gh; pmvso_grid_dims(c, image, &gw, &gh);
  if (ix < 0 || gw <= ix || iy < 0 || gh <= iy) return 0;
  if (c->depth == 0) return 1;
  const int q = c->dp[image][iy * gw + ix];
  if (q < 0) return 1;
  const cam_t* cam = &c->cams[image];
  float ray[4] = {coord[0] - cam->centre[0], coord[1] - cam->centre[1], coord[2] - cam->centre[2], coord[3] - cam->centre[3]};
  unitize4(ray);
  const float* Y = c->s_coords + 4 * q;
  const float d[4] = {coord[0] - Y[0], coord[1] - Y[1], coord[2] - Y[2], coord[3] - Y[3]};
  const float diff = dot4(ray, d);
  const double factor = fmin(2.0, 2.0 + dot4(ray, normal));
  if (diff < get_unit(c, image, coord) * c->csize * strict * factor) return 1;
  return 0;
}

/* isVisible0 + setVImagesVGrids (patchOrganizerS.cpp:420-450, 479-485); strict = _neighborThreshold = 0.5 */
int pmvso_set_vimages(const pmvso_ctx* c, int k, int* vimages, int* vgrids, int cap) {
  unsigned char used[MAXIMG_LOCAL];
  memset(used, 0, sizeof(used));
  for (int e = c->s_img_off[k]; e < c->s_img_off[k + 1]; ++e)
    if (c->s_images[e] < c->tnum) used[c->s_images[e]] = 1;
  const float* X = c->s_coords + 4 * k;
  const float* N = c->s_normals + 4 * k;
  int n = 0;
  for (int image = 0; image < c->tnum; ++image) {
    if (used[image]) continue;
    float ic[3];
    project(c, image, X, c->level, ic);
    const int ix = ((int)floorf(ic[0] + 0.5f)) / c->csize;
    const int iy = ((int)floorf(ic[1] + 0.5f)) / c->csize;
    if (pmvso_is_visible(c, X, N, image, ix, iy, 0.5f) == 0) continue;
    if (n < cap) { vimages[n] = image; vgrids[2 * n] = ix; vgrids[2 * n + 1] = iy; ++n; }
  }
  return n;
}

/* filter.cpp:315-343; strict = _neighborThreshold1 = 1.0 */
int pmvso_filter_exact_safe(const pmvso_ctx* c, int k, int image, int x, int y) {
  int w, h; pmvso_grid_dims(c, image, &w, &h);
  const float* X = c->s_coords + 4 * k;
  const float* N = c->s_normals + 4 * k;
  if (pmvso_is_visible(c, X, N, image, x, y, 1.0f)) return 1;
  if (0 < x && pmvso_is_visible(c, X, N, image, x - 1, y, 1.0f)) return 1;
  if (x < w - 1 && pmvso_is_visible(c, X, N, image, x + 1, y, 1.0f)) return 1;
  if (0 < y && pmvso_is_visible(c, X, N, image, x, y - 1, 1.0f)) return 1;
  if (y < h - 1 && pmvso_is_visible(c, X, N, image, x, y + 1, 1.0f)) return 1;
  return 0;
}

/* CFindMatch::isNeighbor (findMatch.cpp:120-149) */
int pmvso_is_neighbor(const pmvso_ctx* c, int a, int b, float thr) {
  const float* Xa = c->s_coords + 4 * a; const float* Xb = c->s_coords + 4 * b;
  const float* Na = c->s_normals + 4 * a; const float* Nb = c->s_normals + 4 * b;
  const float hunit = (float)((get_unit(c, c->s_images[c->s_img_off[a]], Xa) + get_unit(c, c->s_images[c->s_img_off[b]], Xb)) / 2.0 * c->csize);
  if (dot4(Na, Nb) < cos(120.0 * M_PI / 180.0)) return 0;
  const float diff[4] = {Xb[0] - Xa[0], Xb[1] - Xa[1], Xb[2] - Xa[2], Xb[3] - Xa[3]};
  const float vunit = c->s_dscale[a] + c->s_dscale[b];
  const float f0 = dot4(Na, diff);
  const float f1 = dot4(Nb, diff);
  float ftmp = (float)((fabsf(f0) + fabsf(f1)) / 2.0);
  ftmp /= vunit;
  float t[4];
  for (int k = 0; k < 4; ++k) t[k] = diff[k] * 2 - Na[k] * f0 - Nb[k] * f1;
  const float hsize = (float)(norm4(t) / 2.0 / hunit);
  if (1.0 < hsize) ftmp /= fminf_(2.0f, hsize);
  return ftmp < thr ? 1 : 0;
}

/* CFilter::computeGain (filter.cpp:88-146); neighbour threshold = _neighborThreshold1 = 1.0 */
float pmvso_compute_gain(const pmvso_ctx* c, int k) {
  const float thr = c->ncc_threshold;
  float gain = fmaxf_(0.0f, c->s_ncc[k] - thr) * c->s_timages[k];   /* score2 */
  for (int e = c->s_img_off[k]; e < c->s_img_off[k + 1]; ++e) {
    const int index = c->s_images[e];
    if (c->tnum <= index) continue;
    int gw, gh; pmvso_grid_dims(c, index, &gw, &gh);
    const int cell = c->cell_base[index] + c->s_grids[2 * e + 1] * gw + c->s_grids[2 * e];
    float maxpressure = 0.0f;
    for (int j = c->cell_off[cell]; j < c->cell_off[cell + 1]; ++j) {
      const int q = c->cell_patch[j];
      if (!pmvso_is_neighbor(c, k, q, 1.0f)) maxpressure = fmaxf_(maxpressure, c->s_ncc[q] - thr);
    }
    gain -= maxpressure;
  }
  for (int e = c->s_vimg_off[k]; e < c->s_vimg_off[k + 1]; ++e) {
    const int index = c->s_vimages[e];
    if (c->tnum <= index) continue;
    const cam_t* cam = &c->cams[index];
    const float pdepth = dot4(cam->oaxis, c->s_coords + 4 * k);   /* computeDepth, camera.cpp:445-452 (perspective) */
    int gw, gh; pmvso_grid_dims(c, index, &gw, &gh);
    const int cell = c->cell_base[index] + c->s_vgrids[2 * e + 1] * gw + c->s_vgrids[2 * e];
    float maxpressure = 0.0f;
    for (int j = c->cell_off[cell]; j < c->cell_off[cell + 1]; ++j) {
      const int q = c->cell_patch[j];
      const float bdepth = dot4(cam->oaxis, c->s_coords + 4 * q);
      if (pdepth < bdepth && !pmvso_is_neighbor(c, k, q, 1.0f)) maxpressure = fmaxf_(maxpressure, c->s_ncc[q] - thr);
    }
    gain -= maxpressure;
  }
  return gain;
}

/* ---------------------------------------------------------------------------------------------------------------
 * neighbour searches over the table: findNeighbors, findEmptyBlocks, filterNeighbor / filterQuad
 * --------------------------------------------------------------------------------------------------------------- */
/* Vec4f ortho (include/numeric/vec4.hpp:303-322) */
static void ortho4(const float* z, float* x, float* y) {
  x[0] = x[1] = x[2] = x[3] = 0.0f;
  if (fabsf(z[0]) > 0.5f) { x[0] = z[1]; x[1] = -z[0]; x[2] = 0.0f; }
  else if (fabsf(z[1]) > 0.5f) { x[1] = z[2]; x[2] = -z[1]; x[0] = 0.0f; }
  else { x[2] = z[0]; x[0] = -z[2]; x[1] = 0.0f; }
  unitize4(x);
  y[0] = z[1] * x[2] - z[2] * x[1];
  y[1] = z[2] * x[0] - z[0] * x[2];
  y[2] = z[0] * x[1] - z[1] * x[0];
  y[3] = 0.0f;
}

/* CExpand::computeRadius (expand.cpp:182-198) over COptim::computeUnits (optim.cpp:446-471) */
float pmvso_compute_radius(const pmvso_ctx* c, int k) {
  const float* X = c->s_coords + 4 * k; const float* N = c->s_normals + 4 * k;
  float min1 = INFINITY, min2 = INFINITY;   /* nth_element(.., begin + 1, ..): the second smallest value */
  const int n = c->s_img_off[k + 1] - c->s_img_off[k];
  for (int e = c->s_img_off[k]; e < c->s_img_off[k + 1]; ++e) {
    const int im = c->s_images[e];
    float u = get_unit(c, im, X);
    float ray[4];
    for (int j = 0; j < 4; ++j) ray[j] = c->cams[im].centre[j] - X[j];
    unitize4(ray);
    const float denom = dot4(ray, N);
    if (0.0 < denom) u /= denom; else u = (float)(INT_MAX / 2);
    if (u < min1) { min2 = min1; min1 = u; } else if (u < min2) min2 = u;
  }
  if (n >= 2) return min2 * c->csize;
  return n == 1 ? min1 * c->csize : 0.0f;
}

/* CFindMatch::isNeighborRadius (findMatch.cpp:151-185) */
static int is_neighbor_radius(const pmvso_ctx* c, int a, int b, float hunit, float thr, float radius) {
  const float* Xa = c->s_coords + 4 * a; const float* Xb = c->s_coords + 4 * b;
  const float* Na = c->s_normals + 4 * a; const float* Nb = c->s_normals + 4 * b;
  if (dot4(Na, Nb) < cos(120.0 * M_PI / 180.0)) return 0;
  const float diff[4] = {Xb[0] - Xa[0], Xb[1] - Xa[1], Xb[2] - Xa[2], Xb[3] - Xa[3]};
  const float vunit = c->s_dscale[a] + c->s_dscale[b];
  const float f0 = dot4(Na, diff);
  const float f1 = dot4(Nb, diff);
  float ftmp = (float)((fabsf(f0) + fabsf(f1)) / 2.0);
  ftmp /= vunit;
  float t[4];
  for (int k = 0; k < 4; ++k) t[k] = diff[k] * 2 - Na[k] * f0 - Nb[k] * f1;
  const float hsize = (float)(norm4(t) / 2.0 / hunit);
  if (radius / hunit < hsize) return 0;
  if (1.0 < hsize) ftmp /= fminf_(2.0f, hsize);
  return ftmp < thr ? 1 : 0;
}

static int cmp_int(const void* a, const void* b) { const int x = *(const int*)a, y = *(const int*)b; return (x > y) - (x < y); }

/* CPatchOrganizerS::findNeighbors (patchOrganizerS.cpp:528-651): unique table ids, ascending (the reference sorts
 * by address).  Returns the count; at most cap ids are written. */
int pmvso_find_neighbors(const pmvso_ctx* c, int k, float scale, int margin, int skipvis, int* out, int cap) {
  const float* X = c->s_coords + 4 * k;
  const float radius = (float)(1.5 * margin * pmvso_compute_radius(c, k));
  float unit = 0.0f;
  const int n = c->s_img_off[k + 1] - c->s_img_off[k];
  for (int e = c->s_img_off[k]; e < c->s_img_off[k + 1]; ++e) unit += get_unit(c, c->s_images[e], X);
  unit /= n;
  unit *= c->csize;
  const float thr = 0.5f * scale;   /* _neighborThreshold (findMatch.cpp:96) x scale */
  int cnt = 0, room = 256;
  int* buf = (int*)malloc(sizeof(int) * room);
  for (int pass = 0; pass < (skipvis ? 1 : 2); ++pass) {
    const int e0 = pass ? c->s_vimg_off[k] : c->s_img_off[k], e1 = pass ? c->s_vimg_off[k + 1] : c->s_img_off[k + 1];
    for (int e = e0; e < e1; ++e) {
      const int image = pass ? c->s_vimages[e] : c->s_images[e];
      if (c->tnum <= image) continue;
      const int ix = pass ? c->s_vgrids[2 * e] : c->s_grids[2 * e], iy = pass ? c->s_vgrids[2 * e + 1] : c->s_grids[2 * e + 1];
      int gw, gh; pmvso_grid_dims(c, image, &gw, &gh);
      for (int j = -margin; j <= margin; ++j) {
        const int y = iy + j;
        if (y < 0 || gh <= y) continue;
        for (int i = -margin; i <= margin; ++i) {
          const int x = ix + i;
          if (x < 0 || gw <= x) continue;
          const int cell = c->cell_base[image] + y * gw + x;
          for (int list = 0; list < 2; ++list) {
            const int* off = list ? c->vcell_off : c->cell_off;
            const int* lst = list ? c->vcell_patch : c->cell_patch;
            for (int q = off[cell]; q < off[cell + 1]; ++q)
              if (is_neighbor_radius(c, k, lst[q], unit, thr, radius)) {
                if (cnt == room) { room *= 2; buf = (int*)realloc(buf, sizeof(int) * room); }
                buf[cnt++] = lst[q];
              }
          }
        }
      }
    }
  }
  qsort(buf, cnt, sizeof(int), cmp_int);
  int u = 0;
  for (int i = 0; i < cnt; ++i) if (i == 0 || buf[i] != buf[i - 1]) buf[u++] = buf[i];
  for (int i = 0; i < u && i < cap; ++i) out[i] = buf[i];
  free(buf);
  return u;
}

/* CExpand::findEmptyBlocks (expand.cpp:108-180): bit i = direction i already filled (fill[i] > 0); *radius = computeRadius */
int pmvso_find_empty_blocks(const pmvso_ctx* c, int k, float* radius_out) {
  const float* X = c->s_coords + 4 * k; const float* N = c->s_normals + 4 * k;
  const int dnum = 6;
  float xdir[4], ydir[4];
  ortho4(N, xdir, ydir);
  float fill[6] = {0, 0, 0, 0, 0, 0};
  const float radius = pmvso_compute_radius(c, k);
  const float radiuslow = radius / 6.0f, radiushigh = radius * 2.5f;
  int cap = c->P > 0 ? c->P : 1;
  int* nb = (int*)malloc(sizeof(int) * cap);
  const int n = pmvso_find_neighbors(c, k, 4.0f, 1, 0, nb, cap);
  for (int i = 0; i < n; ++i) {
    const float* Xq = c->s_coords + 4 * nb[i];
    const float diff[4] = {Xq[0] - X[0], Xq[1] - X[1], Xq[2] - X[2], Xq[3] - X[3]};
    float f2[2] = {dot4(diff, xdir), dot4(diff, ydir)};
    const float len = sqrtf(f2[0] * f2[0] + f2[1] * f2[1]);
    if (len < radiuslow || radiushigh < len) continue;
    f2[0] /= len; f2[1] /= len;
    float angle = (float)atan2((double)f2[1], (double)f2[0]);   /* the reference object imports the double atan2 (nm -u) */
    if (angle < 0.0) angle = (float)(angle + 2 * M_PI);
    const float findex = (float)(angle / (2 * M_PI / dnum));
    const int lindex = (int)floor(findex);
    const int hindex = lindex + 1;
    fill[lindex % dnum] += hindex - findex;
    fill[hindex % dnum] += findex - lindex;
  }
  free(nb);
  int mask = 0;
  for (int i = 0; i < dnum; ++i) if (0.0f < fill[i]) mask |= 1 << i;
  if (radius_out) *radius_out = radius;
  return mask;
}

/* CFilter::filterQuad (filter.cpp:394-462) over the neighbour list nb[0..n).  The least-squares solve (Cmylapack::lls ->
 * Eigen jacobiSvd in the reference, absent: PARITY UNPINNED) is the 5x5 normal equations in double with partial pivoting. */
static float quad_residual(const pmvso_ctx* c, int k, const int* nb, int n) {
  const float* X = c->s_coords + 4 * k; const float* N = c->s_normals + 4 * k;
  float xdir[4], ydir[4];
  ortho4(N, xdir, ydir);
  float h = 0.0f;
  for (int i = 0; i < n; ++i) {
    const float* Xq = c->s_coords + 4 * nb[i];
    const float d[4] = {Xq[0] - X[0], Xq[1] - X[1], Xq[2] - X[2], Xq[3] - X[3]};
    h += norm4(d);
  }
  h /= n;
  float* fx = (float*)malloc(sizeof(float) * n); float* fy = (float*)malloc(sizeof(float) * n); float* fz = (float*)malloc(sizeof(float) * n);
  double ATA[5][5] = {{0}}, ATb[5] = {0};
  for (int i = 0; i < n; ++i) {
    const float* Xq = c->s_coords + 4 * nb[i];
    const float d[4] = {Xq[0] - X[0], Xq[1] - X[1], Xq[2] - X[2], Xq[3] - X[3]};
    fx[i] = dot4(d, xdir) / h; fy[i] = dot4(d, ydir) / h; fz[i] = dot4(d, N);
    const double row[5] = {(double)(fx[i] * fx[i]), (double)(fy[i] * fy[i]), (double)(fx[i] * fy[i]), (double)fx[i], (double)fy[i]};
    for (int a = 0; a < 5; ++a) {
      for (int b = 0; b < 5; ++b) ATA[a][b] += row[a] * row[b];
      ATb[a] += row[a] * (double)fz[i];
    }
  }
  double M[5][6], x[5] = {0, 0, 0, 0, 0};
  for (int a = 0; a < 5; ++a) { for (int b = 0; b < 5; ++b) M[a][b] = ATA[a][b]; M[a][5] = ATb[a]; }
  int singular = 0;
  for (int col = 0; col < 5 && !singular; ++col) {
    int piv = col;
    for (int r = col + 1; r < 5; ++r) if (fabs(M[r][col]) > fabs(M[piv][col])) piv = r;
    if (fabs(M[piv][col]) < 1e-300) { singular = 1; break; }
    if (piv != col) for (int j = 0; j < 6; ++j) { const double t = M[col][j]; M[col][j] = M[piv][j]; M[piv][j] = t; }
    for (int r = col + 1; r < 5; ++r) {
      const double f = M[r][col] / M[col][col];
      for (int j = col; j < 6; ++j) M[r][j] -= f * M[col][j];
    }
  }
  if (!singular)
    for (int r = 4; r >= 0; --r) {
      double acc = M[r][5];
      for (int j = r + 1; j < 5; ++j) acc -= M[r][j] * x[j];
      x[r] = acc / M[r][r];
    }
  const float xs[5] = {(float)x[0], (float)x[1], (float)x[2], (float)x[3], (float)x[4]};
  const int nimg = c->s_img_off[k + 1] - c->s_img_off[k];
  const int inum = c->tau < nimg ? c->tau : nimg;
  float unit = 0.0f;
  for (int i = 0; i < inum; ++i) unit += get_unit(c, c->s_images[c->s_img_off[k] + i], X);
  unit /= inum;
  float residual = 0.0f;
  for (int i = 0; i < n; ++i) {
    const float res = xs[0] * (fx[i] * fx[i]) + xs[1] * (fy[i] * fy[i]) + xs[2] * (fx[i] * fy[i]) + xs[3] * fx[i] + xs[4] * fy[i] - fz[i];
    residual += fabsf(res) / unit;
  }
  residual /= (n - 5);
  free(fx); free(fy); free(fz);
  return residual;
}

/* CFilter::filterNeighborThread (filter.cpp:357-392).  Returns 1 = reject; *residual_out = -1 when there are fewer than 6 neighbours. */
int pmvso_filter_neighbor(const pmvso_ctx* c, int k, float quad, float* residual_out, int* ncount_out) {
  int cap = c->P > 0 ? c->P : 1;
  int* nb = (int*)malloc(sizeof(int) * cap);
  const int n = pmvso_find_neighbors(c, k, 4.0f, 2, 1, nb, cap);
  if (ncount_out) *ncount_out = n;
  if (n < 6) { free(nb); if (residual_out) *residual_out = -1.0f; return 1; }
  const float residual = quad_residual(c, k, nb, n);
  free(nb);
  if (residual_out) *residual_out = residual;
  return residual < quad ? 0 : 1;
}

/* COptim::check (optim.cpp:363-383) on table patch k: 1 = reject; *gain_out = computeGain (what check stores in _tmp) */
int pmvso_check(const pmvso_ctx* c, int k, float quad, float* gain_out) {
  const float gain = pmvso_compute_gain(c, k);
  if (gain_out) *gain_out = gain;
  if (gain < 0.0) return 1;
  int cap = c->P > 0 ? c->P : 1;
  int* nb = (int*)malloc(sizeof(int) * cap);
  const int n = pmvso_find_neighbors(c, k, 4.0f, 2, 0, nb, cap);
  int rej = 0;
  if (6 < n && !(quad_residual(c, k, nb, n) < quad)) rej = 1;
  free(nb);
  return rej;
}

/* ---------------------------------------------------------------------------------------------------------------
 * feature detection: CHarris::run (harris.cpp:174-240) and CDifferenceOfGaussians::run
 * (dog.cpp:96-198) on the working-level image, unmasked images (mask and edge empty, detector.hpp:23-94 take the
 * clamping branch).  Same f32 operation order; exp / pow / log are the double libm entry points (nm -u detector.o).
 * --------------------------------------------------------------------------------------------------------------- */
typedef struct { float r; int x, y; long seq; } fpoint_t;

static void gauss_i(float sigma, float* g, int* ntaps) {   /* CDetector::setGaussI (detector.cpp:31-49) */
  const int margin = (int)ceil(2 * sigma);
  const int size = 2 * margin + 1;
  float denom = 0.0f;
  for (int x = 0; x < size; ++x) {
    const int xt = x - margin;
    const float d = (float)exp(-(xt * xt) / (2 * sigma * sigma));
    g[x] = d;
    denom += d;
  }
  for (int x = 0; x < size; ++x) g[x] /= denom;
  *ntaps = size;
}

/* masked overload with an empty mask: coordinates clamp to the image (detector.hpp:26-94) */
static void conv_x(float* img, int w, int h, const float* f, int n, float* buf) {
  const int margin = n / 2;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      float acc = 0.0f;
      for (int j = 0; j < n; ++j) {
        int xt = x + j - margin;
        if (xt < 0) xt = 0; else if (w <= xt) xt = w - 1;
        acc += f[j] * img[(size_t)y * w + xt];
      }
      buf[(size_t)y * w + x] = acc;
    }
  memcpy(img, buf, sizeof(float) * (size_t)w * h);
}
static void conv_y(float* img, int w, int h, const float* f, int n, float* buf) {
  const int margin = n / 2;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      float acc = 0.0f;
      for (int j = 0; j < n; ++j) {
        int yt = y + j - margin;
        if (yt < 0) yt = 0; else if (h <= yt) yt = h - 1;
        acc += f[j] * img[(size_t)yt * w + x];
      }
      buf[(size_t)y * w + x] = acc;
    }
  memcpy(img, buf, sizeof(float) * (size_t)w * h);
}

/* the per-block multiset<CPoint> (ordered by response; equal keys keep insertion order, begin() = smallest, oldest) */
typedef struct { fpoint_t p[5]; int n; } block_t;
static void block_insert(block_t* b, fpoint_t q) {   /* insert after the elements that are <= q, then drop begin() beyond 4 */
  int pos = b->n;
  while (pos > 0 && q.r < b->p[pos - 1].r) { b->p[pos] = b->p[pos - 1]; --pos; }
  b->p[pos] = q;
  b->n++;
  if (b->n > 4) { for (int i = 1; i < b->n; ++i) b->p[i - 1] = b->p[i]; b->n--; }
}
static int cmp_fpoint_desc(const void* a, const void* b) {   /* reverse iteration of the result multiset */
  const fpoint_t* p = (const fpoint_t*)a; const fpoint_t* q = (const fpoint_t*)b;
  if (p->r != q->r) return p->r < q->r ? 1 : -1;
  return p->seq < q->seq ? 1 : (p->seq > q->seq ? -1 : 0);
}
static int emit_blocks(block_t* blocks, int nb, int type, float* xy, float* resp, int* types, int cap, int have) {
  int total = 0;
  for (int i = 0; i < nb; ++i) total += blocks[i].n;
  fpoint_t* all = (fpoint_t*)malloc(sizeof(fpoint_t) * (total ? total : 1));
  long seq = 0;
  int k = 0;
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j < blocks[i].n; ++j) { all[k] = blocks[i].p[j]; all[k].seq = seq++; ++k; }
  qsort(all, total, sizeof(fpoint_t), cmp_fpoint_desc);
  for (int i = 0; i < total && have + i < cap; ++i) {
    xy[2 * (have + i)] = (float)all[i].x; xy[2 * (have + i) + 1] = (float)all[i].y;
    resp[have + i] = all[i].r; types[have + i] = type;
  }
  free(all);
  return total;
}

/* features of image `index` at the working level, Harris points first then DoG, each strongest first
 * (detectFeatures.cpp:77-118).  Returns the total count; at most cap are written. */
int pmvso_detect_features(const pmvso_ctx* c, int index, int gspeedup, float* xy, float* resp, int* types, int cap) {
  const int w = c->w[index * c->nlevels + c->level], h = c->h[index * c->nlevels + c->level];
  const unsigned char* pix = c->pix[index * c->nlevels + c->level];
  const size_t n = (size_t)w * h;
  float* im[3];
  for (int k = 0; k < 3; ++k) {
    im[k] = (float*)malloc(sizeof(float) * n);
    for (size_t i = 0; i < n; ++i) im[k][i] = ((int)pix[3 * i + k]) / 255.0f;
  }
  float* buf = (float*)malloc(sizeof(float) * n);
  const int factor = 2, gridsize = gspeedup * factor;
  const int gw = (w + gridsize - 1) / gridsize, gh = (h + gridsize - 1) / gridsize;
  int have = 0;
  /* ---- Harris, sigma 4 */
  {
    const float sigma = 4.0f;
    float gI[64]; int nI;
    gauss_i(sigma, gI, &nI);
    const float dfilter[3] = {-0.5f, 0.0f, 0.5f};
    const float ifilter[3] = {(float)(1.0 / 3.0), (float)(1.0 / 3.0), (float)(1.0 / 3.0)};
    float* xx = (float*)calloc(n, sizeof(float)); float* yy = (float*)calloc(n, sizeof(float)); float* xyp = (float*)calloc(n, sizeof(float));
    float* dx = (float*)malloc(sizeof(float) * n); float* dy = (float*)malloc(sizeof(float) * n);
    float* sx = (float*)calloc(n, sizeof(float)); float* sy = (float*)calloc(n, sizeof(float)); float* sxy = (float*)calloc(n, sizeof(float));
    /* Vec3f * Vec3f = x x' + y y' + z z' summed left to right (vec3.hpp); accumulate the three channel products in that order */
    for (int k = 0; k < 3; ++k) {
      memcpy(dx, im[k], sizeof(float) * n); memcpy(dy, im[k], sizeof(float) * n);
      conv_x(dx, w, h, dfilter, 3, buf); conv_y(dx, w, h, ifilter, 3, buf);
      conv_x(dy, w, h, ifilter, 3, buf); conv_y(dy, w, h, dfilter, 3, buf);
      for (size_t i = 0; i < n; ++i) {
        if (k == 0) { sx[i] = dx[i] * dx[i]; sy[i] = dy[i] * dy[i]; sxy[i] = dx[i] * dy[i]; }
        else { sx[i] = sx[i] + dx[i] * dx[i]; sy[i] = sy[i] + dy[i] * dy[i]; sxy[i] = sxy[i] + dx[i] * dy[i]; }
      }
    }
    for (size_t i = 0; i < n; ++i) { xx[i] += sx[i]; yy[i] += sy[i]; xyp[i] += sxy[i]; }   /* harris.cpp:76-78 */
    conv_x(xx, w, h, gI, nI, buf); conv_y(xx, w, h, gI, nI, buf);
    conv_x(yy, w, h, gI, nI, buf); conv_y(yy, w, h, gI, nI, buf);
    conv_x(xyp, w, h, gI, nI, buf); conv_y(xyp, w, h, gI, nI, buf);
    float* response = dx;   /* reuse */
    for (size_t i = 0; i < n; ++i) {
      const float D = xx[i] * yy[i] - xyp[i] * xyp[i];
      const float tr = xx[i] + yy[i];
      response[i] = (float)(D - 0.06 * tr * tr);
    }
    float* nms = dy;
    memcpy(nms, response, sizeof(float) * n);
    for (int y = 1; y < h - 1; ++y)
      for (int x = 1; x < w - 1; ++x) {
        const float v = response[(size_t)y * w + x];
        if (v < response[(size_t)y * w + x + 1] || v < response[(size_t)y * w + x - 1] || v < response[(size_t)(y + 1) * w + x] ||
            v < response[(size_t)(y - 1) * w + x]) nms[(size_t)y * w + x] = 0.0f;
      }
    block_t* blocks = (block_t*)calloc((size_t)gw * gh, sizeof(block_t));
    const int margin = (2 * (int)ceil(2 * sigma) + 1) / 2;   /* _gaussD.size() / 2 (detector.cpp:10-12) */
    for (int y = margin; y < h - margin; ++y)
      for (int x = margin; x < w - margin; ++x) {
        const float v = nms[(size_t)y * w + x];
        if (v == 0.0) continue;
        const int x0 = x / gridsize < gw - 1 ? x / gridsize : gw - 1, y0 = y / gridsize < gh - 1 ? y / gridsize : gh - 1;
        block_t* b = &blocks[(size_t)y0 * gw + x0];
        if (b->n < 4 || b->p[0].r < v) { fpoint_t q = {v, x, y, 0}; block_insert(b, q); }
      }
    have += emit_blocks(blocks, gw * gh, 0, xy, resp, types, cap, have);
    free(blocks); free(xx); free(yy); free(xyp); free(dx); free(dy); free(sx); free(sy); free(sxy);
  }
  /* ---- difference of Gaussians, scales 1 .. 3 */
  {
    const float first = 1.0f, last = 3.0f;
    const float scalestep = (float)pow(2.0f, 1 / 2.0f);
    int steps = (int)ceil(log(last / first) / log(scalestep));
    if (steps < 4) steps = 4;
    const int nres = steps + 1;                      /* res[k] at sigma first * step^k (the reference keeps a window of two) */
    float** res = (float**)malloc(sizeof(float*) * nres);
    float* ch = (float*)malloc(sizeof(float) * n);
    for (int k = 0; k < nres; ++k) {
      float sigma;
      if (k == 0) sigma = first; else if (k == 1) sigma = first * scalestep; else if (k == 2) sigma = first * scalestep * scalestep;
      else sigma = (float)(first * pow(scalestep, k));   /* cscale = _firstScale * pow(scalestep, i + 1), k = i + 1 */
      float g[128]; int ng;
      gauss_i(sigma, g, &ng);
      res[k] = (float*)calloc(n, sizeof(float));
      float* acc0 = res[k];
      /* norm(Vec3f) = sqrt(x x + y y + z z) (vec3.hpp:210-214) */
      float* c0 = (float*)malloc(sizeof(float) * n); float* c1 = (float*)malloc(sizeof(float) * n);
      memcpy(c0, im[0], sizeof(float) * n); conv_x(c0, w, h, g, ng, buf); conv_y(c0, w, h, g, ng, buf);
      memcpy(c1, im[1], sizeof(float) * n); conv_x(c1, w, h, g, ng, buf); conv_y(c1, w, h, g, ng, buf);
      memcpy(ch, im[2], sizeof(float) * n); conv_x(ch, w, h, g, ng, buf); conv_y(ch, w, h, g, ng, buf);
      for (size_t i = 0; i < n; ++i) acc0[i] = sqrtf(c0[i] * c0[i] + c1[i] * c1[i] + ch[i] * ch[i]);
      free(c0); free(c1);
    }
    float** dog = (float**)malloc(sizeof(float*) * (nres - 1));
    for (int k = 0; k + 1 < nres; ++k) {
      dog[k] = (float*)malloc(sizeof(float) * n);
      for (size_t i = 0; i < n; ++i) dog[k][i] = res[k + 1][i] - res[k][i];
    }
    unsigned char* seen = (unsigned char*)calloc(n, 1);
    block_t* blocks = (block_t*)calloc((size_t)gw * gh, sizeof(block_t));
    for (int i = 2; i <= steps - 1; ++i) {
      const float cscale = (float)(first * pow(scalestep, i + 1));
      const float* pd = dog[i - 2]; const float* cd = dog[i - 1]; const float* nd = dog[i];
      const int margin = (int)ceil(2 * cscale);
      for (int y = margin; y < h - margin; ++y)
        for (int x = margin; x < w - margin; ++x) {
          const size_t o = (size_t)y * w + x;
          const float v = cd[o];
          if (seen[o] || v == 0.0) continue;
          int ext;
          if (0.0 < v)
            ext = cd[o - w - 1] < v && cd[o - 1] < v && cd[o + w - 1] < v && cd[o - w] < v && cd[o + w] < v && cd[o - w + 1] < v && cd[o + 1] < v &&
                  cd[o + w + 1] < v && pd[o] < v && nd[o] < v;
          else
            ext = cd[o - w - 1] > v && cd[o - 1] > v && cd[o + w - 1] > v && cd[o - w] > v && cd[o + w] > v && cd[o - w + 1] > v && cd[o + 1] > v &&
                  cd[o + w + 1] > v && v < pd[o] && v < nd[o];
          if (!ext) continue;
          seen[o] = 1;
          const int x0 = x / gridsize < gw - 1 ? x / gridsize : gw - 1, y0 = y / gridsize < gh - 1 ? y / gridsize : gh - 1;
          fpoint_t q = {fabsf(v), x, y, 0};
          block_insert(&blocks[(size_t)y0 * gw + x0], q);
        }
    }
    const int got = emit_blocks(blocks, gw * gh, 1, xy, resp, types, cap, have);
    have += got;
    free(blocks); free(seen); free(ch);
    for (int k = 0; k < nres; ++k) free(res[k]);
    for (int k = 0; k + 1 < nres; ++k) free(dog[k]);
    free(res); free(dog);
  }
  free(buf);
  for (int k = 0; k < 3; ++k) free(im[k]);
  return have;
}
