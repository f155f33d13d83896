// cmvs-pmvs_b200/csrc/pmvs_device.cuh
//
// Device-side building blocks of the patch-optimisation path for sm_100a: projection, patch axes,
// per-view window set-up, bilinear texel gathers, warp-cooperative normalise / NCC, and the bounded
// Nelder-Mead that drives them.  One warp owns one patch; lanes own views during set-up and texels
// during sampling.  Written from scratch against the behaviour documented in SURVEY.md (reference
// file:line cited per function; paths relative to /root/reference).
//
// Arithmetic policy: this translation unit is compiled with -fmad=false.  Everything that feeds a
// discontinuous decision (angle gate, pyramid-level pick, bounds test, cell indexes) is evaluated in
// the reference's f32 operation order so those decisions are bit-exact.  Sums over texels use warp
// shuffles, i.e. a different association than the reference's sequential loops; NCC agrees to ~1e-6.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmvsb {

constexpr int kMaxTau = 8;
constexpr int kMaxLevels = 8;
constexpr unsigned kFull = 0xffffffffu;

struct __align__(16) CamDev {  // one per image (Image::CCamera + COptim axes), 128 B
  float P[3][4];           // projection at the working level (include/image/camera.hpp:89-108)
  float centre[4];         // optical centre, w = 1 (source/image/camera.cpp:138-175)
  float oaxis[4];          // optical axis (camera.cpp:112-118)
  float xaxis[3];          // COptim::_xaxes/_yaxes/_zaxes (source/pmvs/optim.cpp:47-53)
  float yaxis[3];
  float zaxis[3];
  float ipscale;           // COptim::_ipscales (optim.cpp:56-63)
  float pad[2];
};

struct LevelDev {          // one per (image, pyramid level); 32 B
  const uchar4* pix;       // RGBA8, row-major, pitch = w pixels; alpha unused
  int w, h;
  float ax1, ay1;          // this level's origin inside the gather atlas, plus one: texel (x, y)'s 2x2 footprint is centred at
                           // (x + ax1, y + ay1) in atlas coordinates (SceneDev::atlas)
  int pad0, pad1;
};

struct SceneDev {
  const CamDev* cams;
  const LevelDev* levels;  // [image * nlevels + level]
  int num, tnum, level, nlevels, csize, wsize, tau, min_image_num;
  float cos_angle1;        // smallest float >= cos(angleThreshold1): weight < cos(..) <=> weight < this
  float level_thr[kMaxLevels];  // ratio >= level_thr[k]  <=>  leveldif >= k - level + 1   (see host code)
  int n_level_thr;         // level + 2
  float ascale;            // (float)(M_PI / 48.0f)   (optim.cpp:590)
  double xtol, step;       // Nelder-Mead knobs (oracle/nm3.h is the written definition)
  int maxeval;
  uint32_t f32_2p23;       // 0x4B000000, passed as data so PRMT keeps its immediate slot for the byte selector
  const uchar4* dummy_pix; // any valid level (image 0, level 0): idle lanes sample texel (0,0) of it instead of branching
  unsigned long long atlas; // cudaTextureObject_t over ONE block-linear RGBA8 array holding every (image, level): point filter,
                            // normalized-float reads, unnormalized coordinates, gather enabled; 0 = not built (does not fit)
  // CImage::_masks / _edges at the working level (255 in, 0 out), one pointer per image; an image without a map has a
  // null entry, a scene without any has a null table (the common case: every gate below is then a uniform early-out)
  const unsigned char* const* mask_lv;
  const unsigned char* const* edge_lv;
  const int32_t* bimages;   // SOption::_bindexes (bounding images, option useBound)
  int n_bimages;
};

// IEEE f32 division / square root.  With PMVS_NOINLINE_DIV the ~13-instruction expansions (60 of them in the
// refine loop) become calls to one shared copy: code size vs call overhead, decided by measurement.
#ifndef PMVS_NOINLINE_DIV
#define PMVS_NOINLINE_DIV 0
#endif
#if PMVS_NOINLINE_DIV
__device__ __noinline__ float fdiv(float a, float b) { return a / b; }
__device__ __noinline__ float fsqrt(float a) { return sqrtf(a); }
#else
__device__ __forceinline__ float fdiv(float a, float b) { return a / b; }
__device__ __forceinline__ float fsqrt(float a) { return sqrtf(a); }
#endif

// ---------------------------------------------------------------------------------------------------
// small f32 helpers in reference operation order (include/numeric/vec3.hpp, vec4.hpp)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dot4(const float* u, const float* v) {
  return u[0] * v[0] + u[1] * v[1] + u[2] * v[2] + u[3] * v[3];
}
__device__ __forceinline__ float dot3(const float* u, const float* v) {
  return u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
}
__device__ __forceinline__ void cross3(const float* u, const float* v, float* o) {
  o[0] = u[1] * v[2] - v[1] * u[2];
  o[1] = -u[0] * v[2] + v[0] * u[2];
  o[2] = u[0] * v[1] - v[0] * u[1];
}
__device__ __forceinline__ void unitize3(float* v) {
  const float l = dot3(v, v);
  if (l != 1.0f && l != 0.0f) {
    const float s = fsqrt(l);
    v[0] = fdiv(v[0], s); v[1] = fdiv(v[1], s); v[2] = fdiv(v[2], s);
  }
}
__device__ __forceinline__ void unitize4(float* v) {
  const float l = dot4(v, v);
  if (l != 1.0f && l != 0.0f) {
    const float s = fsqrt(l);
    v[0] = fdiv(v[0], s); v[1] = fdiv(v[1], s); v[2] = fdiv(v[2], s); v[3] = fdiv(v[3], s);
  }
}
// std::min / std::max semantics (NaN handling differs from fminf/fmaxf)
__device__ __forceinline__ float smin(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float smax(float a, float b) { return a < b ? b : a; }

__device__ __forceinline__ void load_cam(const SceneDev& s, int index, CamDev& c) {
  const float4* src = reinterpret_cast<const float4*>(s.cams + index);
  float4* dst = reinterpret_cast<float4*>(&c);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(CamDev) / 16); ++i) dst[i] = __ldg(src + i);
}

// CCamera::project (include/image/camera.hpp:89-108)
__device__ __forceinline__ void project(const CamDev& cam, const float* X, float* o) {
  o[0] = dot4(cam.P[0], X);
  o[1] = dot4(cam.P[1], X);
  o[2] = dot4(cam.P[2], X);
  if (o[2] <= 0.0f) {
    o[0] = -65535.0f; o[1] = -65535.0f; o[2] = -1.0f;
    return;
  }
  const float z = o[2];
  o[0] = fdiv(o[0], z); o[1] = fdiv(o[1], z); o[2] = 1.0f;  // z / z == 1 exactly for finite z > 0
  const float lim = 2147483648.0f;  // (float)(INT_MAX - 3.0f) and -(float)(INT_MIN + 3.0f)
  o[0] = smax(-lim, smin(lim, o[0]));
  o[1] = smax(-lim, smin(lim, o[1]));
}

// CPhoto::getMask(coord, level) (include/image/photo.hpp:44-49 -> image.hpp:540-565): 1 without a map, 1 outside the image
__device__ __forceinline__ int get_mask_img(const SceneDev& s, int image, const float* X) {
  const unsigned char* m = s.mask_lv ? s.mask_lv[image] : nullptr;
  if (!m) return 1;
  CamDev cam;
  load_cam(s, image, cam);
  float ic[3];
  project(cam, X, ic);
  const LevelDev lv = s.levels[image * s.nlevels + s.level];
  const int ix = (int)floorf(ic[0] + 0.5f), iy = (int)floorf(ic[1] + 0.5f);
  if (ix < 0 || lv.w <= ix || iy < 0 || lv.h <= iy) return 1;
  return m[(size_t)iy * lv.w + ix];
}
// CPhoto::getEdge(coord, level) (photo.hpp:51-59 -> image.hpp:567-592): 1 without a map, 0 outside [0, w-1) x [0, h-1)
__device__ __forceinline__ int get_edge_img(const SceneDev& s, const CamDev& cam, int image, const float* X) {
  const unsigned char* m = s.edge_lv ? s.edge_lv[image] : nullptr;
  if (!m) return 1;
  float ic[3];
  project(cam, X, ic);
  const LevelDev lv = s.levels[image * s.nlevels + s.level];
  if (ic[0] < 0.0f || (float)(lv.w - 1) <= ic[0] || ic[1] < 0.0f || (float)(lv.h - 1) <= ic[1]) return 0;
  const int ix = (int)floorf(ic[0] + 0.5f), iy = (int)floorf(ic[1] + 0.5f);
  if (ix < 0 || lv.w <= ix || iy < 0 || lv.h <= iy) return 1;
  return m[(size_t)iy * lv.w + ix];
}
// one term of CFindMatch::insideBimages (source/pmvs/findMatch.cpp:109-118)
__device__ __forceinline__ int inside_bimage(const SceneDev& s, int image, const float* X) {
  CamDev cam;
  load_cam(s, image, cam);
  float ic[3];
  project(cam, X, ic);
  const LevelDev lv = s.levels[image * s.nlevels + s.level];
  return !(ic[0] < 0.0f || (float)(lv.w - 1) < ic[0] || ic[1] < 0.0f || (float)(lv.h - 1) < ic[1]);
}
// `getMask(coord, level) == 0 || insideBimages(coord) == 0` of postProcess / expandSub / collectCandidates
// (optim.cpp:153, expand.cpp:212, seed.cpp:314): CPhotoSetS::getMask walks EVERY image (photoSetS.hpp:109-116).
// Warp-collective; true = the point passes.
__device__ __forceinline__ bool mask_gate_warp(const SceneDev& s, const float* X, int lane) {
  if (!s.mask_lv && s.n_bimages == 0) return true;
  bool bad = false;
  if (s.mask_lv)
    for (int i = lane; i < s.num; i += 32) bad |= get_mask_img(s, i, X) == 0;
  for (int i = lane; i < s.n_bimages; i += 32) bad |= inside_bimage(s, s.bimages[i], X) == 0;
  return !__any_sync(kFull, bad);
}

// COptim::getUnit (optim.cpp:1116-1124): 2.0 * |X - C| * 2^level / ipscale, evaluated in double
__device__ __forceinline__ float get_unit(const CamDev& cam, int level, const float* X) {
  const float d[4] = {X[0] - cam.centre[0], X[1] - cam.centre[1], X[2] - cam.centre[2], X[3] - cam.centre[3]};
  const float fz = fsqrt(dot4(d, d));
  if (cam.ipscale == 0.0f) return 1.0f;
  return (float)(2.0 * (double)fz * (double)(1 << level) / (double)cam.ipscale);
}

// COptim::getPAxes (optim.cpp:1127-1144)
__device__ __forceinline__ void get_paxes(const CamDev& cam, int level, const float* coord, const float* normal,
                                          float* px, float* py) {
  const float pscale = get_unit(cam, level, coord);
  const float n3[3] = {normal[0], normal[1], normal[2]};
  float y3[3], x3[3];
  cross3(n3, cam.xaxis, y3);
  unitize3(y3);
  cross3(y3, n3, x3);
  px[0] = x3[0] * pscale; px[1] = x3[1] * pscale; px[2] = x3[2] * pscale; px[3] = 0.0f * pscale;
  py[0] = y3[0] * pscale; py[1] = y3[1] * pscale; py[2] = y3[2] * pscale; py[3] = 0.0f * pscale;
  float c0[3], c1[3], t[4], d[3];
  project(cam, coord, c0);
  t[0] = coord[0] + px[0]; t[1] = coord[1] + px[1]; t[2] = coord[2] + px[2]; t[3] = coord[3] + px[3];
  project(cam, t, c1);
  d[0] = c1[0] - c0[0]; d[1] = c1[1] - c0[1]; d[2] = c1[2] - c0[2];
  const float xdis = sqrtf(dot3(d, d));
  t[0] = coord[0] + py[0]; t[1] = coord[1] + py[1]; t[2] = coord[2] + py[2]; t[3] = coord[3] + py[3];
  project(cam, t, c1);
  d[0] = c1[0] - c0[0]; d[1] = c1[1] - c0[1]; d[2] = c1[2] - c0[2];
  const float ydis = sqrtf(dot3(d, d));
#pragma unroll
  for (int k = 0; k < 4; ++k) { px[k] /= xdis; py[k] /= ydis; }
}

// ---------------------------------------------------------------------------------------------------
// per-view sampling window (everything grabTex decides before it touches a texel, optim.cpp:815-846)
// ---------------------------------------------------------------------------------------------------
struct ViewWin {
  float lx, ly;   // top-left sample position at `newlevel`
  float dxx, dxy; // step per sample column
  float dyx, dyy; // step per sample row
  int newlevel;   // pyramid level to sample, -1 = view rejected
};

template <int WSIZE>
__device__ __forceinline__ ViewWin view_window(const SceneDev& s, const CamDev& cam, int index, const float* coord,
                                               const float* px, const float* py, const float* pz) {
  ViewWin w;
  w.newlevel = -1;
  w.lx = w.ly = w.dxx = w.dxy = w.dyx = w.dyy = 0.0f;
  float ray[4] = {cam.centre[0] - coord[0], cam.centre[1] - coord[1], cam.centre[2] - coord[2], cam.centre[3] - coord[3]};
  unitize4(ray);
  const float weight = smax(0.0f, dot4(ray, pz));
  if (weight < s.cos_angle1) return w;  // optim.cpp:823

  float center[3], dx[3], dy[3], t[4], q[3];
  project(cam, coord, center);
  t[0] = coord[0] + px[0]; t[1] = coord[1] + px[1]; t[2] = coord[2] + px[2]; t[3] = coord[3] + px[3];
  project(cam, t, q);
  dx[0] = q[0] - center[0]; dx[1] = q[1] - center[1]; dx[2] = q[2] - center[2];
  t[0] = coord[0] + py[0]; t[1] = coord[1] + py[1]; t[2] = coord[2] + py[2]; t[3] = coord[3] + py[3];
  project(cam, t, q);
  dy[0] = q[0] - center[0]; dy[1] = q[1] - center[1]; dy[2] = q[2] - center[2];

  // leveldif = clamp(floor(log(ratio)/Log2 + 0.5), -level, 2)  (optim.cpp:831-835): monotone in ratio,
  // so the host tabulated the float thresholds with the same libm the reference calls.
  const float ratio = (sqrtf(dot3(dx, dx)) + sqrtf(dot3(dy, dy))) / 2.0f;
  int leveldif = -s.level;
#pragma unroll
  for (int k = 0; k < kMaxLevels; ++k)
    if (k < s.n_level_thr && ratio >= s.level_thr[k]) ++leveldif;
  const int newlevel = s.level + leveldif;
  const float scale = (leveldif >= 0) ? (float)(1 << leveldif) : 1.0f / (float)(1 << (-leveldif));  // MyPow2
  center[0] /= scale; center[1] /= scale;
  dx[0] /= scale; dx[1] /= scale;
  dy[0] /= scale; dy[1] /= scale;

  // grabSafe (optim.cpp:783-805); written so that NaN positions are rejected instead of sampled
  constexpr int margin = WSIZE / 2;
  const float m = (float)margin;
  float minx, maxx, miny, maxy;
  {
    const float tl = center[0] - dx[0] * m - dy[0] * m, tr = center[0] + dx[0] * m - dy[0] * m;
    const float bl = center[0] - dx[0] * m + dy[0] * m, br = center[0] + dx[0] * m + dy[0] * m;
    minx = smin(tl, smin(tr, smin(bl, br)));
    maxx = smax(tl, smax(tr, smax(bl, br)));
  }
  {
    const float tl = center[1] - dx[1] * m - dy[1] * m, tr = center[1] + dx[1] * m - dy[1] * m;
    const float bl = center[1] - dx[1] * m + dy[1] * m, br = center[1] + dx[1] * m + dy[1] * m;
    miny = smin(tl, smin(tr, smin(bl, br)));
    maxy = smax(tl, smax(tr, smax(bl, br)));
  }
  const LevelDev lv = s.levels[index * s.nlevels + newlevel];
  const bool safe = (minx >= 3.0f) && (maxx < (float)(lv.w - 1 - 3)) && (miny >= 3.0f) && (maxy < (float)(lv.h - 1 - 3));
  if (!safe) return w;

  w.lx = center[0] - dx[0] * m - dy[0] * m;
  w.ly = center[1] - dx[1] * m - dy[1] * m;
  w.dxx = dx[0]; w.dxy = dx[1];
  w.dyx = dy[0]; w.dyy = dy[1];
  w.newlevel = newlevel;
  return w;
}

// CImage::getColor, bilinear branch (include/image/image.hpp:435-476) on the RGBA8 layout.
__device__ __forceinline__ void get_color(const LevelDev& lv, float x, float y, float* rgb) {
  const int lx = (int)x;
  const int ly = (int)y;
  const float dx1 = x - (float)lx, dx0 = 1.0f - dx1;
  const float dy1 = y - (float)ly, dy0 = 1.0f - dy1;
  const float f00 = dx0 * dy0, f01 = dx0 * dy1, f10 = dx1 * dy0, f11 = dx1 * dy1;
  const uchar4* p = lv.pix + (size_t)ly * lv.w + lx;
  const uchar4 a = __ldg(p);          // (lx  , ly  )
  const uchar4 b = __ldg(p + 1);      // (lx+1, ly  )
  const uchar4 c = __ldg(p + lv.w);   // (lx  , ly+1)
  const uchar4 d = __ldg(p + lv.w + 1);
  rgb[0] = ((float)a.x * f00 + (float)c.x * f01) + ((float)b.x * f10 + (float)d.x * f11);
  rgb[1] = ((float)a.y * f00 + (float)c.y * f01) + ((float)b.y * f10 + (float)d.y * f11);
  rgb[2] = ((float)a.z * f00 + (float)c.z * f01) + ((float)b.z * f10 + (float)d.z * f11);
}

// sample position of texel (row, col): the reference accumulates left += dy per row and v += dx per
// column (optim.cpp:850-859); the same sequential adds are replayed here.
template <int WSIZE>
__device__ __forceinline__ void sample_pos(const ViewWin& w, int row, int col, float& x, float& y) {
  x = w.lx; y = w.ly;
#pragma unroll
  for (int i = 0; i < WSIZE - 1; ++i)
    if (i < row) { x += w.dyx; y += w.dyy; }
#pragma unroll
  for (int i = 0; i < WSIZE - 1; ++i)
    if (i < col) { x += w.dxx; y += w.dxy; }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// A view's texture held across the warp: texel t = lane + 32*j, 3 channels each.
template <int WSIZE>
struct WarpTex {
  static constexpr int N = WSIZE * WSIZE;
  static constexpr int J = (N + 31) / 32;
  float v[J][3];
};

// grabTex's sampling loop + COptim::normalize (optim.cpp:846-860, 1031-1067), warp-cooperative.
// `w` must be warp-uniform and valid.
template <int WSIZE>
__device__ __forceinline__ void grab_and_normalize(const SceneDev& s, int index, const ViewWin& w, int lane,
                                                   WarpTex<WSIZE>& tex, bool do_normalize = true) {
  constexpr int N = WSIZE * WSIZE;
  constexpr int J = WarpTex<WSIZE>::J;
  const LevelDev lv = s.levels[index * s.nlevels + w.newlevel];
  float sum[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int t = lane + 32 * j;
    if (t < N) {
      float x, y;
      sample_pos<WSIZE>(w, t / WSIZE, t % WSIZE, x, y);
      get_color(lv, x, y, tex.v[j]);
      sum[0] += tex.v[j][0]; sum[1] += tex.v[j][1]; sum[2] += tex.v[j][2];
    } else {
      tex.v[j][0] = tex.v[j][1] = tex.v[j][2] = 0.0f;
    }
  }
  if (!do_normalize) return;
  const float ave0 = warp_sum(sum[0]) / (float)N;
  const float ave1 = warp_sum(sum[1]) / (float)N;
  const float ave2 = warp_sum(sum[2]) / (float)N;
  float sq = 0.0f;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    if (lane + 32 * j < N) {
      const float f0 = ave0 - tex.v[j][0], f1 = ave1 - tex.v[j][1], f2 = ave2 - tex.v[j][2];
      sq += f0 * f0 + f1 * f1 + f2 * f2;
    }
  }
  float sd = sqrtf(warp_sum(sq) / (float)(3 * N));
  if (sd == 0.0f) sd = 1.0f;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    if (lane + 32 * j < N) {
      tex.v[j][0] = (tex.v[j][0] - ave0) / sd;
      tex.v[j][1] = (tex.v[j][1] - ave1) / sd;
      tex.v[j][2] = (tex.v[j][2] - ave2) / sd;
    }
  }
}

// COptim::dot (optim.cpp:1069-1077)
template <int WSIZE>
__device__ __forceinline__ float tex_dot(const WarpTex<WSIZE>& a, const WarpTex<WSIZE>& b) {
  float acc = 0.0f;
#pragma unroll
  for (int j = 0; j < WarpTex<WSIZE>::J; ++j)
    acc += a.v[j][0] * b.v[j][0] + a.v[j][1] * b.v[j][1] + a.v[j][2] * b.v[j][2];
  return warp_sum(acc) / (float)(3 * WSIZE * WSIZE);
}

__device__ __forceinline__ float robustincc(float r) { return r / (1.0f + 3.0f * r); }      // optim.hpp:86-88
__device__ __forceinline__ float unrobustincc(float r) { return r / (1.0f - 3.0f * r); }    // optim.hpp:90-92

// ---------------------------------------------------------------------------------------------------
// per-patch refinement context (what refinePatchBFGS stores per thread, optim.cpp:584-596)
// ---------------------------------------------------------------------------------------------------
struct PatchCtx {
  float centre[4];
  float ray[4];
  float dscale;
  int size;         // min(tau, nimages)
  int nimages;
  int my_image;     // lane v < size: images[v]; else -1
  float my_weight;  // lane v: _weightsT[v]
  int ref;
};

// COptim::decode (optim.cpp:690-707).  Lane 0 evaluates sin/cos of angle1 in double, lane 1 of angle2
// (the reference's sin/cos resolve to the double libm functions); results are broadcast.
__device__ __forceinline__ void decode(const SceneDev& s, const PatchCtx& pc, const CamDev& refcam, const double* x,
                                       int lane, float* coord, float* normal) {
  const double sd = (double)pc.dscale * x[0];
#pragma unroll
  for (int k = 0; k < 4; ++k) coord[k] = pc.centre[k] + (float)((double)pc.ray[k] * sd);
  const float angle1 = (float)(x[1] * (double)s.ascale);
  const float angle2 = (float)(x[2] * (double)s.ascale);
  double sn = 0.0, cs = 0.0;
  if (lane < 2) sincos((double)(lane == 0 ? angle1 : angle2), &sn, &cs);
  const double s1 = __shfl_sync(kFull, sn, 0), c1 = __shfl_sync(kFull, cs, 0);
  const double s2 = __shfl_sync(kFull, sn, 1), c2 = __shfl_sync(kFull, cs, 1);
  const float fx = (float)(s1 * c2);
  const float fy = (float)s2;
  const float fz = (float)(-c1 * c2);
#pragma unroll
  for (int k = 0; k < 3; ++k) normal[k] = refcam.xaxis[k] * fx + refcam.yaxis[k] * fy + refcam.zaxis[k] * fz;
  normal[3] = 0.0f;
}

// COptim::encode (optim.cpp:660-688); uniform across the warp (called once per patch)
template <class Ctx>
__device__ __forceinline__ void encode(const SceneDev& s, const Ctx& pc, const CamDev& refcam, const float* coord,
                                       const float* normal, double* x) {
  const float d[4] = {coord[0] - pc.centre[0], coord[1] - pc.centre[1], coord[2] - pc.centre[2], coord[3] - pc.centre[3]};
  x[0] = (double)(dot4(d, pc.ray) / pc.dscale);
  float n3[3] = {normal[0], normal[1], normal[2]};
  if (normal[3] != 1.0f && normal[3] != 0.0f) { n3[0] /= normal[3]; n3[1] /= normal[3]; n3[2] /= normal[3]; }
  const float fx = dot3(refcam.xaxis, n3);
  const float fy = dot3(refcam.yaxis, n3);
  const float fz = dot3(refcam.zaxis, n3);
  x[2] = asin((double)smax(-1.0f, smin(1.0f, fy)));
  const float cosb = (float)cos(x[2]);
  if (cosb == 0.0f) {
    x[1] = 0.0;
  } else {
    const float sina = fx / cosb;
    const float cosa = -fz / cosb;
    x[1] = acos((double)smax(-1.0f, smin(1.0f, cosa)));
    if (sina < 0.0f) x[1] = -x[1];
  }
  x[1] = x[1] / (double)s.ascale;
  x[2] = x[2] / (double)s.ascale;
}

// Set up the per-patch context.  Lane v (< size) keeps image v and its computeINCC weight
// (COptim::computeUnits + setWeightsT, optim.cpp:446-471, 1146-1152).
__device__ __forceinline__ void patch_ctx_init(const SceneDev& s, PatchCtx& pc, const float* coord, const float* normal,
                                               const int32_t* images, int nimages, float dscale, int lane, CamDev& refcam) {
  pc.nimages = nimages;
  pc.size = nimages < s.tau ? nimages : s.tau;
  pc.dscale = dscale;
  pc.ref = images[0];
  load_cam(s, pc.ref, refcam);
#pragma unroll
  for (int k = 0; k < 4; ++k) { pc.centre[k] = coord[k]; pc.ray[k] = coord[k] - refcam.centre[k]; }
  unitize4(pc.ray);
  pc.my_image = lane < pc.size ? images[lane] : -1;
  float unit = 1.0f;
  if (lane < pc.size) {
    CamDev cam;
    load_cam(s, pc.my_image, cam);
    unit = get_unit(cam, s.level, coord);
    float ray[4] = {cam.centre[0] - coord[0], cam.centre[1] - coord[1], cam.centre[2] - coord[2], cam.centre[3] - coord[3]};
    unitize4(ray);
    const float denom = dot4(ray, normal);
    if (0.0f < denom) unit /= denom; else unit = 1073741824.0f;  // (float)(INT_MAX/2)
  }
  const float u0 = __shfl_sync(kFull, unit, 0);
  pc.my_weight = lane == 0 ? 1.0f : smin(1.0f, u0 / unit);
}

// The photo-consistency evaluation shared by my_f and computeINCC: decode-independent part.
//   mode 0: my_f          mean over valid views of robustincc(1 - NCC), 2.0 when too few (optim.cpp:557-574)
//   mode 1: computeINCC   weighted mean, robust            (optim.cpp:890-937)
//   mode 2: computeINCC   weighted mean, non-robust
template <int WSIZE>
__device__ __forceinline__ double photo_score(const SceneDev& s, const PatchCtx& pc, const CamDev& refcam, const float* coord,
                                              const float* normal, int lane, int mode) {
  float px[4], py[4];
  get_paxes(refcam, s.level, coord, normal, px, py);

  // lane v prepares view v's window
  ViewWin mine;
  mine.newlevel = -1;
  mine.lx = mine.ly = mine.dxx = mine.dxy = mine.dyx = mine.dyy = 0.0f;
  if (lane < pc.size) {
    CamDev cam;
    load_cam(s, pc.my_image, cam);
    mine = view_window<WSIZE>(s, cam, pc.my_image, coord, px, py, normal);
  }
  const unsigned validmask = __ballot_sync(kFull, mine.newlevel >= 0);
  if (!(validmask & 1u)) return 2.0;  // reference texture missing (optim.cpp:557-559, 890-892)

  WarpTex<WSIZE> ref, cur;
  double acc = 0.0;
  float totalweight = 0.0f;
  int denom = 0;
  for (int v = 0; v < pc.size; ++v) {
    if (!((validmask >> v) & 1u)) continue;
    ViewWin w;
    w.lx = __shfl_sync(kFull, mine.lx, v);   w.ly = __shfl_sync(kFull, mine.ly, v);
    w.dxx = __shfl_sync(kFull, mine.dxx, v); w.dxy = __shfl_sync(kFull, mine.dxy, v);
    w.dyx = __shfl_sync(kFull, mine.dyx, v); w.dyy = __shfl_sync(kFull, mine.dyy, v);
    w.newlevel = __shfl_sync(kFull, mine.newlevel, v);
    const int index = __shfl_sync(kFull, pc.my_image, v);
    if (v == 0) {
      grab_and_normalize<WSIZE>(s, index, w, lane, ref);
      continue;
    }
    grab_and_normalize<WSIZE>(s, index, w, lane, cur);
    const float d = tex_dot<WSIZE>(ref, cur);
    const float wv = __shfl_sync(kFull, pc.my_weight, v);
    if (mode == 0) {
      acc += (double)robustincc(1.0f - d);
      ++denom;
    } else if (mode == 1) {
      totalweight += wv;
      acc += (double)(robustincc(1.0f - d) * wv);
    } else {
      totalweight += wv;
      acc += (1.0 - (double)d) * (double)wv;
    }
  }
  if (mode == 0) {
    const int mininum = s.min_image_num < pc.size ? s.min_image_num : pc.size;
    if (denom < mininum - 1) return 2.0;
    return acc / (double)denom;
  }
  if (totalweight == 0.0f) return 2.0;
  return acc / (double)totalweight;
}

// my_f(x) (optim.cpp:507-578) when mode == 0; computeINCC at decode(x) when mode == 1 / 2.
// Kept as ONE call site inside the optimiser loop so the (large, unrolled) sampling code exists once
// per kernel and stays inside the instruction cache.
template <int WSIZE>
__device__ __forceinline__ double objective(const SceneDev& s, const PatchCtx& pc, const CamDev& refcam, const double* x, int lane,
                                            int mode, float* coord, float* normal) {
  decode(s, pc, refcam, x, lane, coord, normal);
  return photo_score<WSIZE>(s, pc, refcam, coord, normal, lane, mode);
}

// ---------------------------------------------------------------------------------------------------
// Bounded Nelder-Mead, n = 3; the written definition is oracle/nm3.h (same steps, same tie rules).
// Written as a state machine with a single evaluation site; the simplex lives in registers (all
// indexes are compile-time after unrolling).  All state is warp-uniform.
// On return: ok = stopped on the x-tolerance; x = best vertex; if ok, `final_score` holds
// computeINCC(robust) at the best vertex and coord/normal its decoded patch (optim.cpp:649-652).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct Simplex3 {
  double p[4][3];
  double f[4];
  // vertex `k` currently holds a new point: move it down past strictly worse predecessors
  // (stable: it stays behind equal values), considering only slots [0, k].
  __device__ __forceinline__ void insert(int k) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, tf = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q == k) { t0 = p[q][0]; t1 = p[q][1]; t2 = p[q][2]; tf = f[q]; }
    int pos = k;
#pragma unroll
    for (int q = 3; q >= 1; --q) {
      if (q == pos && tf < f[q - 1]) {
        p[q][0] = p[q - 1][0]; p[q][1] = p[q - 1][1]; p[q][2] = p[q - 1][2];
        f[q] = f[q - 1];
        pos = q - 1;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q == pos) { p[q][0] = t0; p[q][1] = t1; p[q][2] = t2; f[q] = tf; }
  }
  __device__ __forceinline__ void get(int k, double* x) const {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q == k) { x[0] = p[q][0]; x[1] = p[q][1]; x[2] = p[q][2]; }
  }
  __device__ __forceinline__ void set(int k, const double* x, double fx) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q == k) { p[q][0] = x[0]; p[q][1] = x[1]; p[q][2] = x[2]; f[q] = fx; }
  }
};

template <int WSIZE>
__device__ __forceinline__ bool nelder_mead3(const SceneDev& s, const PatchCtx& pc, const CamDev& refcam, int lane,
                                             double* x, int& evals, double& final_score, float* coord, float* normal) {
  const double lb1 = -23.99999, ub1 = 23.99999;  // optim.cpp:601-602; the depth variable is unbounded
  enum { INIT, REFLECT, EXPAND, CONTRACT, SHRINK, FINAL };
  Simplex3 sx;
  {
    const double x0 = x[0], x1 = clampd(x[1], lb1, ub1), x2 = clampd(x[2], lb1, ub1);
#pragma unroll
    for (int i = 0; i < 4; ++i) { sx.p[i][0] = x0; sx.p[i][1] = x1; sx.p[i][2] = x2; sx.f[i] = 1.0e300; }
    sx.p[1][0] = x0 + s.step;
    sx.p[2][1] = (x1 + s.step > ub1) ? x1 - s.step : x1 + s.step;
    sx.p[3][2] = (x2 + s.step > ub1) ? x2 - s.step : x2 + s.step;
  }
  int state = INIT, idx = 0, cnt = 0;
  double xt[3] = {sx.p[0][0], sx.p[0][1], sx.p[0][2]};
  double c[3] = {0, 0, 0}, xr[3] = {0, 0, 0}, fr = 0.0, fref = 0.0;
  bool ok = false;
  final_score = 2.0;

  for (;;) {
    if (state != FINAL && cnt >= s.maxeval) break;  // budget is checked before each evaluation (nm3.h)
    const double fx = objective<WSIZE>(s, pc, refcam, xt, lane, state == FINAL ? 1 : 0, coord, normal);
    if (state == FINAL) { final_score = fx; ok = true; break; }
    ++cnt;
    bool new_iter = false;
    if (state == INIT) {
      sx.set(idx, xt, fx);
      sx.insert(idx);
      ++idx;
      if (idx <= 3) sx.get(idx, xt); else new_iter = true;
    } else if (state == REFLECT) {
      fr = fx;
      if (fr < sx.f[0]) {
        xt[0] = c[0] + 2.0 * (c[0] - sx.p[3][0]);
        xt[1] = clampd(c[1] + 2.0 * (c[1] - sx.p[3][1]), lb1, ub1);
        xt[2] = clampd(c[2] + 2.0 * (c[2] - sx.p[3][2]), lb1, ub1);
        state = EXPAND;
      } else if (fr < sx.f[2]) {
        sx.set(3, xr, fr);
        sx.insert(3);
        new_iter = true;
      } else {
        if (fr < sx.f[3]) {
#pragma unroll
          for (int j = 0; j < 3; ++j) xt[j] = c[j] + 0.5 * (xr[j] - c[j]);
          fref = fr;
        } else {
#pragma unroll
          for (int j = 0; j < 3; ++j) xt[j] = c[j] + 0.5 * (sx.p[3][j] - c[j]);
          fref = sx.f[3];
        }
        state = CONTRACT;
      }
    } else if (state == EXPAND) {
      if (fx < fr) sx.set(3, xt, fx); else sx.set(3, xr, fr);
      sx.insert(3);
      new_iter = true;
    } else if (state == CONTRACT) {
      if (fx < fref) {
        sx.set(3, xt, fx);
        sx.insert(3);
        new_iter = true;
      } else {
#pragma unroll
        for (int i = 1; i <= 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) sx.p[i][j] = sx.p[0][j] + 0.5 * (sx.p[i][j] - sx.p[0][j]);
        state = SHRINK;
        idx = 1;
        sx.get(1, xt);
      }
    } else {  // SHRINK
      double keep[3];
      sx.get(idx, keep);
      sx.set(idx, keep, fx);
      ++idx;
      if (idx <= 3) {
        sx.get(idx, xt);
      } else {
        sx.insert(1); sx.insert(2); sx.insert(3);
        new_iter = true;
      }
    }
    if (new_iter) {
      double size = 0.0;
#pragma unroll
      for (int i = 1; i <= 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const double d = fabs(sx.p[i][j] - sx.p[0][j]);
          if (d > size) size = d;
        }
      if (size <= s.xtol) {
        state = FINAL;
        xt[0] = sx.p[0][0]; xt[1] = sx.p[0][1]; xt[2] = sx.p[0][2];
      } else {
#pragma unroll
        for (int j = 0; j < 3; ++j) c[j] = ((sx.p[0][j] + sx.p[1][j]) + sx.p[2][j]) / 3.0;
        xr[0] = c[0] + (c[0] - sx.p[3][0]);
        xr[1] = clampd(c[1] + (c[1] - sx.p[3][1]), lb1, ub1);
        xr[2] = clampd(c[2] + (c[2] - sx.p[3][2]), lb1, ub1);
        xt[0] = xr[0]; xt[1] = xr[1]; xt[2] = xr[2];
        state = REFLECT;
      }
    }
  }
  x[0] = sx.p[0][0]; x[1] = sx.p[0][1]; x[2] = sx.p[0][2];
  evals = cnt;
  return ok;
}

}  // namespace pmvsb
