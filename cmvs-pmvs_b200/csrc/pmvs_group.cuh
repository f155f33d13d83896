// cmvs-pmvs_b200/csrc/pmvs_group.cuh
//
// Second-generation mapping of the patch-optimisation path: EIGHT lanes own one patch, so a warp
// refines four patches in lock step.  Within a group, lane c (< wsize) owns column c of the
// wsize x wsize sampling window and walks its rows; lane v (< tau) prepares view v's window.
// Why: with one warp per patch (pmvs_device.cuh) ~35% of the issued instructions were warp-uniform
// bookkeeping (decode, patch axes, window set-up, simplex arithmetic) executed for 32 lanes on behalf
// of one patch, the 49 texels filled 49 of 64 lane slots, and every reduction was a 5-step shuffle
// chain.  Here the uniform work is shared by four patches, 49 of 56 slots are filled and reductions are
// 3 steps (ncu evidence: profiles/).
//
// Arithmetic policy is unchanged for everything that decides an integer (angle gate, level pick,
// grabSafe, window origin and steps: reference f32 operation order, IEEE div/sqrt, no FMA).  The texture
// statistics are evaluated as   NCC = sum_c sum_k (a_k - mean_a)(b_k - mean_b) / (147 sd_a sd_b)
// with two-pass means, i.e. without the reference's per-element division (optim.cpp:1061-1066); that
// differs from the reference by rounding only (~1e-7, bar 1e-4).
#pragma once
#include "pmvs_device.cuh"

namespace pmvsb {

constexpr int kGroup = 8;  // lanes per patch
// first pivot of a patch's reference view: mid-grey in the units of the sampling path (the atlas path reads byte / 255)
#define kPivot0 (TEX ? 0.5f : 127.5f)

// Instruction-count switches of the scoring loop, all ON by default (measured on B200, 262 144 patches x 5 views:
// 76.0 ms with all off, 68.9 ms with FAST_POS + FMA_INTERP, 65.2 ms with FOLD_PIVOT and 8 resident blocks per SM).
// They change roundings inside the texture statistics only (positions by a few ulp, colours by <= 2 ulp), never a
// texel choice by more than the continuous bilinear hand-over, and no integer decision: my_f / computeINCC stay
// within ~1e-6 of the reference (bar 1e-4).  The bit-exact texture path is k_grab_tex (pmvs_device.cuh).
#ifndef PMVS_FAST_POS
#define PMVS_FAST_POS 1
#endif
#ifndef PMVS_FMA_INTERP
#define PMVS_FMA_INTERP 1
#endif
#ifndef PMVS_FOLD_PIVOT
#define PMVS_FOLD_PIVOT 1
#endif

// PMVS_FAST_DIV: inside the optimiser loop (get_paxes_c, view_window_c, the per-view statistics) IEEE f32 division and
// square root (13- and 10-instruction expansions, 12 % of the kernel's issue slots in the round-1 ncu source view:
// profiles/r1_k_refine_g_lines_v3.txt) become MUFU.RCP * x and MUFU.SQRT / MUFU.RSQ (<= 2 ulp), and the dot products of
// project() are FMA chains.  The objective moves by ~1e-6 (bar 1e-4).  An integer decision of the loop (angle gate, level
// pick, grabSafe) can only flip where its operand lies within 2 ulp of the threshold; pre/postProcess, grabTex and every
// kernel with integer outputs keep the exact helpers of pmvs_device.cuh.
#ifndef PMVS_FAST_DIV
#define PMVS_FAST_DIV 1
#endif
#if PMVS_FAST_DIV
// .ftz forms: bare MUFU without the denormal pre/post-scaling (no operand here is ever subnormal: distances, depths,
// texture variances >= 1/147 or exactly 0)
__device__ __forceinline__ float grcp(float a) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float gdiv(float a, float b) { return a * grcp(b); }
__device__ __forceinline__ float gsqrt(float a) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float grsqrt(float a) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float gdot4(const float* u, const float* v) { return fmaf(u[3], v[3], fmaf(u[2], v[2], fmaf(u[1], v[1], u[0] * v[0]))); }
__device__ __forceinline__ float gdot3(const float* u, const float* v) { return fmaf(u[2], v[2], fmaf(u[1], v[1], u[0] * v[0])); }
#else
__device__ __forceinline__ float gdiv(float a, float b) { return fdiv(a, b); }
__device__ __forceinline__ float gsqrt(float a) { return fsqrt(a); }
__device__ __forceinline__ float gdot4(const float* u, const float* v) { return dot4(u, v); }
__device__ __forceinline__ float gdot3(const float* u, const float* v) { return dot3(u, v); }
#endif
// unitize / project / getUnit of pmvs_device.cuh on the g-helpers (same control flow)
__device__ __forceinline__ void gunitize3(float* v) {
  const float l = gdot3(v, v);
  if (l != 1.0f && l != 0.0f) {
#if PMVS_FAST_DIV
    const float r = grsqrt(l);
    v[0] *= r; v[1] *= r; v[2] *= r;
#else
    const float s = fsqrt(l);
    v[0] = fdiv(v[0], s); v[1] = fdiv(v[1], s); v[2] = fdiv(v[2], s);
#endif
  }
}
__device__ __forceinline__ void gunitize4(float* v) {
  const float l = gdot4(v, v);
  if (l != 1.0f && l != 0.0f) {
#if PMVS_FAST_DIV
    const float r = grsqrt(l);
    v[0] *= r; v[1] *= r; v[2] *= r; v[3] *= r;
#else
    const float s = fsqrt(l);
    v[0] = fdiv(v[0], s); v[1] = fdiv(v[1], s); v[2] = fdiv(v[2], s); v[3] = fdiv(v[3], s);
#endif
  }
}
__device__ __forceinline__ void gproject(const CamDev& cam, const float* X, float* o) {
#if PMVS_FAST_DIV
  o[0] = gdot4(cam.P[0], X);
  o[1] = gdot4(cam.P[1], X);
  o[2] = gdot4(cam.P[2], X);
  if (o[2] <= 0.0f) {
    o[0] = -65535.0f; o[1] = -65535.0f; o[2] = -1.0f;
    return;
  }
  const float r = grcp(o[2]);
  const float lim = 2147483648.0f;
  o[0] = smax(-lim, smin(lim, o[0] * r));
  o[1] = smax(-lim, smin(lim, o[1] * r));
  o[2] = 1.0f;
#else
  project(cam, X, o);
#endif
}
__device__ __forceinline__ float gget_unit(const CamDev& cam, int level, const float* X) {
#if PMVS_FAST_DIV
  const float d[4] = {X[0] - cam.centre[0], X[1] - cam.centre[1], X[2] - cam.centre[2], X[3] - cam.centre[3]};
  const float fz = gsqrt(gdot4(d, d));
  if (cam.ipscale == 0.0f) return 1.0f;
  return gdiv(fz * (float)(2 << level), cam.ipscale);
#else
  return get_unit(cam, level, X);
#endif
}

// a / b for a small integer-valued b with y = RN(1 / b): q = RN(a y), r = a - b q (exact in one FMA), q' = RN(q + r y) is the
// correctly rounded quotient (Markstein's final division step; holds for every finite a without over/underflow), i.e. the
// same bits as the IEEE division of oracle/nm3.h in 3 instructions instead of ~30
__device__ __forceinline__ double div_small(double a, double b, double y) {
  const double q = a * y;
  const double r = fma(-b, q, a);
  return fma(r, y, q);
}

// Reference-view scratch of the group kernels, one column per thread of a 128-thread CTA.
// PMVS_REFTEX_SOA = 1: three float planes [channel][row][thread] (12 B per sample, 3 LDS.32 / STS.32 with immediate plane
// offsets) instead of one float4 per sample (16 B, LDS.128): 10.5 KB instead of 14 KB per CTA, which lets 8 CTAs fit the
// 132 KB shared-memory carve-out and leaves ~32 KB more of the SM's 256 KB to the texture cache.
#ifndef PMVS_REFTEX_SOA
#define PMVS_REFTEX_SOA 1
#endif
#if PMVS_REFTEX_SOA && !(PMVS_FMA_INTERP && PMVS_FOLD_PIVOT)
#error "PMVS_REFTEX_SOA needs the default sampling path (PMVS_FMA_INTERP and PMVS_FOLD_PIVOT)"
#endif
template <int WSIZE>
struct RefTex {
#if PMVS_REFTEX_SOA
  static constexpr int kFloats = 3 * WSIZE * 128;
  static constexpr uint32_t kRowBytes = 128 * 4, kPlaneBytes = WSIZE * 128 * 4;
#else
  static constexpr int kFloats = 4 * WSIZE * 128;
  static constexpr uint32_t kRowBytes = 128 * 16, kPlaneBytes = 0;
#endif
  // shared-window address of this thread's column in row 0
  static __device__ __forceinline__ uint32_t column(const float* base) {
#if PMVS_REFTEX_SOA
    return (uint32_t)__cvta_generic_to_shared(base) + threadIdx.x * 4;
#else
    return (uint32_t)__cvta_generic_to_shared(base) + threadIdx.x * 16;
#endif
  }
  static __device__ __forceinline__ float4 load(uint32_t a) {
    float4 v;
#if PMVS_REFTEX_SOA
    asm volatile("ld.shared.f32 %0, [%3];\n\tld.shared.f32 %1, [%3+%4];\n\tld.shared.f32 %2, [%3+%5];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z) : "r"(a), "n"(kPlaneBytes), "n"(2 * kPlaneBytes) : "memory");
    v.w = 0.0f;
#else
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
#endif
    return v;
  }
  static __device__ __forceinline__ void store(uint32_t a, float x, float y, float z) {
#if PMVS_REFTEX_SOA
    asm volatile("st.shared.f32 [%0], %1;\n\tst.shared.f32 [%0+%4], %2;\n\tst.shared.f32 [%0+%5], %3;"
                 :: "r"(a), "f"(x), "f"(y), "f"(z), "n"(kPlaneBytes), "n"(2 * kPlaneBytes) : "memory");
#else
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %3};" :: "r"(a), "f"(x), "f"(y), "f"(z) : "memory");
#endif
  }
};

__device__ __forceinline__ float group_sum(float v) {
  v += __shfl_xor_sync(kFull, v, 4);
  v += __shfl_xor_sync(kFull, v, 2);
  v += __shfl_xor_sync(kFull, v, 1);
  return v;
}

// u8 -> f32 without the conversion (XU) pipe: splice the byte under the exponent of 2^23, subtract 2^23.
// `magic` = 0x4B000000 comes in as data (SceneDev::f32_2p23) so the selector can be PRMT's immediate.
__device__ __forceinline__ float byte_to_float(uint32_t word, uint32_t magic, uint32_t selector) {
  return __uint_as_float(__byte_perm(word, magic, selector)) - 8388608.0f;
}

// CImage::getColor (include/image/image.hpp:435-476) on RGBA8 words; same operation order as get_color().
__device__ __forceinline__ void get_color_fast(const uint32_t* __restrict__ pix, int w, uint32_t magic, float x, float y, float* rgb,
                                               float pv0 = 0.0f, float pv1 = 0.0f, float pv2 = 0.0f) {
  const int lx = (int)x;
  const int ly = (int)y;
#if PMVS_FMA_INTERP
  // the four weights from ONE product: f11 = dx1 dy1, f10 = dx1 - f11, f01 = dy1 - f11, f00 = (1 - dx1) - f01
  // (1 FMUL + 4 FADD instead of 2 FADD + 4 FMUL; each weight within 1 ulp of 1 of the reference's product)
  const float dx1 = x - (float)lx, dy1 = y - (float)ly;
  const float f11 = dx1 * dy1, f10 = dx1 - f11, f01 = dy1 - f11, f00 = (1.0f - dx1) - f01;
#else
  const float dx1 = x - (float)lx, dx0 = 1.0f - dx1;
  const float dy1 = y - (float)ly, dy0 = 1.0f - dy1;
  const float f00 = dx0 * dy0, f01 = dx0 * dy1, f10 = dx1 * dy0, f11 = dx1 * dy1;
#endif
  const uint32_t* p = pix + (ly * w + lx);  // < 2^31 texels per level
  const uint32_t a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + w), d = __ldg(p + w + 1);
#if PMVS_FMA_INTERP
  // contracted: 1 FMUL + 3 FFMA per channel instead of 4 FMUL + 3 FADD; differs from the reference's sum by rounding
  // only (<= 2 ulp of a value in [0, 255]), the texel choice is unchanged
  // (pv = minus the pivot the caller subtracts from every sample: folded into the chain's first FFMA)
  const float pv[3] = {pv0, pv1, pv2};
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const uint32_t sel = 0x7540u + ch;
    float v = fmaf(byte_to_float(d, magic, sel), f11, pv[ch]);
    v = fmaf(byte_to_float(b, magic, sel), f10, v);
    v = fmaf(byte_to_float(c, magic, sel), f01, v);
    rgb[ch] = fmaf(byte_to_float(a, magic, sel), f00, v);
  }
  return;
#endif
  rgb[0] = (byte_to_float(a, magic, 0x7540) * f00 + byte_to_float(c, magic, 0x7540) * f01) + (byte_to_float(b, magic, 0x7540) * f10 + byte_to_float(d, magic, 0x7540) * f11);
  rgb[1] = (byte_to_float(a, magic, 0x7541) * f00 + byte_to_float(c, magic, 0x7541) * f01) + (byte_to_float(b, magic, 0x7541) * f10 + byte_to_float(d, magic, 0x7541) * f11);
  rgb[2] = (byte_to_float(a, magic, 0x7542) * f00 + byte_to_float(c, magic, 0x7542) * f01) + (byte_to_float(b, magic, 0x7542) * f10 + byte_to_float(d, magic, 0x7542) * f11);
}

// The same bilinear sample through the texture unit: ONE tex2Dgather per channel returns the 2x2 footprint already as
// floats (byte / 255, exact; NCC is scale-invariant), so the 4 LDG + 12 PRMT + 12 FADD of get_color_fast become 3 TLD4 and
// the gathers leave the LSU data path.  The footprint is named by its integer centre (lxf + 1 + origin: exact in f32, no
// dependence on the sampler's fixed-point rounding), the weights come from the level-local coordinates as before.
// Component order on sm_100a (tools/probe/tex_gather_probe.cu): w = (x0,y0), z = (x1,y0), x = (x0,y1), y = (x1,y1).
__device__ __forceinline__ void get_color_gather(unsigned long long atlas, float ax1, float ay1, float x, float y, float* rgb,
                                                 float pv0, float pv1, float pv2) {
  const float lxf = truncf(x), lyf = truncf(y);   // x, y >= 3 (grabSafe): trunc == floor
  const float dx1 = x - lxf, dy1 = y - lyf;
  const float f11 = dx1 * dy1, f10 = dx1 - f11, f01 = dy1 - f11, f00 = (1.0f - dx1) - f01;
  const float gx = lxf + ax1, gy = lyf + ay1;
  const float pv[3] = {pv0, pv1, pv2};
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float4 t = tex2Dgather<float4>((cudaTextureObject_t)atlas, gx, gy, ch);
    float v = fmaf(t.y, f11, pv[ch]);
    v = fmaf(t.z, f10, v);
    v = fmaf(t.x, f01, v);
    rgb[ch] = fmaf(t.w, f00, v);
  }
}

// Split form for software pipelining: issue the three gathers of a sample now, interpolate later (the kernel is bound by
// TLD4 latency at 8 warps per scheduler, so the next row's gathers are put in flight before this row is consumed).
struct Foot { float4 r, g, b; float dx1, dy1; };
__device__ __forceinline__ void gather_issue(unsigned long long atlas, float ax1, float ay1, float x, float y, Foot& f) {
  const float lxf = truncf(x), lyf = truncf(y);
  f.dx1 = x - lxf; f.dy1 = y - lyf;
  const float gx = lxf + ax1, gy = lyf + ay1;
  f.r = tex2Dgather<float4>((cudaTextureObject_t)atlas, gx, gy, 0);
  f.g = tex2Dgather<float4>((cudaTextureObject_t)atlas, gx, gy, 1);
  f.b = tex2Dgather<float4>((cudaTextureObject_t)atlas, gx, gy, 2);
}
__device__ __forceinline__ void gather_finish(const Foot& f, float* rgb, float pv0, float pv1, float pv2) {
  const float f11 = f.dx1 * f.dy1, f10 = f.dx1 - f11, f01 = f.dy1 - f11, f00 = (1.0f - f.dx1) - f01;
  rgb[0] = fmaf(f.r.w, f00, fmaf(f.r.x, f01, fmaf(f.r.z, f10, fmaf(f.r.y, f11, pv0))));
  rgb[1] = fmaf(f.g.w, f00, fmaf(f.g.x, f01, fmaf(f.g.z, f10, fmaf(f.g.y, f11, pv1))));
  rgb[2] = fmaf(f.b.w, f00, fmaf(f.b.x, f01, fmaf(f.b.z, f10, fmaf(f.b.y, f11, pv2))));
}
#ifndef PMVS_TEX_PIPELINE
#define PMVS_TEX_PIPELINE 1
#endif

// ---- compact (code-size conscious) variants: the refine loop must stay inside the instruction cache ----
// Same operations in the same order as get_paxes() / view_window() of pmvs_device.cuh; the two axis
// projections run as a 2-trip loop around ONE inlined copy of project().
__device__ __forceinline__ void get_paxes_c(const CamDev& cam, int level, const float* coord, const float* normal,
                                            float* px, float* py) {
  const float pscale = gget_unit(cam, level, coord);
  const float n3[3] = {normal[0], normal[1], normal[2]};
  float y3[3], x3[3];
  cross3(n3, cam.xaxis, y3);
  gunitize3(y3);
  cross3(y3, n3, x3);
  px[0] = x3[0] * pscale; px[1] = x3[1] * pscale; px[2] = x3[2] * pscale; px[3] = 0.0f * pscale;
  py[0] = y3[0] * pscale; py[1] = y3[1] * pscale; py[2] = y3[2] * pscale; py[3] = 0.0f * pscale;
  float c0[3];
  gproject(cam, coord, c0);
  float dis0 = 1.0f, dis1 = 1.0f;
#pragma unroll 1
  for (int a = 0; a < 2; ++a) {
    float t[4], c1[3];
#pragma unroll
    for (int k = 0; k < 4; ++k) t[k] = coord[k] + (a == 0 ? px[k] : py[k]);
    gproject(cam, t, c1);
    const float d[3] = {c1[0] - c0[0], c1[1] - c0[1], c1[2] - c0[2]};
    const float dis = gsqrt(gdot3(d, d));
    if (a == 0) dis0 = dis; else dis1 = dis;
  }
#pragma unroll
#if PMVS_FAST_DIV
  const float r0 = grcp(dis0), r1 = grcp(dis1);
#pragma unroll
  for (int k = 0; k < 4; ++k) { px[k] *= r0; py[k] *= r1; }
#else
  for (int k = 0; k < 4; ++k) { px[k] = fdiv(px[k], dis0); py[k] = fdiv(py[k], dis1); }
#endif
}

template <int WSIZE>
__device__ __forceinline__ ViewWin view_window_c(const SceneDev& s, const CamDev& cam, int index, const float* coord,
                                                 const float* px, const float* py, const float* pz, float2* atlas_origin) {
  ViewWin w;
  *atlas_origin = make_float2(1.0f, 1.0f);
  w.newlevel = -1;
  w.lx = w.ly = w.dxx = w.dxy = w.dyx = w.dyy = 0.0f;
  float ray[4] = {cam.centre[0] - coord[0], cam.centre[1] - coord[1], cam.centre[2] - coord[2], cam.centre[3] - coord[3]};
  gunitize4(ray);
  const float weight = smax(0.0f, gdot4(ray, pz));
  if (weight < s.cos_angle1) return w;  // optim.cpp:823

  float center[3], dx[3] = {0, 0, 0}, dy[3] = {0, 0, 0};
  gproject(cam, coord, center);
  float nrm = 0.0f;
#pragma unroll 1
  for (int a = 0; a < 2; ++a) {
    float t[4], q[3];
#pragma unroll
    for (int k = 0; k < 4; ++k) t[k] = coord[k] + (a == 0 ? px[k] : py[k]);
    gproject(cam, t, q);
    const float d[3] = {q[0] - center[0], q[1] - center[1], q[2] - center[2]};
    const float len = gsqrt(gdot3(d, d));
    if (a == 0) { dx[0] = d[0]; dx[1] = d[1]; nrm = len; }
    else { dy[0] = d[0]; dy[1] = d[1]; nrm = nrm + len; }   // norm(dx) + norm(dy), optim.cpp:831
  }
  const float ratio = nrm / 2.0f;
  int leveldif = -s.level;
#pragma unroll 1
  for (int k = 0; k < s.n_level_thr; ++k)
    if (ratio >= s.level_thr[k]) ++leveldif;
  const int newlevel = s.level + leveldif;
  // MyPow2(leveldif) is a power of two: multiplying by its exact reciprocal 2^-leveldif (built from the exponent field)
  // gives the bits of the reference's division
  const float iscale = __int_as_float((127 - leveldif) << 23);
  center[0] *= iscale; center[1] *= iscale;
  dx[0] *= iscale; dx[1] *= iscale;
  dy[0] *= iscale; dy[1] *= iscale;

  constexpr float m = (float)(WSIZE / 2);
  float lo[2], hi[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float tl = center[k] - dx[k] * m - dy[k] * m, tr = center[k] + dx[k] * m - dy[k] * m;
    const float bl = center[k] - dx[k] * m + dy[k] * m, br = center[k] + dx[k] * m + dy[k] * m;
    lo[k] = smin(tl, smin(tr, smin(bl, br)));
    hi[k] = smax(tl, smax(tr, smax(bl, br)));
  }
  const LevelDev lv = s.levels[index * s.nlevels + newlevel];
  *atlas_origin = make_float2(lv.ax1, lv.ay1);   // travels with the window: no level-table load in the per-view loop
  const bool safe = (lo[0] >= 3.0f) && (hi[0] < (float)(lv.w - 1 - 3)) && (lo[1] >= 3.0f) && (hi[1] < (float)(lv.h - 1 - 3));
  if (!safe) return w;
  w.lx = center[0] - dx[0] * m - dy[0] * m;
  w.ly = center[1] - dx[1] * m - dy[1] * m;
  w.dxx = dx[0]; w.dxy = dx[1];
  w.dyx = dy[0]; w.dyy = dy[1];
  w.newlevel = newlevel;
  return w;
}

// Column offsets of one view for this lane.  The reference walks a row with `vftmp += dx` (optim.cpp:858); lane
// `gl` replays its first `gl` additions and pads with +0.0f (exact), so every row costs 2*(WSIZE-1) plain FADDs.
template <int WSIZE>
struct ColSteps {
  float sx[WSIZE - 1], sy[WSIZE - 1];
  __device__ __forceinline__ void set(const ViewWin& w, int gl) {
#if PMVS_FAST_POS
    sx[0] = w.dxx; sx[1] = w.dxy; sy[0] = (float)gl;   // step vector and this lane's column
#else
#pragma unroll
    for (int i = 0; i < WSIZE - 1; ++i) { sx[i] = i < gl ? w.dxx : 0.0f; sy[i] = i < gl ? w.dxy : 0.0f; }
#endif
  }
};

// One sample of grabTex's loop (optim.cpp:850-859): (bx, by) is the row's first sample (the reference's running
// `left`).  Idle lanes pass pix = SceneDev::dummy_pix and sample texel (0,0) of it: no branch in the row loop.
template <int WSIZE>
__device__ __forceinline__ void sample_row(const uint32_t* __restrict__ pix, int w, uint32_t magic, const ColSteps<WSIZE>& cs,
                                           float bx, float by, float* rgb, float pv0 = 0.0f, float pv1 = 0.0f, float pv2 = 0.0f) {
#if PMVS_FAST_POS
  // closed form bx + gl * dx (one FFMA per coordinate) instead of replaying the reference's running sum: the position
  // differs by a few ulp, bilinear interpolation is continuous across texel borders, so the colour moves by ~1e-6
  const float x = fmaf(cs.sx[0], cs.sy[0], bx), y = fmaf(cs.sx[1], cs.sy[0], by);
#else
  float x = bx, y = by;
#pragma unroll
  for (int i = 0; i < WSIZE - 1; ++i) { x += cs.sx[i]; y += cs.sy[i]; }
#endif
  get_color_fast(pix, w, magic, x, y, rgb, pv0, pv1, pv2);
}
template <int WSIZE>
__device__ __forceinline__ void sample_row_tex(unsigned long long atlas, float ax1, float ay1, const ColSteps<WSIZE>& cs, float bx, float by,
                                               float* rgb, float pv0, float pv1, float pv2) {
  const float x = fmaf(cs.sx[0], cs.sy[0], bx), y = fmaf(cs.sx[1], cs.sy[0], by);
  get_color_gather(atlas, ax1, ay1, x, y, rgb, pv0, pv1, pv2);
}

// per-group patch context; every lane of the group holds the same values except my_image / my_weight
struct GroupCtx {
  float centre[4];
  float ray[4];
  float dscale;
  int size;         // min(tau, nimages); 0 = group idle
  int nimages;
  int ref;
  int my_image;     // lane v < size: images[v]; else -1
  float my_weight;  // lane v: _weightsT[v]
};

// refinePatchBFGS's per-thread set-up (optim.cpp:584-596).  Called by one group at a time (divergent),
// so shuffles name only the group's lanes.
__device__ __forceinline__ void group_ctx_init(const SceneDev& s, GroupCtx& gc, const float* coord, const float* normal,
                                               const int32_t* images, int nimages, float dscale, int gl, unsigned gmask) {
  gc.nimages = nimages;
  gc.size = nimages < s.tau ? nimages : s.tau;
  gc.dscale = dscale;
  gc.ref = images[0];
  CamDev cam;
  load_cam(s, gc.ref, cam);
#pragma unroll
  for (int k = 0; k < 4; ++k) { gc.centre[k] = coord[k]; gc.ray[k] = coord[k] - cam.centre[k]; }
  unitize4(gc.ray);
  gc.my_image = gl < gc.size ? images[gl] : -1;
  float unit = 1.0f;
  if (gl < gc.size) {
    load_cam(s, gc.my_image, cam);
    unit = get_unit(cam, s.level, coord);
    float ray[4] = {cam.centre[0] - coord[0], cam.centre[1] - coord[1], cam.centre[2] - coord[2], cam.centre[3] - coord[3]};
    unitize4(ray);
    const float denom = dot4(ray, normal);
    if (0.0f < denom) unit /= denom; else unit = 1073741824.0f;  // (float)(INT_MAX/2)
  }
  const float u0 = __shfl_sync(gmask, unit, 0, kGroup);
  gc.my_weight = gl == 0 ? 1.0f : smin(1.0f, u0 / unit);
}

#ifndef PMVS_CUSTOM_SINCOS
#define PMVS_CUSTOM_SINCOS 0
#endif
#ifndef PMVS_ROW_UNROLL
#define PMVS_ROW_UNROLL 1
#endif
#define PMVS_STR2(x) #x
#define PMVS_STR(x) PMVS_STR2(x)
#define PMVS_UNROLL_ROWS _Pragma(PMVS_STR(unroll PMVS_ROW_UNROLL))

// sin and cos in double for |a| <= pi/2 (the optimiser's angles are boxed to +-23.99999 * pi/48): one
// quadrant fold plus the classic degree-13/12 minimax kernels (fdlibm's __kernel_sin / __kernel_cos
// coefficients), < 2 ulp.  libm-style sincos() costs ~10x the code for range reduction that can never
// trigger here; code size matters because the refine loop has to stay inside the instruction cache.
__device__ __forceinline__ void sincos_quadrant(double a, double* sn, double* cs) {
  const double pio4 = 0.78539816339744828, pio2_hi = 1.57079632673412561417e+00, pio2_lo = 6.07710050650619224932e-11;
  double r = a;
  int q = 0;
  if (a > pio4) { r = (a - pio2_hi) - pio2_lo; q = 1; }
  else if (a < -pio4) { r = (a + pio2_hi) + pio2_lo; q = -1; }
  const double z = r * r;
  double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = fma(z, ps, 2.75573137070700676789e-06);
  ps = fma(z, ps, -1.98412698298579493134e-04);
  ps = fma(z, ps, 8.33333333332248946124e-03);
  ps = fma(z, ps, -1.66666666666666324348e-01);
  const double sr = fma(z * r, ps, r);
  double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = fma(z, pc, -2.75573143513906633035e-07);
  pc = fma(z, pc, 2.48015872894767294178e-05);
  pc = fma(z, pc, -1.38888888888741095749e-03);
  pc = fma(z, pc, 4.16666666666666019037e-02);
  const double cr = 1.0 - (0.5 * z - z * z * pc);
  if (q == 0) { *sn = sr; *cs = cr; }
  else if (q == 1) { *sn = cr; *cs = -sr; }
  else { *sn = -cr; *cs = sr; }
}

// COptim::decode (optim.cpp:690-707) for four groups at once; lanes 0/1 of each group evaluate the
// double-precision sincos of angle1/angle2.
__device__ __forceinline__ void group_decode(const SceneDev& s, const GroupCtx& gc, const float* xaxis, const float* yaxis,
                                             const float* zaxis, const double* x, int gl, float* coord, float* normal) {
  const double sd = (double)gc.dscale * x[0];
#pragma unroll
  for (int k = 0; k < 4; ++k) coord[k] = gc.centre[k] + (float)((double)gc.ray[k] * sd);
  const float angle1 = (float)(x[1] * (double)s.ascale);
  const float angle2 = (float)(x[2] * (double)s.ascale);
  double sn = 0.0, cs = 0.0;
  if (gl < 2) {
    const double a = (double)(gl == 0 ? angle1 : angle2);
#if PMVS_CUSTOM_SINCOS
    if (fabs(a) <= 1.5707963267948968) sincos_quadrant(a, &sn, &cs); else sincos(a, &sn, &cs);
#else
    sincos(a, &sn, &cs);
#endif
  }
  const double s1 = __shfl_sync(kFull, sn, 0, kGroup), c1 = __shfl_sync(kFull, cs, 0, kGroup);
  const double s2 = __shfl_sync(kFull, sn, 1, kGroup), c2 = __shfl_sync(kFull, cs, 1, kGroup);
  const float fx = (float)(s1 * c2);
  const float fy = (float)s2;
  const float fz = (float)(-c1 * c2);
#pragma unroll
  for (int k = 0; k < 3; ++k) normal[k] = xaxis[k] * fx + yaxis[k] * fy + zaxis[k] * fz;
  normal[3] = 0.0f;
}

// my_f's scoring (mode 0, optim.cpp:530-574) or computeINCC (mode 1 robust / 2 plain, optim.cpp:865-938) of the
// patch (coord, normal) for the four groups of a warp.  Must be called by all 32 lanes; groups with
// gc.size == 0 idle through.
//
// Texture statistics are STREAMED so that no texture lives in registers and the row loop can stay rolled
// (code size: the refine loop has to fit the instruction cache; registers: occupancy):
//   reference view : one pass: samples the 7 rows pivoted by the previous evaluation's reference means (`ra_state`),
//                    stores a' in shared memory, accumulates sum a' and sum a'^2;
//   other views    : one pass; each sample b is pivoted by THIS evaluation's reference means (all views see the
//                    same surface, so b' = b - mean_ref is almost centred and the raw-moment variance
//                    sum b'^2 - (sum b')^2/N does not cancel); sum a' b' - (sum a')(sum b')/N is the covariance.
// `reftex` is the CTA's RefTex<WSIZE> scratch in shared memory (kFloats floats; blockDim.x = 128).
template <int WSIZE, bool TEX = false>
__device__ __forceinline__ double group_photo_score(const SceneDev& s, const GroupCtx& gc, const CamDev& refcam, const float* coord,
                                                    const float* normal, int gl, int g, int mode, float* reftex,
                                                    float* ra_state) {
  const bool live = gc.size > 0;
  float px[4], py[4];
  get_paxes_c(refcam, s.level, coord, normal, px, py);
  ViewWin mine;
  mine.newlevel = -1;
  mine.lx = mine.ly = mine.dxx = mine.dxy = mine.dyx = mine.dyy = 0.0f;
  float2 mine_origin = make_float2(1.0f, 1.0f);
  if (gl < gc.size) {
    CamDev cam;
    load_cam(s, gc.my_image, cam);
    mine = view_window_c<WSIZE>(s, cam, gc.my_image, coord, px, py, normal, &mine_origin);
  }
  const unsigned validmask = (__ballot_sync(kFull, mine.newlevel >= 0) >> (g * kGroup)) & 0xffu;
  const bool have_ref = live && (validmask & 1u);

  constexpr float N = (float)(WSIZE * WSIZE);
  constexpr float N3 = (float)(3 * WSIZE * WSIZE);
  // The reference view is pivoted by the reference means of the group's PREVIOUS evaluation (ra_state; the window moves
  // by a fraction of a pixel between evaluations, so the stored values a' = a - pivot are centred to ~1 %) and every
  // statistic is a raw moment of pivoted data: sum a'^2 - (sum a')^2 / N for the variance, sum a' b' - (sum a')(sum b') / N
  // for the covariance.  That is normalize()'s two passes (optim.cpp:1036-1053) without the second pass over the rows.
  float ra0 = ra_state[0], ra1 = ra_state[1], ra2 = ra_state[2];   // pivot, then this evaluation's reference means
  float rd0 = 0.f, rd1 = 0.f, rd2 = 0.f;   // sum of the stored (pivoted) reference values per channel
  float sq_ref = 1.0f;
  double acc = 0.0;
  float totalweight = 0.0f;
  int denom = 0;
  const bool col = gl < WSIZE;
  const int vmax = s.tau;
#pragma unroll 1
  for (int v = 0; v < vmax; ++v) {
    const bool on = have_ref && ((validmask >> v) & 1u);
    if (!__any_sync(kFull, on)) continue;
    ViewWin w;
    w.lx = __shfl_sync(kFull, mine.lx, v, kGroup);   w.ly = __shfl_sync(kFull, mine.ly, v, kGroup);
    w.dxx = __shfl_sync(kFull, mine.dxx, v, kGroup); w.dxy = __shfl_sync(kFull, mine.dxy, v, kGroup);
    w.dyx = __shfl_sync(kFull, mine.dyx, v, kGroup); w.dyy = __shfl_sync(kFull, mine.dyy, v, kGroup);
    float ax1 = 1.0f, ay1 = 1.0f;   // TEX: off groups gather the footprint at the atlas origin
    int index = 0;
    if (TEX) {
      ax1 = __shfl_sync(kFull, mine_origin.x, v, kGroup); ay1 = __shfl_sync(kFull, mine_origin.y, v, kGroup);
      w.newlevel = 0;
    } else {
      w.newlevel = __shfl_sync(kFull, mine.newlevel, v, kGroup);
      index = __shfl_sync(kFull, gc.my_image, v, kGroup);
    }
    const float wv = __shfl_sync(kFull, gc.my_weight, v, kGroup);
    const bool smp = on && col;
    // groups whose view is off (rejected view, idle group) read texel (0,0) of a dummy level.  The 8th lane of a live
    // group re-samples column WSIZE-1 (same addresses as its neighbour: no extra L1 wavefront, ncu: the dummy texel cost
    // one of 5.7 tag requests per gather) and is masked out of the sums by `smp`.
    const uint32_t* pix = reinterpret_cast<const uint32_t*>(s.dummy_pix);
    int lw = 0;
    float bx = 0.0f, by = 0.0f;
    ColSteps<WSIZE> cs;
    ViewWin wz = w;
    // (clamped texture addressing makes any coordinate safe: only the global-load path has to pin idle lanes to the dummy texel;
    //  an invalid view's window is all zeros anyway)
    if (!TEX && !on) { wz.dxx = wz.dxy = wz.dyx = wz.dyy = 0.0f; }
    if (on) {
      if (!TEX) {
        const LevelDev lv = s.levels[index * s.nlevels + w.newlevel];
        pix = reinterpret_cast<const uint32_t*>(lv.pix); lw = lv.w;
      }
      bx = w.lx; by = w.ly;
    }
    cs.set(wz, gl < WSIZE ? gl : WSIZE - 1);
    // ONE sampling loop for every view (a second copy of the loop body costs instruction-cache misses):
    // the reference view (v == 0) runs it with ra = 0, i.e. b = raw sample, and additionally stores the row.
    const bool is_ref = v == 0;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, q = 0.f, cr = 0.f;
    uint32_t rt = RefTex<WSIZE>::column(reftex);
#if PMVS_TEX_PIPELINE && PMVS_FMA_INTERP && PMVS_FOLD_PIVOT && PMVS_FAST_POS
    if (TEX) {
      static_assert(WSIZE % 2 == 1, "the pipelined loop consumes rows in pairs plus one");
      const float ox = cs.sx[0] * cs.sy[0], oy = cs.sx[1] * cs.sy[0];   // this lane's column offset along the row
      float px = bx + ox, py = by + oy;
      auto consume = [&](const Foot& f) {
        float rgb[3];
        gather_finish(f, rgb, -ra0, -ra1, -ra2);
        const float4 d = RefTex<WSIZE>::load(rt);
        if (is_ref) RefTex<WSIZE>::store(rt, rgb[0], rgb[1], rgb[2]);
        rt += RefTex<WSIZE>::kRowBytes;
        s0 += rgb[0]; s1 += rgb[1]; s2 += rgb[2];
        q = fmaf(rgb[0], rgb[0], q); q = fmaf(rgb[1], rgb[1], q); q = fmaf(rgb[2], rgb[2], q);
        cr = fmaf(d.x, rgb[0], cr); cr = fmaf(d.y, rgb[1], cr); cr = fmaf(d.z, rgb[2], cr);
      };
      Foot A, B;
      gather_issue(s.atlas, ax1, ay1, px, py, A);
#pragma unroll 1
      for (int row = 0; row < WSIZE - 1; row += 2) {
        px += wz.dyx; py += wz.dyy;
        gather_issue(s.atlas, ax1, ay1, px, py, B);
        consume(A);
        px += wz.dyx; py += wz.dyy;
        gather_issue(s.atlas, ax1, ay1, px, py, A);
        consume(B);
      }
      consume(A);
    } else
#endif
    PMVS_UNROLL_ROWS
    for (int row = 0; row < WSIZE; ++row) {
      float rgb[3];
#if PMVS_FMA_INTERP && PMVS_FOLD_PIVOT
      // rgb = sample - reference mean (ra = 0 for the reference view)
      if (TEX) sample_row_tex<WSIZE>(s.atlas, ax1, ay1, cs, bx, by, rgb, -ra0, -ra1, -ra2);
      else sample_row<WSIZE>(pix, lw, s.f32_2p23, cs, bx, by, rgb, -ra0, -ra1, -ra2);
      // explicit 32-bit shared addressing: one live register, no generic-to-shared rematerialisation in the loop
      const float4 d = RefTex<WSIZE>::load(rt);
      if (is_ref) RefTex<WSIZE>::store(rt, rgb[0], rgb[1], rgb[2]);
      rt += RefTex<WSIZE>::kRowBytes;
      const float b0 = rgb[0], b1 = rgb[1], b2 = rgb[2];
#else
      sample_row<WSIZE>(pix, lw, s.f32_2p23, cs, bx, by, rgb);
      float4* col = reinterpret_cast<float4*>(reftex) + threadIdx.x;   // (PMVS_REFTEX_SOA == 0 here)
      const float4 d = col[row * 128];
      const float b0 = rgb[0] - ra0, b1 = rgb[1] - ra1, b2 = rgb[2] - ra2;
      if (is_ref) col[row * 128] = make_float4(b0, b1, b2, 0.0f);
#endif
      s0 += b0; s1 += b1; s2 += b2;
      q = fmaf(b0, b0, q); q = fmaf(b1, b1, q); q = fmaf(b2, b2, q);
      cr = fmaf(d.x, b0, cr); cr = fmaf(d.y, b1, cr); cr = fmaf(d.z, b2, cr);
      bx += wz.dyx; by += wz.dyy;
    }
    s0 = group_sum(smp ? s0 : 0.0f); s1 = group_sum(smp ? s1 : 0.0f); s2 = group_sum(smp ? s2 : 0.0f);
    if (is_ref) {
      q = group_sum(smp ? q : 0.0f);
      rd0 = s0; rd1 = s1; rd2 = s2;
      sq_ref = fmaxf(fmaf(-(1.0f / N), fmaf(s2, s2, fmaf(s1, s1, s0 * s0)), q), 0.0f);
      if (sq_ref <= 1.0e-6f * q) sq_ref = 0.0f;   // constant texture up to the rounding of the raw moments: sd = 0 -> 1 (optim.cpp:1058)
      ra0 = fmaf(s0, 1.0f / N, ra0); ra1 = fmaf(s1, 1.0f / N, ra1); ra2 = fmaf(s2, 1.0f / N, ra2);   // the true channel means
      continue;
    }
    q = group_sum(smp ? q : 0.0f); cr = group_sum(smp ? cr : 0.0f);
    if (on) {
      // sd = sqrt(sum dev^2 / 147), 0 -> 1 (optim.cpp:1055-1059); dot = sum(t_ref t_cur) / 147 (optim.cpp:1069-1077)
#if PMVS_FAST_DIV
      // d = cross / (147 sd_a sd_b) with sd = sqrt(sq / 147), 0 -> 1, i.e. cross * rsqrt(A B) with A = sq or 147
      float sq_cur = fmaxf(fmaf(-(1.0f / N), fmaf(s2, s2, fmaf(s1, s1, s0 * s0)), q), 0.0f);
      if (sq_cur <= 1.0e-6f * q) sq_cur = 0.0f;
      const float cross = fmaf(-(1.0f / N), fmaf(s2, rd2, fmaf(s1, rd1, s0 * rd0)), cr);
      const float A = sq_ref == 0.0f ? N3 : sq_ref, B = sq_cur == 0.0f ? N3 : sq_cur;
      const float d = cross * grsqrt(A * B);
      const float r1 = 1.0f - d;
      const float rob = gdiv(r1, fmaf(3.0f, r1, 1.0f));
#else
      const float sq_cur = fmaxf(q - (s0 * s0 + s1 * s1 + s2 * s2) / N, 0.0f);
      const float cross = cr - (s0 * rd0 + s1 * rd1 + s2 * rd2) / N;
      float sda = fsqrt(fdiv(sq_ref, N3)), sdb = fsqrt(fdiv(sq_cur, N3));
      if (sda == 0.0f) sda = 1.0f;
      if (sdb == 0.0f) sdb = 1.0f;
      const float d = fdiv(fdiv(cross, sda * sdb), N3);
      const float rob = robustincc(1.0f - d);
#endif
      if (mode == 0) {
        acc += (double)rob;
        ++denom;
      } else if (mode == 1) {
        totalweight += wv;
        acc += (double)(rob * wv);
      } else {
        totalweight += wv;
        acc += (1.0 - (double)d) * (double)wv;
      }
    }
  }
  if (have_ref) { ra_state[0] = ra0; ra_state[1] = ra1; ra_state[2] = ra2; }
  if (!have_ref) return 2.0;
  if (mode == 0) {
    const int mininum = s.min_image_num < gc.size ? s.min_image_num : gc.size;
    if (denom < mininum - 1) return 2.0;
    if (denom >= 1 && denom <= 7) {   // tau - 1 views at most in practice: same bits as the division, without its expansion
      const double y = denom == 1 ? 1.0 : denom == 2 ? 0.5 : denom == 3 ? 1.0 / 3.0 : denom == 4 ? 0.25 : denom == 5 ? 0.2 : denom == 6 ? 1.0 / 6.0 : 1.0 / 7.0;
      return div_small(acc, (double)denom, y);
    }
    return acc / (double)denom;
  }
  if (gc.nimages < 2) return 2.0;
  if (totalweight == 0.0f) return 2.0;
  return acc / (double)totalweight;
}

// my_f(x) / computeINCC at decode(x); coord/normal receive the decoded patch.
template <int WSIZE, bool TEX = false>
__device__ __forceinline__ double group_objective(const SceneDev& s, const GroupCtx& gc, const double* x, int gl, int g, int mode,
                                                  float* coord, float* normal, float* reftex, float* ra_state) {
  CamDev refcam;
  load_cam(s, gc.size > 0 ? gc.ref : 0, refcam);  // 128 B, L1-resident; not kept in registers across the loop
  group_decode(s, gc, refcam.xaxis, refcam.yaxis, refcam.zaxis, x, gl, coord, normal);
  return group_photo_score<WSIZE, TEX>(s, gc, refcam, coord, normal, gl, g, mode, reftex, ra_state);
}

// The Nelder-Mead state of one group, kept in SHARED memory (216 B per patch) and advanced by the group's
// leader lane only: the simplex costs no registers in the sampling loop and its bookkeeping is a few
// dozen instructions of compact, loop-based code.  Same steps and tie rules as oracle/nm3.h.
struct NMShared {
  double p[4][3];
  double f[4];
  double xt[3], c[3], xr[3];
  double fr, fref;
  int state, idx, cnt, pad;
};
enum { NM_INIT, NM_REFLECT, NM_EXPAND, NM_CONTRACT, NM_SHRINK, NM_FINAL, NM_DONE_OK };

__device__ __forceinline__ void nm_start(NMShared& n, const double* x, double step) {
  const double lb1 = -23.99999, ub1 = 23.99999;  // optim.cpp:601-602; depth is unbounded
  const double x0[3] = {x[0], clampd(x[1], lb1, ub1), clampd(x[2], lb1, ub1)};
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 3; ++j) n.p[i][j] = x0[j];
    n.f[i] = 1.0e300;
  }
  n.p[1][0] = x0[0] + step;
  n.p[2][1] = (x0[1] + step > ub1) ? x0[1] - step : x0[1] + step;
  n.p[3][2] = (x0[2] + step > ub1) ? x0[2] - step : x0[2] + step;
  n.state = NM_INIT; n.idx = 0; n.cnt = 0;
  for (int j = 0; j < 3; ++j) { n.xt[j] = x0[j]; n.c[j] = 0.0; n.xr[j] = 0.0; }
  n.fr = 0.0; n.fref = 0.0;
}

// consume f(xt) and choose the next point (leader lane only).  Inlined into the kernel so that `n` is known to
// be shared memory (32-bit LDS/STS), with ONE copy of the insertion loop: every branch only decides which
// vertices [ins_lo, ins_hi] hold new points that must be moved down past strictly worse predecessors (stable).
__device__ __forceinline__ void nm_advance(NMShared& n, double fx, double xtol) {
  const double lb1 = -23.99999, ub1 = 23.99999;
  if (n.state == NM_FINAL) { n.fr = fx; n.state = NM_DONE_OK; return; }
  ++n.cnt;
  int ins_lo = 1, ins_hi = 0;   // empty range
  bool new_iter = false;
  const int st = n.state;
  if (st == NM_INIT) {
    n.f[n.idx] = fx;
    ins_lo = ins_hi = n.idx;
    ++n.idx;
    if (n.idx <= 3) { for (int j = 0; j < 3; ++j) n.xt[j] = n.p[n.idx][j]; }
    else new_iter = true;
  } else if (st == NM_REFLECT) {
    n.fr = fx;
    if (fx < n.f[0]) {
      n.xt[0] = n.c[0] + 2.0 * (n.c[0] - n.p[3][0]);
      n.xt[1] = clampd(n.c[1] + 2.0 * (n.c[1] - n.p[3][1]), lb1, ub1);
      n.xt[2] = clampd(n.c[2] + 2.0 * (n.c[2] - n.p[3][2]), lb1, ub1);
      n.state = NM_EXPAND;
    } else if (fx < n.f[2]) {
      for (int j = 0; j < 3; ++j) n.p[3][j] = n.xr[j];
      n.f[3] = fx;
      ins_lo = ins_hi = 3;
      new_iter = true;
    } else {
      const bool outside = fx < n.f[3];
      for (int j = 0; j < 3; ++j) n.xt[j] = n.c[j] + 0.5 * ((outside ? n.xr[j] : n.p[3][j]) - n.c[j]);
      n.fref = outside ? fx : n.f[3];
      n.state = NM_CONTRACT;
    }
  } else if (st == NM_EXPAND || st == NM_CONTRACT) {
    const bool take_xt = st == NM_EXPAND ? (fx < n.fr) : (fx < n.fref);
    if (take_xt || st == NM_EXPAND) {
      for (int j = 0; j < 3; ++j) n.p[3][j] = take_xt ? n.xt[j] : n.xr[j];
      n.f[3] = take_xt ? fx : n.fr;
      ins_lo = ins_hi = 3;
      new_iter = true;
    } else {  // failed contraction: shrink towards the best vertex, re-evaluate vertices 1..3 in order
      for (int i = 1; i <= 3; ++i)
        for (int j = 0; j < 3; ++j) n.p[i][j] = n.p[0][j] + 0.5 * (n.p[i][j] - n.p[0][j]);
      n.state = NM_SHRINK;
      n.idx = 1;
      for (int j = 0; j < 3; ++j) n.xt[j] = n.p[1][j];
    }
  } else {  // NM_SHRINK
    n.f[n.idx] = fx;
    ++n.idx;
    if (n.idx <= 3) { for (int j = 0; j < 3; ++j) n.xt[j] = n.p[n.idx][j]; }
    else { ins_lo = 1; ins_hi = 3; new_iter = true; }
  }
#pragma unroll 1
  for (int k = ins_lo; k <= ins_hi; ++k) {
    const double t0 = n.p[k][0], t1 = n.p[k][1], t2 = n.p[k][2], tf = n.f[k];
    int q = k;
#pragma unroll 1
    while (q > 0 && tf < n.f[q - 1]) {
      n.p[q][0] = n.p[q - 1][0]; n.p[q][1] = n.p[q - 1][1]; n.p[q][2] = n.p[q - 1][2];
      n.f[q] = n.f[q - 1];
      --q;
    }
    n.p[q][0] = t0; n.p[q][1] = t1; n.p[q][2] = t2; n.f[q] = tf;
  }
  if (new_iter) {
    double size = 0.0;
#pragma unroll 1
    for (int i = 1; i <= 3; ++i)
      for (int j = 0; j < 3; ++j) {
        const double d = fabs(n.p[i][j] - n.p[0][j]);
        if (d > size) size = d;
      }
    if (size <= xtol) {
      n.state = NM_FINAL;
      for (int j = 0; j < 3; ++j) n.xt[j] = n.p[0][j];
    } else {
      for (int j = 0; j < 3; ++j) n.c[j] = div_small((n.p[0][j] + n.p[1][j]) + n.p[2][j], 3.0, 1.0 / 3.0);   // == sum / 3.0
      n.xr[0] = n.c[0] + (n.c[0] - n.p[3][0]);
      n.xr[1] = clampd(n.c[1] + (n.c[1] - n.p[3][1]), lb1, ub1);
      n.xr[2] = clampd(n.c[2] + (n.c[2] - n.p[3][2]), lb1, ub1);
      for (int j = 0; j < 3; ++j) n.xt[j] = n.xr[j];
      n.state = NM_REFLECT;
    }
  }
}

// nm_advance spread over lanes 0..2 of the group: lane j owns coordinate j of every vertex (its column of p, xt, c, xr), all
// three hold the scalars (f values, state) in registers and take identical decisions, so the per-coordinate loops of
// nm_advance become one operation and the only exchange is the simplex extent (max over the three columns).  Same steps,
// same tie rules, same bits as nm_advance / oracle/nm3.h (checked bit for bit against the leader-only form on the GPU).
// The stable insertion runs as a fixed compare-exchange network on registers: vertex k in [ins_lo, ins_hi] moves down
// while it is strictly better than its predecessor.
// Measured on B200 (262 144 patches): 41.0 ms against 40.7 ms for the leader-only form, results bit-identical -- the three
// lanes' loads / stores of the state and the always-executed network cost what the shorter paths save.  Kept off.
#ifndef PMVS_NM_LANES
#define PMVS_NM_LANES 0
#endif
__device__ __forceinline__ void cswapd(bool sw, double& a, double& b) {
  const double t = sw ? b : a;
  b = sw ? a : b;
  a = t;
}
__device__ __forceinline__ void nm_advance_lanes(NMShared& n, double fx, double xtol, int j, unsigned mask3, int lane_base) {
  const double lb1 = -23.99999, ub1 = 23.99999;
  const bool boxed = j > 0;   // depth is unbounded, the two angles are boxed (optim.cpp:601-602)
  int state = n.state;
  if (state == NM_FINAL) {
    if (j == 0) { n.fr = fx; n.state = NM_DONE_OK; }
    return;
  }
  int idx = n.idx;
  const int cnt = n.cnt + 1;
  double f[4] = {n.f[0], n.f[1], n.f[2], n.f[3]};
  double p[4] = {n.p[0][j], n.p[1][j], n.p[2][j], n.p[3][j]};
  double fr = n.fr, fref = n.fref;
  double xt = n.xt[j], c = n.c[j], xr = n.xr[j];
  int ins_lo = 1, ins_hi = 0;   // empty range
  bool new_iter = false;
  if (state == NM_INIT || state == NM_SHRINK) {
#pragma unroll
    for (int k = 0; k < 4; ++k) f[k] = idx == k ? fx : f[k];
    if (state == NM_INIT) ins_lo = ins_hi = idx;
    ++idx;
    if (idx <= 3) xt = idx == 1 ? p[1] : (idx == 2 ? p[2] : p[3]);
    else { new_iter = true; if (state == NM_SHRINK) { ins_lo = 1; ins_hi = 3; } }
  } else if (state == NM_REFLECT) {
    fr = fx;
    if (fx < f[0]) {
      const double e = c + 2.0 * (c - p[3]);
      xt = boxed ? clampd(e, lb1, ub1) : e;
      state = NM_EXPAND;
    } else if (fx < f[2]) {
      p[3] = xr; f[3] = fx;
      ins_lo = ins_hi = 3;
      new_iter = true;
    } else {
      const bool outside = fx < f[3];
      xt = c + 0.5 * ((outside ? xr : p[3]) - c);
      fref = outside ? fx : f[3];
      state = NM_CONTRACT;
    }
  } else {  // NM_EXPAND or NM_CONTRACT
    const bool take_xt = state == NM_EXPAND ? (fx < fr) : (fx < fref);
    if (take_xt || state == NM_EXPAND) {
      p[3] = take_xt ? xt : xr;
      f[3] = take_xt ? fx : fr;
      ins_lo = ins_hi = 3;
      new_iter = true;
    } else {  // failed contraction: shrink towards the best vertex, re-evaluate vertices 1..3 in order
#pragma unroll
      for (int i = 1; i <= 3; ++i) p[i] = p[0] + 0.5 * (p[i] - p[0]);
      state = NM_SHRINK;
      idx = 1;
      xt = p[1];
    }
  }
#pragma unroll
  for (int k = 1; k <= 3; ++k) {
    bool moving = ins_lo <= k && k <= ins_hi;
#pragma unroll
    for (int q = k; q >= 1; --q) {
      const bool sw = moving && f[q] < f[q - 1];
      cswapd(sw, f[q], f[q - 1]);
      cswapd(sw, p[q], p[q - 1]);
      moving = sw;
    }
  }
  if (new_iter) {
    double mine = fabs(p[1] - p[0]);
    mine = fmax(mine, fabs(p[2] - p[0]));
    mine = fmax(mine, fabs(p[3] - p[0]));
    const double s0 = __shfl_sync(mask3, mine, lane_base), s1 = __shfl_sync(mask3, mine, lane_base + 1), s2 = __shfl_sync(mask3, mine, lane_base + 2);
    const double size = fmax(s0, fmax(s1, s2));
    if (size <= xtol) {
      state = NM_FINAL;
      xt = p[0];
    } else {
      c = div_small((p[0] + p[1]) + p[2], 3.0, 1.0 / 3.0);   // == sum / 3.0
      const double r = c + (c - p[3]);
      xr = boxed ? clampd(r, lb1, ub1) : r;
      xt = xr;
      state = NM_REFLECT;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) n.p[i][j] = p[i];
  n.xt[j] = xt; n.c[j] = c; n.xr[j] = xr;
  if (j == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) n.f[i] = f[i];
    n.fr = fr; n.fref = fref;
    n.state = state; n.idx = idx; n.cnt = cnt;
  }
}

}  // namespace pmvsb
