// cmvs-pmvs_b200/csrc/pmvs_b200.cu -- kernels + C ABI (include/pmvs_b200.h) for sm_100a.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false (see __graft_entry__.build).
// No CPU fallback: every entry point needs a live CUDA context.
#include "../../include/pmvs_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sched.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pmvs_device.cuh"
#include "pmvs_group.cuh"
#include "pmvs_select.cuh"
#include "pmvs_filter.cuh"
#include "pmvs_cells.cuh"
#include "pmvs_table.cuh"
#include "pmvs_seed.cuh"
#include "pmvs_features.cuh"

#ifndef PMVS_MINBLOCKS
#define PMVS_MINBLOCKS 8
#endif

using namespace pmvsb;

// =====================================================================================================
// kernels
// =====================================================================================================
namespace {

// interleaved RGB (host layout, CImage::_images) -> RGBA8 words
__global__ void k_rgb_to_rgba(const uint8_t* __restrict__ rgb, uchar4* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint8_t* p = rgb + 3 * i;
    out[i] = make_uchar4(p[0], p[1], p[2], 0);
  }
}
__global__ void k_rgba_to_rgb(const uchar4* __restrict__ in, uint8_t* __restrict__ rgb, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uchar4 v = in[i];
    rgb[3 * i] = v.x; rgb[3 * i + 1] = v.y; rgb[3 * i + 2] = v.z;
  }
}

// K1: one pyramid level (CImage::buildImage, source/image/image.cpp:228-325, filter 0).
// Weights [1 3 3 1]^2/64 in double with renormalisation over the taps inside the source; all partial
// sums are exact dyadic rationals, so (unsigned char)(int)floor(sum/denom + 0.5f) == (2S + D) / (2D)
// in integers.  One thread per output pixel; rows of the source are read as coalesced uchar4.
__global__ void k_pyr_down(const uchar4* __restrict__ src, int sw, int sh, uchar4* __restrict__ dst, int dw, int dh) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dw || y >= dh) return;
  const int wt[4] = {1, 3, 3, 1};
  int S0 = 0, S1 = 0, S2 = 0, D = 0;
#pragma unroll
  for (int j = -1; j < 3; ++j) {
    const int yt = 2 * y + j;
    if (yt < 0 || sh - 1 < yt) continue;
#pragma unroll
    for (int i = -1; i < 3; ++i) {
      const int xt = 2 * x + i;
      if (xt < 0 || sw - 1 < xt) continue;
      const int k = wt[j + 1] * wt[i + 1];
      const uchar4 p = __ldg(src + (size_t)yt * sw + xt);
      S0 += k * p.x; S1 += k * p.y; S2 += k * p.z;
      D += k;
    }
  }
  dst[(size_t)y * dw + x] = make_uchar4((unsigned char)((2 * S0 + D) / (2 * D)), (unsigned char)((2 * S1 + D) / (2 * D)),
                                        (unsigned char)((2 * S2 + D) / (2 * D)), 0);
}

__global__ void k_project(SceneDev s, int n, const float* __restrict__ coords, const int32_t* __restrict__ image, int level,
                          float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  CamDev cam;
  load_cam(s, image[i], cam);
  // s.cams holds P at the working level; other levels scale rows 0,1 by exact powers of two (camera.cpp:56-68)
  const int dl = level - s.level;
  const float sc = dl >= 0 ? 1.0f / (float)(1 << dl) : (float)(1 << (-dl));
#pragma unroll
  for (int k = 0; k < 4; ++k) { cam.P[0][k] *= sc; cam.P[1][k] *= sc; }
  const float X[4] = {coords[4 * i], coords[4 * i + 1], coords[4 * i + 2], coords[4 * i + 3]};
  float o[3];
  project(cam, X, o);
  out[3 * i] = o[0]; out[3 * i + 1] = o[1]; out[3 * i + 2] = o[2];
}

__device__ __forceinline__ void load_patch(const float* coords, const float* normals, int p, float* coord, float* normal) {
  const float4 c = __ldg(reinterpret_cast<const float4*>(coords) + p);
  const float4 n = __ldg(reinterpret_cast<const float4*>(normals) + p);
  coord[0] = c.x; coord[1] = c.y; coord[2] = c.z; coord[3] = c.w;
  normal[0] = n.x; normal[1] = n.y; normal[2] = n.z; normal[3] = n.w;
}

// parity hook: raw textures per (patch, view).  One warp per patch.
template <int WSIZE>
__global__ void k_grab_tex(SceneDev s, int P, int stride, const float* __restrict__ coords, const float* __restrict__ normals,
                           const int32_t* __restrict__ images, const int32_t* __restrict__ nimages, float* __restrict__ tex,
                           int32_t* __restrict__ flag, int32_t* __restrict__ newlevel) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= P) return;
  const int p = warp;
  float coord[4], normal[4];
  load_patch(coords, normals, p, coord, normal);
  const int n = nimages ? min(nimages[p], stride) : stride;
  const int32_t* im = images + (size_t)p * stride;
  CamDev refcam;
  load_cam(s, im[0], refcam);
  float px[4], py[4];
  get_paxes(refcam, s.level, coord, normal, px, py);
  constexpr int N = WSIZE * WSIZE;
  for (int v = 0; v < n; ++v) {
    CamDev cam;
    load_cam(s, im[v], cam);
    const ViewWin w = view_window<WSIZE>(s, cam, im[v], coord, px, py, normal);
    if (lane == 0) {
      flag[(size_t)p * stride + v] = w.newlevel >= 0 ? 0 : 1;
      newlevel[(size_t)p * stride + v] = w.newlevel;
    }
    if (w.newlevel < 0) continue;
    WarpTex<WSIZE> t;
    grab_and_normalize<WSIZE>(s, im[v], w, lane, t, false);
    float* o = tex + ((size_t)p * stride + v) * (3 * N);
#pragma unroll
    for (int j = 0; j < WarpTex<WSIZE>::J; ++j) {
      const int ti = lane + 32 * j;
      if (ti < N) { o[3 * ti] = t.v[j][0]; o[3 * ti + 1] = t.v[j][1]; o[3 * ti + 2] = t.v[j][2]; }
    }
  }
}

// my_f / computeINCC hooks.  mode 0: my_f(x); 1/2: computeINCC robust / plain.
template <int WSIZE>
__global__ void k_score(SceneDev s, int P, int stride, const float* __restrict__ coords, const float* __restrict__ normals,
                        const int32_t* __restrict__ images, const int32_t* __restrict__ nimages, const float* __restrict__ dscales,
                        const double* __restrict__ xs, int mode, double* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= P) return;
  const int p = warp;
  float coord[4], normal[4];
  load_patch(coords, normals, p, coord, normal);
  const int n = nimages ? min(nimages[p], stride) : stride;
  PatchCtx pc;
  CamDev refcam;
  patch_ctx_init(s, pc, coord, normal, images + (size_t)p * stride, n, dscales ? dscales[p] : 1.0f, lane, refcam);
  double f;
  if (mode == 0) {
    const double x[3] = {xs[3 * p], xs[3 * p + 1], xs[3 * p + 2]};
    float c2[4], n2[4];
    f = objective<WSIZE>(s, pc, refcam, x, lane, 0, c2, n2);
  } else {
    f = n < 2 ? 2.0 : photo_score<WSIZE>(s, pc, refcam, coord, normal, lane, mode);
  }
  if (lane == 0) out[p] = f;
}

// COptim::setINCCs, vector form (optim.cpp:709-744) over ALL images of the patch (not capped at tau).
template <int WSIZE>
__global__ void k_set_inccs(SceneDev s, int P, int stride, const float* __restrict__ coords, const float* __restrict__ normals,
                            const int32_t* __restrict__ images, const int32_t* __restrict__ nimages, int robust,
                            float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= P) return;
  const int p = warp;
  float coord[4], normal[4];
  load_patch(coords, normals, p, coord, normal);
  const int n = nimages ? min(nimages[p], stride) : stride;
  const int32_t* im = images + (size_t)p * stride;
  float* o = out + (size_t)p * stride;
  if (n <= 0) return;
  CamDev refcam;
  load_cam(s, im[0], refcam);
  float px[4], py[4];
  get_paxes(refcam, s.level, coord, normal, px, py);
  WarpTex<WSIZE> ref, cur;
  const ViewWin w0 = view_window<WSIZE>(s, refcam, im[0], coord, px, py, normal);
  if (w0.newlevel < 0) {
    for (int v = lane; v < n; v += 32) o[v] = 2.0f;
    return;
  }
  grab_and_normalize<WSIZE>(s, im[0], w0, lane, ref);
  if (lane == 0) o[0] = 0.0f;
  for (int v = 1; v < n; ++v) {
    CamDev cam;
    load_cam(s, im[v], cam);
    const ViewWin w = view_window<WSIZE>(s, cam, im[v], coord, px, py, normal);
    float r = 2.0f;
    if (w.newlevel >= 0) {
      grab_and_normalize<WSIZE>(s, im[v], w, lane, cur);
      const float d = tex_dot<WSIZE>(ref, cur);
      r = robust ? robustincc(1.0f - d) : 1.0f - d;
    }
    if (lane == 0) o[v] = r;
  }
}

// CPatchOrganizerS::setScales (patchOrganizerS.cpp:663-684); one thread per patch.
__global__ void k_set_scales(SceneDev s, int P, int stride, const float* __restrict__ coords, const int32_t* __restrict__ images,
                             const int32_t* __restrict__ nimages, float* __restrict__ dscale, float* __restrict__ ascale) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float coord[4] = {coords[4 * p], coords[4 * p + 1], coords[4 * p + 2], coords[4 * p + 3]};
  const int n = nimages ? min(nimages[p], stride) : stride;
  const int32_t* im = images + (size_t)p * stride;
  CamDev cam;
  load_cam(s, im[0], cam);
  const float unit = get_unit(cam, s.level, coord);
  const float unit2 = 2.0f * unit;
  float ray[4] = {coord[0] - cam.centre[0], coord[1] - cam.centre[1], coord[2] - cam.centre[2], coord[3] - cam.centre[3]};
  unitize4(ray);
  const int inum = s.tau < n ? s.tau : n;
  float ds = 0.0f;
  for (int i = 1; i < inum; ++i) {
    load_cam(s, im[i], cam);
    float a[3], b[3];
    project(cam, coord, a);
    const float t[4] = {coord[0] - ray[0] * unit2, coord[1] - ray[1] * unit2, coord[2] - ray[2] * unit2, coord[3] - ray[3] * unit2};
    project(cam, t, b);
    const float d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    ds += sqrtf(dot3(d, d));
  }
  ds /= (float)(inum - 1);
  ds = unit2 / ds;
  dscale[p] = ds;
  ascale[p] = (float)atan((double)(ds / (unit * (float)s.wsize / 2.0f)));
}

// K3: COptim::refinePatch for a whole frontier.  Persistent warps pull patches from a global counter
// (evaluation counts differ 2x between patches, so static assignment would leave SMs idle at the tail).
template <int WSIZE>
__global__ void __launch_bounds__(128) k_refine(SceneDev s, int P, int stride, float* __restrict__ coords, float* __restrict__ normals,
                                                const int32_t* __restrict__ images, const int32_t* __restrict__ nimages,
                                                const float* __restrict__ dscales, float* __restrict__ ncc_out,
                                                int32_t* __restrict__ evals_out, uint8_t* __restrict__ ok_out,
                                                int* __restrict__ counter, const int32_t* __restrict__ order) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    int p = 0;
    if (lane == 0) p = atomicAdd(counter, 1);
    p = __shfl_sync(kFull, p, 0);
    if (p >= P) return;
    if (order) p = order[p];

    float coord[4], normal[4];
    load_patch(coords, normals, p, coord, normal);
    const int n = nimages ? min(nimages[p], stride) : stride;
    PatchCtx pc;
    CamDev refcam;
    patch_ctx_init(s, pc, coord, normal, images + (size_t)p * stride, n, dscales[p], lane, refcam);

    double x[3];
    encode(s, pc, refcam, coord, normal, x);
    // NLOPT refuses a start outside the box; the reference clamps first (optim.cpp:629-634)
    x[1] = clampd(x[1], -23.99999, 23.99999);
    x[2] = clampd(x[2], -23.99999, 23.99999);
    int evals = 0;
    double incc = 2.0;
    float rc[4], rn[4];
    const bool ok = nelder_mead3<WSIZE>(s, pc, refcam, lane, x, evals, incc, rc, rn);
    float ncc = -1.0f;
    if (ok) {
      if (n < 2) incc = 2.0;  // computeINCC's early exit (optim.cpp:866)
      ncc = (float)(1.0 - (double)unrobustincc((float)incc));  // optim.cpp:652
      if (lane == 0) {
        reinterpret_cast<float4*>(coords)[p] = make_float4(rc[0], rc[1], rc[2], rc[3]);
        reinterpret_cast<float4*>(normals)[p] = make_float4(rn[0], rn[1], rn[2], rn[3]);
      }
    }
    if (lane == 0) {
      ncc_out[p] = ncc;
      evals_out[p] = evals;
      ok_out[p] = ok ? 1 : 0;
    }
  }
}


// ---- second-generation kernels: 8 lanes per patch, 4 patches per warp (pmvs_group.cuh) ---------------
template <int WSIZE, bool TEX>
__global__ void __launch_bounds__(128) k_score_g(SceneDev s, int P, int stride, const float* __restrict__ coords,
                                                 const float* __restrict__ normals, const int32_t* __restrict__ images,
                                                 const int32_t* __restrict__ nimages, const float* __restrict__ dscales,
                                                 const double* __restrict__ xs, int mode, double* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, g = lane >> 3, gl = lane & 7;
  const unsigned gmask = 0xffu << (g * kGroup);
  const int p = warp * 4 + g;
  GroupCtx gc;
  gc.size = 0; gc.nimages = 0; gc.ref = 0; gc.my_image = -1; gc.my_weight = 0.f; gc.dscale = 1.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) { gc.centre[k] = 0.f; gc.ray[k] = 0.f; }
  float coord[4] = {0, 0, 0, 1}, normal[4] = {0, 0, 1, 0};
  double x[3] = {0, 0, 0};
  if (p < P) {
    load_patch(coords, normals, p, coord, normal);
    const int n = nimages ? min(nimages[p], stride) : stride;
    group_ctx_init(s, gc, coord, normal, images + (size_t)p * stride, n, dscales ? dscales[p] : 1.0f, gl, gmask);
    if (mode == 0) { x[0] = xs[3 * p]; x[1] = xs[3 * p + 1]; x[2] = xs[3 * p + 2]; }
  }
  __syncwarp();
  __shared__ __align__(16) float reftex[RefTex<WSIZE>::kFloats];
  double f;
  float ra_state[3] = {kPivot0, kPivot0, kPivot0};
  if (mode == 0) {
    float c2[4], n2[4];
    f = group_objective<WSIZE, TEX>(s, gc, x, gl, g, 0, c2, n2, reftex, ra_state);
  } else {
    CamDev refcam;
    load_cam(s, gc.size > 0 ? gc.ref : 0, refcam);
    f = group_photo_score<WSIZE, TEX>(s, gc, refcam, coord, normal, gl, g, mode, reftex, ra_state);
  }
  if (p < P && gl == 0) out[p] = f;
}

// Processing order of a refine batch: counting sort of the patches by (reference image, 8x8-pixel tile of the projection at
// the working level).  Only the ORDER in which patches are handed out depends on it (results are written by patch index and
// patches are independent), and the order inside a tile is whatever the atomics give.
__global__ void k_order_count(SceneDev s, int P, int stride, const float* __restrict__ coords, const int32_t* __restrict__ images,
                              int tw, int th, int32_t* __restrict__ bin_of, int32_t* __restrict__ counts) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  int im = images[(size_t)p * stride];
  im = im < 0 ? 0 : (im >= s.num ? s.num - 1 : im);   // a bad index is reported by the refine kernel; here it only needs a bin
  const float4 c = __ldg(reinterpret_cast<const float4*>(coords) + p);
  const float X[4] = {c.x, c.y, c.z, c.w};
  CamDev cam;
  load_cam(s, im, cam);
  float o[3];
  project(cam, X, o);
  const int tx = (int)fminf(fmaxf(o[0] * 0.125f, 0.0f), (float)(tw - 1));   // fmaxf(NaN, 0) = 0
  const int ty = (int)fminf(fmaxf(o[1] * 0.125f, 0.0f), (float)(th - 1));
  const int bin = (im * th + ty) * tw + tx;
  bin_of[p] = bin;
  atomicAdd(counts + bin, 1);
}
__global__ void k_order_fill(int P, const int32_t* __restrict__ bin_of, int32_t* __restrict__ cursor, int32_t* __restrict__ order) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  order[atomicAdd(cursor + bin_of[p], 1)] = p;
}

// K3 (v2): COptim::refinePatch for a whole frontier.  Each 8-lane group pulls patches from a global counter
// and runs its own Nelder-Mead (state in shared memory, advanced by the group leader); the four groups of a
// warp evaluate their objectives in lock step.
// GATHER: the all-gather of the refined records is part of the kernel.  When a patch leaves the optimiser its group leader stores
// the 48-byte record (coord, normal, ncc, ok, evaluations) into this rank's slot of EVERY rank's mailbox -- its own memory and,
// through the CUDA IPC mappings, the other GPUs' memory over NVLink -- so the exchange trickles out underneath the ~125
// evaluations per patch of the groups still working and there is no collective after the kernel, only a flag (pmvsb_refine_batch_dev_gather).
struct RecDst {
  float4* rec[PMVSB_MAX_RANKS];   // this rank's record array in every rank's mailbox
  int n;
};
// Out of line and fed from memory on purpose: the call sits on the once-per-patch path, and neither the 16 mailbox pointers nor
// the record may cost the optimiser loop a register (the loop runs at the 64-register cap).  The leader lane has just written
// the patch's results, so it reads its own stores back.
__device__ __noinline__ void post_record(const RecDst* __restrict__ rd, int p, const float* coords, const float* normals, const float* ncc_out,
                                         const int32_t* evals_out, const uint8_t* ok_out) {
  const float4 c = reinterpret_cast<const float4*>(coords)[p], n = reinterpret_cast<const float4*>(normals)[p];
  const float4 t = make_float4(ncc_out[p], ok_out[p] ? 1.0f : 0.0f, (float)evals_out[p], 0.0f);
  const int nd = rd->n;
  for (int d = 0; d < nd; ++d) {
    float4* o = rd->rec[d] + (size_t)3 * p;
    o[0] = c; o[1] = n; o[2] = t;
  }
}
template <int WSIZE, bool TEX, bool GATHER = false>
__global__ void __launch_bounds__(128, PMVS_MINBLOCKS) k_refine_g(SceneDev s, int P, int stride, float* __restrict__ coords, float* __restrict__ normals,
                                                     const int32_t* __restrict__ images, const int32_t* __restrict__ nimages,
                                                     const float* __restrict__ dscales, float* __restrict__ ncc_out,
                                                     int32_t* __restrict__ evals_out, uint8_t* __restrict__ ok_out,
                                                     int* __restrict__ counter, const int32_t* __restrict__ order, const RecDst* __restrict__ rd = nullptr) {
  // counter[0] = next entry of `order` to hand out, counter[1] = set to 1 when a patch names an image outside [0, num).
  // `order` = the patches sorted by reference image and 8x8-pixel tile (k_order_*); nullptr = identity.  Groups draw single
  // entries from the global counter (evaluation counts differ 2x between patches: any coarser or static hand-out loses more
  // at the tail than it gains -- per-CTA chunks of 4 / 16 entries measured 6 % / 36 % slower on 262 144 patches), so the
  // ~19 000 patches in flight are a contiguous run of the order: neighbouring windows, L2- and partly L1-resident.
  __shared__ NMShared nms[4][4];  // [warp in CTA][group in warp]
  __shared__ __align__(16) float reftex[RefTex<WSIZE>::kFloats];  // pivoted reference-view samples (pmvs_group.cuh: RefTex)
  const int lane = threadIdx.x & 31, g = lane >> 3, gl = lane & 7;
  const unsigned gmask = 0xffu << (g * kGroup);
  NMShared& nm = nms[threadIdx.x >> 5][g];
  GroupCtx gc;
  gc.size = 0; gc.nimages = 0; gc.ref = 0; gc.my_image = -1; gc.my_weight = 0.f; gc.dscale = 1.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) { gc.centre[k] = 0.f; gc.ray[k] = 0.f; }
  if (gl == 0) {
    const double z[3] = {0.0, 0.0, 0.0};
    nm_start(nm, z, 1.0);
  }
  int p = -1;
  bool have = false, exhausted = false;
  float ra_state[3] = {kPivot0, kPivot0, kPivot0};   // reference-view pivot carried from one evaluation of a patch to the next
  for (;;) {
    if (!have && !exhausted) {  // this group needs a patch (divergent between groups; shuffles name the group only)
      int q = 0;
      if (gl == 0) q = atomicAdd(counter, 1);
      q = __shfl_sync(gmask, q, 0, kGroup);
      if (q >= P) {
        exhausted = true;
        gc.size = 0;
        if (gl == 0) { nm.xt[0] = nm.xt[1] = nm.xt[2] = 0.0; nm.state = NM_INIT; }
      } else {
        p = order ? order[q] : q;
        float coord[4], normal[4];
        load_patch(coords, normals, p, coord, normal);
        const int n = nimages ? min(nimages[p], stride) : stride;
        // image indexes come from the caller: a bad one must not turn into a wild load
        bool bad = n < 1;
        for (int k = gl; k < n && k < s.tau; k += kGroup) {
          const int im = images[(size_t)p * stride + k];
          bad |= (im < 0 || im >= s.num);
        }
        if (__any_sync(gmask, bad)) {
          if (gl == 0) {
            atomicExch(counter + 1, 1); ncc_out[p] = -1.0f; evals_out[p] = 0; ok_out[p] = 0;
            if (GATHER) post_record(rd, p, coords, normals, ncc_out, evals_out, ok_out);
          }
          // this group stays without a patch for one trip and asks again on the next
        } else {
          group_ctx_init(s, gc, coord, normal, images + (size_t)p * stride, n, dscales[p], gl, gmask);
          if (gl == 0) {
            CamDev refcam;
            load_cam(s, gc.ref, refcam);
            double x[3];
            encode(s, gc, refcam, coord, normal, x);
            nm_start(nm, x, s.step);  // clamps the start into the box as optim.cpp:629-634 does
          }
          have = true;
          ra_state[0] = ra_state[1] = ra_state[2] = kPivot0;   // results must not depend on the group's previous patch
        }
      }
    }
    __syncwarp();
    if (!__any_sync(kFull, have)) {
      if (__all_sync(kFull, exhausted)) break;
      continue;
    }

    const int mode = (have && nm.state == NM_FINAL) ? 1 : 0;
    const double xt[3] = {nm.xt[0], nm.xt[1], nm.xt[2]};
    float rc[4], rn[4];
    const double fx = group_objective<WSIZE, TEX>(s, gc, xt, gl, g, mode, rc, rn, reftex, ra_state);
    __syncwarp();
#if PMVS_NM_LANES
    if (have && gl < 3) nm_advance_lanes(nm, fx, s.xtol, gl, 0x7u << (g * kGroup), g * kGroup);
#else
    if (have && gl == 0) nm_advance(nm, fx, s.xtol);
#endif
    __syncwarp();
    if (have) {
      const int st = nm.state;
      const bool ok = st == NM_DONE_OK;
      if (ok || (st != NM_FINAL && nm.cnt >= s.maxeval)) {  // budget is checked before each evaluation (nm3.h)
        if (gl == 0) {
          float ncc = -1.0f;
          if (ok) {
            ncc = (float)(1.0 - (double)unrobustincc((float)nm.fr));  // optim.cpp:652
            reinterpret_cast<float4*>(coords)[p] = make_float4(rc[0], rc[1], rc[2], rc[3]);
            reinterpret_cast<float4*>(normals)[p] = make_float4(rn[0], rn[1], rn[2], rn[3]);
          }
          ncc_out[p] = ncc;
          evals_out[p] = nm.cnt;
          ok_out[p] = ok ? 1 : 0;
          if (GATHER) post_record(rd, p, coords, normals, ncc_out, evals_out, ok_out);   // (a failed patch keeps its input coordinates)
        }
        have = false;
        gc.size = 0;
      }
    }
    __syncwarp();
  }
}


__global__ void k_iota(int32_t* __restrict__ out, int first, int n, int value0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[first + i] = value0 + i;
}

// ---- masks / edges (CImage::_masks, _edges) -------------------------------------------------------------
// CImage::alloc's binarisation (source/image/image.cpp:146-176): masks keep 127 < v, edge files keep 1 < v
__global__ void k_map_binarise(const uint8_t* __restrict__ in, size_t n, int above, uint8_t* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = above < (int)in[i] ? 255 : 0;
}
// CImage::buildMask / buildEdge (image.cpp:326-393): a pixel is "in" when any of its (clamped) 2x2 parents is
__global__ void k_map_down(const uint8_t* __restrict__ src, int sw, int sh, uint8_t* __restrict__ dst, int dw, int dh) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dw || y >= dh) return;
  const int y0 = 2 * y, y1 = min(sh - 1, 2 * y + 1), x0 = 2 * x, x1 = min(sw - 1, 2 * x + 1);
  const int in = (src[(size_t)y0 * sw + x0] != 0) + (src[(size_t)y0 * sw + x1] != 0) + (src[(size_t)y1 * sw + x0] != 0) + (src[(size_t)y1 * sw + x1] != 0);
  dst[(size_t)y * dw + x] = 0 < in ? 255 : 0;
}
// CImage::setEdge (image.cpp:407-471): squared central differences summed over the channels (integers, exact in f32) ...
__global__ void k_edge_grad(const uchar4* __restrict__ pix, int w, int h, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  float v = 0.0f;
  if (1 <= y && y < h - 1 && 1 <= x && x < w - 1) {
    const uchar4 r = pix[(size_t)y * w + x + 1], l = pix[(size_t)y * w + x - 1], t = pix[(size_t)(y - 1) * w + x], b = pix[(size_t)(y + 1) * w + x];
    const int d[6] = {abs((int)r.x - (int)l.x), abs((int)b.x - (int)t.x), abs((int)r.y - (int)l.y), abs((int)b.y - (int)t.y),
                      abs((int)r.z - (int)l.z), abs((int)b.z - (int)t.z)};
#pragma unroll
    for (int k = 0; k < 6; ++k) v += (float)(d[k] * d[k]);
  }
  out[(size_t)y * w + x] = v;
}
// ... smoothed by CImage::filterG (image.cpp:1013-1060): taps outside the image skipped, sum divided by the weights used;
// sequential f32 multiply-add in tap order (-fmad=false)
template <bool VERTICAL>
__global__ void k_edge_smooth(const float* __restrict__ src, float* __restrict__ dst, int w, int h, const float* __restrict__ taps, int margin) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  float acc = 0.0f, denom = 0.0f;
  for (int j = -margin; j <= margin; ++j) {
    const int xt = VERTICAL ? x : x + j, yt = VERTICAL ? y + j : y;
    if (xt < 0 || w <= xt || yt < 0 || h <= yt) continue;
    const float f = __ldg(taps + j + margin);
    acc += f * src[(size_t)yt * w + xt];
    denom += f;
  }
  dst[(size_t)y * w + x] = fdiv(acc, denom);
}
__global__ void k_edge_threshold(const float* __restrict__ v, size_t n, float thr, uint8_t* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = thr < v[i] ? 255 : 0;
}
// the gate of postProcess / expandSub / collectCandidates for a batch of points: one warp per point
__global__ void k_mask_gate(SceneDev s, int n, const float* __restrict__ coords, uint8_t* __restrict__ inside) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float4 c4 = __ldg(reinterpret_cast<const float4*>(coords) + warp);
  const float X[4] = {c4.x, c4.y, c4.z, c4.w};
  const bool ok = mask_gate_warp(s, X, lane);
  if (lane == 0) inside[warp] = ok ? 1 : 0;
}
// COptim::removeImagesEdge (source/pmvs/optim.cpp:385-396): one thread per patch, list kept in order
__global__ void k_remove_images_edge(SceneDev s, int P, int stride, const float* __restrict__ coords, int32_t* __restrict__ images,
                                     int32_t* __restrict__ nimages) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float4 c4 = __ldg(reinterpret_cast<const float4*>(coords) + p);
  const float X[4] = {c4.x, c4.y, c4.z, c4.w};
  int32_t* im = images + (size_t)p * stride;
  const int n = min(nimages[p], stride);
  int out = 0;
  for (int i = 0; i < n; ++i) {
    const int image = im[i];
    CamDev cam;
    load_cam(s, image, cam);
    if (get_edge_img(s, cam, image, X)) im[out++] = image;
  }
  nimages[p] = out;
}

// ---- K4: visible-image-set selection (pmvs_select.cuh): one warp (= one CTA) per patch ----------------
template <int WSIZE, int MAXV>
__global__ void __launch_bounds__(32) k_pre_process(SceneDev s, SelectParams sp, int P, int stride, const float* __restrict__ coords,
                                                    const float* __restrict__ normals, int32_t* __restrict__ images,
                                                    int32_t* __restrict__ nimages, float* __restrict__ dscale,
                                                    float* __restrict__ ascale, int32_t* __restrict__ verdict) {
  __shared__ SelScratch<WSIZE, MAXV> sc;   // MAXV >= stride: the scratch is the occupancy limit (dispatch by stride)
  const int p = blockIdx.x, lane = threadIdx.x;
  if (p >= P) return;
  float coord[4], normal[4];
  load_patch(coords, normals, p, coord, normal);
  const int cap = min(stride, MAXV);
  int n = min(nimages[p], cap);
  for (int i = lane; i < n; i += 32) sc.images[i] = images[(size_t)p * stride + i];
  __syncwarp();
  float ds = 0.0f, as = 0.0f;
  int v = 1;
  if (n > 0) v = sel_pre_process<WSIZE>(s, sp, sc, n, cap, lane, coord, normal, ds, as);
  __syncwarp();
  for (int i = lane; i < n; i += 32) images[(size_t)p * stride + i] = sc.images[i];
  if (lane == 0) { nimages[p] = n; dscale[p] = ds; ascale[p] = as; verdict[p] = v; }
}

template <int WSIZE, int MAXV>
__global__ void __launch_bounds__(32) k_post_process(SceneDev s, SelectParams sp, int P, int stride, const float* __restrict__ coords,
                                                     const float* __restrict__ normals, const float* __restrict__ ncc,
                                                     int32_t* __restrict__ images, int32_t* __restrict__ nimages,
                                                     int32_t* __restrict__ grids, int32_t* __restrict__ timages,
                                                     float* __restrict__ tmp, int32_t* __restrict__ verdict) {
  __shared__ SelScratch<WSIZE, MAXV> sc;   // MAXV >= stride: the scratch is the occupancy limit (dispatch by stride)
  const int p = blockIdx.x, lane = threadIdx.x;
  if (p >= P) return;
  float coord[4], normal[4];
  load_patch(coords, normals, p, coord, normal);
  const int cap = min(stride, MAXV);
  int n = min(nimages[p], cap);
  for (int i = lane; i < n; i += 32) sc.images[i] = images[(size_t)p * stride + i];
  __syncwarp();
  int t = 0;
  float tm = 0.0f;
  const int v = sel_post_process<WSIZE>(s, sp, sc, n, cap, lane, coord, normal, ncc[p], grids + (size_t)2 * p * stride, t, tm);
  __syncwarp();
  for (int i = lane; i < n; i += 32) images[(size_t)p * stride + i] = sc.images[i];
  if (lane == 0) { nimages[p] = n; timages[p] = t; tmp[p] = tm; verdict[p] = v; }
}


// COptim::setRefImage + setGrids for a batch (filterExact's tail, filter.cpp:277-280)
// One launch per view-capacity class (MAXV = 16 / 32 / 64): a block serves its patch only if the list length falls in
// (LO, MAXV]; the small classes hold 4x / 2x more resident warps per SM (the scratch is the occupancy limit).
template <int WSIZE, int MAXV, int LO>
__global__ void __launch_bounds__(32) k_set_ref_image(SceneDev s, SelectParams sp, int P, int stride, const float* __restrict__ coords,
                                                      const float* __restrict__ normals, int32_t* __restrict__ images,
                                                      int32_t* __restrict__ nimages, int32_t* __restrict__ grids) {
  __shared__ SelScratch<WSIZE, MAXV> sc;
  const int p = blockIdx.x, lane = threadIdx.x;
  if (p >= P) return;
  const int cap = min(stride, kSelMaxViews);
  int n = min(nimages[p], cap);
  if (n > MAXV || (n <= LO && !(LO == 0 && n <= 0))) return;   // another class's patch (n <= 0 is handled by the first class)
  float coord[4], normal[4];
  load_patch(coords, normals, p, coord, normal);
  for (int i = lane; i < n; i += 32) sc.images[i] = images[(size_t)p * stride + i];
  __syncwarp();
  if (n > 0) n = sel_set_ref_image<WSIZE>(s, sc, n, lane, coord, normal);
  __syncwarp();
  if (n > 0) sel_set_grids(s, sc, n, lane, coord, grids + (size_t)2 * p * stride);
  for (int i = lane; i < n; i += 32) images[(size_t)p * stride + i] = sc.images[i];
  if (lane == 0) nimages[p] = n;
}

}  // namespace

// =====================================================================================================
// host side
// =====================================================================================================
struct HostCam {
  float P0[3][4];
  float centre[4], oaxis[4], xaxis[3], yaxis[3], zaxis[3], ipscale;
  bool set = false;
};
struct HostImage {
  std::vector<uchar4*> levels;
  std::vector<int> w, h;
  uint8_t* maps[2] = {nullptr, nullptr};   // working-level mask / edge map (CImage::_masks[level], _edges[level]), null = none
  bool set = false;
};

// grow-only device array; the table's fields keep their capacity across uploads and appends
template <typename T>
struct DVec {
  T* p = nullptr;
  size_t cap = 0;
};
struct StoreBufs {
  DVec<float> coords, normals, ncc, dscale;
  DVec<int32_t> timages, img_off, images, grids, entry_patch, vimg_off, vimages, vgrids, ventry_patch;
  DVec<int32_t> alt_voff, alt_vimages, alt_vgrids;                  // double buffers of pmvsb_store_update_vimages
  DVec<int32_t> cell_base, gw, gh, cell_off, cell_patch, vcell_off, vcell_patch, cursor, tile_sums, counts;
  DVec<unsigned long long> dp;
  // device-side reorganisation (pmvs_table.cuh): creation sequence numbers, second copies of every field, sort scratch
  DVec<int32_t> seq, alt_seq, alt_timages, alt_img_off, alt_images, alt_grids, first_cell, perm, bucket_off, rows, rows_n, row_cells;
  DVec<float> alt_coords, alt_normals, alt_ncc, alt_dscale;
};

struct pmvsb_ctx {
  int device = 0;
  int num = 0, tnum = 0, level = 1, csize = 2, wsize = 7, min_image_num = 3, tau = 0, nlevels = 0;
  float threshold = 0.7f, ncc_threshold = 0.7f, ncc_threshold_before = 0.4f;
  float angle_threshold0 = 0, angle_threshold1 = 0, max_angle_threshold = 0;
  double xtol = 1.0e-3, step = 1.0;   // the optimiser definition shared with oracle/nm3.h's callers
  int maxeval = 1000;
  std::vector<HostCam> cams;
  std::vector<HostImage> images;
  std::vector<std::vector<int32_t>> visdata2;
  CamDev* d_cams = nullptr;
  LevelDev* d_levels = nullptr;
  cudaArray_t atlas_array = nullptr;        // every (image, level) in one block-linear RGBA8 array (texture gather)
  cudaTextureObject_t atlas_tex = 0;
  int atlas_w = 0, atlas_h = 0;
  bool atlas_enabled = true;                // PMVSB_NO_ATLAS=1 in the environment: global-load gathers (A/B measurements)
  int* d_counter = nullptr;
  DVec<int32_t> order_bin, order_counts, order_idx;   // processing order of a refine batch (k_order_*), grow-only
  bool order_enabled = true;                // PMVSB_NO_ORDER=1: hand patches out in index order (A/B measurements)
  int32_t* d_vis_off = nullptr;
  const unsigned char** d_map_tab[2] = {nullptr, nullptr};   // per-image pointers to the working-level masks / edges
  // features of every image binned by cell (CSeed::_ppoints), for the seed candidate kernel
  std::vector<std::vector<float>> feat_xy;
  std::vector<std::vector<int32_t>> feat_type;
  bool feat_dirty = true;
  std::vector<int32_t> h_feat_base, h_fcell_base, h_fcell_off, h_flist;
  DVec<float> d_fxy;
  DVec<int32_t> d_ftype, d_fcell_base, d_fcell_off, d_flist, d_fgw, d_fgh;
  DVec<uint8_t> d_blocked;
  std::vector<uint8_t> h_blocked_shadow;   // what d_blocked holds (pmvsb_seed_candidates uploads only the blocks that changed)
  DVec<SeedHit> d_seed_out;
  // results of the last pmvsb_evaluate_batch, kept on the device until pmvsb_evaluate_fetch
  struct EvalOut {
    int P = 0, A = 0, E = 0, VE = 0, refined = 0;
    DVec<int32_t> verdict, index, timages, img_off, images, grids, vimg_off, vimages, vgrids;
    DVec<float> coords, normals, scal;   // scal = (ncc, dscale, ascale, tmp) per accepted candidate
  } ev, ev_all;                           // ev_all: the second set pmvsb_evaluate_allgather unpacks the whole wave into
  double exchanged_bytes = 0.0;
  std::vector<int32_t> bimages;
  int32_t* d_bimages = nullptr;
  // filter-stage patch table
  StoreDev store;
  StoreBufs sb;                     // grow-only device arrays behind `store`
  bool store_appended = false;      // coords-only append (pmvsb_depth_maps_add): lists are stale
  int store_entries = 0, store_ventries = 0, store_cells = 0;
  std::vector<int32_t> h_gw, h_gh, h_base;
  int depth_flag = 0;
  bool store_set = false, depth_built = false;
  int seq_next = 0;                 // next creation sequence number handed out by pmvsb_store_append
  std::vector<std::pair<size_t, void*>> pool;   // idle scratch blocks (size, pointer)
  char* arena = nullptr;          // grow-only device staging for the host-pointer entry points
  size_t arena_cap = 0, arena_used = 0;
  int32_t* d_vis_idx = nullptr;
  SelectParams select;
  bool finalized = false;
  ncclComm_t comm = nullptr;        // wave exchange across GPUs (one process per GPU)
  int comm_rank = 0, comm_world = 1;
  char* comm_buf = nullptr;
  size_t comm_cap = 0;
  // peer-memory wave exchange: one mailbox per rank (cudaMalloc), mapped into every other rank's address space through CUDA IPC;
  // ranks STORE their messages straight into each other's mailboxes over NVLink (pmvsb_peer_export / pmvsb_peer_open)
  struct PeerBox {
    bool on = false;
    char* own = nullptr;                       // this rank's mailbox: [kPeerHeader bytes of flags and sizes][world slots of slot_bytes]
    char* base[PMVSB_MAX_RANKS] = {nullptr};   // every rank's mailbox as seen from here (base[rank] == own)
    size_t slot_bytes = 0, needed = 0;
    uint32_t seq = 0;                          // wave number, the value the flags take
    int* d_done = nullptr;                     // block counter of k_wave_post
    void* d_rec = nullptr;                     // two RecDst tables of pmvsb_refine_batch_dev_gather (by call parity)
    uint32_t* h_board = nullptr;               // pinned copy of the mailbox header for the polls
  } peer;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool refine_timed = false;
  int sm_count = 148;
  int refine_blocks_per_sm = 0;
  uint64_t launches = 0;
  std::string err;
  SceneDev scene;
};

namespace {

int fail(pmvsb_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ctx, PMVSB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));           \
  } while (0)

// ---- host camera maths (CCamera::updateCamera, getOpticalCenter; COptim::setAxesScales) ----------
float h_dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
float h_dot4(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]; }
void h_cross3(const float* u, const float* v, float* o) {
  o[0] = u[1] * v[2] - v[1] * u[2];
  o[1] = -u[0] * v[2] + v[0] * u[2];
  o[2] = u[0] * v[1] - v[0] * u[1];
}
void h_unitize3(float* v) {
  const float l = h_dot3(v, v);
  if (l != 1.0f && l != 0.0f) {
    const float s = std::sqrt(l);
    v[0] /= s; v[1] /= s; v[2] /= s;
  }
}
void d_cross3(const double* u, const double* v, double* o) {
  o[0] = u[1] * v[2] - v[1] * u[2];
  o[1] = -u[0] * v[2] + v[0] * u[2];
  o[2] = u[0] * v[1] - v[0] * u[1];
}

void derive_camera(HostCam& c) {
  // optical axis: third row, xyz normalised, w scaled alike (source/image/camera.cpp:112-118)
  float oa[4] = {c.P0[2][0], c.P0[2][1], c.P0[2][2], 0.0f};
  const float len = std::sqrt(h_dot4(oa, oa));
  oa[3] = c.P0[2][3];
  for (int k = 0; k < 4; ++k) c.oaxis[k] = oa[k] / len;
  // optical centre (camera.cpp:138-175)
  if (c.P0[2][0] == 0.0f && c.P0[2][1] == 0.0f && c.P0[2][2] == 0.0f) {
    float v2[3];
    h_cross3(c.P0[0], c.P0[1], v2);
    h_unitize3(v2);
    c.centre[0] = v2[0]; c.centre[1] = v2[1]; c.centre[2] = v2[2]; c.centre[3] = 0.0f;
  } else {
    double A[3][3], b[3], adj[3][3];
    for (int y = 0; y < 3; ++y) {
      for (int x = 0; x < 3; ++x) A[y][x] = c.P0[y][x];
      b[y] = -c.P0[y][3];
    }
    d_cross3(A[1], A[2], adj[0]);
    d_cross3(A[2], A[0], adj[1]);
    d_cross3(A[0], A[1], adj[2]);
    const double det = adj[0][0] * A[0][0] + adj[0][1] * A[0][1] + adj[0][2] * A[0][2];
    for (int y = 0; y < 3; ++y) {
      const double r0 = adj[0][y] / det, r1 = adj[1][y] / det, r2 = adj[2][y] / det;  // row y of inverse
      c.centre[y] = (float)(r0 * b[0] + r1 * b[1] + r2 * b[2]);
    }
    c.centre[3] = 1.0f;
  }
  // optimiser axes and image-plane scale (source/pmvs/optim.cpp:43-64)
  c.zaxis[0] = c.oaxis[0]; c.zaxis[1] = c.oaxis[1]; c.zaxis[2] = c.oaxis[2];
  const float xa[3] = {c.P0[0][0], c.P0[0][1], c.P0[0][2]};
  h_cross3(c.zaxis, xa, c.yaxis);
  h_unitize3(c.yaxis);
  h_cross3(c.yaxis, c.zaxis, c.xaxis);
  const float xe[4] = {c.xaxis[0], c.xaxis[1], c.xaxis[2], 0.0f};
  const float ye[4] = {c.yaxis[0], c.yaxis[1], c.yaxis[2], 0.0f};
  c.ipscale = h_dot4(xe, c.P0[0]) + h_dot4(ye, c.P0[1]);
}

// the reference's pyramid-level decision for one ratio (optim.cpp:813, 831-835), host libm
int leveldif_of(float ratio, int level) {
  static const float Log2 = (float)std::log(2.0);
  const double lv = std::floor(std::log((double)ratio) / (double)Log2 + 0.5);
  int ld;
  if (!(lv > -1.0e9) || lv > 1.0e9) ld = INT32_MIN;  // what cvttsd2si yields for -inf / NaN / overflow
  else ld = (int)lv;
  ld = ld < 2 ? ld : 2;
  ld = -level < ld ? ld : -level;
  return ld;
}
// smallest positive float whose decision is >= target (decision is monotone in ratio)
float level_threshold(int target, int level) {
  uint32_t lo = 0x00000001u, hi = 0x7f7fffffu;  // positive finite floats are ordered like their bits
  auto val = [](uint32_t b) { float f; std::memcpy(&f, &b, 4); return f; };
  if (leveldif_of(val(hi), level) < target) return INFINITY;
  while (lo < hi) {
    const uint32_t mid = lo + (hi - lo) / 2;
    if (leveldif_of(val(mid), level) >= target) hi = mid; else lo = mid + 1;
  }
  return val(lo);
}


// ---- float bisection helpers: integer decisions that the reference takes through libm are tabulated -------
uint32_t float_key(float f) {  // order-preserving map float -> uint32
  uint32_t b; std::memcpy(&b, &f, 4);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
float key_float(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  float f; std::memcpy(&f, &b, 4);
  return f;
}
// the reference's angle of a clamped dot product (source/image/photoSetS.cpp:176-177)
float angle_of(float d) { return (float)std::acos((double)d); }

void fill_select(pmvsb_ctx* c) {
  SelectParams& sp = c->select;
  sp.vis_off = c->d_vis_off; sp.vis_idx = c->d_vis_idx;
  sp.overflow = c->d_counter + 2;
  sp.cos_angle0_f = (float)std::cos((double)c->angle_threshold0);          // optim.cpp:416
  sp.sort_threshold = (float)(1.0f - std::cos(10.0 * M_PI / 180.0));      // optim.cpp:287
  sp.ncc_threshold = c->ncc_threshold; sp.ncc_threshold_before = c->ncc_threshold_before;
  // checkAngles(minAngle = maxAngleThreshold, maxAngle = angleThreshold1) (optim.cpp:114): angle_of is non-increasing in d
  const float minA = c->max_angle_threshold, maxA = c->angle_threshold1;
  const uint32_t klo = float_key(-1.0f), khi = float_key(1.0f);
  {  // hi = largest d in [-1,1] with angle_of(d) > minA
    uint32_t lo = klo, hi = khi;
    if (!(angle_of(key_float(lo)) > minA)) sp.angle_dot_hi = -2.0f;  // empty
    else { while (lo < hi) { const uint32_t mid = lo + (hi - lo + 1) / 2; if (angle_of(key_float(mid)) > minA) lo = mid; else hi = mid - 1; } sp.angle_dot_hi = key_float(lo); }
  }
  {  // lo = smallest d in [-1,1] with angle_of(d) < maxA
    uint32_t lo = klo, hi = khi;
    if (!(angle_of(key_float(hi)) < maxA)) sp.angle_dot_lo = 2.0f;   // empty
    else { while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2; if (angle_of(key_float(mid)) < maxA) hi = mid; else lo = mid + 1; } sp.angle_dot_lo = key_float(lo); }
  }
}

void fill_scene(pmvsb_ctx* c) {
  SceneDev& s = c->scene;
  s.cams = c->d_cams;
  s.levels = c->d_levels;
  s.num = c->num; s.tnum = c->tnum; s.level = c->level; s.nlevels = c->nlevels; s.csize = c->csize;
  s.wsize = c->wsize; s.tau = c->tau; s.min_image_num = c->min_image_num;
  // weight < cos(angleThreshold1) with a double right-hand side <=> weight < smallest float >= that double
  const double ca = std::cos((double)c->angle_threshold1);
  float cf = (float)ca;
  if ((double)cf < ca) cf = std::nextafterf(cf, INFINITY);
  s.cos_angle1 = cf;
  s.n_level_thr = c->level + 2;
  for (int k = 0; k < kMaxLevels; ++k) s.level_thr[k] = INFINITY;
  for (int k = 0; k < s.n_level_thr && k < kMaxLevels; ++k) s.level_thr[k] = level_threshold(-c->level + k + 1, c->level);
  s.ascale = (float)(M_PI / 48.0f);
  s.xtol = c->xtol; s.step = c->step; s.maxeval = c->maxeval;
  s.f32_2p23 = 0x4B000000u;
  s.dummy_pix = c->images.empty() || c->images[0].levels.empty() ? nullptr : c->images[0].levels[0];
  s.atlas = (unsigned long long)c->atlas_tex;
  s.mask_lv = c->d_map_tab[0]; s.edge_lv = c->d_map_tab[1];
  s.bimages = c->d_bimages; s.n_bimages = c->d_bimages ? (int)c->bimages.size() : 0;
}

// Scratch device buffers of the host-pointer entry points come from a per-context pool: cudaMalloc / cudaFree cost
// ~0.1-1 ms each and a pipeline makes hundreds of small batched calls.  Every entry point synchronises its stream
// before returning, so a block released by one call is idle when the next call takes it.
thread_local pmvsb_ctx* g_current = nullptr;   // context of the entry point running on this thread

// Device memory comes from the device's stream-ordered pool with an unlimited release threshold (set in pmvsb_create): once
// the pool has grown, cudaMallocAsync / cudaFreeAsync are sub-microsecond list operations, while cudaMalloc / cudaFree go to the
// driver every time (tens of microseconds alone, milliseconds each when another process holds a context on the same GPU) and
// cudaFree synchronises the device.  A pipeline run makes thousands of scratch allocations.
cudaError_t dev_malloc(pmvsb_ctx* ctx, void** p, size_t bytes) {
  return cudaMallocAsync(p, bytes ? bytes : 1, ctx ? ctx->stream : (cudaStream_t)0);
}
void dev_free(pmvsb_ctx* ctx, void* p) {
  if (p) cudaFreeAsync(p, ctx ? ctx->stream : (cudaStream_t)0);
}

void* pool_take(pmvsb_ctx* ctx, size_t bytes, size_t& granted) {
  size_t want = 256;
  while (want < bytes) want <<= 1;
  granted = want;
  if (ctx) {
    for (size_t i = 0; i < ctx->pool.size(); ++i)
      if (ctx->pool[i].first == want) { void* p = ctx->pool[i].second; ctx->pool.erase(ctx->pool.begin() + i); return p; }
  }
  void* p = nullptr;
  if (dev_malloc(ctx, &p, want) != cudaSuccess) return nullptr;
  return p;
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t bytes = 0;
  pmvsb_ctx* owner = nullptr;
  ~DevBuf() {
    if (!p) return;
    if (owner && owner->pool.size() < 64) owner->pool.push_back({bytes, (void*)p}); else dev_free(owner, p);
  }
  cudaError_t alloc(size_t n) {
    owner = g_current;
    p = (T*)pool_take(owner, sizeof(T) * (n ? n : 1), bytes);
    return p ? cudaSuccess : cudaErrorMemoryAllocation;
  }
};

// peer-memory exchange (pmvsb_peer_*): unmap the peers' mailboxes / free this rank's
static void peer_unmap(pmvsb_ctx* ctx) {
  pmvsb_ctx::PeerBox& pb = ctx->peer;
  for (int k = 0; k < PMVSB_MAX_RANKS; ++k) {
    if (pb.base[k] && pb.base[k] != pb.own) cudaIpcCloseMemHandle(pb.base[k]);
    pb.base[k] = nullptr;
  }
  pb.on = false;
}
static void peer_free(pmvsb_ctx* ctx) {
  peer_unmap(ctx);
  pmvsb_ctx::PeerBox& pb = ctx->peer;
  cudaFree(pb.own); pb.own = nullptr;
  cudaFree(pb.d_done); pb.d_done = nullptr;
  cudaFree(pb.d_rec); pb.d_rec = nullptr;
  if (pb.h_board) cudaFreeHost(pb.h_board);
  pb.h_board = nullptr;
}

int check_ready(pmvsb_ctx* ctx) {
  if (!ctx) return PMVSB_EINVAL;
  g_current = ctx;
  if (!ctx->finalized) return fail(ctx, PMVSB_ESTATE, "scene not finalised: call pmvsb_finalize_scene first");
  if (ctx->wsize != 7 && ctx->wsize != 5 && ctx->wsize != 9) return fail(ctx, PMVSB_EINVAL, "wsize must be 5, 7 or 9");
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) return fail(ctx, PMVSB_ECUDA, cudaGetErrorString(e));
  return PMVSB_OK;
}

// group kernels (8 lanes per patch) cover wsize <= 8; wsize 9 falls back to the warp-per-patch kernels
// (the group kernels gather through the texture atlas when the scene has one, else through global loads)
#define DISPATCH_GROUP(ctx, KG, KW, gridg, gridw, block, ...)                                   \
  do {                                                                                          \
    const bool tex_ = (ctx)->scene.atlas != 0;                                                  \
    switch ((ctx)->wsize) {                                                                     \
      case 5: if (tex_) KG<5, true><<<gridg, block, 0, (ctx)->stream>>>(__VA_ARGS__);           \
              else KG<5, false><<<gridg, block, 0, (ctx)->stream>>>(__VA_ARGS__); break;        \
      case 9: KW<9><<<gridw, block, 0, (ctx)->stream>>>(__VA_ARGS__); break;                    \
      default: if (tex_) KG<7, true><<<gridg, block, 0, (ctx)->stream>>>(__VA_ARGS__);          \
               else KG<7, false><<<gridg, block, 0, (ctx)->stream>>>(__VA_ARGS__); break;       \
    }                                                                                           \
    ++(ctx)->launches;                                                                          \
  } while (0)

// selection kernels: scratch capacity = smallest of 16 / 32 / 48 / 64 views that holds `stride` (larger strides are capped at 64)
#define DISPATCH_VIEWS(ctx, stride, LAUNCH)                                                     \
  do {                                                                                          \
    if ((ctx)->wsize == 5) {                                                                    \
      if ((stride) <= 16) LAUNCH(5, 16); else if ((stride) <= 32) LAUNCH(5, 32);                \
      else if ((stride) <= 48) LAUNCH(5, 48); else LAUNCH(5, 64);                               \
    } else if ((ctx)->wsize == 9) {   /* 243 floats per view: 32 views fill the 48 KB of static shared memory */ \
      if ((stride) <= 16) LAUNCH(9, 16); else LAUNCH(9, 32);                                    \
    } else {                                                                                    \
      if ((stride) <= 16) LAUNCH(7, 16); else if ((stride) <= 32) LAUNCH(7, 32);                \
      else if ((stride) <= 48) LAUNCH(7, 48); else LAUNCH(7, 64);                               \
    }                                                                                           \
  } while (0)

#define DISPATCH_WSIZE(ctx, KERNEL, grid, block, ...)                                           \
  do {                                                                                          \
    switch ((ctx)->wsize) {                                                                     \
      case 5: KERNEL<5><<<grid, block, 0, (ctx)->stream>>>(__VA_ARGS__); break;                 \
      case 9: KERNEL<9><<<grid, block, 0, (ctx)->stream>>>(__VA_ARGS__); break;                 \
      default: KERNEL<7><<<grid, block, 0, (ctx)->stream>>>(__VA_ARGS__); break;                \
    }                                                                                           \
    ++(ctx)->launches;                                                                          \
  } while (0)

}  // namespace

// grow-only device arena: the batched host-pointer calls stage through it instead of cudaMalloc/cudaFree per call
static int arena_reserve(pmvsb_ctx* ctx, size_t bytes) {
  ctx->arena_used = 0;
  if (bytes <= ctx->arena_cap) return PMVSB_OK;
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->arena);
  ctx->arena = nullptr; ctx->arena_cap = 0;
  const size_t want = bytes + bytes / 4 + (1u << 20);
  CK(dev_malloc(ctx, (void**)&ctx->arena, want));
  ctx->arena_cap = want;
  return PMVSB_OK;
}
template <typename T>
static T* arena_take(pmvsb_ctx* ctx, size_t n) {
  const size_t off = (ctx->arena_used + 255) & ~(size_t)255;
  ctx->arena_used = off + sizeof(T) * n;
  return reinterpret_cast<T*>(ctx->arena + off);
}

// ---- NCCL, bound at run time: a single-GPU run never needs the library, and inside a Python process the copy that
// torch already loaded (same soname) is the one that gets used
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  const char* (*GetErrorString)(ncclResult_t);
};
static NcclApi* nccl_api() {
  static NcclApi api;
  static int state = 0;   // 0 untried, 1 ok, -1 unavailable
  if (state == 0) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    state = -1;
    if (h) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
      api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
      if (api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy && api.GetErrorString) state = 1;
    }
  }
  return state == 1 ? &api : nullptr;
}

// ---- resident patch table: grow-only arrays, cell lists built on the device -----------------------------------
template <typename T>
static int dvec_reserve(pmvsb_ctx* ctx, DVec<T>& v, size_t n, size_t keep = 0) {
  if (n <= v.cap && v.p) return PMVSB_OK;
  const size_t want = 2 * n + 4096;
  T* np = nullptr;
  CK(dev_malloc(ctx, (void**)&np, sizeof(T) * want));
  if (keep && v.p) CK(cudaMemcpyAsync(np, v.p, sizeof(T) * keep, cudaMemcpyDeviceToDevice, ctx->stream));
  dev_free(ctx, v.p);   // stream-ordered: after the copy above
  v.p = np; v.cap = want;
  return PMVSB_OK;
}
template <typename T>
static int dvec_put(pmvsb_ctx* ctx, DVec<T>& v, size_t at, const T* src, size_t n) {
  int r = dvec_reserve(ctx, v, at + n, at);
  if (r) return r;
  if (n) CK(cudaMemcpyAsync(v.p + at, src, sizeof(T) * n, cudaMemcpyHostToDevice, ctx->stream));
  return PMVSB_OK;
}
static void store_free(pmvsb_ctx* ctx) {
  StoreBufs& b = ctx->sb;
  for (DVec<float>* v : {&b.coords, &b.normals, &b.ncc, &b.dscale, &b.alt_coords, &b.alt_normals, &b.alt_ncc, &b.alt_dscale}) { cudaFree(v->p); v->p = nullptr; v->cap = 0; }
  for (DVec<int32_t>* v : {&b.seq, &b.alt_seq, &b.alt_timages, &b.alt_img_off, &b.alt_images, &b.alt_grids, &b.first_cell, &b.perm, &b.bucket_off, &b.rows,
                           &b.rows_n, &b.row_cells}) { cudaFree(v->p); v->p = nullptr; v->cap = 0; }
  for (DVec<int32_t>* v : {&b.timages, &b.img_off, &b.images, &b.grids, &b.entry_patch, &b.vimg_off, &b.vimages, &b.vgrids, &b.ventry_patch,
                           &b.alt_voff, &b.alt_vimages, &b.alt_vgrids, &b.cell_base, &b.gw, &b.gh, &b.cell_off, &b.cell_patch, &b.vcell_off,
                           &b.vcell_patch, &b.cursor, &b.tile_sums, &b.counts}) { cudaFree(v->p); v->p = nullptr; v->cap = 0; }
  cudaFree(b.dp.p); b.dp.p = nullptr; b.dp.cap = 0;
}
static void store_view(pmvsb_ctx* ctx) {   // device pointers may have moved: refresh the struct the kernels take
  StoreDev& st = ctx->store;
  const StoreBufs& b = ctx->sb;
  st.coords = b.coords.p; st.normals = b.normals.p; st.ncc = b.ncc.p; st.dscale = b.dscale.p; st.timages = b.timages.p;
  st.img_off = b.img_off.p; st.images = b.images.p; st.grids = b.grids.p; st.entry_patch = b.entry_patch.p;
  st.vimg_off = b.vimg_off.p; st.vimages = b.vimages.p; st.vgrids = b.vgrids.p;
  st.cell_base = b.cell_base.p; st.gw = b.gw.p; st.gh = b.gh.p;
  st.cell_off = b.cell_off.p; st.cell_patch = b.cell_patch.p; st.vcell_off = b.vcell_off.p; st.vcell_patch = b.vcell_patch.p;
  st.dp = b.dp.p;
}
// exclusive scan of n int32 in place (n includes the trailing total slot)
static int device_scan(pmvsb_ctx* ctx, int32_t* data, int n) {
  const int tiles = (n + kScanTile - 1) / kScanTile;
  int r = dvec_reserve(ctx, ctx->sb.tile_sums, (size_t)tiles);
  if (r) return r;
  k_scan_tile_sums<<<tiles, kScanThreads, 0, ctx->stream>>>(data, n, ctx->sb.tile_sums.p);
  k_scan_spine<<<1, kScanThreads, 0, ctx->stream>>>(ctx->sb.tile_sums.p, tiles);
  k_scan_apply<<<tiles, kScanThreads, 0, ctx->stream>>>(data, n, ctx->sb.tile_sums.p, data);
  ctx->launches += 3;
  CK(cudaGetLastError());
  return PMVSB_OK;
}
// _pgrids (visible = false) or _vpgrids (true) of the whole table as CSR over the flattened cells
static int build_cell_lists(pmvsb_ctx* ctx, bool visible) {
  StoreBufs& b = ctx->sb;
  const int cells = ctx->store_cells, E = visible ? ctx->store_ventries : ctx->store_entries;
  DVec<int32_t>& off = visible ? b.vcell_off : b.cell_off;
  DVec<int32_t>& lst = visible ? b.vcell_patch : b.cell_patch;
  const int32_t* images = visible ? b.vimages.p : b.images.p;
  const int32_t* grids = visible ? b.vgrids.p : b.grids.p;
  const int32_t* owner = visible ? b.ventry_patch.p : b.entry_patch.p;
  int r;
  if ((r = dvec_reserve(ctx, off, (size_t)cells + 1))) return r;
  if ((r = dvec_reserve(ctx, lst, (size_t)std::max(E, 1)))) return r;
  if ((r = dvec_reserve(ctx, b.cursor, (size_t)cells + 1))) return r;
  CK(cudaMemsetAsync(off.p, 0, sizeof(int32_t) * ((size_t)cells + 1), ctx->stream));
  CK(cudaMemsetAsync(b.cursor.p, 0, sizeof(int32_t) * ((size_t)cells + 1), ctx->stream));
  if (E > 0) {
    k_cells_count<<<(E + 255) / 256, 256, 0, ctx->stream>>>(ctx->tnum, E, images, grids, b.cell_base.p, b.gw.p, off.p);
    ++ctx->launches;
  }
  if ((r = device_scan(ctx, off.p, cells + 1))) return r;
  if (E > 0) {
    k_cells_fill<<<(E + 255) / 256, 256, 0, ctx->stream>>>(ctx->tnum, E, images, grids, owner, b.cell_base.p, b.gw.p, off.p, b.cursor.p, lst.p);
    k_cells_sort<<<(cells + 255) / 256, 256, 0, ctx->stream>>>(cells, off.p, lst.p);
    ctx->launches += 2;
  }
  CK(cudaGetLastError());
  store_view(ctx);
  return PMVSB_OK;
}
// validates the lists of `count` patches (offsets relative to off[0]); the kernels trust every index afterwards
static int check_lists(pmvsb_ctx* ctx, const char* who, int count, const int32_t* off, const int32_t* images, const int32_t* grids, bool visible) {
  for (int p = 0; p < count; ++p) {
    if (off[p + 1] < off[p]) return fail(ctx, PMVSB_EINVAL, std::string(who) + ": offsets not monotone");
    for (int e = off[p] - off[0]; e < off[p + 1] - off[0]; ++e) {
      const int im = images[e];
      if (im < 0 || im >= (visible ? ctx->tnum : ctx->num)) return fail(ctx, PMVSB_EINVAL, std::string(who) + (visible ? ": vimage index out of range" : ": image index out of range"));
      if (im < ctx->tnum) {
        const int x = grids[2 * e], y = grids[2 * e + 1];
        if (x < 0 || x >= ctx->h_gw[im] || y < 0 || y >= ctx->h_gh[im]) return fail(ctx, PMVSB_EINVAL, std::string(who) + (visible ? ": vgrid cell out of range" : ": grid cell out of range"));
      }
    }
  }
  return PMVSB_OK;
}

extern "C" {

const char* pmvsb_version(void) { return "pmvs-b200 0.2 (sm_100a)"; }

int pmvsb_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

const char* pmvsb_last_error(const pmvsb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int pmvsb_create(pmvsb_ctx** out, int device, int num_images, int num_target, int level, int csize, int wsize,
                 int min_image_num, float threshold, float max_angle_deg) {
  if (!out) return PMVSB_EINVAL;
  *out = nullptr;
  if (num_images < 1 || num_target < 1 || num_target > num_images || level < 0 || level > 5 || csize < 1 ||
      min_image_num < 2 || (wsize != 5 && wsize != 7 && wsize != 9))
    return PMVSB_EINVAL;
  // PMVSB_TRACE_INIT=1: milliseconds spent in each CUDA start-up step (the first runtime call brings the driver and the device up)
  const bool trace = std::getenv("PMVSB_TRACE_INIT") != nullptr;
  const auto t_init = std::chrono::steady_clock::now();
  auto stamp = [&](const char* what) {
    if (trace) std::fprintf(stderr, "init %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_init).count());
  };
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return PMVSB_ECUDA;
  stamp("cudaGetDeviceCount");
  pmvsb_ctx* ctx = new pmvsb_ctx();
  if (const char* v = std::getenv("PMVSB_NO_ATLAS")) ctx->atlas_enabled = !(v[0] && v[0] != '0');
  if (const char* v = std::getenv("PMVSB_NO_ORDER")) ctx->order_enabled = !(v[0] && v[0] != '0');
  ctx->device = device;
  ctx->num = num_images; ctx->tnum = num_target; ctx->level = level; ctx->csize = csize; ctx->wsize = wsize;
  ctx->min_image_num = min_image_num;
  ctx->tau = std::min(min_image_num * 2, num_images);  // findMatch.cpp:56
  if (ctx->tau > kMaxTau) { delete ctx; return PMVSB_EINVAL; }
  ctx->nlevels = level + 3;                            // findMatch.cpp:72
  ctx->threshold = threshold;
  ctx->ncc_threshold = threshold;
  ctx->ncc_threshold_before = threshold - 0.3f;       // findMatch.cpp:104
  ctx->angle_threshold0 = 60.0f * M_PI / 180.0f;      // findMatch.cpp:92-93
  ctx->angle_threshold1 = 60.0f * M_PI / 180.0f;
  ctx->max_angle_threshold = max_angle_deg;
  ctx->max_angle_threshold *= M_PI / 180.0f;          // option.cpp:105-106
  ctx->cams.resize(num_images);
  ctx->images.resize(num_images);
  ctx->visdata2.resize(num_images);
  for (int i = 0; i < num_images; ++i)
    for (int j = 0; j < num_images; ++j)
      if (j != i) ctx->visdata2[i].push_back(j);
  cudaError_t e = cudaSetDevice(device);
  stamp("cudaSetDevice");
  if (e == cudaSuccess) e = cudaFree(nullptr);   // forces the primary context into existence here
  stamp("primary context");
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
  ctx->stream = ctx->own_stream;
  if (e == cudaSuccess) {   // keep freed blocks in the device's stream-ordered pool (see dev_malloc)
    cudaMemPool_t mp = nullptr;
    uint64_t keep_all = UINT64_MAX;
    if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep_all);
    cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
  if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_counter, 4 * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(ctx->d_counter, 0, 4 * sizeof(int));
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  stamp("stream, events, counters");
  if (e != cudaSuccess) { delete ctx; return PMVSB_ECUDA; }
  *out = ctx;
  return PMVSB_OK;
}

static void atlas_free(pmvsb_ctx* ctx);

int pmvsb_destroy(pmvsb_ctx* ctx) {
  if (!ctx) return PMVSB_EINVAL;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto& im : ctx->images) {
    for (auto* p : im.levels) cudaFree(p);
    cudaFree(im.maps[0]); cudaFree(im.maps[1]);
  }
  cudaFree(ctx->d_map_tab[0]); cudaFree(ctx->d_map_tab[1]); cudaFree(ctx->d_bimages);
  cudaFree(ctx->d_cams); cudaFree(ctx->d_levels); cudaFree(ctx->d_counter); cudaFree(ctx->d_vis_off); cudaFree(ctx->d_vis_idx);
  atlas_free(ctx);
  cudaFree(ctx->order_bin.p); cudaFree(ctx->order_counts.p); cudaFree(ctx->order_idx.p);
  cudaFree(ctx->arena);
  cudaFree(ctx->d_fxy.p); cudaFree(ctx->d_ftype.p); cudaFree(ctx->d_fcell_base.p); cudaFree(ctx->d_fcell_off.p); cudaFree(ctx->d_flist.p);
  cudaFree(ctx->d_fgw.p); cudaFree(ctx->d_fgh.p); cudaFree(ctx->d_blocked.p); cudaFree(ctx->d_seed_out.p);
  store_free(ctx);
  if (ctx->comm && nccl_api()) nccl_api()->CommDestroy(ctx->comm);
  peer_free(ctx);
  cudaFree(ctx->comm_buf);
  for (auto& b : ctx->pool) cudaFree(b.second);
  g_current = nullptr;
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return PMVSB_OK;
}

int pmvsb_upload_camera(pmvsb_ctx* ctx, int index, const float* P) {
  if (!ctx || !P || index < 0 || index >= ctx->num) return fail(ctx, PMVSB_EINVAL, "upload_camera: bad index or pointer");
  HostCam& c = ctx->cams[index];
  std::memcpy(c.P0, P, sizeof(float) * 12);
  derive_camera(c);
  c.set = true;
  ctx->finalized = false;
  return PMVSB_OK;
}

int pmvsb_upload_image(pmvsb_ctx* ctx, int index, int width, int height, const uint8_t* rgb) {
  if (!ctx || !rgb || index < 0 || index >= ctx->num || width < 1 || height < 1)
    return fail(ctx, PMVSB_EINVAL, "upload_image: bad argument");
  g_current = ctx;
  CK(cudaSetDevice(ctx->device));
  HostImage& im = ctx->images[index];
  for (auto* p : im.levels) cudaFree(p);
  cudaFree(im.maps[0]); cudaFree(im.maps[1]);
  im.maps[0] = im.maps[1] = nullptr;
  im.levels.assign(ctx->nlevels, nullptr);
  im.w.assign(ctx->nlevels, 0);
  im.h.assign(ctx->nlevels, 0);
  const size_t n0 = (size_t)width * height;
  DevBuf<uint8_t> staging;
  CK(staging.alloc(n0 * 3));
  CK(cudaMemcpyAsync(staging.p, rgb, n0 * 3, cudaMemcpyHostToDevice, ctx->stream));
  im.w[0] = width; im.h[0] = height;
  CK(dev_malloc(ctx, (void**)&im.levels[0], n0 * sizeof(uchar4)));
  const int blocks = (int)std::min<size_t>((n0 + 255) / 256, (size_t)ctx->sm_count * 16);
  k_rgb_to_rgba<<<blocks, 256, 0, ctx->stream>>>(staging.p, im.levels[0], n0);
  ++ctx->launches;
  for (int l = 1; l < ctx->nlevels; ++l) {
    im.w[l] = im.w[l - 1] / 2;  // image.cpp:136-139
    im.h[l] = im.h[l - 1] / 2;
    const size_t nl = (size_t)std::max(im.w[l], 1) * std::max(im.h[l], 1);
    CK(dev_malloc(ctx, (void**)&im.levels[l], nl * sizeof(uchar4)));
    if (im.w[l] > 0 && im.h[l] > 0) {
      dim3 block(32, 8), grid((im.w[l] + 31) / 32, (im.h[l] + 7) / 8);
      k_pyr_down<<<grid, block, 0, ctx->stream>>>(im.levels[l - 1], im.w[l - 1], im.h[l - 1], im.levels[l], im.w[l], im.h[l]);
      ++ctx->launches;
    }
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  im.set = true;
  ctx->finalized = false;
  return PMVSB_OK;
}

// level-0 map (already 255 / 0) -> the working level through the 2x2 "any parent in" pyramid; replaces im.maps[which]
static int map_to_working_level(pmvsb_ctx* ctx, HostImage& im, int which, uint8_t* d_level0) {
  uint8_t* cur = d_level0;
  for (int l = 1; l <= ctx->level; ++l) {
    uint8_t* next = nullptr;
    const size_t nl = (size_t)std::max(im.w[l], 1) * std::max(im.h[l], 1);
    CK(dev_malloc(ctx, (void**)&next, nl));
    if (im.w[l] > 0 && im.h[l] > 0) {
      k_map_down<<<dim3((im.w[l] + 127) / 128, im.h[l]), 128, 0, ctx->stream>>>(cur, im.w[l - 1], im.h[l - 1], next, im.w[l], im.h[l]);
      ++ctx->launches;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(cur);
    cur = next;
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(im.maps[which]);
  im.maps[which] = cur;
  ctx->finalized = false;
  return PMVSB_OK;
}

int pmvsb_upload_mask(pmvsb_ctx* ctx, int index, int which, int width, int height, const uint8_t* gray) {
  if (!ctx || !gray || index < 0 || index >= ctx->num || (which != 0 && which != 1)) return fail(ctx, PMVSB_EINVAL, "upload_mask: bad argument");
  g_current = ctx;
  CK(cudaSetDevice(ctx->device));
  HostImage& im = ctx->images[index];
  if (!im.set) return fail(ctx, PMVSB_ESTATE, "upload_mask: upload the image first");
  if (width != im.w[0] || height != im.h[0]) return fail(ctx, PMVSB_EINVAL, "upload_mask: the map must have the size of its image");
  const size_t n0 = (size_t)width * height;
  DevBuf<uint8_t> staging;
  CK(staging.alloc(n0));
  CK(cudaMemcpyAsync(staging.p, gray, n0, cudaMemcpyHostToDevice, ctx->stream));
  uint8_t* level0 = nullptr;
  CK(dev_malloc(ctx, (void**)&level0, n0));
  k_map_binarise<<<(unsigned)((n0 + 255) / 256), 256, 0, ctx->stream>>>(staging.p, n0, which == 0 ? 127 : 1, level0);
  ++ctx->launches;
  return map_to_working_level(ctx, im, which, level0);
}

int pmvsb_set_edge(pmvsb_ctx* ctx, float threshold) {
  if (!ctx) return PMVSB_EINVAL;
  g_current = ctx;
  CK(cudaSetDevice(ctx->device));
  const float sigma = 3.f, sigma2 = 2.f * sigma * sigma;
  const int margin = (int)std::floor(2 * sigma);
  std::vector<float> taps(2 * margin + 1);
  for (int i = -margin; i <= margin; ++i) taps[i + margin] = expf(-i * i / sigma2);   // image.cpp:446-452 (the object imports expf)
  const float new_threshold = threshold * threshold * (2 * margin + 1) * (2 * margin + 1) / 3.0f;
  DevBuf<float> d_taps;
  CK(d_taps.alloc(taps.size()));
  CK(cudaMemcpyAsync(d_taps.p, taps.data(), sizeof(float) * taps.size(), cudaMemcpyHostToDevice, ctx->stream));
  for (int index = 0; index < ctx->num; ++index) {
    HostImage& im = ctx->images[index];
    if (!im.set) return fail(ctx, PMVSB_ESTATE, "set_edge: upload every image first");
    const int w = im.w[0], h = im.h[0];
    const size_t n0 = (size_t)w * h;
    DevBuf<float> a, b;
    CK(a.alloc(n0)); CK(b.alloc(n0));
    const dim3 grid((w + 127) / 128, h);
    k_edge_grad<<<grid, 128, 0, ctx->stream>>>(im.levels[0], w, h, a.p);
    k_edge_smooth<true><<<grid, 128, 0, ctx->stream>>>(a.p, b.p, w, h, d_taps.p, margin);
    k_edge_smooth<false><<<grid, 128, 0, ctx->stream>>>(b.p, a.p, w, h, d_taps.p, margin);
    uint8_t* level0 = nullptr;
    CK(dev_malloc(ctx, (void**)&level0, n0));
    k_edge_threshold<<<(unsigned)((n0 + 255) / 256), 256, 0, ctx->stream>>>(a.p, n0, new_threshold, level0);
    ctx->launches += 4;
    const int r = map_to_working_level(ctx, im, 1, level0);
    if (r) return r;
  }
  return PMVSB_OK;
}

int pmvsb_set_bimages(pmvsb_ctx* ctx, const int32_t* list, int n) {
  if (!ctx || n < 0 || (n > 0 && !list)) return fail(ctx, PMVSB_EINVAL, "set_bimages: bad argument");
  for (int i = 0; i < n; ++i)
    if (list[i] < 0 || list[i] >= ctx->num) return fail(ctx, PMVSB_EINVAL, "set_bimages: image index out of range");
  ctx->bimages.assign(list, list + n);
  ctx->finalized = false;
  return PMVSB_OK;
}

int pmvsb_download_mask(pmvsb_ctx* ctx, int index, int which, uint8_t* out, int* present) {
  if (!ctx || index < 0 || index >= ctx->num || (which != 0 && which != 1) || !present) return fail(ctx, PMVSB_EINVAL, "download_mask: bad argument");
  CK(cudaSetDevice(ctx->device));
  const HostImage& im = ctx->images[index];
  *present = im.maps[which] ? 1 : 0;
  if (im.maps[which] && out) {
    CK(cudaMemcpyAsync(out, im.maps[which], (size_t)im.w[ctx->level] * im.h[ctx->level], cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return PMVSB_OK;
}

int pmvsb_set_visdata2(pmvsb_ctx* ctx, int index, const int32_t* list, int n) {
  if (!ctx || index < 0 || index >= ctx->num || n < 0 || (n > 0 && !list)) return fail(ctx, PMVSB_EINVAL, "set_visdata2: bad argument");
  for (int i = 0; i < n; ++i)
    if (list[i] < 0 || list[i] >= ctx->num) return fail(ctx, PMVSB_EINVAL, "set_visdata2: image index out of range");
  ctx->visdata2[index].assign(list, list + n);
  ctx->finalized = false;
  return PMVSB_OK;
}

// One block-linear RGBA8 CUDA array holding every (image, pyramid level), read by the group kernels with tex2Dgather.
// A warp's four groups sample different images and levels in one instruction and the texture handle of TLD4 must be
// warp-uniform, hence ONE texture for the whole scene (per-level handles would make the compiler serialise the gather
// per distinct handle).  Shelf packing: level l occupies a block of rows, its images laid out left to right, top to
// bottom in cells of that level's largest width x height.  Limits: cudaDeviceProp::maxTexture2DGather (32768 x 32768 on
// B200); a scene that does not fit keeps atlas_tex = 0 and the kernels gather through global loads.
static void atlas_free(pmvsb_ctx* ctx) {
  if (ctx->atlas_tex) cudaDestroyTextureObject(ctx->atlas_tex);
  if (ctx->atlas_array) cudaFreeArray(ctx->atlas_array);
  ctx->atlas_tex = 0; ctx->atlas_array = nullptr; ctx->atlas_w = ctx->atlas_h = 0;
}

static int atlas_build(pmvsb_ctx* ctx, std::vector<LevelDev>& hl) {
  for (LevelDev& d : hl) { d.ax1 = 1.0f; d.ay1 = 1.0f; d.pad0 = d.pad1 = 0; }
  if (!ctx->atlas_enabled || ctx->num < 1) { atlas_free(ctx); return PMVSB_OK; }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, ctx->device));
  const int max_w = prop.maxTexture2DGather[0], max_h = prop.maxTexture2DGather[1];
  std::vector<int> cw(ctx->nlevels, 0), ch(ctx->nlevels, 0);   // cell size per level
  for (int i = 0; i < ctx->num; ++i)
    for (int l = 0; l < ctx->nlevels; ++l) { cw[l] = std::max(cw[l], ctx->images[i].w[l]); ch[l] = std::max(ch[l], ctx->images[i].h[l]); }
  if (cw[0] < 1 || ch[0] < 1 || cw[0] > max_w) { atlas_free(ctx); return PMVSB_OK; }
  // roughly square: columns of level 0 chosen so that width ~ height of the level-0 block
  int cols0 = (int)std::ceil(std::sqrt((double)ctx->num * ch[0] / cw[0]));
  cols0 = std::max(1, std::min(cols0, max_w / cw[0]));
  const int W = cols0 * cw[0];
  std::vector<int> cols(ctx->nlevels, 1), row0(ctx->nlevels, 0);
  long long H = 0;
  for (int l = 0; l < ctx->nlevels; ++l) {
    if (cw[l] < 1 || ch[l] < 1) continue;
    cols[l] = std::max(1, W / cw[l]);
    row0[l] = (int)H;
    H += (long long)((ctx->num + cols[l] - 1) / cols[l]) * ch[l];
  }
  if (H > max_h) { atlas_free(ctx); return PMVSB_OK; }
  if (!ctx->atlas_array || ctx->atlas_w != W || ctx->atlas_h != (int)H) {
    atlas_free(ctx);
    const cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
    if (cudaMallocArray(&ctx->atlas_array, &cd, (size_t)W, (size_t)H, cudaArrayTextureGather) != cudaSuccess) {
      cudaGetLastError();   // no memory for the second copy: stay on the global-load path
      ctx->atlas_array = nullptr;
      return PMVSB_OK;
    }
    ctx->atlas_w = W; ctx->atlas_h = (int)H;
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray; rd.res.array.array = ctx->atlas_array;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(&ctx->atlas_tex, &rd, &td, nullptr));
  }
  for (int i = 0; i < ctx->num; ++i)
    for (int l = 0; l < ctx->nlevels; ++l) {
      const int w = ctx->images[i].w[l], h = ctx->images[i].h[l];
      if (w < 1 || h < 1) continue;
      const int ox = (i % cols[l]) * cw[l], oy = row0[l] + (i / cols[l]) * ch[l];
      CK(cudaMemcpy2DToArrayAsync(ctx->atlas_array, (size_t)ox * sizeof(uchar4), (size_t)oy, ctx->images[i].levels[l], (size_t)w * sizeof(uchar4),
                                  (size_t)w * sizeof(uchar4), (size_t)h, cudaMemcpyDeviceToDevice, ctx->stream));
      LevelDev& d = hl[(size_t)i * ctx->nlevels + l];
      d.ax1 = (float)(ox + 1); d.ay1 = (float)(oy + 1);   // integers < 2^15 + texel index < 2^15: exact in f32
    }
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

int pmvsb_finalize_scene(pmvsb_ctx* ctx) {
  if (!ctx) return PMVSB_EINVAL;
  for (int i = 0; i < ctx->num; ++i) {
    if (!ctx->cams[i].set) return fail(ctx, PMVSB_ESTATE, "finalize_scene: camera " + std::to_string(i) + " missing");
    if (!ctx->images[i].set) return fail(ctx, PMVSB_ESTATE, "finalize_scene: image " + std::to_string(i) + " missing");
  }
  CK(cudaSetDevice(ctx->device));
  std::vector<CamDev> hc(ctx->num);
  std::vector<LevelDev> hl((size_t)ctx->num * ctx->nlevels);
  const float sc = 1.0f / (float)(1 << ctx->level);  // rows 0,1 halved per level: exact (camera.cpp:56-68)
  for (int i = 0; i < ctx->num; ++i) {
    const HostCam& c = ctx->cams[i];
    std::memset(&hc[i], 0, sizeof(CamDev));
    for (int k = 0; k < 4; ++k) {
      hc[i].P[0][k] = c.P0[0][k] * sc;
      hc[i].P[1][k] = c.P0[1][k] * sc;
      hc[i].P[2][k] = c.P0[2][k];
    }
    std::memcpy(hc[i].centre, c.centre, 16); std::memcpy(hc[i].oaxis, c.oaxis, 16);
    std::memcpy(hc[i].xaxis, c.xaxis, 12); std::memcpy(hc[i].yaxis, c.yaxis, 12); std::memcpy(hc[i].zaxis, c.zaxis, 12);
    hc[i].ipscale = c.ipscale;
    for (int l = 0; l < ctx->nlevels; ++l) {
      LevelDev& d = hl[(size_t)i * ctx->nlevels + l];
      d.pix = ctx->images[i].levels[l]; d.w = ctx->images[i].w[l]; d.h = ctx->images[i].h[l];
    }
  }
  {
    const int r = atlas_build(ctx, hl);
    if (r) return r;
  }
  cudaFree(ctx->d_cams); cudaFree(ctx->d_levels);
  ctx->d_cams = nullptr; ctx->d_levels = nullptr;
  CK(cudaMalloc((void**)&ctx->d_cams, sizeof(CamDev) * hc.size()));
  CK(cudaMalloc((void**)&ctx->d_levels, sizeof(LevelDev) * hl.size()));
  CK(cudaMemcpy(ctx->d_cams, hc.data(), sizeof(CamDev) * hc.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->d_levels, hl.data(), sizeof(LevelDev) * hl.size(), cudaMemcpyHostToDevice));
  {
    std::vector<int32_t> off(ctx->num + 1, 0), idx;
    for (int i = 0; i < ctx->num; ++i) {
      idx.insert(idx.end(), ctx->visdata2[i].begin(), ctx->visdata2[i].end());
      off[i + 1] = (int32_t)idx.size();
    }
    cudaFree(ctx->d_vis_off); cudaFree(ctx->d_vis_idx);
    ctx->d_vis_off = nullptr; ctx->d_vis_idx = nullptr;
    CK(cudaMalloc((void**)&ctx->d_vis_off, sizeof(int32_t) * off.size()));
    CK(cudaMalloc((void**)&ctx->d_vis_idx, sizeof(int32_t) * (idx.size() ? idx.size() : 1)));
    CK(cudaMemcpy(ctx->d_vis_off, off.data(), sizeof(int32_t) * off.size(), cudaMemcpyHostToDevice));
    if (!idx.empty()) CK(cudaMemcpy(ctx->d_vis_idx, idx.data(), sizeof(int32_t) * idx.size(), cudaMemcpyHostToDevice));
  }
  for (int which = 0; which < 2; ++which) {
    cudaFree(ctx->d_map_tab[which]);
    ctx->d_map_tab[which] = nullptr;
    std::vector<const unsigned char*> tab(ctx->num, nullptr);
    bool any = false;
    for (int i = 0; i < ctx->num; ++i) { tab[i] = ctx->images[i].maps[which]; any |= tab[i] != nullptr; }
    if (!any) continue;
    CK(cudaMalloc((void**)&ctx->d_map_tab[which], sizeof(void*) * tab.size()));
    CK(cudaMemcpy(ctx->d_map_tab[which], tab.data(), sizeof(void*) * tab.size(), cudaMemcpyHostToDevice));
  }
  cudaFree(ctx->d_bimages);
  ctx->d_bimages = nullptr;
  if (!ctx->bimages.empty()) {
    CK(cudaMalloc((void**)&ctx->d_bimages, sizeof(int32_t) * ctx->bimages.size()));
    CK(cudaMemcpy(ctx->d_bimages, ctx->bimages.data(), sizeof(int32_t) * ctx->bimages.size(), cudaMemcpyHostToDevice));
  }
  fill_scene(ctx);
  fill_select(ctx);
  ctx->finalized = true;
  return PMVSB_OK;
}

int pmvsb_set_thresholds(pmvsb_ctx* ctx, float ncc_threshold, float ncc_threshold_before) {
  if (!ctx) return PMVSB_EINVAL;
  ctx->ncc_threshold = ncc_threshold;
  ctx->ncc_threshold_before = ncc_threshold_before;
  ctx->select.ncc_threshold = ncc_threshold;
  ctx->select.ncc_threshold_before = ncc_threshold_before;
  return PMVSB_OK;
}

int pmvsb_set_optimizer(pmvsb_ctx* ctx, double xtol, double step, int maxeval) {
  if (!ctx || !(xtol > 0.0) || !(step > 0.0) || maxeval < 1) return fail(ctx, PMVSB_EINVAL, "set_optimizer: bad argument");
  ctx->xtol = xtol; ctx->step = step; ctx->maxeval = maxeval;
  ctx->scene.xtol = xtol; ctx->scene.step = step; ctx->scene.maxeval = maxeval;
  return PMVSB_OK;
}

int pmvsb_image_dims(pmvsb_ctx* ctx, int index, int level, int* width, int* height) {
  if (!ctx || index < 0 || index >= ctx->num || level < 0 || level >= ctx->nlevels || !ctx->images[index].set)
    return fail(ctx, PMVSB_EINVAL, "image_dims: bad argument");
  *width = ctx->images[index].w[level];
  *height = ctx->images[index].h[level];
  return PMVSB_OK;
}

int pmvsb_download_image(pmvsb_ctx* ctx, int index, int level, uint8_t* rgb) {
  if (!ctx || !rgb || index < 0 || index >= ctx->num || level < 0 || level >= ctx->nlevels || !ctx->images[index].set)
    return fail(ctx, PMVSB_EINVAL, "download_image: bad argument");
  g_current = ctx;
  CK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)ctx->images[index].w[level] * ctx->images[index].h[level];
  if (n == 0) return PMVSB_OK;
  DevBuf<uint8_t> tmp;
  CK(tmp.alloc(n * 3));
  const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 16);
  k_rgba_to_rgb<<<blocks, 256, 0, ctx->stream>>>(ctx->images[index].levels[level], tmp.p, n);
  ++ctx->launches;
  CK(cudaMemcpyAsync(rgb, tmp.p, n * 3, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

int pmvsb_get_camera(pmvsb_ctx* ctx, int index, int level, float* P, float* centre, float* oaxis, float* xaxis, float* yaxis,
                     float* zaxis, float* ipscale) {
  if (!ctx || index < 0 || index >= ctx->num || level < 0 || level >= ctx->nlevels || !ctx->cams[index].set)
    return fail(ctx, PMVSB_EINVAL, "get_camera: bad argument");
  const HostCam& c = ctx->cams[index];
  const float sc = 1.0f / (float)(1 << level);
  for (int k = 0; k < 4; ++k) { P[k] = c.P0[0][k] * sc; P[4 + k] = c.P0[1][k] * sc; P[8 + k] = c.P0[2][k]; }
  std::memcpy(centre, c.centre, 16); std::memcpy(oaxis, c.oaxis, 16);
  std::memcpy(xaxis, c.xaxis, 12); std::memcpy(yaxis, c.yaxis, 12); std::memcpy(zaxis, c.zaxis, 12);
  *ipscale = c.ipscale;
  return PMVSB_OK;
}

// ---- batched calls with host pointers: stage, launch, copy back ------------------------------------
struct PatchStage {
  DevBuf<float> coords, normals, dscales;
  DevBuf<int32_t> images, nimages;
};

// every LISTED image index of a stride-padded batch must name a scene image (the kernels trust them).  Only the listed
// entries are visited: a pipeline run pushes ~1e8 padded slots through here, and a division per slot was ~10 % of its wall time.
static int check_image_rows(pmvsb_ctx* ctx, int P, int stride, const int32_t* images, const int32_t* nimages) {
  const unsigned num = (unsigned)ctx->num;
  for (int p = 0; p < P; ++p) {
    const int32_t* row = images + (size_t)p * stride;
    const int n = nimages ? std::min(nimages[p], stride) : stride;
    unsigned bad = 0;
    for (int k = 0; k < n; ++k) bad |= (unsigned)((unsigned)row[k] >= num);
    if (bad) return fail(ctx, PMVSB_EINVAL, "image index out of range in patch batch");
  }
  return PMVSB_OK;
}

static int stage_patches(pmvsb_ctx* ctx, PatchStage& st, int P, int stride, const float* coords, const float* normals,
                         const int32_t* images, const int32_t* nimages, const float* dscales) {
  if (P < 0 || stride < 1 || !coords || !images) return fail(ctx, PMVSB_EINVAL, "bad patch batch");
  CK(st.coords.alloc((size_t)4 * P));
  CK(cudaMemcpyAsync(st.coords.p, coords, sizeof(float) * 4 * P, cudaMemcpyHostToDevice, ctx->stream));
  if (normals) {
    CK(st.normals.alloc((size_t)4 * P));
    CK(cudaMemcpyAsync(st.normals.p, normals, sizeof(float) * 4 * P, cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(st.images.alloc((size_t)stride * P));
  CK(cudaMemcpyAsync(st.images.p, images, sizeof(int32_t) * (size_t)stride * P, cudaMemcpyHostToDevice, ctx->stream));
  if (nimages) {
    CK(st.nimages.alloc(P));
    CK(cudaMemcpyAsync(st.nimages.p, nimages, sizeof(int32_t) * P, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (dscales) {
    CK(st.dscales.alloc(P));
    CK(cudaMemcpyAsync(st.dscales.p, dscales, sizeof(float) * P, cudaMemcpyHostToDevice, ctx->stream));
  }
  // image indexes are trusted by the kernels: validate on the host
  return check_image_rows(ctx, P, stride, images, nimages);
}

int pmvsb_project_batch(pmvsb_ctx* ctx, int n, const float* coords, const int32_t* image, int level, float* out) {
  int r = check_ready(ctx);
  if (r) return r;
  if (n < 0 || !coords || !image || !out || level < 0 || level >= ctx->nlevels) return fail(ctx, PMVSB_EINVAL, "project_batch: bad argument");
  if (n == 0) return PMVSB_OK;
  for (int i = 0; i < n; ++i)
    if (image[i] < 0 || image[i] >= ctx->num) return fail(ctx, PMVSB_EINVAL, "project_batch: image index out of range");
  DevBuf<float> dc, dout;
  DevBuf<int32_t> di;
  CK(dc.alloc((size_t)4 * n)); CK(di.alloc(n)); CK(dout.alloc((size_t)3 * n));
  CK(cudaMemcpyAsync(dc.p, coords, sizeof(float) * 4 * n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(di.p, image, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
  k_project<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->scene, n, dc.p, di.p, level, dout.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, dout.p, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

int pmvsb_grab_tex_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals, const int32_t* images,
                         const int32_t* nimages, float* tex, int32_t* flag, int32_t* newlevel) {
  int r = check_ready(ctx);
  if (r) return r;
  if (!normals || !tex || !flag || !newlevel) return fail(ctx, PMVSB_EINVAL, "grab_tex_batch: null pointer");
  if (P == 0) return PMVSB_OK;
  PatchStage st;
  r = stage_patches(ctx, st, P, stride, coords, normals, images, nimages, nullptr);
  if (r) return r;
  const size_t tsz = (size_t)3 * ctx->wsize * ctx->wsize;
  DevBuf<float> dt;
  DevBuf<int32_t> df, dl;
  CK(dt.alloc((size_t)P * stride * tsz)); CK(df.alloc((size_t)P * stride)); CK(dl.alloc((size_t)P * stride));
  CK(cudaMemsetAsync(dt.p, 0, sizeof(float) * (size_t)P * stride * tsz, ctx->stream));
  CK(cudaMemsetAsync(df.p, 0xff, sizeof(int32_t) * (size_t)P * stride, ctx->stream));
  CK(cudaMemsetAsync(dl.p, 0xff, sizeof(int32_t) * (size_t)P * stride, ctx->stream));
  const int blocks = (P + 3) / 4;
  DISPATCH_WSIZE(ctx, k_grab_tex, blocks, 128, ctx->scene, P, stride, st.coords.p, st.normals.p, st.images.p, st.nimages.p, dt.p, df.p, dl.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(tex, dt.p, sizeof(float) * (size_t)P * stride * tsz, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(flag, df.p, sizeof(int32_t) * (size_t)P * stride, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(newlevel, dl.p, sizeof(int32_t) * (size_t)P * stride, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

static int score_common(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals, const int32_t* images,
                        const int32_t* nimages, const float* dscales, const double* x, int mode, double* out) {
  int r = check_ready(ctx);
  if (r) return r;
  if (!normals || !out || (mode == 0 && (!x || !dscales))) return fail(ctx, PMVSB_EINVAL, "score: null pointer");
  if (P == 0) return PMVSB_OK;
  PatchStage st;
  r = stage_patches(ctx, st, P, stride, coords, normals, images, nimages, dscales);
  if (r) return r;
  DevBuf<double> dx, dout;
  CK(dout.alloc(P));
  if (x) {
    CK(dx.alloc((size_t)3 * P));
    CK(cudaMemcpyAsync(dx.p, x, sizeof(double) * 3 * P, cudaMemcpyHostToDevice, ctx->stream));
  }
  const int blocks = (P + 3) / 4;       // warp per patch
  const int blocksg = (P + 15) / 16;    // 8 lanes per patch
  DISPATCH_GROUP(ctx, k_score_g, k_score, blocksg, blocks, 128, ctx->scene, P, stride, st.coords.p, st.normals.p, st.images.p, st.nimages.p,
                 st.dscales.p, dx.p, mode, dout.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, dout.p, sizeof(double) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

int pmvsb_eval_objective_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals, const int32_t* images,
                               const int32_t* nimages, const float* dscales, const double* x, double* f) {
  return score_common(ctx, P, stride, coords, normals, images, nimages, dscales, x, 0, f);
}

int pmvsb_compute_incc_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals, const int32_t* images,
                             const int32_t* nimages, int robust, double* out) {
  return score_common(ctx, P, stride, coords, normals, images, nimages, nullptr, nullptr, robust ? 1 : 2, out);
}

int pmvsb_set_inccs_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals, const int32_t* images,
                          const int32_t* nimages, int robust, float* out) {
  int r = check_ready(ctx);
  if (r) return r;
  if (!normals || !out) return fail(ctx, PMVSB_EINVAL, "set_inccs_batch: null pointer");
  if (P == 0) return PMVSB_OK;
  PatchStage st;
  r = stage_patches(ctx, st, P, stride, coords, normals, images, nimages, nullptr);
  if (r) return r;
  DevBuf<float> dout;
  CK(dout.alloc((size_t)P * stride));
  CK(cudaMemsetAsync(dout.p, 0, sizeof(float) * (size_t)P * stride, ctx->stream));
  const int blocks = (P + 3) / 4;
  DISPATCH_WSIZE(ctx, k_set_inccs, blocks, 128, ctx->scene, P, stride, st.coords.p, st.normals.p, st.images.p, st.nimages.p, robust, dout.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, dout.p, sizeof(float) * (size_t)P * stride, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

int pmvsb_set_scales_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const int32_t* images, const int32_t* nimages,
                           float* dscale, float* ascale) {
  int r = check_ready(ctx);
  if (r) return r;
  if (!dscale || !ascale) return fail(ctx, PMVSB_EINVAL, "set_scales_batch: null pointer");
  if (P == 0) return PMVSB_OK;
  PatchStage st;
  r = stage_patches(ctx, st, P, stride, coords, nullptr, images, nimages, nullptr);
  if (r) return r;
  DevBuf<float> dd, da;
  CK(dd.alloc(P)); CK(da.alloc(P));
  k_set_scales<<<(P + 127) / 128, 128, 0, ctx->stream>>>(ctx->scene, P, stride, st.coords.p, st.images.p, st.nimages.p, dd.p, da.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(dscale, dd.p, sizeof(float) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ascale, da.p, sizeof(float) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

// ---- filter stage -----------------------------------------------------------------------------------------
int pmvsb_set_depth(pmvsb_ctx* ctx, int depth) {
  if (!ctx || depth < 0) return fail(ctx, PMVSB_EINVAL, "set_depth: bad argument");
  ctx->depth_flag = depth;
  ctx->store.depth_flag = depth;
  return PMVSB_OK;
}

int pmvsb_grid_dims(pmvsb_ctx* ctx, int image, int* gwidth, int* gheight) {
  if (!ctx || image < 0 || image >= ctx->num || !ctx->images[image].set || !gwidth || !gheight)
    return fail(ctx, PMVSB_EINVAL, "grid_dims: bad argument");
  *gwidth = (ctx->images[image].w[ctx->level] + ctx->csize - 1) / ctx->csize;
  *gheight = (ctx->images[image].h[ctx->level] + ctx->csize - 1) / ctx->csize;
  return PMVSB_OK;
}

static void grid_geometry(pmvsb_ctx* ctx);

int pmvsb_store_upload(pmvsb_ctx* ctx, int P, const float* coords, const float* normals, const float* ncc, const float* dscale,
                       const int32_t* img_off, const int32_t* images, const int32_t* grids, const int32_t* vimg_off,
                       const int32_t* vimages, const int32_t* vgrids, const int32_t* timages) {
  int r = check_ready(ctx);
  if (r) return r;
  if (P < 0 || !coords || !normals || !ncc || !dscale || !img_off || !vimg_off || !timages) return fail(ctx, PMVSB_EINVAL, "store_upload: null pointer");
  const int E = img_off[P], VE = vimg_off[P];
  if (img_off[0] != 0 || vimg_off[0] != 0) return fail(ctx, PMVSB_EINVAL, "store_upload: offsets must start at 0");
  if ((E > 0 && (!images || !grids)) || (VE > 0 && (!vimages || !vgrids))) return fail(ctx, PMVSB_EINVAL, "store_upload: null list");
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->store_set = false; ctx->depth_built = false; ctx->store_appended = false;
  StoreBufs& b = ctx->sb;
  if (!b.cell_base.p) {   // grid geometry (patchOrganizerS.cpp:72-77), once per scene
    grid_geometry(ctx);
    if ((r = dvec_put(ctx, b.cell_base, 0, ctx->h_base.data(), ctx->h_base.size()))) return r;
    if ((r = dvec_put(ctx, b.gw, 0, ctx->h_gw.data(), ctx->h_gw.size()))) return r;
    if ((r = dvec_put(ctx, b.gh, 0, ctx->h_gh.data(), ctx->h_gh.size()))) return r;
    if ((r = dvec_reserve(ctx, b.dp, (size_t)std::max(ctx->h_base[ctx->tnum], 1)))) return r;
  }
  const int cells = ctx->h_base[ctx->tnum];
  if ((r = check_lists(ctx, "store_upload", P, img_off, images, grids, false))) return r;
  if ((r = check_lists(ctx, "store_upload", P, vimg_off, vimages, vgrids, true))) return r;
  if ((r = dvec_put(ctx, b.coords, 0, coords, (size_t)4 * P))) return r;
  if ((r = dvec_put(ctx, b.normals, 0, normals, (size_t)4 * P))) return r;
  if ((r = dvec_put(ctx, b.ncc, 0, ncc, (size_t)P))) return r;
  if ((r = dvec_put(ctx, b.dscale, 0, dscale, (size_t)P))) return r;
  if ((r = dvec_put(ctx, b.timages, 0, timages, (size_t)P))) return r;
  if ((r = dvec_put(ctx, b.img_off, 0, img_off, (size_t)P + 1))) return r;
  if ((r = dvec_put(ctx, b.images, 0, images, (size_t)E))) return r;
  if ((r = dvec_put(ctx, b.grids, 0, grids, (size_t)2 * E))) return r;
  if ((r = dvec_put(ctx, b.vimg_off, 0, vimg_off, (size_t)P + 1))) return r;
  if ((r = dvec_put(ctx, b.vimages, 0, vimages, (size_t)VE))) return r;
  if ((r = dvec_put(ctx, b.vgrids, 0, vgrids, (size_t)2 * VE))) return r;
  if ((r = dvec_reserve(ctx, b.entry_patch, (size_t)std::max(E, 1)))) return r;
  if ((r = dvec_reserve(ctx, b.ventry_patch, (size_t)std::max(VE, 1)))) return r;
  if ((r = dvec_reserve(ctx, b.seq, (size_t)std::max(P, 1)))) return r;
  if (P > 0) { k_iota<<<(P + 255) / 256, 256, 0, ctx->stream>>>(b.seq.p, 0, P, 0); ++ctx->launches; }   // default creation order = table order
  StoreDev& st = ctx->store;
  st.P = P;
  st.depth_flag = ctx->depth_flag;
  st.ncc_threshold = ctx->ncc_threshold;
  const double c120 = std::cos(120.0 * M_PI / 180.0);   // findMatch.cpp:126
  float cf = (float)c120;
  if ((double)cf < c120) cf = std::nextafterf(cf, INFINITY);
  st.cos120_f = cf;
  ctx->store_entries = E; ctx->store_ventries = VE; ctx->store_cells = cells;
  store_view(ctx);
  if (P > 0) {
    k_entry_owner<<<(P + 255) / 256, 256, 0, ctx->stream>>>(0, P, b.img_off.p, b.entry_patch.p);
    k_entry_owner<<<(P + 255) / 256, 256, 0, ctx->stream>>>(0, P, b.vimg_off.p, b.ventry_patch.p);
    ctx->launches += 2;
  }
  if ((r = build_cell_lists(ctx, false))) return r;
  if ((r = build_cell_lists(ctx, true))) return r;
  CK(cudaStreamSynchronize(ctx->stream));   // the caller's host buffers are free again
  ctx->store_set = true;
  ctx->seq_next = P;
  return PMVSB_OK;
}

static int need_store(pmvsb_ctx* ctx, bool depth);

int pmvsb_store_append(pmvsb_ctx* ctx, int n, const float* coords, const float* normals, const float* ncc, const float* dscale,
                       const int32_t* img_off, const int32_t* images, const int32_t* grids, const int32_t* vimg_off,
                       const int32_t* vimages, const int32_t* vgrids, const int32_t* timages) {
  int r = need_store(ctx, false);
  if (r) return r;
  if (n < 0 || (n > 0 && (!coords || !normals || !ncc || !dscale || !img_off || !vimg_off || !timages))) return fail(ctx, PMVSB_EINVAL, "store_append: bad argument");
  if (n == 0) return PMVSB_OK;
  const int dE = img_off[n] - img_off[0], dVE = vimg_off[n] - vimg_off[0];
  if ((dE > 0 && (!images || !grids)) || (dVE > 0 && (!vimages || !vgrids))) return fail(ctx, PMVSB_EINVAL, "store_append: null list");
  if ((r = check_lists(ctx, "store_append", n, img_off, images, grids, false))) return r;
  if ((r = check_lists(ctx, "store_append", n, vimg_off, vimages, vgrids, true))) return r;
  StoreBufs& b = ctx->sb;
  const int P = ctx->store.P, E = ctx->store_entries, VE = ctx->store_ventries;
  std::vector<int32_t> off(n), voff(n);
  for (int i = 0; i < n; ++i) { off[i] = E + img_off[i + 1] - img_off[0]; voff[i] = VE + vimg_off[i + 1] - vimg_off[0]; }
  if ((r = dvec_put(ctx, b.coords, (size_t)4 * P, coords, (size_t)4 * n))) return r;
  if ((r = dvec_put(ctx, b.normals, (size_t)4 * P, normals, (size_t)4 * n))) return r;
  if ((r = dvec_put(ctx, b.ncc, (size_t)P, ncc, (size_t)n))) return r;
  if ((r = dvec_put(ctx, b.dscale, (size_t)P, dscale, (size_t)n))) return r;
  if ((r = dvec_put(ctx, b.timages, (size_t)P, timages, (size_t)n))) return r;
  if ((r = dvec_put(ctx, b.img_off, (size_t)P + 1, off.data(), (size_t)n))) return r;
  if ((r = dvec_put(ctx, b.images, (size_t)E, images, (size_t)dE))) return r;
  if ((r = dvec_put(ctx, b.grids, (size_t)2 * E, grids, (size_t)2 * dE))) return r;
  if ((r = dvec_put(ctx, b.vimg_off, (size_t)P + 1, voff.data(), (size_t)n))) return r;
  if ((r = dvec_put(ctx, b.vimages, (size_t)VE, vimages, (size_t)dVE))) return r;
  if ((r = dvec_put(ctx, b.vgrids, (size_t)2 * VE, vgrids, (size_t)2 * dVE))) return r;
  if ((r = dvec_reserve(ctx, b.entry_patch, (size_t)E + dE, (size_t)E))) return r;
  if ((r = dvec_reserve(ctx, b.ventry_patch, (size_t)std::max(VE + dVE, 1), (size_t)VE))) return r;
  if ((r = dvec_reserve(ctx, b.seq, (size_t)P + n, (size_t)P))) return r;
  k_iota<<<(n + 255) / 256, 256, 0, ctx->stream>>>(b.seq.p, P, n, ctx->seq_next);   // appended patches are the youngest
  ++ctx->launches;
  ctx->seq_next += n;
  store_view(ctx);
  k_entry_owner<<<(n + 255) / 256, 256, 0, ctx->stream>>>(P, n, b.img_off.p, b.entry_patch.p);
  k_entry_owner<<<(n + 255) / 256, 256, 0, ctx->stream>>>(P, n, b.vimg_off.p, b.ventry_patch.p);
  ctx->launches += 2;
  ctx->store_entries = E + dE; ctx->store_ventries = VE + dVE;
  if ((r = build_cell_lists(ctx, false))) return r;
  if ((r = build_cell_lists(ctx, true))) return r;
  if (ctx->depth_built) {   // CPatchOrganizerS::updateDepthMaps (patchOrganizerS.cpp:351-381)
    const long long t = (long long)n * ctx->tnum;
    k_depth_maps<<<(unsigned)((t + 255) / 256), 256, 0, ctx->stream>>>(ctx->scene, ctx->store, P, n);
    ++ctx->launches;
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->store.P = P + n;
  return PMVSB_OK;
}

static int need_store(pmvsb_ctx* ctx, bool depth) {
  int r = check_ready(ctx);
  if (r) return r;
  if (!ctx->store_set) return fail(ctx, PMVSB_ESTATE, "no patch table: call pmvsb_store_upload first");
  if (depth && ctx->depth_flag != 0 && !ctx->depth_built) return fail(ctx, PMVSB_ESTATE, "depth maps not built: call pmvsb_build_depth_maps first");
  if (!depth && ctx->store_appended) return fail(ctx, PMVSB_ESTATE, "table was extended by pmvsb_depth_maps_add (coordinates only): upload it again first");
  ctx->store.depth_flag = ctx->depth_flag;
  ctx->store.ncc_threshold = ctx->ncc_threshold;
  return PMVSB_OK;
}

int pmvsb_build_depth_maps(pmvsb_ctx* ctx) {
  int r = need_store(ctx, false);
  if (r) return r;
  CK(cudaMemsetAsync(ctx->store.dp, 0xff, sizeof(unsigned long long) * (ctx->store_cells ? ctx->store_cells : 1), ctx->stream));
  const long long n = (long long)ctx->store.P * ctx->tnum;
  if (n > 0) {
    k_depth_maps<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->scene, ctx->store, 0, ctx->store.P);
    ++ctx->launches;
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->depth_built = true;
  return PMVSB_OK;
}

int pmvsb_depth_maps_add(pmvsb_ctx* ctx, int n, const float* coords) {
  int r = need_store(ctx, true);
  if (r) return r;
  if (n < 0 || (n > 0 && !coords)) return fail(ctx, PMVSB_EINVAL, "depth_maps_add: bad argument");
  if (!ctx->depth_built) return fail(ctx, PMVSB_ESTATE, "depth maps not built: call pmvsb_build_depth_maps first");
  if (n == 0) return PMVSB_OK;
  if ((r = dvec_put(ctx, ctx->sb.coords, (size_t)4 * ctx->store.P, coords, (size_t)4 * n))) return r;
  store_view(ctx);
  const long long t = (long long)n * ctx->tnum;
  k_depth_maps<<<(unsigned)((t + 255) / 256), 256, 0, ctx->stream>>>(ctx->scene, ctx->store, ctx->store.P, n);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->store.P += n;
  ctx->store_appended = true;
  return PMVSB_OK;
}

int pmvsb_download_depth_map(pmvsb_ctx* ctx, int image, int32_t* patch_id) {
  int r = need_store(ctx, true);
  if (r) return r;
  if (!ctx->depth_built || image < 0 || image >= ctx->tnum || !patch_id) return fail(ctx, PMVSB_EINVAL, "download_depth_map: bad argument or maps not built");
  int gw, gh;
  pmvsb_grid_dims(ctx, image, &gw, &gh);
  std::vector<int32_t> base(ctx->tnum + 1, 0);
  for (int i = 0; i < ctx->tnum; ++i) { int a, b; pmvsb_grid_dims(ctx, i, &a, &b); base[i + 1] = base[i] + a * b; }
  std::vector<unsigned long long> keys((size_t)gw * gh);
  CK(cudaMemcpyAsync(keys.data(), ctx->store.dp + base[image], sizeof(unsigned long long) * keys.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < keys.size(); ++i) patch_id[i] = keys[i] == ~0ull ? -1 : (int32_t)(keys[i] & 0xffffffffull);
  return PMVSB_OK;
}

int pmvsb_set_vimages_store(pmvsb_ctx* ctx, int vcap, int32_t* vimages, int32_t* vgrids, int32_t* nv) {
  int r = need_store(ctx, true);
  if (r) return r;
  if (vcap < 1 || !vimages || !vgrids || !nv) return fail(ctx, PMVSB_EINVAL, "set_vimages_store: bad argument");
  const int P = ctx->store.P;
  if (P == 0) return PMVSB_OK;
  DevBuf<int32_t> dv, dg, dn;
  CK(dv.alloc((size_t)vcap * P)); CK(dg.alloc((size_t)2 * vcap * P)); CK(dn.alloc(P));
  CK(cudaMemsetAsync(dv.p, 0xff, sizeof(int32_t) * (size_t)vcap * P, ctx->stream));
  CK(cudaMemsetAsync(dg.p, 0xff, sizeof(int32_t) * (size_t)2 * vcap * P, ctx->stream));
  k_set_vimages<<<(P + 3) / 4, 128, 0, ctx->stream>>>(ctx->scene, ctx->store, vcap, dv.p, dg.p, dn.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(vimages, dv.p, sizeof(int32_t) * (size_t)vcap * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(vgrids, dg.p, sizeof(int32_t) * (size_t)2 * vcap * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(nv, dn.p, sizeof(int32_t) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

int pmvsb_filter_exact_store(pmvsb_ctx* ctx, uint8_t* safe) {
  int r = need_store(ctx, true);
  if (r) return r;
  if (!safe) return fail(ctx, PMVSB_EINVAL, "filter_exact_store: null pointer");
  const int E = ctx->store_entries;
  if (E == 0) return PMVSB_OK;
  DevBuf<uint8_t> ds;
  CK(ds.alloc(E));
  k_filter_exact<<<(E + 127) / 128, 128, 0, ctx->stream>>>(ctx->scene, ctx->store, E, ds.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(safe, ds.p, E, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

int pmvsb_compute_gains_store(pmvsb_ctx* ctx, float* gains) {
  int r = need_store(ctx, false);
  if (r) return r;
  if (!gains) return fail(ctx, PMVSB_EINVAL, "compute_gains_store: null pointer");
  const int P = ctx->store.P;
  if (P == 0) return PMVSB_OK;
  DevBuf<float> dg;
  CK(dg.alloc(P));
  k_gains<<<(P + 127) / 128, 128, 0, ctx->stream>>>(ctx->scene, ctx->store, dg.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(gains, dg.p, sizeof(float) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

static int store_update_vimages_impl(pmvsb_ctx* ctx, int additive, int32_t* total);

int pmvsb_store_update_vimages(pmvsb_ctx* ctx, int additive, int32_t* total) {
  int r = need_store(ctx, true);
  if (r) return r;
  return store_update_vimages_impl(ctx, additive, total);
}

static int store_update_vimages_impl(pmvsb_ctx* ctx, int additive, int32_t* total) {
  int r = 0;
  if (ctx->store_appended) return fail(ctx, PMVSB_ESTATE, "table was extended by pmvsb_depth_maps_add (coordinates only): upload it again first");
  StoreBufs& b = ctx->sb;
  const int P = ctx->store.P;
  if (total) *total = 0;
  if (P == 0) return PMVSB_OK;
  if ((r = dvec_reserve(ctx, b.alt_voff, (size_t)P + 1))) return r;
  CK(cudaMemsetAsync(b.alt_voff.p + P, 0, sizeof(int32_t), ctx->stream));
  k_store_vimages<false><<<(P + 3) / 4, 128, 0, ctx->stream>>>(ctx->scene, ctx->store, additive, b.alt_voff.p, nullptr, nullptr, nullptr);
  ++ctx->launches;
  if ((r = device_scan(ctx, b.alt_voff.p, P + 1))) return r;
  int32_t VE = 0;
  CK(cudaMemcpyAsync(&VE, b.alt_voff.p + P, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if ((r = dvec_reserve(ctx, b.alt_vimages, (size_t)std::max(VE, 1)))) return r;
  if ((r = dvec_reserve(ctx, b.alt_vgrids, (size_t)2 * std::max(VE, 1)))) return r;
  k_store_vimages<true><<<(P + 3) / 4, 128, 0, ctx->stream>>>(ctx->scene, ctx->store, additive, nullptr, b.alt_voff.p, b.alt_vimages.p, b.alt_vgrids.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  std::swap(b.vimg_off, b.alt_voff); std::swap(b.vimages, b.alt_vimages); std::swap(b.vgrids, b.alt_vgrids);
  ctx->store_ventries = VE;
  if ((r = dvec_reserve(ctx, b.ventry_patch, (size_t)std::max(VE, 1)))) return r;
  store_view(ctx);
  k_entry_owner<<<(P + 255) / 256, 256, 0, ctx->stream>>>(0, P, b.vimg_off.p, b.ventry_patch.p);
  ++ctx->launches;
  if ((r = build_cell_lists(ctx, true))) return r;
  CK(cudaStreamSynchronize(ctx->stream));
  if (total) *total = VE;
  return PMVSB_OK;
}

int pmvsb_store_download_vimages(pmvsb_ctx* ctx, int32_t* vimg_off, int32_t* vimages, int32_t* vgrids) {
  int r = need_store(ctx, false);
  if (r) return r;
  if (!vimg_off || ((!vimages || !vgrids) && ctx->store_ventries > 0)) return fail(ctx, PMVSB_EINVAL, "store_download_vimages: null pointer");
  const int P = ctx->store.P, VE = ctx->store_ventries;
  CK(cudaMemcpyAsync(vimg_off, ctx->sb.vimg_off.p, sizeof(int32_t) * ((size_t)P + 1), cudaMemcpyDeviceToHost, ctx->stream));
  if (VE > 0) {
    CK(cudaMemcpyAsync(vimages, ctx->sb.vimages.p, sizeof(int32_t) * (size_t)VE, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(vgrids, ctx->sb.vgrids.p, sizeof(int32_t) * (size_t)2 * VE, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}



// ---- seed candidate enumeration (pmvs_seed.cuh) ------------------------------------------------------------------
int pmvsb_set_features(pmvsb_ctx* ctx, int index, int n, const float* xy, const int32_t* type) {
  if (!ctx || index < 0 || index >= ctx->num || n < 0 || (n > 0 && (!xy || !type))) return fail(ctx, PMVSB_EINVAL, "set_features: bad argument");
  if (ctx->feat_xy.empty()) { ctx->feat_xy.resize(ctx->num); ctx->feat_type.resize(ctx->num); }
  ctx->feat_xy[index].assign(xy, xy + (size_t)2 * n);
  ctx->feat_type[index].assign(type, type + n);
  ctx->feat_dirty = true;
  return PMVSB_OK;
}

static void grid_geometry(pmvsb_ctx* ctx) {   // patchOrganizerS.cpp:72-77
  if (!ctx->h_gw.empty()) return;
  ctx->h_gw.resize(ctx->num); ctx->h_gh.resize(ctx->num); ctx->h_base.assign(ctx->tnum + 1, 0);
  for (int i = 0; i < ctx->num; ++i) {
    ctx->h_gw[i] = (ctx->images[i].w[ctx->level] + ctx->csize - 1) / ctx->csize;
    ctx->h_gh[i] = (ctx->images[i].h[ctx->level] + ctx->csize - 1) / ctx->csize;
  }
  for (int i = 0; i < ctx->tnum; ++i) ctx->h_base[i + 1] = ctx->h_base[i] + ctx->h_gw[i] * ctx->h_gh[i];
}

// CSeed::readPoints (seed.cpp:23-36): features binned by cell, in the order they were handed over
static int upload_feature_bins(pmvsb_ctx* ctx) {
  if (!ctx->feat_dirty) return PMVSB_OK;
  grid_geometry(ctx);
  if (ctx->feat_xy.empty()) { ctx->feat_xy.resize(ctx->num); ctx->feat_type.resize(ctx->num); }
  const int num = ctx->num;
  ctx->h_feat_base.assign(num + 1, 0); ctx->h_fcell_base.assign(num + 1, 0);
  for (int i = 0; i < num; ++i) {
    ctx->h_feat_base[i + 1] = ctx->h_feat_base[i] + (int)ctx->feat_type[i].size();
    ctx->h_fcell_base[i + 1] = ctx->h_fcell_base[i] + ctx->h_gw[i] * ctx->h_gh[i];
  }
  const int nf = ctx->h_feat_base[num], cells = ctx->h_fcell_base[num];
  std::vector<float> xy((size_t)2 * std::max(nf, 1));
  std::vector<int32_t> type(std::max(nf, 1)), cell_of(std::max(nf, 1), -1);
  ctx->h_fcell_off.assign((size_t)cells + 1, 0);
  ctx->h_flist.assign(std::max(nf, 1), 0);
  for (int i = 0; i < num; ++i)
    for (int k = 0; k < (int)ctx->feat_type[i].size(); ++k) {
      const int f = ctx->h_feat_base[i] + k;
      const float x = ctx->feat_xy[i][2 * k], y = ctx->feat_xy[i][2 * k + 1];
      xy[2 * f] = x; xy[2 * f + 1] = y; type[f] = ctx->feat_type[i][k];
      const int ix = ((int)std::floor(x + 0.5f)) / ctx->csize, iy = ((int)std::floor(y + 0.5f)) / ctx->csize;
      if (ix < 0 || ix >= ctx->h_gw[i] || iy < 0 || iy >= ctx->h_gh[i]) continue;   // cannot happen for detector output
      cell_of[f] = ctx->h_fcell_base[i] + iy * ctx->h_gw[i] + ix;
      ++ctx->h_fcell_off[cell_of[f] + 1];
    }
  for (int c = 0; c < cells; ++c) ctx->h_fcell_off[c + 1] += ctx->h_fcell_off[c];
  std::vector<int32_t> cur(ctx->h_fcell_off.begin(), ctx->h_fcell_off.end() - 1);
  for (int f = 0; f < nf; ++f)
    if (cell_of[f] >= 0) ctx->h_flist[cur[cell_of[f]]++] = f;
  int r;
  if ((r = dvec_put(ctx, ctx->d_fxy, 0, xy.data(), xy.size())) || (r = dvec_put(ctx, ctx->d_ftype, 0, type.data(), type.size())) ||
      (r = dvec_put(ctx, ctx->d_fcell_base, 0, ctx->h_fcell_base.data(), ctx->h_fcell_base.size())) ||
      (r = dvec_put(ctx, ctx->d_fcell_off, 0, ctx->h_fcell_off.data(), ctx->h_fcell_off.size())) ||
      (r = dvec_put(ctx, ctx->d_flist, 0, ctx->h_flist.data(), ctx->h_flist.size())) ||
      (r = dvec_put(ctx, ctx->d_fgw, 0, ctx->h_gw.data(), ctx->h_gw.size())) || (r = dvec_put(ctx, ctx->d_fgh, 0, ctx->h_gh.data(), ctx->h_gh.size())))
    return r;
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->feat_dirty = false;
  return PMVSB_OK;
}

// Image::setF (include/image/camera.hpp:129-151) at the working level: double copies of the float projection rows, 4x4
// determinants through the triple cross product of numeric/vec4.hpp:216-231 (the reference's operation order)
static double det4_rows(const double* a, const double* b, const double* c, const double* d) {
  const double d1 = (c[2] * d[3]) - (c[3] * d[2]), d2 = (c[1] * d[3]) - (c[3] * d[1]), d3 = (c[1] * d[2]) - (c[2] * d[1]);
  const double d4 = (c[0] * d[3]) - (c[3] * d[0]), d5 = (c[0] * d[2]) - (c[2] * d[0]), d6 = (c[0] * d[1]) - (c[1] * d[0]);
  const double x0 = -b[1] * d1 + b[2] * d2 - b[3] * d3, x1 = b[0] * d1 - b[2] * d4 + b[3] * d5;
  const double x2 = -b[0] * d2 + b[1] * d4 - b[3] * d6, x3 = b[0] * d3 - b[1] * d5 + b[2] * d6;
  return a[0] * x0 + a[1] * x1 + a[2] * x2 + a[3] * x3;
}
static void level_projection(const pmvsb_ctx* ctx, int image, double P[3][4]) {
  const float sc = 1.0f / (float)(1 << ctx->level);
  for (int k = 0; k < 4; ++k) {
    P[0][k] = (double)(ctx->cams[image].P0[0][k] * sc);
    P[1][k] = (double)(ctx->cams[image].P0[1][k] * sc);
    P[2][k] = (double)ctx->cams[image].P0[2][k];
  }
}
static void set_F(const pmvsb_ctx* ctx, int lhs, int rhs, double* F) {
  double a[3][4], b[3][4];
  level_projection(ctx, lhs, a);
  level_projection(ctx, rhs, b);
  F[0] = det4_rows(a[1], a[2], b[1], b[2]); F[1] = det4_rows(a[1], a[2], b[2], b[0]); F[2] = det4_rows(a[1], a[2], b[0], b[1]);
  F[3] = det4_rows(a[2], a[0], b[1], b[2]); F[4] = det4_rows(a[2], a[0], b[2], b[0]); F[5] = det4_rows(a[2], a[0], b[0], b[1]);
  F[6] = det4_rows(a[0], a[1], b[1], b[2]); F[7] = det4_rows(a[0], a[1], b[2], b[0]); F[8] = det4_rows(a[0], a[1], b[0], b[1]);
}

int pmvsb_seed_candidates(pmvsb_ctx* ctx, int index, int nviews, const int32_t* views, const uint8_t* blocked, int cap_ref, int32_t* nref,
                          int32_t* ref_feature, int32_t* ref_cell, int32_t* ref_start, int32_t* ref_count, int cap, int32_t* total, float* coords,
                          int32_t* other_image, int32_t* other_feature, float* resp) {
  int r = check_ready(ctx);
  if (r) return r;
  if (index < 0 || index >= ctx->num || nviews < 0 || nviews > kMaxTau || (nviews > 0 && !views) || !blocked || !nref || !total || cap_ref < 0 || cap < 0)
    return fail(ctx, PMVSB_EINVAL, "seed_candidates: bad argument");
  for (int k = 0; k < nviews; ++k)
    if (views[k] < 0 || views[k] >= ctx->num || views[k] == index) return fail(ctx, PMVSB_EINVAL, "seed_candidates: bad view list");
  if ((r = upload_feature_bins(ctx))) return r;
  // the reference features: cells of `index` in row-major order that may take a patch, their features in bin order (seed.cpp:143-151)
  std::vector<int32_t> feats, cells;
  const int gw = ctx->h_gw[index], gh = ctx->h_gh[index], cb = ctx->h_fcell_base[index];
  for (int c = 0; c < gw * gh; ++c) {
    if (blocked[cb + c]) continue;
    for (int j = ctx->h_fcell_off[cb + c]; j < ctx->h_fcell_off[cb + c + 1]; ++j) { feats.push_back(ctx->h_flist[j]); cells.push_back(c); }
  }
  const int R = (int)feats.size();
  *nref = R; *total = 0;
  if (R == 0 || nviews == 0) { *nref = nviews == 0 ? 0 : R; return PMVSB_OK; }
  if (R > cap_ref) return PMVSB_OK;   // the caller sizes its arrays from *nref and calls again
  if (!ref_feature || !ref_cell || !ref_start || !ref_count) return fail(ctx, PMVSB_EINVAL, "seed_candidates: null output");
  SeedParams sp;
  std::memset(&sp, 0, sizeof(sp));
  sp.index = index; sp.nviews = nviews; sp.ep_threshold = 2.0f;   // findMatch.cpp:106
  for (int k = 0; k < nviews; ++k) { sp.view[k].image = views[k]; set_F(ctx, index, views[k], sp.view[k].F); }
  SeedDev sd;
  sd.fxy = ctx->d_fxy.p; sd.ftype = ctx->d_ftype.p; sd.fcell_base = ctx->d_fcell_base.p; sd.fcell_off = ctx->d_fcell_off.p; sd.flist = ctx->d_flist.p;
  sd.gw = ctx->d_fgw.p; sd.gh = ctx->d_fgh.p;
  const int allcells = ctx->h_fcell_base[ctx->num];
  // `blocked` covers every cell of every image (24 MB for a 32-view 4000x3000 cluster) and changes in a few places between two
  // calls of a seed round: the device copy is brought up to date from a host shadow, 16 KB blocks that differ travelling as runs
  {
    std::vector<uint8_t>& shadow = ctx->h_blocked_shadow;
    if (shadow.size() != (size_t)allcells || !ctx->d_blocked.p || ctx->d_blocked.cap < (size_t)allcells) {
      if ((r = dvec_put(ctx, ctx->d_blocked, 0, blocked, (size_t)allcells))) return r;
      shadow.assign(blocked, blocked + allcells);
    } else {
      const size_t B = 16384;
      size_t run_lo = 0, run_hi = 0;   // pending run [run_lo, run_hi)
      for (size_t at = 0; at < (size_t)allcells; at += B) {
        const size_t n = std::min(B, (size_t)allcells - at);
        if (std::memcmp(shadow.data() + at, blocked + at, n) != 0) {
          std::memcpy(shadow.data() + at, blocked + at, n);
          if (run_hi == at && run_hi > run_lo) run_hi = at + n;
          else {
            if (run_hi > run_lo) CK(cudaMemcpyAsync(ctx->d_blocked.p + run_lo, shadow.data() + run_lo, run_hi - run_lo, cudaMemcpyHostToDevice, ctx->stream));
            run_lo = at; run_hi = at + n;
          }
        }
      }
      if (run_hi > run_lo) CK(cudaMemcpyAsync(ctx->d_blocked.p + run_lo, shadow.data() + run_lo, run_hi - run_lo, cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  DevBuf<int32_t> dfeat, dcount, dstart, dmisc;
  CK(dfeat.alloc(R)); CK(dcount.alloc(R)); CK(dstart.alloc(R)); CK(dmisc.alloc(4));
  CK(cudaMemcpyAsync(dfeat.p, feats.data(), sizeof(int32_t) * (size_t)R, cudaMemcpyHostToDevice, ctx->stream));
  size_t capacity = std::max<size_t>(ctx->d_seed_out.cap, (size_t)1 << 18);
  int32_t misc[4] = {0, 0, 0, 0};   // cursor, features over kSeedCap, features that did not fit the buffer, -
  for (int attempt = 0; attempt < 2; ++attempt) {
    if ((r = dvec_reserve(ctx, ctx->d_seed_out, capacity))) return r;
    capacity = ctx->d_seed_out.cap;
    CK(cudaMemsetAsync(dmisc.p, 0, sizeof(int32_t) * 4, ctx->stream));
    k_seed_candidates<<<(R + kSeedWarps - 1) / kSeedWarps, kSeedWarps * 32, 0, ctx->stream>>>(ctx->scene, sd, sp, ctx->d_blocked.p, R, dfeat.p, dcount.p, dstart.p,
                                                                                              dmisc.p, (int)std::min<size_t>(capacity, 0x7fffffff), ctx->d_seed_out.p,
                                                                                              dmisc.p + 1);
    ++ctx->launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(misc, dmisc.p, sizeof(misc), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (misc[2] == 0) break;
    capacity = (size_t)misc[0] + 1024;   // the cursor counted every hit: the second pass fits
  }
  if (misc[1] != 0) return fail(ctx, PMVSB_ERANGE, "seed_candidates: a feature has more than " + std::to_string(kSeedCap) + " epipolar candidates");
  if (misc[2] != 0) return fail(ctx, PMVSB_ENOMEM, "seed_candidates: candidate buffer too small");
  *total = misc[0];
  CK(cudaMemcpyAsync(ref_count, dcount.p, sizeof(int32_t) * (size_t)R, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ref_start, dstart.p, sizeof(int32_t) * (size_t)R, cudaMemcpyDeviceToHost, ctx->stream));
  for (int k = 0; k < R; ++k) { ref_feature[k] = feats[k] - ctx->h_feat_base[index]; ref_cell[k] = cells[k]; }
  if (misc[0] > cap) { CK(cudaStreamSynchronize(ctx->stream)); return PMVSB_OK; }   // sizes known now: call again with room for *total
  if (misc[0] > 0) {
    if (!coords || !other_image || !other_feature || !resp) return fail(ctx, PMVSB_EINVAL, "seed_candidates: null output");
    std::vector<SeedHit> hits((size_t)misc[0]);
    CK(cudaMemcpyAsync(hits.data(), ctx->d_seed_out.p, sizeof(SeedHit) * hits.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < hits.size(); ++i) {
      const SeedHit& h = hits[i];
      coords[4 * i] = h.x; coords[4 * i + 1] = h.y; coords[4 * i + 2] = h.z; coords[4 * i + 3] = 1.0f;
      other_image[i] = h.other_image; other_feature[i] = h.other_feature - ctx->h_feat_base[h.other_image]; resp[i] = h.resp;
    }
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

// ---- the table reorganised on the device (pmvs_table.cuh) -------------------------------------------------------
int pmvsb_store_set_seq(pmvsb_ctx* ctx, int first, int n, const int32_t* seq) {
  int r = need_store(ctx, false);
  if (r) return r;
  if (first < 0 || n < 0 || first + n > ctx->store.P || (n > 0 && !seq)) return fail(ctx, PMVSB_EINVAL, "store_set_seq: bad range");
  if (n == 0) return PMVSB_OK;
  CK(cudaMemcpyAsync(ctx->sb.seq.p + first, seq, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; ++i) ctx->seq_next = std::max(ctx->seq_next, seq[i] + 1);
  return PMVSB_OK;
}

int pmvsb_store_counts(pmvsb_ctx* ctx, int32_t* P, int32_t* entries, int32_t* ventries) {
  int r = need_store(ctx, false);
  if (r) return r;
  if (P) *P = ctx->store.P;
  if (entries) *entries = ctx->store_entries;
  if (ventries) *ventries = ctx->store_ventries;
  return PMVSB_OK;
}

static int store_update_vimages_impl(pmvsb_ctx* ctx, int additive, int32_t* total);

int pmvsb_store_rebuild(pmvsb_ctx* ctx, const uint8_t* keep, int additive, int32_t* new_count, int32_t* perm_out) {
  int r = need_store(ctx, false);
  if (r) return r;
  StoreBufs& b = ctx->sb;
  const int P = ctx->store.P, cells = ctx->store_cells;
  if (new_count) *new_count = 0;
  DevBuf<uint8_t> dkeep;
  if (keep && P > 0) {
    CK(dkeep.alloc(P));
    CK(cudaMemcpyAsync(dkeep.p, keep, (size_t)P, cudaMemcpyHostToDevice, ctx->stream));
  }
  if ((r = dvec_reserve(ctx, b.first_cell, (size_t)std::max(P, 1)))) return r;
  if ((r = dvec_reserve(ctx, b.perm, (size_t)std::max(P, 1)))) return r;
  if ((r = dvec_reserve(ctx, b.bucket_off, (size_t)cells + 1))) return r;
  if ((r = dvec_reserve(ctx, b.cursor, (size_t)cells + 1))) return r;
  CK(cudaMemsetAsync(b.bucket_off.p, 0, sizeof(int32_t) * ((size_t)cells + 1), ctx->stream));
  CK(cudaMemsetAsync(b.cursor.p, 0, sizeof(int32_t) * ((size_t)cells + 1), ctx->stream));
  int32_t newP = 0;
  if (P > 0) {
    k_tab_first_cell<<<(P + 255) / 256, 256, 0, ctx->stream>>>(ctx->store, ctx->tnum, keep ? dkeep.p : nullptr, b.first_cell.p, b.bucket_off.p);
    ++ctx->launches;
    if ((r = device_scan(ctx, b.bucket_off.p, cells + 1))) return r;
    k_tab_bucket_fill<<<(P + 255) / 256, 256, 0, ctx->stream>>>(P, b.first_cell.p, b.bucket_off.p, b.cursor.p, b.perm.p);
    k_tab_bucket_sort<<<(cells + 255) / 256, 256, 0, ctx->stream>>>(cells, b.bucket_off.p, b.perm.p, b.seq.p);
    ctx->launches += 2;
    CK(cudaMemcpyAsync(&newP, b.bucket_off.p + cells, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  // scalar fields and the image lists through the permutation, into the second set of buffers
  const size_t n1 = (size_t)std::max(newP, 1);
  if ((r = dvec_reserve(ctx, b.alt_coords, 4 * n1)) || (r = dvec_reserve(ctx, b.alt_normals, 4 * n1)) || (r = dvec_reserve(ctx, b.alt_ncc, n1)) ||
      (r = dvec_reserve(ctx, b.alt_dscale, n1)) || (r = dvec_reserve(ctx, b.alt_timages, n1)) || (r = dvec_reserve(ctx, b.alt_seq, n1)) ||
      (r = dvec_reserve(ctx, b.alt_img_off, n1 + 1)) || (r = dvec_reserve(ctx, b.alt_voff, n1 + 1)))
    return r;
  int32_t E = 0, VE = 0;
  if (newP > 0) {
    k_tab_gather_fields<<<(newP + 255) / 256, 256, 0, ctx->stream>>>(newP, b.perm.p, reinterpret_cast<const float4*>(b.coords.p),
        reinterpret_cast<const float4*>(b.normals.p), b.ncc.p, b.dscale.p, b.timages.p, b.seq.p, reinterpret_cast<float4*>(b.alt_coords.p),
        reinterpret_cast<float4*>(b.alt_normals.p), b.alt_ncc.p, b.alt_dscale.p, b.alt_timages.p, b.alt_seq.p);
    k_tab_gather_len<<<(newP + 256) / 256, 256, 0, ctx->stream>>>(newP, b.perm.p, b.img_off.p, b.alt_img_off.p);
    ctx->launches += 2;
    if ((r = device_scan(ctx, b.alt_img_off.p, newP + 1))) return r;
    CK(cudaMemcpyAsync(&E, b.alt_img_off.p + newP, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (additive) {
      k_tab_gather_len<<<(newP + 256) / 256, 256, 0, ctx->stream>>>(newP, b.perm.p, b.vimg_off.p, b.alt_voff.p);
      ++ctx->launches;
      if ((r = device_scan(ctx, b.alt_voff.p, newP + 1))) return r;
      CK(cudaMemcpyAsync(&VE, b.alt_voff.p + newP, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
      CK(cudaMemsetAsync(b.alt_voff.p, 0, sizeof(int32_t) * ((size_t)newP + 1), ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    if ((r = dvec_reserve(ctx, b.alt_images, (size_t)std::max(E, 1))) || (r = dvec_reserve(ctx, b.alt_grids, (size_t)2 * std::max(E, 1))) ||
        (r = dvec_reserve(ctx, b.alt_vimages, (size_t)std::max(VE, 1))) || (r = dvec_reserve(ctx, b.alt_vgrids, (size_t)2 * std::max(VE, 1))))
      return r;
    const int gl = (int)(((size_t)newP * 32 + 255) / 256);
    k_tab_gather_lists<<<gl, 256, 0, ctx->stream>>>(newP, b.perm.p, b.img_off.p, b.images.p, b.grids.p, b.alt_img_off.p, b.alt_images.p, b.alt_grids.p);
    ++ctx->launches;
    if (additive && VE > 0) {
      k_tab_gather_lists<<<gl, 256, 0, ctx->stream>>>(newP, b.perm.p, b.vimg_off.p, b.vimages.p, b.vgrids.p, b.alt_voff.p, b.alt_vimages.p, b.alt_vgrids.p);
      ++ctx->launches;
    }
  } else {
    CK(cudaMemsetAsync(b.alt_img_off.p, 0, sizeof(int32_t), ctx->stream));
    CK(cudaMemsetAsync(b.alt_voff.p, 0, sizeof(int32_t), ctx->stream));
  }
  CK(cudaGetLastError());
  if (perm_out && newP > 0) CK(cudaMemcpyAsync(perm_out, b.perm.p, sizeof(int32_t) * (size_t)newP, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  std::swap(b.coords, b.alt_coords); std::swap(b.normals, b.alt_normals); std::swap(b.ncc, b.alt_ncc); std::swap(b.dscale, b.alt_dscale);
  std::swap(b.timages, b.alt_timages); std::swap(b.seq, b.alt_seq); std::swap(b.img_off, b.alt_img_off); std::swap(b.images, b.alt_images);
  std::swap(b.grids, b.alt_grids); std::swap(b.vimg_off, b.alt_voff); std::swap(b.vimages, b.alt_vimages); std::swap(b.vgrids, b.alt_vgrids);
  ctx->store.P = newP;
  ctx->store_entries = E; ctx->store_ventries = VE;
  ctx->store_appended = false;
  if ((r = dvec_reserve(ctx, b.entry_patch, (size_t)std::max(E, 1)))) return r;
  if ((r = dvec_reserve(ctx, b.ventry_patch, (size_t)std::max(VE, 1)))) return r;
  store_view(ctx);
  if (newP > 0) {
    k_entry_owner<<<(newP + 255) / 256, 256, 0, ctx->stream>>>(0, newP, b.img_off.p, b.entry_patch.p);
    ++ctx->launches;
  }
  if ((r = build_cell_lists(ctx, false))) return r;
  // CFilter::setDepthMaps, then setVImagesVGrids for every patch and _vpgrids (filter.cpp:734-783)
  CK(cudaMemsetAsync(ctx->store.dp, 0xff, sizeof(unsigned long long) * (cells ? cells : 1), ctx->stream));
  const long long nt = (long long)newP * ctx->tnum;
  if (nt > 0) {
    k_depth_maps<<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(ctx->scene, ctx->store, 0, newP);
    ++ctx->launches;
  }
  CK(cudaGetLastError());
  ctx->depth_built = true;
  if (newP > 0 && additive != 2) {
    if ((r = store_update_vimages_impl(ctx, additive, nullptr))) return r;
  } else {
    if (newP > 0) {
      k_entry_owner<<<(newP + 255) / 256, 256, 0, ctx->stream>>>(0, newP, b.vimg_off.p, b.ventry_patch.p);
      ++ctx->launches;
    }
    if ((r = build_cell_lists(ctx, true))) return r;
    CK(cudaStreamSynchronize(ctx->stream));
  }
  if (new_count) *new_count = newP;
  return PMVSB_OK;
}

int pmvsb_filter_exact_apply_store(pmvsb_ctx* ctx, uint8_t* keep) {
  int r = need_store(ctx, true);
  if (r) return r;
  if (!keep) return fail(ctx, PMVSB_EINVAL, "filter_exact_apply_store: null pointer");
  if (ctx->wsize == 9 && ctx->num > 32) return fail(ctx, PMVSB_EINVAL, "filter_exact_apply_store: wsize 9 keeps at most 32 images per patch in the selection kernels");
  StoreBufs& b = ctx->sb;
  const int P = ctx->store.P, E = ctx->store_entries;
  if (P == 0) return PMVSB_OK;
  const int stride = std::min(ctx->num, PMVSB_MAX_VIEWS);
  DevBuf<uint8_t> safe, dkeep;
  CK(safe.alloc((size_t)std::max(E, 1))); CK(dkeep.alloc(P));
  if ((r = dvec_reserve(ctx, b.rows, (size_t)stride * P)) || (r = dvec_reserve(ctx, b.rows_n, (size_t)P + 1)) ||
      (r = dvec_reserve(ctx, b.row_cells, (size_t)2 * stride * P)) || (r = dvec_reserve(ctx, b.alt_img_off, (size_t)P + 1)))
    return r;
  if (E > 0) {
    k_filter_exact<<<(E + 127) / 128, 128, 0, ctx->stream>>>(ctx->scene, ctx->store, E, safe.p);
    ++ctx->launches;
  }
  k_tab_prune<<<(P + 127) / 128, 128, 0, ctx->stream>>>(ctx->store, ctx->tnum, ctx->min_image_num, safe.p, stride, b.rows.p, b.rows_n.p, b.timages.p);
  ++ctx->launches;
  // setRefImage + setGrids for the survivors (filter.cpp:277-280), one launch per view-capacity class
#define LAUNCH_SET_REF(W, MAXV, LO) \
  k_set_ref_image<W, MAXV, LO><<<P, 32, 0, ctx->stream>>>(ctx->scene, ctx->select, P, stride, b.coords.p, b.normals.p, b.rows.p, b.rows_n.p, b.row_cells.p); ++ctx->launches
  if (ctx->wsize == 5) {
    LAUNCH_SET_REF(5, 16, 0);
    if (stride > 16) { LAUNCH_SET_REF(5, 32, 16); }
    if (stride > 32) { LAUNCH_SET_REF(5, 64, 32); }
  } else if (ctx->wsize == 9) {
    LAUNCH_SET_REF(9, 16, 0);
    if (stride > 16) { LAUNCH_SET_REF(9, 32, 16); }
  } else {
    LAUNCH_SET_REF(7, 16, 0);
    if (stride > 16) { LAUNCH_SET_REF(7, 32, 16); }
    if (stride > 32) { LAUNCH_SET_REF(7, 64, 32); }
  }
#undef LAUNCH_SET_REF
  k_tab_flags_from_len<<<(P + 255) / 256, 256, 0, ctx->stream>>>(P, b.rows_n.p, dkeep.p);
  CK(cudaMemcpyAsync(b.alt_img_off.p, b.rows_n.p, sizeof(int32_t) * (size_t)P, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemsetAsync(b.alt_img_off.p + P, 0, sizeof(int32_t), ctx->stream));
  ++ctx->launches;
  if ((r = device_scan(ctx, b.alt_img_off.p, P + 1))) return r;
  int32_t newE = 0;
  CK(cudaMemcpyAsync(&newE, b.alt_img_off.p + P, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(keep, dkeep.p, (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if ((r = dvec_reserve(ctx, b.alt_images, (size_t)std::max(newE, 1))) || (r = dvec_reserve(ctx, b.alt_grids, (size_t)2 * std::max(newE, 1)))) return r;
  k_tab_rows_to_lists<<<(P + 127) / 128, 128, 0, ctx->stream>>>(P, stride, b.rows.p, b.row_cells.p, b.alt_img_off.p, b.alt_images.p, b.alt_grids.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  std::swap(b.img_off, b.alt_img_off); std::swap(b.images, b.alt_images); std::swap(b.grids, b.alt_grids);
  ctx->store_entries = newE;
  if ((r = dvec_reserve(ctx, b.entry_patch, (size_t)std::max(newE, 1)))) return r;
  store_view(ctx);
  k_entry_owner<<<(P + 255) / 256, 256, 0, ctx->stream>>>(0, P, b.img_off.p, b.entry_patch.p);
  ++ctx->launches;
  if ((r = build_cell_lists(ctx, false))) return r;   // the cells of the dropped images lose the patch (filter.cpp:240-252)
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

int pmvsb_small_group_edges_store(pmvsb_ctx* ctx, float neighbor_threshold, int32_t* adj_off, int32_t* adj, int cap, int32_t* total) {
  int r = need_store(ctx, false);
  if (r) return r;
  if (!adj_off || !total || cap < 0 || (cap > 0 && !adj)) return fail(ctx, PMVSB_EINVAL, "small_group_edges_store: bad argument");
  const int P = ctx->store.P;
  *total = 0;
  adj_off[0] = 0;
  if (P == 0) return PMVSB_OK;
  DevBuf<int32_t> doff, dadj;
  CK(doff.alloc((size_t)P + 1));
  k_tab_group_edges<false><<<(P + 128) / 128, 128, 0, ctx->stream>>>(ctx->scene, ctx->store, neighbor_threshold, doff.p, nullptr, nullptr);
  ++ctx->launches;
  if ((r = device_scan(ctx, doff.p, P + 1))) return r;
  CK(cudaMemcpyAsync(adj_off, doff.p, sizeof(int32_t) * ((size_t)P + 1), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *total = adj_off[P];
  if (*total > cap) return PMVSB_OK;   // the caller sizes adj from *total and calls again
  if (*total == 0) return PMVSB_OK;
  CK(dadj.alloc((size_t)*total));
  k_tab_group_edges<true><<<(P + 128) / 128, 128, 0, ctx->stream>>>(ctx->scene, ctx->store, neighbor_threshold, nullptr, doff.p, dadj.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(adj, dadj.p, sizeof(int32_t) * (size_t)*total, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

int pmvsb_filter_small_groups_store(pmvsb_ctx* ctx, float neighbor_threshold, uint8_t* keep, int32_t* group_threshold) {
  int r = need_store(ctx, false);
  if (r) return r;
  if (!keep) return fail(ctx, PMVSB_EINVAL, "filter_small_groups_store: null pointer");
  const int P = ctx->store.P;
  if (group_threshold) *group_threshold = std::max(20, P / 10000);
  if (P == 0) return PMVSB_OK;
  std::vector<int32_t> off((size_t)P + 1), adj;
  int32_t total = 0;
  if ((r = pmvsb_small_group_edges_store(ctx, neighbor_threshold, off.data(), nullptr, 0, &total))) return r;
  adj.resize((size_t)std::max(total, 1));
  if ((r = pmvsb_small_group_edges_store(ctx, neighbor_threshold, off.data(), adj.data(), total, &total))) return r;
  // the labelling walk of filterSmallGroups (filter.cpp:541-563): sequential by definition (a patch takes the label of the
  // first walk that reaches it), integer work over the adjacency the device built
  std::vector<int32_t> label((size_t)P, -1), queue;
  queue.reserve(P);
  int id = -1;
  for (int start = 0; start < P; ++start) {
    if (label[start] != -1) continue;
    label[start] = ++id;
    queue.clear();
    queue.push_back(start);
    for (size_t head = 0; head < queue.size(); ++head) {
      const int k = queue[head];
      for (int j = off[k]; j < off[k + 1]; ++j) {
        const int q = adj[j];
        if (label[q] != -1) continue;
        label[q] = id;
        queue.push_back(q);
      }
    }
  }
  std::vector<int32_t> size((size_t)id + 1, 0);
  for (int p = 0; p < P; ++p) ++size[label[p]];
  const int threshold = std::max(20, P / 10000);
  for (int p = 0; p < P; ++p) keep[p] = size[label[p]] < threshold ? 0 : 1;
  return PMVSB_OK;
}

int pmvsb_store_download_lists(pmvsb_ctx* ctx, int32_t* seq, int32_t* timages, int32_t* img_off, int32_t* images, int32_t* grids) {
  int r = need_store(ctx, false);
  if (r) return r;
  const int P = ctx->store.P, E = ctx->store_entries;
  const StoreBufs& b = ctx->sb;
  if (seq && P > 0) CK(cudaMemcpyAsync(seq, b.seq.p, sizeof(int32_t) * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
  if (timages && P > 0) CK(cudaMemcpyAsync(timages, b.timages.p, sizeof(int32_t) * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
  if (img_off) CK(cudaMemcpyAsync(img_off, b.img_off.p, sizeof(int32_t) * ((size_t)P + 1), cudaMemcpyDeviceToHost, ctx->stream));
  if (images && E > 0) CK(cudaMemcpyAsync(images, b.images.p, sizeof(int32_t) * (size_t)E, cudaMemcpyDeviceToHost, ctx->stream));
  if (grids && E > 0) CK(cudaMemcpyAsync(grids, b.grids.p, sizeof(int32_t) * (size_t)2 * E, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

int pmvsb_download_cell_lists(pmvsb_ctx* ctx, int visible, int32_t* cell_off, int32_t* cell_patch) {
  int r = need_store(ctx, false);
  if (r) return r;
  if (!cell_off) return fail(ctx, PMVSB_EINVAL, "download_cell_lists: null pointer");
  const int cells = ctx->store_cells;
  const DVec<int32_t>& off = visible ? ctx->sb.vcell_off : ctx->sb.cell_off;
  const DVec<int32_t>& lst = visible ? ctx->sb.vcell_patch : ctx->sb.cell_patch;
  CK(cudaMemcpyAsync(cell_off, off.p, sizeof(int32_t) * ((size_t)cells + 1), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (cell_patch && cell_off[cells] > 0) {
    CK(cudaMemcpyAsync(cell_patch, lst.p, sizeof(int32_t) * (size_t)cell_off[cells], cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return PMVSB_OK;
}

int pmvsb_find_empty_blocks_store(pmvsb_ctx* ctx, int n, const int32_t* ids, uint8_t* mask, float* radius) {
  int r = need_store(ctx, false);
  if (r) return r;
  if (n < 0 || (n > 0 && (!ids || !mask || !radius))) return fail(ctx, PMVSB_EINVAL, "find_empty_blocks_store: bad argument");
  if (n == 0) return PMVSB_OK;
  for (int i = 0; i < n; ++i)
    if (ids[i] < 0 || ids[i] >= ctx->store.P) return fail(ctx, PMVSB_EINVAL, "find_empty_blocks_store: patch index out of range");
  DevBuf<int32_t> di;
  DevBuf<uint8_t> dm;
  DevBuf<float> dr;
  CK(di.alloc(n)); CK(dm.alloc(n)); CK(dr.alloc(n));
  CK(cudaMemcpyAsync(di.p, ids, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  k_find_empty_blocks<<<(n + 3) / 4, 128, 0, ctx->stream>>>(ctx->scene, ctx->store, n, di.p, dm.p, dr.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(mask, dm.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(radius, dr.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

int pmvsb_filter_neighbor_store(pmvsb_ctx* ctx, float quad, uint8_t* reject, float* residual, int32_t* ncount, int32_t* overflow) {
  int r = need_store(ctx, false);
  if (r) return r;
  if (!reject) return fail(ctx, PMVSB_EINVAL, "filter_neighbor_store: null pointer");
  if (overflow) *overflow = 0;
  const int P = ctx->store.P;
  if (P == 0) return PMVSB_OK;
  DevBuf<uint8_t> dj;
  DevBuf<float> dres;
  DevBuf<int32_t> dn, dov;
  CK(dj.alloc(P)); CK(dres.alloc(P)); CK(dn.alloc(P)); CK(dov.alloc(1));
  CK(cudaMemsetAsync(dov.p, 0, sizeof(int32_t), ctx->stream));
  k_filter_neighbor<<<(P + kNbWarps - 1) / kNbWarps, kNbWarps * 32, 0, ctx->stream>>>(ctx->scene, ctx->store, quad, ctx->tau, dj.p, dres.p, dn.p, dov.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(reject, dj.p, (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
  if (residual) CK(cudaMemcpyAsync(residual, dres.p, sizeof(float) * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
  if (ncount) CK(cudaMemcpyAsync(ncount, dn.p, sizeof(int32_t) * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
  int32_t ov = 0;
  CK(cudaMemcpyAsync(&ov, dov.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (overflow) *overflow = ov;
  return PMVSB_OK;
}

int pmvsb_check_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals, const float* ncc, const float* dscale,
                      const int32_t* timages, const int32_t* images, const int32_t* nimages, const int32_t* grids, int vstride,
                      const int32_t* vimages, const int32_t* nv, const int32_t* vgrids, float quad, float* gain, uint8_t* reject, int32_t* overflow) {
  int r = need_store(ctx, false);
  if (r) return r;
  if (P < 0 || stride < 1 || vstride < 1 || (P > 0 && (!coords || !normals || !ncc || !dscale || !timages || !images || !nimages || !grids ||
                                                      !vimages || !nv || !vgrids || !gain || !reject)))
    return fail(ctx, PMVSB_EINVAL, "check_batch: bad argument");
  if (overflow) *overflow = 0;
  if (P == 0) return PMVSB_OK;
  for (int p = 0; p < P; ++p) {   // the kernel trusts image indexes; cells are range-checked on the device
    for (int i = 0; i < std::min(nimages[p], stride); ++i)
      if (images[(size_t)p * stride + i] < 0 || images[(size_t)p * stride + i] >= ctx->num) return fail(ctx, PMVSB_EINVAL, "check_batch: image index out of range");
    for (int i = 0; i < std::min(nv[p], vstride); ++i)
      if (vimages[(size_t)p * vstride + i] < 0 || vimages[(size_t)p * vstride + i] >= ctx->tnum) return fail(ctx, PMVSB_EINVAL, "check_batch: vimage index out of range");
  }
  DevBuf<float> dc, dn, dncc, dds, dgain;
  DevBuf<int32_t> dti, dim, dni, dgr, dvi, dnv, dvg, dov;
  DevBuf<uint8_t> drej;
  CK(dc.alloc((size_t)4 * P)); CK(dn.alloc((size_t)4 * P)); CK(dncc.alloc(P)); CK(dds.alloc(P)); CK(dgain.alloc(P)); CK(dti.alloc(P));
  CK(dim.alloc((size_t)stride * P)); CK(dni.alloc(P)); CK(dgr.alloc((size_t)2 * stride * P)); CK(dvi.alloc((size_t)vstride * P)); CK(dnv.alloc(P));
  CK(dvg.alloc((size_t)2 * vstride * P)); CK(dov.alloc(1)); CK(drej.alloc(P));
  auto up = [&](void* d, const void* h, size_t bytes) { return cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream); };
  CK(up(dc.p, coords, sizeof(float) * 4 * (size_t)P)); CK(up(dn.p, normals, sizeof(float) * 4 * (size_t)P));
  CK(up(dncc.p, ncc, sizeof(float) * (size_t)P)); CK(up(dds.p, dscale, sizeof(float) * (size_t)P)); CK(up(dti.p, timages, sizeof(int32_t) * (size_t)P));
  CK(up(dim.p, images, sizeof(int32_t) * (size_t)stride * P)); CK(up(dni.p, nimages, sizeof(int32_t) * (size_t)P));
  CK(up(dgr.p, grids, sizeof(int32_t) * 2 * (size_t)stride * P)); CK(up(dvi.p, vimages, sizeof(int32_t) * (size_t)vstride * P));
  CK(up(dnv.p, nv, sizeof(int32_t) * (size_t)P)); CK(up(dvg.p, vgrids, sizeof(int32_t) * 2 * (size_t)vstride * P));
  CK(cudaMemsetAsync(dov.p, 0, sizeof(int32_t), ctx->stream));
  k_check_batch<<<(P + kNbWarps - 1) / kNbWarps, kNbWarps * 32, 0, ctx->stream>>>(ctx->scene, ctx->store, P, stride, vstride, dc.p, dn.p, dncc.p, dds.p,
                                                                                 dti.p, dim.p, dni.p, dgr.p, dvi.p, dnv.p, dvg.p, quad, ctx->tau, dgain.p,
                                                                                 drej.p, dov.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(gain, dgain.p, sizeof(float) * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(reject, drej.p, (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
  int32_t ov = 0;
  CK(cudaMemcpyAsync(&ov, dov.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (overflow) *overflow = ov;
  return PMVSB_OK;
}

int pmvsb_set_vimages_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals, const int32_t* images,
                            const int32_t* nimages, int vstride, int32_t* vimages, int32_t* nv, int32_t* vgrids) {
  int r = need_store(ctx, true);
  if (r) return r;
  if (!normals || !nimages || vstride < 1 || !vimages || !nv || !vgrids) return fail(ctx, PMVSB_EINVAL, "set_vimages_batch: bad argument");
  if (P == 0) return PMVSB_OK;
  for (int p = 0; p < P; ++p)
    for (int e = 0; e < std::min(nv[p], vstride); ++e)
      if (vimages[(size_t)p * vstride + e] < 0 || vimages[(size_t)p * vstride + e] >= ctx->tnum) return fail(ctx, PMVSB_EINVAL, "set_vimages_batch: vimage out of range");
  PatchStage st;
  r = stage_patches(ctx, st, P, stride, coords, normals, images, nimages, nullptr);
  if (r) return r;
  DevBuf<int32_t> dv, dg, dn;
  CK(dv.alloc((size_t)vstride * P)); CK(dg.alloc((size_t)2 * vstride * P)); CK(dn.alloc(P));
  CK(cudaMemcpyAsync(dv.p, vimages, sizeof(int32_t) * (size_t)vstride * P, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dg.p, vgrids, sizeof(int32_t) * (size_t)2 * vstride * P, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dn.p, nv, sizeof(int32_t) * P, cudaMemcpyHostToDevice, ctx->stream));
  k_set_vimages_batch<<<(P + 3) / 4, 128, 0, ctx->stream>>>(ctx->scene, ctx->store, P, stride, st.coords.p, st.normals.p, st.images.p,
                                                             st.nimages.p, vstride, dv.p, dn.p, dg.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(vimages, dv.p, sizeof(int32_t) * (size_t)vstride * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(vgrids, dg.p, sizeof(int32_t) * (size_t)2 * vstride * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(nv, dn.p, sizeof(int32_t) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

int pmvsb_set_ref_image_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals, int32_t* images,
                              int32_t* nimages, int32_t* grids) {
  int r = check_ready(ctx);
  if (r) return r;
  if (!normals || !nimages || !grids) return fail(ctx, PMVSB_EINVAL, "set_ref_image_batch: null pointer");
  if (ctx->wsize == 9 && std::min(stride, ctx->num) > 32) return fail(ctx, PMVSB_EINVAL, "set_ref_image_batch: wsize 9 keeps at most 32 images per patch in the selection kernels");
  if (P == 0) return PMVSB_OK;
  PatchStage st;
  r = stage_patches(ctx, st, P, stride, coords, normals, images, nimages, nullptr);
  if (r) return r;
  DevBuf<int32_t> dg;
  CK(dg.alloc((size_t)2 * stride * P));
  CK(cudaMemsetAsync(dg.p, 0xff, sizeof(int32_t) * (size_t)2 * stride * P, ctx->stream));
#define LAUNCH_SET_REF(W, MAXV, LO) \
  k_set_ref_image<W, MAXV, LO><<<P, 32, 0, ctx->stream>>>(ctx->scene, ctx->select, P, stride, st.coords.p, st.normals.p, st.images.p, st.nimages.p, dg.p); ++ctx->launches
  if (ctx->wsize == 5) {
    LAUNCH_SET_REF(5, 16, 0);
    if (stride > 16) { LAUNCH_SET_REF(5, 32, 16); }
    if (stride > 32) { LAUNCH_SET_REF(5, 64, 32); }
  } else if (ctx->wsize == 9) {
    LAUNCH_SET_REF(9, 16, 0);
    if (stride > 16) { LAUNCH_SET_REF(9, 32, 16); }
  } else {
    LAUNCH_SET_REF(7, 16, 0);
    if (stride > 16) { LAUNCH_SET_REF(7, 32, 16); }
    if (stride > 32) { LAUNCH_SET_REF(7, 64, 32); }
  }
#undef LAUNCH_SET_REF
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(images, st.images.p, sizeof(int32_t) * (size_t)stride * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(nimages, st.nimages.p, sizeof(int32_t) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(grids, dg.p, sizeof(int32_t) * (size_t)2 * stride * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

int pmvsb_patch_colors_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const int32_t* images, const int32_t* nimages,
                             uint8_t* rgb) {
  int r = check_ready(ctx);
  if (r) return r;
  if (!nimages || !rgb) return fail(ctx, PMVSB_EINVAL, "patch_colors_batch: null pointer");
  if (P == 0) return PMVSB_OK;
  PatchStage st;
  r = stage_patches(ctx, st, P, stride, coords, nullptr, images, nimages, nullptr);
  if (r) return r;
  DevBuf<uint8_t> dc;
  CK(dc.alloc((size_t)3 * P));
  k_patch_colors<<<(P + 127) / 128, 128, 0, ctx->stream>>>(ctx->scene, P, stride, st.coords.p, st.images.p, st.nimages.p, dc.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(rgb, dc.p, (size_t)3 * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

// addImages keeps at most min(stride, PMVSB_MAX_VIEWS) images per patch while the reference keeps all: a list that
// would have grown past the capacity is an ERROR, not a truncation (call after the stream has been synchronised)
static int check_view_overflow(pmvsb_ctx* ctx, const char* who, int stride) {
  int flag = 0;
  CK(cudaMemcpy(&flag, ctx->d_counter + 2, sizeof(int), cudaMemcpyDeviceToHost));
  if (!flag) return PMVSB_OK;
  CK(cudaMemset(ctx->d_counter + 2, 0, sizeof(int)));
  return fail(ctx, PMVSB_ERANGE, std::string(who) + ": a patch is visible in more images than the list capacity (" +
                                     std::to_string(std::min(stride, PMVSB_MAX_VIEWS)) + "); pass stride = number of images, at most " +
                                     std::to_string(PMVSB_MAX_VIEWS) + " (cluster larger scenes)");
}

int pmvsb_mask_gate_batch(pmvsb_ctx* ctx, int n, const float* coords, uint8_t* inside) {
  int r = check_ready(ctx);
  if (r) return r;
  if (n < 0 || (n > 0 && (!coords || !inside))) return fail(ctx, PMVSB_EINVAL, "mask_gate_batch: bad argument");
  if (n == 0) return PMVSB_OK;
  if (!ctx->scene.mask_lv && ctx->scene.n_bimages == 0) { std::memset(inside, 1, (size_t)n); return PMVSB_OK; }   // no map, no bounding image: nothing to launch
  DevBuf<float> dc;
  DevBuf<uint8_t> di;
  CK(dc.alloc((size_t)4 * n)); CK(di.alloc(n));
  CK(cudaMemcpyAsync(dc.p, coords, sizeof(float) * 4 * n, cudaMemcpyHostToDevice, ctx->stream));
  k_mask_gate<<<(n + 3) / 4, 128, 0, ctx->stream>>>(ctx->scene, n, dc.p, di.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(inside, di.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

int pmvsb_remove_images_edge_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, int32_t* images, int32_t* nimages) {
  int r = check_ready(ctx);
  if (r) return r;
  if (!nimages) return fail(ctx, PMVSB_EINVAL, "remove_images_edge_batch: null pointer");
  if (P == 0) return PMVSB_OK;
  PatchStage st;
  r = stage_patches(ctx, st, P, stride, coords, nullptr, images, nimages, nullptr);
  if (r) return r;
  if (!ctx->scene.edge_lv) return PMVSB_OK;   // getEdge == 1 everywhere: the lists stay as they are
  k_remove_images_edge<<<(P + 127) / 128, 128, 0, ctx->stream>>>(ctx->scene, P, stride, st.coords.p, st.images.p, st.nimages.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(images, st.images.p, sizeof(int32_t) * (size_t)stride * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(nimages, st.nimages.p, sizeof(int32_t) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

int pmvsb_pre_process_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals, int32_t* images,
                            int32_t* nimages, float* dscale, float* ascale, int32_t* verdict) {
  int r = check_ready(ctx);
  if (r) return r;
  if (!normals || !nimages || !dscale || !ascale || !verdict) return fail(ctx, PMVSB_EINVAL, "pre_process_batch: null pointer");
  if (ctx->wsize == 9 && std::min(stride, ctx->num) > 32) return fail(ctx, PMVSB_EINVAL, "pre_process_batch: wsize 9 keeps at most 32 images per patch in the selection kernels");
  if (P == 0) return PMVSB_OK;
  PatchStage st;
  r = stage_patches(ctx, st, P, stride, coords, normals, images, nimages, nullptr);
  if (r) return r;
  DevBuf<float> dd, da;
  DevBuf<int32_t> dv;
  CK(dd.alloc(P)); CK(da.alloc(P)); CK(dv.alloc(P));
#define LAUNCH_PRE(W, MAXV) k_pre_process<W, MAXV><<<P, 32, 0, ctx->stream>>>(ctx->scene, ctx->select, P, stride, st.coords.p, st.normals.p, st.images.p, st.nimages.p, dd.p, da.p, dv.p)
  DISPATCH_VIEWS(ctx, stride, LAUNCH_PRE);
#undef LAUNCH_PRE
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(images, st.images.p, sizeof(int32_t) * (size_t)stride * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(nimages, st.nimages.p, sizeof(int32_t) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(dscale, dd.p, sizeof(float) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ascale, da.p, sizeof(float) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(verdict, dv.p, sizeof(int32_t) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return check_view_overflow(ctx, "pre/post_process_batch", stride);
}

int pmvsb_post_process_batch(pmvsb_ctx* ctx, int P, int stride, const float* coords, const float* normals, const float* ncc,
                             int32_t* images, int32_t* nimages, int32_t* grids, int32_t* timages, float* tmp, int32_t* verdict) {
  int r = check_ready(ctx);
  if (r) return r;
  if (!normals || !ncc || !nimages || !grids || !timages || !tmp || !verdict) return fail(ctx, PMVSB_EINVAL, "post_process_batch: null pointer");
  if (ctx->wsize == 9 && std::min(stride, ctx->num) > 32) return fail(ctx, PMVSB_EINVAL, "post_process_batch: wsize 9 keeps at most 32 images per patch in the selection kernels");
  if (P == 0) return PMVSB_OK;
  PatchStage st;
  r = stage_patches(ctx, st, P, stride, coords, normals, images, nimages, nullptr);
  if (r) return r;
  DevBuf<float> dn, dt;
  DevBuf<int32_t> dg, dti, dv;
  CK(dn.alloc(P)); CK(dt.alloc(P)); CK(dg.alloc((size_t)2 * stride * P)); CK(dti.alloc(P)); CK(dv.alloc(P));
  CK(cudaMemcpyAsync(dn.p, ncc, sizeof(float) * P, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(dg.p, 0xff, sizeof(int32_t) * (size_t)2 * stride * P, ctx->stream));
#define LAUNCH_POST(W, MAXV) k_post_process<W, MAXV><<<P, 32, 0, ctx->stream>>>(ctx->scene, ctx->select, P, stride, st.coords.p, st.normals.p, dn.p, st.images.p, st.nimages.p, dg.p, dti.p, dt.p, dv.p)
  DISPATCH_VIEWS(ctx, stride, LAUNCH_POST);
#undef LAUNCH_POST
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(images, st.images.p, sizeof(int32_t) * (size_t)stride * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(nimages, st.nimages.p, sizeof(int32_t) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(grids, dg.p, sizeof(int32_t) * (size_t)2 * stride * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(timages, dti.p, sizeof(int32_t) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(tmp, dt.p, sizeof(float) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(verdict, dv.p, sizeof(int32_t) * P, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return check_view_overflow(ctx, "pre/post_process_batch", stride);
}

// gather != nullptr: the kernel also stores every patch's record into the mailboxes named there (pmvsb_refine_batch_dev_gather)
static int refine_dev_impl(pmvsb_ctx* ctx, int P, int stride, float* d_coords, float* d_normals, const int32_t* d_images,
                           const int32_t* d_nimages, const float* d_dscales, float* d_ncc, int32_t* d_evals, uint8_t* d_ok, const RecDst* gather) {
  int r = check_ready(ctx);
  if (r) return r;
  if (P < 0 || stride < 1 || !d_coords || !d_normals || !d_images || !d_dscales || !d_ncc || !d_evals || !d_ok)
    return fail(ctx, PMVSB_EINVAL, "refine_batch_dev: bad argument");
  if (P == 0) return PMVSB_OK;
  if (ctx->refine_blocks_per_sm == 0) {
    int nb = 0;
    switch (ctx->wsize) {
      case 5: CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_refine_g<5, true>, 128, 0)); break;
      case 9: CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_refine<9>, 128, 0)); break;
      default: CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_refine_g<7, true>, 128, 0)); break;
    }
    ctx->refine_blocks_per_sm = nb > 0 ? nb : 1;
  }
  CK(cudaMemsetAsync(ctx->d_counter, 0, 2 * sizeof(int), ctx->stream));
  CK(cudaEventRecord(ctx->ev0, ctx->stream));   // the ordering kernels are part of the timed call
  // processing order: patches sorted by (reference image, 8x8 tile), so that the patches in flight are neighbours
  const int32_t* d_order = nullptr;
  if (ctx->order_enabled && P >= 2048) {
    const int L = ctx->level;
    int mw = 1, mh = 1;
    for (int i = 0; i < ctx->num; ++i) { mw = std::max(mw, ctx->images[i].w[L]); mh = std::max(mh, ctx->images[i].h[L]); }
    const int tw = (mw + 7) / 8, th = (mh + 7) / 8;
    const long long bins = (long long)ctx->num * tw * th;
    if (bins < (1ll << 28)) {
      if ((r = dvec_reserve(ctx, ctx->order_bin, (size_t)P))) return r;
      if ((r = dvec_reserve(ctx, ctx->order_idx, (size_t)P))) return r;
      if ((r = dvec_reserve(ctx, ctx->order_counts, (size_t)bins + 1))) return r;
      CK(cudaMemsetAsync(ctx->order_counts.p, 0, sizeof(int32_t) * ((size_t)bins + 1), ctx->stream));
      k_order_count<<<(P + 255) / 256, 256, 0, ctx->stream>>>(ctx->scene, P, stride, d_coords, d_images, tw, th, ctx->order_bin.p, ctx->order_counts.p);
      ++ctx->launches;
      if ((r = device_scan(ctx, ctx->order_counts.p, (int)bins + 1))) return r;
      k_order_fill<<<(P + 255) / 256, 256, 0, ctx->stream>>>(P, ctx->order_bin.p, ctx->order_counts.p, ctx->order_idx.p);
      ++ctx->launches;
      d_order = ctx->order_idx.p;
    }
  }
  // persistent grid: a whole number of resident CTAs per SM (148 SMs on B200)
  int grid = ctx->sm_count * ctx->refine_blocks_per_sm;
  const int per_block = ctx->wsize == 9 ? 4 : 16;  // patches a CTA works on at a time
  const int needed = (P + per_block - 1) / per_block;
  if (grid > needed) grid = needed;
  if (gather) {
    if (ctx->scene.atlas == 0 || (ctx->wsize != 7 && ctx->wsize != 5))
      return fail(ctx, PMVSB_EINVAL, "refine_batch_dev_gather: needs the texture-atlas path and wsize 5 or 7");
    if (ctx->wsize == 5)
      k_refine_g<5, true, true><<<grid, 128, 0, ctx->stream>>>(ctx->scene, P, stride, d_coords, d_normals, d_images, d_nimages, d_dscales, d_ncc, d_evals, d_ok,
                                                                ctx->d_counter, d_order, gather);
    else
      k_refine_g<7, true, true><<<grid, 128, 0, ctx->stream>>>(ctx->scene, P, stride, d_coords, d_normals, d_images, d_nimages, d_dscales, d_ncc, d_evals, d_ok,
                                                                ctx->d_counter, d_order, gather);
    ++ctx->launches;
  } else
  DISPATCH_GROUP(ctx, k_refine_g, k_refine, grid, grid, 128, ctx->scene, P, stride, d_coords, d_normals, d_images, d_nimages, d_dscales,
                 d_ncc, d_evals, d_ok, ctx->d_counter, d_order);
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->refine_timed = true;
  CK(cudaGetLastError());
  return PMVSB_OK;
}

int pmvsb_refine_batch_dev(pmvsb_ctx* ctx, int P, int stride, float* d_coords, float* d_normals, const int32_t* d_images,
                           const int32_t* d_nimages, const float* d_dscales, float* d_ncc, int32_t* d_evals, uint8_t* d_ok) {
  return refine_dev_impl(ctx, P, stride, d_coords, d_normals, d_images, d_nimages, d_dscales, d_ncc, d_evals, d_ok, nullptr);
}


// ---- the in-process contract of the hot path, batched ------------------------------------------------------------
//   if (preProcess(patch, id, seed)) fail;  refinePatch(patch, id, 100);  if (postProcess(patch, id, seed)) fail;
// (source/pmvs/seed.cpp:397-409, expand.cpp:225-237) for a whole wave of candidates without leaving the device between the
// stages: pre -> compaction -> refine -> post (incl. setVImagesVGrids at _depth >= 1 and check at _depth >= 2) -> compaction.
__global__ void k_pack_scal(int n, const float* __restrict__ ncc, const float* __restrict__ dscale, const float* __restrict__ ascale, const float* __restrict__ tmp,
                            float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float4(ncc[i], dscale[i], ascale[i], tmp[i]);
}

int pmvsb_evaluate_batch(pmvsb_ctx* ctx, int P, const float* coords, const float* normals, const int32_t* img_off, const int32_t* images, float quad,
                         int32_t* accepted, int32_t* entries, int32_t* ventries, int32_t* refined) {
  int r = check_ready(ctx);
  if (r) return r;
  if (P < 0 || !accepted || !entries || !ventries || (P > 0 && (!coords || !normals || !img_off || !images)))
    return fail(ctx, PMVSB_EINVAL, "evaluate_batch: bad argument");
  if (ctx->depth_flag >= 1 && (r = need_store(ctx, true))) return r;
  pmvsb_ctx::EvalOut& ev = ctx->ev;
  ev.P = P; ev.A = ev.E = ev.VE = ev.refined = 0;
  *accepted = *entries = *ventries = 0;
  if (refined) *refined = 0;
  if (P == 0) return PMVSB_OK;
  const int Ein = img_off[P];
  for (int e = 0; e < Ein; ++e)
    if ((unsigned)images[e] >= (unsigned)ctx->num) return fail(ctx, PMVSB_EINVAL, "evaluate_batch: image index out of range");
  const int stride = std::min(ctx->num, PMVSB_MAX_VIEWS), vs = std::max(ctx->tnum, 1);
  if (ctx->wsize == 9 && stride > 32) return fail(ctx, PMVSB_EINVAL, "evaluate_batch: wsize 9 keeps at most 32 images per patch in the selection kernels");
  cudaStream_t q = ctx->stream;
  // ---- stage 0: candidates to the device, lists into stride-padded rows
  DevBuf<float> c0, n0, ds0, as0;
  DevBuf<int32_t> off0, im0, rows0, cnt0, v0, pos0;
  CK(c0.alloc((size_t)4 * P)); CK(n0.alloc((size_t)4 * P)); CK(ds0.alloc(P)); CK(as0.alloc(P)); CK(off0.alloc((size_t)P + 1)); CK(im0.alloc((size_t)std::max(Ein, 1)));
  CK(rows0.alloc((size_t)stride * P)); CK(cnt0.alloc(P)); CK(v0.alloc(P)); CK(pos0.alloc((size_t)P + 1));
  CK(cudaMemcpyAsync(c0.p, coords, sizeof(float) * 4 * (size_t)P, cudaMemcpyHostToDevice, q));
  CK(cudaMemcpyAsync(n0.p, normals, sizeof(float) * 4 * (size_t)P, cudaMemcpyHostToDevice, q));
  CK(cudaMemcpyAsync(off0.p, img_off, sizeof(int32_t) * ((size_t)P + 1), cudaMemcpyHostToDevice, q));
  if (Ein > 0) CK(cudaMemcpyAsync(im0.p, images, sizeof(int32_t) * (size_t)Ein, cudaMemcpyHostToDevice, q));
  k_rows_from_csr<<<(P + 255) / 256, 256, 0, q>>>(P, stride, off0.p, im0.p, rows0.p, cnt0.p);
  // ---- preProcess
#define LAUNCH_PRE(W, MAXV) k_pre_process<W, MAXV><<<P, 32, 0, q>>>(ctx->scene, ctx->select, P, stride, c0.p, n0.p, rows0.p, cnt0.p, ds0.p, as0.p, v0.p)
  DISPATCH_VIEWS(ctx, stride, LAUNCH_PRE);
#undef LAUNCH_PRE
  k_flag_zero<<<(P + 256) / 256, 256, 0, q>>>(P, v0.p, pos0.p);
  ctx->launches += 3;
  if ((r = device_scan(ctx, pos0.p, P + 1))) return r;
  int32_t L = 0;
  CK(cudaMemcpyAsync(&L, pos0.p + P, sizeof(int32_t), cudaMemcpyDeviceToHost, q));
  CK(cudaStreamSynchronize(q));
  if ((r = dvec_reserve(ctx, ev.verdict, (size_t)P))) return r;
  ev.refined = L;
  if (refined) *refined = L;
  if (L == 0) {
    CK(cudaMemcpyAsync(ev.verdict.p, v0.p, sizeof(int32_t) * (size_t)P, cudaMemcpyDeviceToDevice, q));   // all 1
    CK(cudaStreamSynchronize(q));
    return check_view_overflow(ctx, "evaluate_batch", stride);
  }
  // ---- survivors to the front; refinePatch; postProcess
  DevBuf<float> c1, n1, ds1, as1, ncc1, tmp1;
  DevBuf<int32_t> idx1, rows1, cnt1, ev1, gr1, ti1, v1, pos1;
  DevBuf<uint8_t> ok1;
  CK(c1.alloc((size_t)4 * L)); CK(n1.alloc((size_t)4 * L)); CK(ds1.alloc(L)); CK(as1.alloc(L)); CK(ncc1.alloc(L)); CK(tmp1.alloc(L)); CK(idx1.alloc(L));
  CK(rows1.alloc((size_t)stride * L)); CK(cnt1.alloc(L)); CK(ev1.alloc(L)); CK(gr1.alloc((size_t)2 * stride * L)); CK(ti1.alloc(L)); CK(v1.alloc(L));
  CK(pos1.alloc((size_t)L + 1)); CK(ok1.alloc(L));
  k_compact_patches<<<(int)(((size_t)P * 32 + 255) / 256), 256, 0, q>>>(P, stride, v0.p, pos0.p, nullptr, reinterpret_cast<const float4*>(c0.p),
      reinterpret_cast<const float4*>(n0.p), rows0.p, cnt0.p, ds0.p, as0.p, idx1.p, reinterpret_cast<float4*>(c1.p), reinterpret_cast<float4*>(n1.p), rows1.p,
      cnt1.p, ds1.p, as1.p);
  ++ctx->launches;
  if ((r = pmvsb_refine_batch_dev(ctx, L, stride, c1.p, n1.p, rows1.p, cnt1.p, ds1.p, ncc1.p, ev1.p, ok1.p))) return r;
  CK(cudaMemsetAsync(gr1.p, 0xff, sizeof(int32_t) * (size_t)2 * stride * L, q));
#define LAUNCH_POST(W, MAXV) k_post_process<W, MAXV><<<L, 32, 0, q>>>(ctx->scene, ctx->select, L, stride, c1.p, n1.p, ncc1.p, rows1.p, cnt1.p, gr1.p, ti1.p, tmp1.p, v1.p)
  DISPATCH_VIEWS(ctx, stride, LAUNCH_POST);
#undef LAUNCH_POST
  k_scatter_verdict<<<(L + 255) / 256, 256, 0, q>>>(L, idx1.p, v1.p, nullptr, 2, v0.p);
  k_flag_zero<<<(L + 256) / 256, 256, 0, q>>>(L, v1.p, pos1.p);
  ctx->launches += 3;
  if ((r = device_scan(ctx, pos1.p, L + 1))) return r;
  int32_t A = 0;
  CK(cudaMemcpyAsync(&A, pos1.p + L, sizeof(int32_t), cudaMemcpyDeviceToHost, q));
  CK(cudaStreamSynchronize(q));
  // ---- accepted so far to the front
  const size_t A1 = (size_t)std::max(A, 1);
  DevBuf<float> c2, n2, ds2, as2, ncc2, tmp2, gain2;
  DevBuf<int32_t> idx2, rows2, cnt2, gr2, ti2, vim2, nv2, vgr2, v2, pos2, dov;
  DevBuf<uint8_t> rej2;
  CK(c2.alloc(4 * A1)); CK(n2.alloc(4 * A1)); CK(ds2.alloc(A1)); CK(as2.alloc(A1)); CK(ncc2.alloc(A1)); CK(tmp2.alloc(A1)); CK(gain2.alloc(A1)); CK(idx2.alloc(A1));
  CK(rows2.alloc((size_t)stride * A1)); CK(cnt2.alloc(A1)); CK(gr2.alloc((size_t)2 * stride * A1)); CK(ti2.alloc(A1)); CK(vim2.alloc((size_t)vs * A1));
  CK(nv2.alloc(A1)); CK(vgr2.alloc((size_t)2 * vs * A1)); CK(v2.alloc(A1)); CK(pos2.alloc(A1 + 1)); CK(dov.alloc(1)); CK(rej2.alloc(A1));
  int32_t A2 = A;
  if (A > 0) {
    const int gl = (int)(((size_t)L * 32 + 255) / 256);
    k_compact_patches<<<gl, 256, 0, q>>>(L, stride, v1.p, pos1.p, idx1.p, reinterpret_cast<const float4*>(c1.p), reinterpret_cast<const float4*>(n1.p), rows1.p,
                                         cnt1.p, ds1.p, as1.p, idx2.p, reinterpret_cast<float4*>(c2.p), reinterpret_cast<float4*>(n2.p), rows2.p, cnt2.p, ds2.p, as2.p);
    k_compact_extras<<<gl, 256, 0, q>>>(L, stride, v1.p, pos1.p, cnt1.p, gr1.p, ncc1.p, ok1.p, ti1.p, tmp1.p, gr2.p, ncc2.p, ti2.p, tmp2.p);
    ctx->launches += 2;
    CK(cudaMemsetAsync(nv2.p, 0, sizeof(int32_t) * A1, q));
    if (ctx->depth_flag >= 1) {   // setVImagesVGrids against the current depth maps (optim.cpp:184-186)
      k_set_vimages_batch<<<(A + 3) / 4, 128, 0, q>>>(ctx->scene, ctx->store, A, stride, c2.p, n2.p, rows2.p, cnt2.p, vs, vim2.p, nv2.p, vgr2.p);
      ++ctx->launches;
    }
    if (ctx->depth_flag >= 2) {   // COptim::check (optim.cpp:363-383): gain against the table's cells, then the quadric fit
      CK(cudaMemsetAsync(dov.p, 0, sizeof(int32_t), q));
      k_check_batch<<<(A + kNbWarps - 1) / kNbWarps, kNbWarps * 32, 0, q>>>(ctx->scene, ctx->store, A, stride, vs, c2.p, n2.p, ncc2.p, ds2.p, ti2.p, rows2.p, cnt2.p,
                                                                           gr2.p, vim2.p, nv2.p, vgr2.p, quad, ctx->tau, gain2.p, rej2.p, dov.p);
      k_merge_reject<<<(A + 255) / 256, 256, 0, q>>>(A, rej2.p, gain2.p, v2.p, tmp2.p);
      k_scatter_verdict<<<(A + 255) / 256, 256, 0, q>>>(A, idx2.p, v2.p, nullptr, 2, v0.p);
      k_flag_zero<<<(A + 256) / 256, 256, 0, q>>>(A, v2.p, pos2.p);
      ctx->launches += 4;
      if ((r = device_scan(ctx, pos2.p, A + 1))) return r;
      CK(cudaMemcpyAsync(&A2, pos2.p + A, sizeof(int32_t), cudaMemcpyDeviceToHost, q));
      CK(cudaStreamSynchronize(q));
    }
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ev.verdict.p, v0.p, sizeof(int32_t) * (size_t)P, cudaMemcpyDeviceToDevice, q));
  // ---- the accepted candidates' records, lists as CSR
  const size_t F1 = (size_t)std::max(A2, 1);
  if ((r = dvec_reserve(ctx, ev.index, F1)) || (r = dvec_reserve(ctx, ev.timages, F1)) || (r = dvec_reserve(ctx, ev.coords, 4 * F1)) ||
      (r = dvec_reserve(ctx, ev.normals, 4 * F1)) || (r = dvec_reserve(ctx, ev.scal, 4 * F1)) || (r = dvec_reserve(ctx, ev.img_off, F1 + 1)) ||
      (r = dvec_reserve(ctx, ev.vimg_off, F1 + 1)))
    return r;
  int32_t E = 0, VE = 0;
  if (A2 > 0) {
    const float *fc = c2.p, *fn = n2.p, *fncc = ncc2.p, *fds = ds2.p, *fas = as2.p, *ftmp = tmp2.p;
    const int32_t *fidx = idx2.p, *frows = rows2.p, *fcnt = cnt2.p, *fgr = gr2.p, *fti = ti2.p, *fvim = vim2.p, *fnv = nv2.p, *fvgr = vgr2.p;
    DevBuf<float> c3, n3, ds3, as3, ncc3, tmp3;
    DevBuf<int32_t> idx3, rows3, cnt3, gr3, ti3, vim3, nv3, vgr3;
    if (A2 != A) {   // check rejected some: one more compaction (image rows, cells, visible-image rows)
      CK(c3.alloc(4 * F1)); CK(n3.alloc(4 * F1)); CK(ds3.alloc(F1)); CK(as3.alloc(F1)); CK(ncc3.alloc(F1)); CK(tmp3.alloc(F1)); CK(idx3.alloc(F1));
      CK(rows3.alloc((size_t)stride * F1)); CK(cnt3.alloc(F1)); CK(gr3.alloc((size_t)2 * stride * F1)); CK(ti3.alloc(F1)); CK(vim3.alloc((size_t)vs * F1));
      CK(nv3.alloc(F1)); CK(vgr3.alloc((size_t)2 * vs * F1));
      DevBuf<uint8_t> ones;
      CK(ones.alloc(A1));
      CK(cudaMemsetAsync(ones.p, 1, A1, q));
      const int gl = (int)(((size_t)A * 32 + 255) / 256);
      k_compact_patches<<<gl, 256, 0, q>>>(A, stride, v2.p, pos2.p, idx2.p, reinterpret_cast<const float4*>(c2.p), reinterpret_cast<const float4*>(n2.p), rows2.p,
                                           cnt2.p, ds2.p, as2.p, idx3.p, reinterpret_cast<float4*>(c3.p), reinterpret_cast<float4*>(n3.p), rows3.p, cnt3.p, ds3.p, as3.p);
      k_compact_extras<<<gl, 256, 0, q>>>(A, stride, v2.p, pos2.p, cnt2.p, gr2.p, ncc2.p, ones.p, ti2.p, tmp2.p, gr3.p, ncc3.p, ti3.p, tmp3.p);
      // visible-image rows: the same kernels with vs as the stride (rows = vimages, cells = vgrids)
      DevBuf<float> dumpf;
      DevBuf<int32_t> dumpi;
      CK(dumpf.alloc(4 * F1)); CK(dumpi.alloc(F1));
      k_compact_patches<<<gl, 256, 0, q>>>(A, vs, v2.p, pos2.p, idx2.p, reinterpret_cast<const float4*>(c2.p), reinterpret_cast<const float4*>(n2.p), vim2.p,
                                           nv2.p, nullptr, nullptr, dumpi.p, reinterpret_cast<float4*>(dumpf.p), reinterpret_cast<float4*>(dumpf.p), vim3.p, nv3.p,
                                           nullptr, nullptr);
      k_compact_extras<<<gl, 256, 0, q>>>(A, vs, v2.p, pos2.p, nv2.p, vgr2.p, ncc2.p, ones.p, ti2.p, tmp2.p, vgr3.p, dumpf.p, dumpi.p, dumpf.p);
      ctx->launches += 4;
      CK(cudaStreamSynchronize(q));   // dumpf / dumpi / ones go back to the pool here
      fc = c3.p; fn = n3.p; fncc = ncc3.p; fds = ds3.p; fas = as3.p; ftmp = tmp3.p;
      fidx = idx3.p; frows = rows3.p; fcnt = cnt3.p; fgr = gr3.p; fti = ti3.p; fvim = vim3.p; fnv = nv3.p; fvgr = vgr3.p;
    }
    k_copy_len<<<(A2 + 256) / 256, 256, 0, q>>>(A2, fcnt, ev.img_off.p);
    k_copy_len<<<(A2 + 256) / 256, 256, 0, q>>>(A2, fnv, ev.vimg_off.p);
    ctx->launches += 2;
    if ((r = device_scan(ctx, ev.img_off.p, A2 + 1))) return r;
    if ((r = device_scan(ctx, ev.vimg_off.p, A2 + 1))) return r;
    CK(cudaMemcpyAsync(&E, ev.img_off.p + A2, sizeof(int32_t), cudaMemcpyDeviceToHost, q));
    CK(cudaMemcpyAsync(&VE, ev.vimg_off.p + A2, sizeof(int32_t), cudaMemcpyDeviceToHost, q));
    CK(cudaStreamSynchronize(q));
    if ((r = dvec_reserve(ctx, ev.images, (size_t)std::max(E, 1))) || (r = dvec_reserve(ctx, ev.grids, (size_t)2 * std::max(E, 1))) ||
        (r = dvec_reserve(ctx, ev.vimages, (size_t)std::max(VE, 1))) || (r = dvec_reserve(ctx, ev.vgrids, (size_t)2 * std::max(VE, 1))))
      return r;
    k_tab_rows_to_lists<<<(A2 + 127) / 128, 128, 0, q>>>(A2, stride, frows, fgr, ev.img_off.p, ev.images.p, ev.grids.p);
    if (VE > 0) k_tab_rows_to_lists<<<(A2 + 127) / 128, 128, 0, q>>>(A2, vs, fvim, fvgr, ev.vimg_off.p, ev.vimages.p, ev.vgrids.p);
    k_pack_scal<<<(A2 + 255) / 256, 256, 0, q>>>(A2, fncc, fds, fas, ftmp, reinterpret_cast<float4*>(ev.scal.p));
    ctx->launches += 3;
    CK(cudaMemcpyAsync(ev.coords.p, fc, sizeof(float) * 4 * (size_t)A2, cudaMemcpyDeviceToDevice, q));
    CK(cudaMemcpyAsync(ev.normals.p, fn, sizeof(float) * 4 * (size_t)A2, cudaMemcpyDeviceToDevice, q));
    CK(cudaMemcpyAsync(ev.index.p, fidx, sizeof(int32_t) * (size_t)A2, cudaMemcpyDeviceToDevice, q));
    CK(cudaMemcpyAsync(ev.timages.p, fti, sizeof(int32_t) * (size_t)A2, cudaMemcpyDeviceToDevice, q));
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(q));
  ev.A = A2; ev.E = E; ev.VE = VE;
  *accepted = A2; *entries = E; *ventries = VE;
  return check_view_overflow(ctx, "evaluate_batch", stride);
}

int pmvsb_evaluate_fetch(pmvsb_ctx* ctx, int32_t* verdict, int32_t* index, float* coords, float* normals, float* scal, int32_t* timages, int32_t* img_off,
                         int32_t* images, int32_t* grids, int32_t* vimg_off, int32_t* vimages, int32_t* vgrids) {
  int r = check_ready(ctx);
  if (r) return r;
  const pmvsb_ctx::EvalOut& ev = ctx->ev;
  cudaStream_t q = ctx->stream;
  auto down = [&](void* h, const void* d, size_t bytes) { return (h && bytes) ? cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, q) : cudaSuccess; };
  CK(down(verdict, ev.verdict.p, sizeof(int32_t) * (size_t)ev.P));
  if (ev.A > 0) {
    CK(down(index, ev.index.p, sizeof(int32_t) * (size_t)ev.A)); CK(down(coords, ev.coords.p, sizeof(float) * 4 * (size_t)ev.A));
    CK(down(normals, ev.normals.p, sizeof(float) * 4 * (size_t)ev.A)); CK(down(scal, ev.scal.p, sizeof(float) * 4 * (size_t)ev.A));
    CK(down(timages, ev.timages.p, sizeof(int32_t) * (size_t)ev.A)); CK(down(img_off, ev.img_off.p, sizeof(int32_t) * ((size_t)ev.A + 1)));
    CK(down(images, ev.images.p, sizeof(int32_t) * (size_t)ev.E)); CK(down(grids, ev.grids.p, sizeof(int32_t) * 2 * (size_t)ev.E));
    CK(down(vimg_off, ev.vimg_off.p, sizeof(int32_t) * ((size_t)ev.A + 1))); CK(down(vimages, ev.vimages.p, sizeof(int32_t) * (size_t)ev.VE));
    CK(down(vgrids, ev.vgrids.p, sizeof(int32_t) * 2 * (size_t)ev.VE));
  } else {
    if (img_off) img_off[0] = 0;
    if (vimg_off) vimg_off[0] = 0;
  }
  CK(cudaStreamSynchronize(q));
  return PMVSB_OK;
}

int pmvsb_refine_batch(pmvsb_ctx* ctx, int P, int stride, float* coords, float* normals, const int32_t* images, const int32_t* nimages,
                       const float* dscales, float* ncc, int32_t* evals, uint8_t* ok) {
  int r = check_ready(ctx);
  if (r) return r;
  if (!normals || !dscales || !ncc || !evals || !ok) return fail(ctx, PMVSB_EINVAL, "refine_batch: null pointer");
  if (P == 0) return PMVSB_OK;
  if (stride < 1 || !coords || !images) return fail(ctx, PMVSB_EINVAL, "refine_batch: bad patch batch");
  if (ctx->wsize == 9) {  // the warp-per-patch fallback trusts its indexes: validate here
    if ((r = check_image_rows(ctx, P, stride, images, nimages))) return r;
  }
  const size_t nP = (size_t)P;
  r = arena_reserve(ctx, nP * (16 + 16 + 4 * (size_t)stride + 4 + 4 + 4 + 4 + 1) + 16 * 256);
  if (r) return r;
  float* d_coords = arena_take<float>(ctx, 4 * nP);
  float* d_normals = arena_take<float>(ctx, 4 * nP);
  int32_t* d_images = arena_take<int32_t>(ctx, (size_t)stride * nP);
  int32_t* d_nimages = nimages ? arena_take<int32_t>(ctx, nP) : nullptr;
  float* d_dscales = arena_take<float>(ctx, nP);
  float* d_ncc = arena_take<float>(ctx, nP);
  int32_t* d_evals = arena_take<int32_t>(ctx, nP);
  uint8_t* d_ok = arena_take<uint8_t>(ctx, nP);
  CK(cudaMemcpyAsync(d_coords, coords, sizeof(float) * 4 * nP, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_normals, normals, sizeof(float) * 4 * nP, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_images, images, sizeof(int32_t) * (size_t)stride * nP, cudaMemcpyHostToDevice, ctx->stream));
  if (nimages) CK(cudaMemcpyAsync(d_nimages, nimages, sizeof(int32_t) * nP, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_dscales, dscales, sizeof(float) * nP, cudaMemcpyHostToDevice, ctx->stream));
  r = pmvsb_refine_batch_dev(ctx, P, stride, d_coords, d_normals, d_images, d_nimages, d_dscales, d_ncc, d_evals, d_ok);
  if (r) return r;
  int flags[2] = {0, 0};
  CK(cudaMemcpyAsync(flags, ctx->d_counter, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(coords, d_coords, sizeof(float) * 4 * nP, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(normals, d_normals, sizeof(float) * 4 * nP, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ncc, d_ncc, sizeof(float) * nP, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(evals, d_evals, sizeof(int32_t) * nP, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ok, d_ok, sizeof(uint8_t) * nP, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (flags[1]) return fail(ctx, PMVSB_EINVAL, "image index out of range in patch batch (those patches were skipped, ok = 0)");
  return PMVSB_OK;
}

// ---- feature detection -------------------------------------------------------------------------------------------
namespace {
// CDetector::setGaussI (source/pmvs/detector.cpp:31-49); exp is the double libm entry point in the reference's object
std::vector<float> gauss_taps(float sigma) {
  const int margin = (int)std::ceil(2 * sigma);
  std::vector<float> g(2 * margin + 1);
  float denom = 0.0f;
  for (int x = 0; x < (int)g.size(); ++x) {
    const int xt = x - margin;
    const float d = (float)std::exp((double)(-(xt * xt) / (2 * sigma * sigma)));
    g[x] = d;
    denom += d;
  }
  for (float& v : g) v /= denom;
  return g;
}
struct FeatPoint { float r; int x, y; long seq; };
}  // namespace

int pmvsb_detect_features(pmvsb_ctx* ctx, int index, int gspeedup, int cap, float* xy, float* response, int32_t* type, int32_t* count) {
  int r = check_ready(ctx);
  if (r) return r;
  if (index < 0 || index >= ctx->num || gspeedup < 1 || cap < 0 || !count || (cap > 0 && (!xy || !response || !type)))
    return fail(ctx, PMVSB_EINVAL, "detect_features: bad argument");
  const HostImage& hi = ctx->images[index];
  const int w = hi.w[ctx->level], h = hi.h[ctx->level];
  const int n = w * h;
  LevelDev lv;
  lv.pix = hi.levels[ctx->level]; lv.w = w; lv.h = h;
  const int factor = 2, gridsize = gspeedup * factor;
  const int gw = (w + gridsize - 1) / gridsize, gh = (h + gridsize - 1) / gridsize, nb = gw * gh;
  // filters: derivative, box, Gaussians (Harris sigma 4; DoG 1 * sqrt(2)^k)
  const float first = 1.0f, last = 3.0f;
  const float scalestep = (float)std::pow((double)2.0f, (double)(1 / 2.0f));
  int steps = (int)std::ceil(std::log((double)(last / first)) / std::log((double)scalestep));
  steps = std::max(4, steps);
  const int nres = steps + 1;
  std::vector<std::vector<float>> filters;
  filters.push_back({-0.5f, 0.0f, 0.5f});
  filters.push_back({(float)(1.0 / 3.0), (float)(1.0 / 3.0), (float)(1.0 / 3.0)});
  filters.push_back(gauss_taps(4.0f));
  std::vector<float> sigmas(nres);
  for (int k = 0; k < nres; ++k) {
    if (k == 0) sigmas[k] = first; else if (k == 1) sigmas[k] = first * scalestep; else if (k == 2) sigmas[k] = first * scalestep * scalestep;
    else sigmas[k] = (float)((double)first * std::pow((double)scalestep, (double)k));   // _firstScale * pow(scalestep, i + 1)
    filters.push_back(gauss_taps(sigmas[k]));
  }
  std::vector<float> flat;
  std::vector<int> foff;
  for (const auto& f : filters) { foff.push_back((int)flat.size()); flat.insert(flat.end(), f.begin(), f.end()); }
  DevBuf<float> d_taps, img, ta, tb, tc, resb, dogb, outr;
  DevBuf<int32_t> outxy, outn;
  DevBuf<unsigned char> seen;
  CK(d_taps.alloc(flat.size())); CK(img.alloc((size_t)3 * n)); CK(ta.alloc((size_t)3 * n)); CK(tb.alloc((size_t)3 * n)); CK(tc.alloc((size_t)3 * n));
  CK(resb.alloc((size_t)nres * n)); CK(dogb.alloc((size_t)(nres - 1) * n)); CK(outr.alloc((size_t)nb * 4)); CK(outxy.alloc((size_t)nb * 8));
  CK(outn.alloc(nb)); CK(seen.alloc(n));
  CK(cudaMemcpyAsync(d_taps.p, flat.data(), sizeof(float) * flat.size(), cudaMemcpyHostToDevice, ctx->stream));
  const dim3 cb(128), cg3((w + 127) / 128, h, 3), cg1((w + 127) / 128, h, 1);
  const int tb1 = 256, gb1 = (n + 255) / 256;
  // the detectors' _mask: the image's mask and / or edge map at the working level (detectFeatures.cpp:88-92)
  DevBuf<unsigned char> maskbuf;
  const unsigned char* fmask = nullptr;
  if (hi.maps[0] || hi.maps[1]) {
    CK(maskbuf.alloc(n));
    k_feat_mask<<<gb1, tb1, 0, ctx->stream>>>(hi.maps[0], hi.maps[1], n, maskbuf.p);
    ++ctx->launches;
    fmask = maskbuf.p;
  }
  auto conv = [&](bool vertical, const float* src, float* dst, int filt, dim3 grid) {
    const int nt = (int)filters[filt].size();
    if (vertical) k_feat_conv<true><<<grid, cb, 0, ctx->stream>>>(src, dst, w, h, d_taps.p + foff[filt], nt, fmask);
    else k_feat_conv<false><<<grid, cb, 0, ctx->stream>>>(src, dst, w, h, d_taps.p + foff[filt], nt, fmask);
    ++ctx->launches;
  };
  std::vector<FeatPoint> feats;
  std::vector<float> h_r((size_t)nb * 4);
  std::vector<int32_t> h_xy((size_t)nb * 8), h_n(nb);
  int total = 0;
  auto collect = [&](int ftype) -> int {   // reverse iteration of the result multiset: strongest first, later insertion first among equals
    if (cudaMemcpyAsync(h_r.data(), outr.p, sizeof(float) * h_r.size(), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
    if (cudaMemcpyAsync(h_xy.data(), outxy.p, sizeof(int32_t) * h_xy.size(), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
    if (cudaMemcpyAsync(h_n.data(), outn.p, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
    feats.clear();
    long seq = 0;
    for (int b = 0; b < nb; ++b)
      for (int k = 0; k < h_n[b]; ++k) feats.push_back({h_r[(size_t)b * 4 + k], h_xy[((size_t)b * 4 + k) * 2], h_xy[((size_t)b * 4 + k) * 2 + 1], seq++});
    std::sort(feats.begin(), feats.end(), [](const FeatPoint& a, const FeatPoint& b) { return a.r != b.r ? a.r > b.r : a.seq > b.seq; });
    for (const FeatPoint& f : feats) {
      if (total < cap) { xy[2 * total] = (float)f.x; xy[2 * total + 1] = (float)f.y; response[total] = f.r; type[total] = ftype; }
      ++total;
    }
    return 0;
  };
  k_feat_planes<<<gb1, tb1, 0, ctx->stream>>>(lv, img.p);
  ++ctx->launches;
  // ---- Harris (harris.cpp:112-137, 60-110, 139-172)
  conv(false, img.p, ta.p, 0, cg3); conv(true, ta.p, tb.p, 1, cg3);    // tb = dI/dx
  conv(false, img.p, ta.p, 1, cg3); conv(true, ta.p, tc.p, 0, cg3);    // tc = dI/dy
  k_feat_products<<<gb1, tb1, 0, ctx->stream>>>(tb.p, tc.p, n, ta.p, fmask);   // ta = (xx, yy, xy)
  ++ctx->launches;
  conv(false, ta.p, tb.p, 2, cg3); conv(true, tb.p, ta.p, 2, cg3);
  k_feat_harris_response<<<gb1, tb1, 0, ctx->stream>>>(ta.p, n, tb.p, fmask);
  k_feat_nms<<<cg1, cb, 0, ctx->stream>>>(tb.p, w, h, tc.p);
  {
    const int margin = (2 * (int)std::ceil(2 * 4.0f) + 1) / 2;   // _gaussD.size() / 2
    k_feat_select<0><<<(nb + 3) / 4, 128, 0, ctx->stream>>>(tc.p, nullptr, nullptr, nullptr, w, h, gridsize, gw, gh, margin, margin, nullptr,
                                                            outr.p, outxy.p, outn.p);
  }
  ctx->launches += 3;
  CK(cudaGetLastError());
  if (collect(0)) return fail(ctx, PMVSB_ECUDA, "detect_features: copy failed");
  // ---- difference of Gaussians (dog.cpp:122-183)
  for (int k = 0; k < nres; ++k) {
    conv(false, img.p, ta.p, 3 + k, cg3); conv(true, ta.p, tb.p, 3 + k, cg3);
    k_feat_norm<<<gb1, tb1, 0, ctx->stream>>>(tb.p, n, resb.p + (size_t)k * n);
    ++ctx->launches;
  }
  for (int k = 0; k + 1 < nres; ++k) {
    k_feat_sub<<<gb1, tb1, 0, ctx->stream>>>(resb.p + (size_t)(k + 1) * n, resb.p + (size_t)k * n, n, dogb.p + (size_t)k * n);
    ++ctx->launches;
  }
  CK(cudaMemsetAsync(seen.p, 0, n, ctx->stream));
  CK(cudaMemsetAsync(outn.p, 0, sizeof(int32_t) * nb, ctx->stream));
  // the reference walks i = 2 .. steps-1 with (pdog, cdog, ndog) = dog[i-2 .. i]; the select kernel takes two scales per
  // launch and keeps the block's multiset in registers, so steps == 4 (scales 1 .. 3, the only values pmvs uses) is one launch
  if (steps != 4) return fail(ctx, PMVSB_EINVAL, "detect_features: unsupported scale range");
  {
    const int ma = (int)std::ceil(2 * (float)((double)first * std::pow((double)scalestep, 3.0)));
    const int mb = (int)std::ceil(2 * (float)((double)first * std::pow((double)scalestep, 4.0)));
    k_feat_select<1><<<(nb + 3) / 4, 128, 0, ctx->stream>>>(dogb.p, dogb.p + (size_t)n, dogb.p + (size_t)2 * n, dogb.p + (size_t)3 * n, w, h, gridsize,
                                                            gw, gh, ma, mb, seen.p, outr.p, outxy.p, outn.p);
    ++ctx->launches;
  }
  CK(cudaGetLastError());
  if (collect(1)) return fail(ctx, PMVSB_ECUDA, "detect_features: copy failed");
  *count = total;
  return PMVSB_OK;
}

// ---- multi-GPU: one process (and one context) per GPU; the only exchange of the path is the all-gather of a wave's results
int pmvsb_comm_unique_id(pmvsb_ctx* ctx, uint8_t* id128) {
  if (!ctx || !id128) return fail(ctx, PMVSB_EINVAL, "comm_unique_id: null pointer");
  NcclApi* api = nccl_api();
  if (!api) return fail(ctx, PMVSB_ESTATE, "comm_unique_id: libnccl.so.2 could not be loaded");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  const ncclResult_t r = api->GetUniqueId(&id);
  if (r != ncclSuccess) return fail(ctx, PMVSB_ECUDA, std::string("ncclGetUniqueId: ") + api->GetErrorString(r));
  std::memcpy(id128, &id, 128);
  return PMVSB_OK;
}

int pmvsb_comm_init(pmvsb_ctx* ctx, int rank, int world, const uint8_t* id128) {
  if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return fail(ctx, PMVSB_EINVAL, "comm_init: bad argument");
  if (ctx->comm) return fail(ctx, PMVSB_ESTATE, "comm_init: communicator already initialised");
  NcclApi* api = nccl_api();
  if (!api) return fail(ctx, PMVSB_ESTATE, "comm_init: libnccl.so.2 could not be loaded");
  CK(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  const ncclResult_t r = api->CommInitRank(&ctx->comm, world, id, rank);
  if (r != ncclSuccess) { ctx->comm = nullptr; return fail(ctx, PMVSB_ECUDA, std::string("ncclCommInitRank: ") + api->GetErrorString(r)); }
  ctx->comm_rank = rank; ctx->comm_world = world;
  return PMVSB_OK;
}

int pmvsb_allgather(pmvsb_ctx* ctx, const void* send, size_t bytes, void* recv) {
  if (!ctx || (bytes > 0 && (!send || !recv))) return fail(ctx, PMVSB_EINVAL, "allgather: null pointer");
  if (bytes == 0) return PMVSB_OK;
  if (ctx->comm_world == 1) { std::memcpy(recv, send, bytes); return PMVSB_OK; }
  if (!ctx->comm) return fail(ctx, PMVSB_ESTATE, "allgather: call pmvsb_comm_init first");
  NcclApi* api = nccl_api();
  CK(cudaSetDevice(ctx->device));
  const size_t need = bytes * ((size_t)ctx->comm_world + 1);
  if (need > ctx->comm_cap) {
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->comm_buf); ctx->comm_buf = nullptr; ctx->comm_cap = 0;
    CK(cudaMalloc((void**)&ctx->comm_buf, need + need / 2));
    ctx->comm_cap = need + need / 2;
  }
  char* d_send = ctx->comm_buf;
  char* d_recv = ctx->comm_buf + bytes;
  CK(cudaMemcpyAsync(d_send, send, bytes, cudaMemcpyHostToDevice, ctx->stream));
  const ncclResult_t r = api->AllGather(d_send, d_recv, bytes, ncclChar, ctx->comm, ctx->stream);
  if (r != ncclSuccess) return fail(ctx, PMVSB_ECUDA, std::string("ncclAllGather: ") + api->GetErrorString(r));
  CK(cudaMemcpyAsync(recv, d_recv, bytes * (size_t)ctx->comm_world, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}


// ---- the wave exchange of a multi-GPU run, device to device ----------------------------------------------------
// Each rank has evaluated a contiguous shard of the wave (pmvsb_evaluate_batch); its results are still on the device.  One packed
// message per rank -- verdicts of the shard, then the ACCEPTED candidates' records with their lists at their real lengths --
// goes through ONE ncclAllGather straight from device memory (NVLink / NVSwitch between the GPUs, no host staging), and is
// unpacked into the result arrays of the whole wave, in rank order = candidate order.  pmvsb_evaluate_fetch then returns the
// wave as if this rank had evaluated all of it.
__global__ void k_offsets_to_len(int n, const int32_t* __restrict__ off, int32_t* __restrict__ len) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) len[i] = off[i + 1] - off[i];
}
__global__ void k_add_scalar(int n, int32_t* __restrict__ a, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] += v;
}

// ---- peer-memory exchange (pmvsb_peer_*): mailboxes mapped between the ranks through CUDA IPC -------------------------
// Mailbox of a rank: a header of kPeerHeader bytes -- 64 bytes per SOURCE rank: [0] sizes flag, [1] data flag, [2..5] the source's
// (P, A, E, VE) of the current wave -- followed by one slot of slot_bytes per source rank.  Flags carry the wave number.
constexpr size_t kPeerHeader = 4096;
constexpr int kPeerEntryWords = 16;
static_assert(PMVSB_MAX_RANKS * kPeerEntryWords * sizeof(uint32_t) <= kPeerHeader, "mailbox header too small");
static_assert(sizeof(cudaIpcMemHandle_t) == PMVSB_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
static inline uint32_t* peer_entry(char* box, int src) { return reinterpret_cast<uint32_t*>(box) + (size_t)src * kPeerEntryWords; }
static inline int32_t* peer_slot(char* box, size_t slot_bytes, int src) { return reinterpret_cast<int32_t*>(box + kPeerHeader + (size_t)src * slot_bytes); }

struct PeerDst {
  uint32_t* entry[PMVSB_MAX_RANKS];   // this rank's header entry in every rank's mailbox
  int32_t* slot[PMVSB_MAX_RANKS];     // this rank's slot in every rank's mailbox
  int n;
};
// sizes of this rank's message into every mailbox, then the flag (system-scope fence in between: a reader that sees the flag sees the sizes)
__global__ void k_peer_post_sizes(PeerDst d, int p, int a, int e, int ve, uint32_t seq) {
  const int k = threadIdx.x;
  if (k >= d.n) return;
  volatile uint32_t* en = d.entry[k];
  en[2] = (uint32_t)p; en[3] = (uint32_t)a; en[4] = (uint32_t)e; en[5] = (uint32_t)ve;
  __threadfence_system();
  en[0] = seq;
}
// The wave's message, packed from the result arrays on the fly and STORED into every rank's mailbox (own memory for this rank,
// peer memory over NVLink for the others) by one kernel; the last block to finish raises the data flag in every mailbox.
// Message = the segments back to back; xform 1 adds `arg` (shard-local candidate index -> wave-wide), xform 2 turns CSR offsets
// into lengths (the receiver re-scans them over the whole wave).
struct WaveSeg { const int32_t* src; unsigned words; int xform; int arg; };
struct WavePost {
  WaveSeg seg[12];
  unsigned start[13];
  int nseg;
  unsigned total;
  uint32_t seq;
  int* done;
  PeerDst dst;
};
__global__ void __launch_bounds__(256) k_wave_post(const WavePost w) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < w.total; i += gridDim.x * blockDim.x) {
    int sg = 0;
#pragma unroll 1
    while (sg + 1 < w.nseg && i >= w.start[sg + 1]) ++sg;
    const WaveSeg& g = w.seg[sg];
    const unsigned j = i - w.start[sg];
    int32_t v = __ldg(g.src + j);
    if (g.xform == 1) v += g.arg;
    else if (g.xform == 2) v = __ldg(g.src + j + 1) - v;
    for (int d = 0; d < w.dst.n; ++d) w.dst.slot[d][i] = v;
  }
  __threadfence_system();   // this thread's stores are visible system-wide before its block is counted
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(w.done, 1);
    if (t == (int)gridDim.x - 1) {
      *w.done = 0;
      __threadfence_system();
      for (int d = 0; d < w.dst.n; ++d) *(volatile uint32_t*)(w.dst.entry[d] + 1) = w.seq;
    }
  }
}

int pmvsb_peer_export(pmvsb_ctx* ctx, int rank, int world, size_t slot_bytes, uint8_t* handle64) {
  if (!ctx || !handle64 || world < 1 || world > PMVSB_MAX_RANKS || rank < 0 || rank >= world || slot_bytes < 64)
    return fail(ctx, PMVSB_EINVAL, "peer_export: bad argument (at most PMVSB_MAX_RANKS ranks)");
  CK(cudaSetDevice(ctx->device));
  pmvsb_ctx::PeerBox& pb = ctx->peer;
  for (int k = 0; k < PMVSB_MAX_RANKS; ++k)
    if (pb.base[k] && pb.base[k] != pb.own) return fail(ctx, PMVSB_ESTATE, "peer_export: peers are still mapped (pmvsb_peer_close first)");
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(pb.own); pb.own = nullptr;
  slot_bytes = (slot_bytes + 255) & ~(size_t)255;
  CK(cudaMalloc((void**)&pb.own, kPeerHeader + slot_bytes * (size_t)world));   // plain cudaMalloc: pool memory has no legacy IPC handle
  CK(cudaMemset(pb.own, 0, kPeerHeader));
  if (!pb.d_done) { CK(cudaMalloc((void**)&pb.d_done, sizeof(int))); CK(cudaMemset(pb.d_done, 0, sizeof(int))); }
  if (!pb.h_board) CK(cudaHostAlloc((void**)&pb.h_board, kPeerHeader, cudaHostAllocDefault));
  CK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, pb.own));
  std::memcpy(handle64, &h, sizeof(h));
  pb.slot_bytes = slot_bytes;
  ctx->comm_rank = rank; ctx->comm_world = world;
  return PMVSB_OK;
}

int pmvsb_peer_open(pmvsb_ctx* ctx, const uint8_t* handles) {
  if (!ctx || !handles) return fail(ctx, PMVSB_EINVAL, "peer_open: null pointer");
  pmvsb_ctx::PeerBox& pb = ctx->peer;
  if (!pb.own) return fail(ctx, PMVSB_ESTATE, "peer_open: call pmvsb_peer_export first");
  CK(cudaSetDevice(ctx->device));
  peer_unmap(ctx);
  for (int k = 0; k < ctx->comm_world; ++k) {
    if (k == ctx->comm_rank) { pb.base[k] = pb.own; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + (size_t)k * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      peer_unmap(ctx);
      return fail(ctx, PMVSB_ECUDA, std::string("peer_open: cudaIpcOpenMemHandle (rank ") + std::to_string(k) + "): " + cudaGetErrorString(e));
    }
    pb.base[k] = (char*)ptr;
  }
  pb.on = ctx->comm_world > 1;
  return PMVSB_OK;
}

int pmvsb_peer_close(pmvsb_ctx* ctx) {
  if (!ctx) return PMVSB_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  peer_unmap(ctx);
  return PMVSB_OK;
}

size_t pmvsb_peer_needed(const pmvsb_ctx* ctx) { return ctx ? ctx->peer.needed : 0; }

// waits until header word `word` (0 = sizes, 1 = data) of every source rank has reached `seq`; the host polls a copy of its OWN
// mailbox header (no kernel spins on the device, so two ranks may share a GPU, and a lost peer ends in an error, not a hang)
static int peer_wait(pmvsb_ctx* ctx, int word, uint32_t seq) {
  pmvsb_ctx::PeerBox& pb = ctx->peer;
  const int W = ctx->comm_world;
  static const double limit = [] { const char* v = std::getenv("PMVSB_PEER_TIMEOUT_S"); return v && *v ? std::atof(v) : 120.0; }();
  const auto t0 = std::chrono::steady_clock::now();
  for (long it = 0;; ++it) {
    CK(cudaMemcpyAsync(pb.h_board, pb.own, sizeof(uint32_t) * kPeerEntryWords * (size_t)W, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    bool all = true;
    for (int k = 0; k < W; ++k) all = all && (int32_t)(pb.h_board[(size_t)k * kPeerEntryWords + word] - seq) >= 0;
    if (all) return PMVSB_OK;
    if ((it & 31) == 31 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit)
      return fail(ctx, PMVSB_ESTATE, "evaluate_allgather: a peer rank did not post its message in time (PMVSB_PEER_TIMEOUT_S)");
    if (it > 64) sched_yield();
  }
}

// ---- refine + all-gather of the refined records in one kernel (the multi-GPU microbench step and any caller that wants every
// rank to hold every rank's refined patches).  Each mailbox slot is used as two halves, by wave parity: a rank waits for the
// flags of ALL ranks of wave k before it starts wave k + 1, so a peer is at most one wave ahead and writes the other half.
constexpr int kRecordFloats = 12;   // coord[4], normal[4], ncc, ok, evaluations, 0
constexpr int kPeerWordRecords = 6; // header word of the records flag
__global__ void k_peer_flag(PeerDst d, int word, uint32_t seq) {
  const int k = threadIdx.x;
  if (k >= d.n) return;
  __threadfence_system();
  *(volatile uint32_t*)(d.entry[k] + word) = seq;
}
// The default form of the exchange: ONE kernel after the refine kernel packs the records from the result arrays and stores them
// into every rank's mailbox (thread i writes float4 i of this rank's record array in all mailboxes: coalesced 16-byte peer
// stores), the last block raises the flags.  Measured on B200 (1 048 576 patches, profiles/r2_peer_exchange.txt): the in-kernel
// form (k_refine_g<., ., GATHER>) costs the refine kernel +1.0 % (125.6 vs 124.4 ms: the once-per-patch call perturbs the register
// allocation of a loop that runs at the 64-register cap), more than the whole exchange takes as a kernel of its own.
__global__ void __launch_bounds__(256) k_records_post(int P, const float4* __restrict__ coords, const float4* __restrict__ normals, const float* __restrict__ ncc,
                                                      const uint8_t* __restrict__ ok, const int32_t* __restrict__ evals, const RecDst rd, PeerDst flags,
                                                      int word, uint32_t seq, int* done) {
  const int total = 3 * P;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int p = i / 3, part = i - 3 * p;
    float4 v;
    if (part == 0) v = coords[p];
    else if (part == 1) v = normals[p];
    else v = make_float4(ncc[p], ok[p] ? 1.0f : 0.0f, (float)evals[p], 0.0f);
    for (int d = 0; d < rd.n; ++d) rd.rec[d][i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(done, 1);
    if (t == (int)gridDim.x - 1) {
      *done = 0;
      __threadfence_system();
      for (int d = 0; d < flags.n; ++d) *(volatile uint32_t*)(flags.entry[d] + word) = seq;
    }
  }
}
// cuStreamWaitValue32 through the runtime's driver entry point (no link against libcuda): the wait for the peers' flags is an
// operation of the stream, the host does not block
typedef int (*StreamWaitValue32Fn)(cudaStream_t, unsigned long long, uint32_t, unsigned int);
static StreamWaitValue32Fn stream_wait_value32() {
  static StreamWaitValue32Fn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) { cudaGetLastError(); f = nullptr; }
    return (StreamWaitValue32Fn)f;
  }();
  return fn;
}

int pmvsb_refine_batch_dev_gather(pmvsb_ctx* ctx, int P, int stride, float* d_coords, float* d_normals, const int32_t* d_images,
                                  const int32_t* d_nimages, const float* d_dscales, float* d_ncc, int32_t* d_evals, uint8_t* d_ok,
                                  const float** records, size_t* rank_stride_floats) {
  int r = check_ready(ctx);
  if (r) return r;
  pmvsb_ctx::PeerBox& pb = ctx->peer;
  if (!pb.own || (ctx->comm_world > 1 && !pb.on)) return fail(ctx, PMVSB_ESTATE, "refine_batch_dev_gather: call pmvsb_peer_export and pmvsb_peer_open first");
  if (P < 0) return fail(ctx, PMVSB_EINVAL, "refine_batch_dev_gather: bad argument");
  const size_t half = (pb.slot_bytes / 2) & ~(size_t)255;
  if ((size_t)P * kRecordFloats * sizeof(float) > half) {
    pb.needed = 2 * ((size_t)P * kRecordFloats * sizeof(float) + 256);
    return fail(ctx, PMVSB_EGROW, "refine_batch_dev_gather: the records do not fit half a mailbox slot");
  }
  const int W = ctx->comm_world;
  const uint32_t seq = ++pb.seq;
  const size_t off = (seq & 1u) ? half : 0;
  RecDst rd;
  PeerDst dst;
  rd.n = dst.n = W;
  for (int k = 0; k < PMVSB_MAX_RANKS; ++k) {
    char* box = k < W ? (k == ctx->comm_rank ? pb.own : pb.base[k]) : nullptr;
    rd.rec[k] = box ? reinterpret_cast<float4*>(reinterpret_cast<char*>(peer_slot(box, pb.slot_bytes, ctx->comm_rank)) + off) : nullptr;
    dst.entry[k] = box ? peer_entry(box, ctx->comm_rank) : nullptr;
    dst.slot[k] = nullptr;
  }
  const char* in_kernel_env = std::getenv("PMVSB_GATHER_IN_KERNEL");
  const bool in_kernel = in_kernel_env && *in_kernel_env == '1';
  if (in_kernel) {
    // A/B form: the refine kernel itself stores each finished patch's record (k_refine_g<., ., GATHER>), then one flag kernel
    if (!pb.d_rec) CK(cudaMalloc((void**)&pb.d_rec, 2 * sizeof(RecDst)));
    RecDst* d_rd = reinterpret_cast<RecDst*>(pb.d_rec) + (seq & 1u);   // (the previous call's kernel may still be reading the other one)
    CK(cudaMemcpyAsync(d_rd, &rd, sizeof(RecDst), cudaMemcpyHostToDevice, ctx->stream));
    if ((r = refine_dev_impl(ctx, P, stride, d_coords, d_normals, d_images, d_nimages, d_dscales, d_ncc, d_evals, d_ok, d_rd))) return r;
    k_peer_flag<<<1, 32, 0, ctx->stream>>>(dst, kPeerWordRecords, seq);
  } else {
    if ((r = refine_dev_impl(ctx, P, stride, d_coords, d_normals, d_images, d_nimages, d_dscales, d_ncc, d_evals, d_ok, nullptr))) return r;
    if (!pb.d_done) { CK(cudaMalloc((void**)&pb.d_done, sizeof(int))); CK(cudaMemsetAsync(pb.d_done, 0, sizeof(int), ctx->stream)); }
    const int blocks = (int)std::max<long long>(1, std::min<long long>((3ll * P + 255) / 256, (long long)ctx->sm_count * 8));
    k_records_post<<<blocks, 256, 0, ctx->stream>>>(P, reinterpret_cast<const float4*>(d_coords), reinterpret_cast<const float4*>(d_normals), d_ncc, d_ok, d_evals,
                                                    rd, dst, kPeerWordRecords, seq, pb.d_done);
  }
  ++ctx->launches;
  CK(cudaGetLastError());
  if (W > 1) {
    StreamWaitValue32Fn wait = stream_wait_value32();
    if (wait) {
      for (int k = 0; k < W; ++k) {
        if (k == ctx->comm_rank) continue;
        const int e = wait(ctx->stream, (unsigned long long)(uintptr_t)(peer_entry(pb.own, k) + kPeerWordRecords), seq, 0u /* CU_STREAM_WAIT_VALUE_GEQ: (int32)(*addr - seq) >= 0 */);
        if (e != 0) return fail(ctx, PMVSB_ECUDA, "refine_batch_dev_gather: cuStreamWaitValue32 failed (" + std::to_string(e) + ")");
      }
    } else if ((r = peer_wait(ctx, kPeerWordRecords, seq))) return r;
  }
  if (records) *records = reinterpret_cast<const float*>(reinterpret_cast<char*>(peer_slot(pb.own, pb.slot_bytes, 0)) + off);
  if (rank_stride_floats) *rank_stride_floats = pb.slot_bytes / sizeof(float);
  return PMVSB_OK;
}

int pmvsb_evaluate_allgather(pmvsb_ctx* ctx, int shard_lo, int total_candidates) {
  int r = check_ready(ctx);
  if (r) return r;
  pmvsb_ctx::EvalOut& ev = ctx->ev;
  if (shard_lo < 0 || total_candidates < shard_lo + ev.P) return fail(ctx, PMVSB_EINVAL, "evaluate_allgather: shard outside the wave");
  if (ctx->comm_world == 1) return PMVSB_OK;
  const bool peer = ctx->peer.on;
  if (!peer && !ctx->comm) return fail(ctx, PMVSB_ESTATE, "evaluate_allgather: call pmvsb_peer_open or pmvsb_comm_init first");
  NcclApi* api = peer ? nullptr : nccl_api();
  cudaStream_t q = ctx->stream;
  const int W = ctx->comm_world;
  pmvsb_ctx::PeerBox& pb = ctx->peer;
  PeerDst dst;
  dst.n = W;
  for (int k = 0; k < PMVSB_MAX_RANKS; ++k) {
    dst.entry[k] = peer && k < W ? peer_entry(pb.base[k], ctx->comm_rank) : nullptr;
    dst.slot[k] = peer && k < W ? peer_slot(pb.base[k], pb.slot_bytes, ctx->comm_rank) : nullptr;
  }
  // ---- sizes of every rank's message
  std::vector<int32_t> meta((size_t)4 * W);
  DevBuf<int32_t> dmeta;
  if (peer) {
    ++pb.seq;
    k_peer_post_sizes<<<1, 32, 0, q>>>(dst, ev.P, ev.A, ev.E, ev.VE, pb.seq);
    ++ctx->launches;
    CK(cudaGetLastError());
    if ((r = peer_wait(ctx, 0, pb.seq))) return r;
    for (int k = 0; k < W; ++k)
      for (int c = 0; c < 4; ++c) meta[4 * k + c] = (int32_t)pb.h_board[(size_t)k * kPeerEntryWords + 2 + c];
  } else {
    CK(dmeta.alloc((size_t)4 * (W + 1)));
    const int32_t mine[4] = {ev.P, ev.A, ev.E, ev.VE};
    CK(cudaMemcpyAsync(dmeta.p, mine, sizeof(mine), cudaMemcpyHostToDevice, q));
    ncclResult_t nr = api->AllGather(dmeta.p, dmeta.p + 4, 4, ncclInt32, ctx->comm, q);
    if (nr != ncclSuccess) return fail(ctx, PMVSB_ECUDA, std::string("ncclAllGather: ") + api->GetErrorString(nr));
    CK(cudaMemcpyAsync(meta.data(), dmeta.p + 4, sizeof(int32_t) * meta.size(), cudaMemcpyDeviceToHost, q));
    CK(cudaStreamSynchronize(q));
  }
  auto words = [](int P, int A, int E, int VE) { return (size_t)P + (size_t)4 * A + (size_t)12 * A + (size_t)3 * E + (size_t)3 * VE; };
  size_t longest = 0;
  long long tP = 0, tA = 0, tE = 0, tVE = 0;
  for (int k = 0; k < W; ++k) {
    longest = std::max(longest, words(meta[4 * k], meta[4 * k + 1], meta[4 * k + 2], meta[4 * k + 3]));
    tP += meta[4 * k]; tA += meta[4 * k + 1]; tE += meta[4 * k + 2]; tVE += meta[4 * k + 3];
  }
  if (tP != total_candidates) return fail(ctx, PMVSB_EINVAL, "evaluate_allgather: the shards do not add up to the wave");
  longest = (longest + 3) & ~(size_t)3;
  const int32_t* recv = nullptr;   // rank k's message starts at recv + k * recv_stride
  size_t recv_stride = 0;
  if (peer) {
    if (sizeof(int32_t) * longest > pb.slot_bytes) {   // every rank sees the same sizes, so every rank takes this exit
      pb.needed = sizeof(int32_t) * longest;
      return fail(ctx, PMVSB_EGROW, "evaluate_allgather: the wave's message outgrew the peer mailbox slots");
    }
    // ---- pack + scatter: one kernel stores this rank's message into every rank's mailbox, the last block raises the flags
    WavePost w;
    const int P = ev.P, A = ev.A, E = ev.E, VE = ev.VE;
    int n = 0;
    unsigned at = 0;
    auto seg = [&](const void* src, size_t cnt, int xform, int arg) {
      if (cnt == 0) return;
      w.seg[n] = {reinterpret_cast<const int32_t*>(src), (unsigned)cnt, xform, arg};
      w.start[n] = at;
      at += (unsigned)cnt;
      ++n;
    };
    seg(ev.verdict.p, (size_t)P, 0, 0);
    if (A > 0) {
      seg(ev.index.p, (size_t)A, 1, shard_lo); seg(ev.timages.p, (size_t)A, 0, 0);
      seg(ev.img_off.p, (size_t)A, 2, 0); seg(ev.vimg_off.p, (size_t)A, 2, 0);
      seg(ev.coords.p, (size_t)4 * A, 0, 0); seg(ev.normals.p, (size_t)4 * A, 0, 0); seg(ev.scal.p, (size_t)4 * A, 0, 0);
      seg(ev.images.p, (size_t)E, 0, 0); seg(ev.grids.p, (size_t)2 * E, 0, 0); seg(ev.vimages.p, (size_t)VE, 0, 0); seg(ev.vgrids.p, (size_t)2 * VE, 0, 0);
    }
    for (int k = n; k < 12; ++k) w.seg[k] = {nullptr, 0u, 0, 0};
    for (int k = n; k < 13; ++k) w.start[k] = at;
    w.nseg = n > 0 ? n : 1; w.total = at; w.seq = pb.seq; w.done = pb.d_done; w.dst = dst;
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>(((size_t)at + 255) / 256, (size_t)ctx->sm_count * 8));
    k_wave_post<<<blocks, 256, 0, q>>>(w);
    ++ctx->launches;
    CK(cudaGetLastError());
    if ((r = peer_wait(ctx, 1, pb.seq))) return r;
    recv = peer_slot(pb.own, pb.slot_bytes, 0);
    recv_stride = pb.slot_bytes / sizeof(int32_t);
  } else {
    const size_t need = sizeof(int32_t) * longest * ((size_t)W + 1);
    if (need > ctx->comm_cap) {
      CK(cudaStreamSynchronize(q));
      cudaFree(ctx->comm_buf); ctx->comm_buf = nullptr; ctx->comm_cap = 0;
      CK(cudaMalloc((void**)&ctx->comm_buf, need + need / 2));
      ctx->comm_cap = need + need / 2;
    }
    int32_t* send = reinterpret_cast<int32_t*>(ctx->comm_buf);
    // ---- pack: [verdict P][index A][timages A][ilen A][vlen A][coords 4A][normals 4A][scal 4A][images E][grids 2E][vimages VE][vgrids 2VE]
    {
      const int P = ev.P, A = ev.A, E = ev.E, VE = ev.VE;
      int32_t* w = send;
      auto put = [&](const void* src, size_t n) { cudaError_t e = n ? cudaMemcpyAsync(w, src, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, q) : cudaSuccess; w += n; return e; };
      CK(put(ev.verdict.p, (size_t)P));
      if (A > 0) {
        k_add_scalar<<<(A + 255) / 256, 256, 0, q>>>(A, ev.index.p, shard_lo);
        CK(put(ev.index.p, (size_t)A)); CK(put(ev.timages.p, (size_t)A));
        k_offsets_to_len<<<(A + 255) / 256, 256, 0, q>>>(A, ev.img_off.p, w);
        k_offsets_to_len<<<(A + 255) / 256, 256, 0, q>>>(A, ev.vimg_off.p, w + A);
        ctx->launches += 3;
        w += 2 * (size_t)A;
        CK(put(ev.coords.p, (size_t)4 * A)); CK(put(ev.normals.p, (size_t)4 * A)); CK(put(ev.scal.p, (size_t)4 * A));
        CK(put(ev.images.p, (size_t)E)); CK(put(ev.grids.p, (size_t)2 * E)); CK(put(ev.vimages.p, (size_t)VE)); CK(put(ev.vgrids.p, (size_t)2 * VE));
      }
    }
    ncclResult_t nr = api->AllGather(send, send + longest, longest, ncclInt32, ctx->comm, q);
    if (nr != ncclSuccess) return fail(ctx, PMVSB_ECUDA, std::string("ncclAllGather: ") + api->GetErrorString(nr));
    recv = send + longest;
    recv_stride = longest;
  }
  // ---- unpack into the arrays of the whole wave
  const size_t A1 = (size_t)std::max<long long>(tA, 1), E1 = (size_t)std::max<long long>(tE, 1), V1 = (size_t)std::max<long long>(tVE, 1);
  pmvsb_ctx::EvalOut& o = ctx->ev_all;
  if ((r = dvec_reserve(ctx, o.verdict, (size_t)std::max<long long>(tP, 1))) || (r = dvec_reserve(ctx, o.index, A1)) || (r = dvec_reserve(ctx, o.timages, A1)) ||
      (r = dvec_reserve(ctx, o.img_off, A1 + 1)) || (r = dvec_reserve(ctx, o.vimg_off, A1 + 1)) || (r = dvec_reserve(ctx, o.coords, 4 * A1)) ||
      (r = dvec_reserve(ctx, o.normals, 4 * A1)) || (r = dvec_reserve(ctx, o.scal, 4 * A1)) || (r = dvec_reserve(ctx, o.images, E1)) ||
      (r = dvec_reserve(ctx, o.grids, 2 * E1)) || (r = dvec_reserve(ctx, o.vimages, V1)) || (r = dvec_reserve(ctx, o.vgrids, 2 * V1)))
    return r;
  size_t aP = 0, aA = 0, aE = 0, aV = 0;
  for (int k = 0; k < W; ++k) {
    const int P = meta[4 * k], A = meta[4 * k + 1], E = meta[4 * k + 2], VE = meta[4 * k + 3];
    const int32_t* w = recv + (size_t)k * recv_stride;
    auto get = [&](void* dst, size_t n) { cudaError_t e = n ? cudaMemcpyAsync(dst, w, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, q) : cudaSuccess; w += n; return e; };
    CK(get(o.verdict.p + aP, (size_t)P));
    CK(get(o.index.p + aA, (size_t)A)); CK(get(o.timages.p + aA, (size_t)A));
    CK(get(o.img_off.p + aA, (size_t)A)); CK(get(o.vimg_off.p + aA, (size_t)A));     // lengths for now
    CK(get(o.coords.p + 4 * aA, (size_t)4 * A)); CK(get(o.normals.p + 4 * aA, (size_t)4 * A)); CK(get(o.scal.p + 4 * aA, (size_t)4 * A));
    CK(get(o.images.p + aE, (size_t)E)); CK(get(o.grids.p + 2 * aE, (size_t)2 * E)); CK(get(o.vimages.p + aV, (size_t)VE)); CK(get(o.vgrids.p + 2 * aV, (size_t)2 * VE));
    aP += P; aA += A; aE += E; aV += VE;
  }
  CK(cudaMemsetAsync(o.img_off.p + tA, 0, sizeof(int32_t), q));
  CK(cudaMemsetAsync(o.vimg_off.p + tA, 0, sizeof(int32_t), q));
  if ((r = device_scan(ctx, o.img_off.p, (int)tA + 1))) return r;
  if ((r = device_scan(ctx, o.vimg_off.p, (int)tA + 1))) return r;
  CK(cudaStreamSynchronize(q));
  o.P = (int)tP; o.A = (int)tA; o.E = (int)tE; o.VE = (int)tVE; o.refined = ev.refined;
  std::swap(ctx->ev, ctx->ev_all);
  ctx->exchanged_bytes += (double)sizeof(int32_t) * (peer ? (double)(words((int)tP, (int)tA, (int)tE, (int)tVE)) : (double)longest * W);
  return PMVSB_OK;
}

int pmvsb_evaluate_counts(pmvsb_ctx* ctx, int32_t* candidates, int32_t* accepted, int32_t* entries, int32_t* ventries) {
  if (!ctx) return PMVSB_EINVAL;
  if (candidates) *candidates = ctx->ev.P;
  if (accepted) *accepted = ctx->ev.A;
  if (entries) *entries = ctx->ev.E;
  if (ventries) *ventries = ctx->ev.VE;
  return PMVSB_OK;
}

double pmvsb_exchanged_bytes(const pmvsb_ctx* ctx) { return ctx ? ctx->exchanged_bytes : 0.0; }

int pmvsb_sync(pmvsb_ctx* ctx) {
  if (!ctx) return PMVSB_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return PMVSB_OK;
}

void* pmvsb_stream(pmvsb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int pmvsb_set_stream(pmvsb_ctx* ctx, void* cuda_stream) {
  if (!ctx) return PMVSB_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return PMVSB_OK;
}

uint64_t pmvsb_launch_count(const pmvsb_ctx* ctx) { return ctx ? ctx->launches : 0; }

float pmvsb_last_refine_ms(pmvsb_ctx* ctx) {
  if (!ctx || !ctx->refine_timed) return -1.0f;
  float ms = -1.0f;
  if (cudaEventSynchronize(ctx->ev1) != cudaSuccess) return -1.0f;
  if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) != cudaSuccess) return -1.0f;
  return ms;
}

}  // extern "C"
