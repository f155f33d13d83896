// cmvs-pmvs_b200/csrc/pmvs_filter.cuh
//
// Filter-stage kernels (K5 / K6): nearest-patch depth maps, the depth-consistency visibility test and the
// filterOutside gain, over a patch table that the host hands over as arrays (the cell bookkeeping itself --
// CPatchOrganizerS -- stays on the host, BASELINE.json north_star).  All results are integers or f32 values in
// the reference's operation order: bit-exact.  Reference: /root/reference/source/pmvs/filter.cpp,
// source/pmvs/patchOrganizerS.cpp, source/pmvs/findMatch.cpp (lines cited per function).
#pragma once
#include "pmvs_device.cuh"

namespace pmvsb {

struct StoreDev {               // the patch table (CPatch fields the filter stage reads), SoA + CSR
  int P;
  const float* coords;          // float4 per patch
  const float* normals;
  const float* ncc;
  const float* dscale;
  const int32_t* img_off;       // CSR of CPatch::_images / _grids
  const int32_t* images;
  const int32_t* grids;         // 2 per entry
  const int32_t* entry_patch;   // COO companion of img_off
  const int32_t* vimg_off;      // CSR of CPatch::_vimages / _vgrids
  const int32_t* vimages;
  const int32_t* vgrids;
  const int32_t* timages;
  const int32_t* cell_base;     // per target image: first cell of its grid in the flattened cell arrays
  const int32_t* gw;            // grid width / height per image (patchOrganizerS.cpp:72-77)
  const int32_t* gh;
  const int32_t* cell_off;      // _pgrids as CSR over flattened cells
  const int32_t* cell_patch;
  const int32_t* vcell_off;     // _vpgrids likewise (patches that only SEE the cell, patchOrganizerS.cpp:333-346)
  const int32_t* vcell_patch;
  unsigned long long* dp;       // depth map: (orderable depth << 32 | patch id) per flattened cell, ~0 = empty
  int depth_flag;               // CFindMatch::_depth
  float ncc_threshold;
  float cos120_f;               // smallest float >= cos(120 deg): `n.n < cos(..)` is a double compare (findMatch.cpp:126)
};

__device__ __forceinline__ unsigned int float_order(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// K5: CFilter::setDepthMapsThread (filter.cpp:687-732).  One thread per (patch, target image); the reference walks
// the patches in table order and replaces a cell only by a STRICTLY nearer patch, i.e. the winner is the minimum of
// (depth, patch id) -- a 64-bit atomicMin.
__global__ void k_depth_maps(SceneDev s, StoreDev st, int p0, int count) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)count * s.tnum) return;
  const int p = p0 + (int)(t / s.tnum), image = (int)(t % s.tnum);
  CamDev cam;
  load_cam(s, image, cam);
  const float4 c4 = __ldg(reinterpret_cast<const float4*>(st.coords) + p);
  const float X[4] = {c4.x, c4.y, c4.z, c4.w};
  float ic[3];
  project(cam, X, ic);
  const float fx = ic[0] / (float)s.csize, fy = ic[1] / (float)s.csize;
  const int xs[2] = {(int)floorf(fx), (int)ceilf(fx)};
  const int ys[2] = {(int)floorf(fy), (int)ceilf(fy)};
  const float depth = dot4(cam.oaxis, X);
  const unsigned long long key = ((unsigned long long)float_order(depth) << 32) | (unsigned int)p;
  const int gw = st.gw[image], gh = st.gh[image];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (xs[i] < 0 || gw <= xs[i] || ys[j] < 0 || gh <= ys[j]) continue;
      if (j == 1 && ys[1] == ys[0]) continue;   // floor == ceil: the reference visits the same cell twice, harmless
      if (i == 1 && xs[1] == xs[0]) continue;
      atomicMin(st.dp + st.cell_base[image] + ys[j] * gw + xs[i], key);
    }
}

// CPatchOrganizerS::isVisible (patchOrganizerS.cpp:487-526)
__device__ __forceinline__ int is_visible(const SceneDev& s, const StoreDev& st, const CamDev& cam, const float* coord,
                                          const float* normal, int image, int ix, int iy, float strict) {
  const int gw = st.gw[image], gh = st.gh[image];
  if (ix < 0 || gw <= ix || iy < 0 || gh <= iy) return 0;
  if (st.depth_flag == 0) return 1;
  const unsigned long long key = st.dp[st.cell_base[image] + iy * gw + ix];
  if (key == ~0ull) return 1;
  const int q = (int)(key & 0xffffffffull);
  float ray[4] = {coord[0] - cam.centre[0], coord[1] - cam.centre[1], coord[2] - cam.centre[2], coord[3] - cam.centre[3]};
  unitize4(ray);
  const float4 y4 = __ldg(reinterpret_cast<const float4*>(st.coords) + q);
  const float d[4] = {coord[0] - y4.x, coord[1] - y4.y, coord[2] - y4.z, coord[3] - y4.w};
  const float diff = dot4(ray, d);
  const double factor = fmin(2.0, 2.0 + (double)dot4(ray, normal));
  const float lim = get_unit(cam, s.level, coord) * (float)s.csize * strict;
  return ((double)diff < (double)lim * factor) ? 1 : 0;
}

// CPatchOrganizerS::setVImagesVGrids (patchOrganizerS.cpp:420-450) for every store patch, starting from an empty
// _vimages (what setDepthMapsVGridsVPGridsAddPatchV(0) does, filter.cpp:752-762).  One warp per patch; lanes walk
// the target images 32 at a time and append in image order.  strict = _neighborThreshold = 0.5.
__global__ void k_set_vimages(SceneDev s, StoreDev st, int vcap, int32_t* __restrict__ vimages, int32_t* __restrict__ vgrids,
                              int32_t* __restrict__ nv) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= st.P) return;
  const int p = warp;
  const float4 c4 = __ldg(reinterpret_cast<const float4*>(st.coords) + p);
  const float4 n4 = __ldg(reinterpret_cast<const float4*>(st.normals) + p);
  const float X[4] = {c4.x, c4.y, c4.z, c4.w}, N[4] = {n4.x, n4.y, n4.z, n4.w};
  const int e0 = st.img_off[p], e1 = st.img_off[p + 1];
  int n = 0;
  for (int base = 0; base < s.tnum; base += 32) {
    const int image = base + lane;
    bool ok = false;
    int ix = 0, iy = 0;
    if (image < s.tnum) {
      bool used = false;
      for (int e = e0; e < e1; ++e) used |= (st.images[e] == image);
      if (!used) {
        CamDev cam;
        load_cam(s, image, cam);
        float ic[3];
        project(cam, X, ic);
        ix = ((int)floorf(ic[0] + 0.5f)) / s.csize;
        iy = ((int)floorf(ic[1] + 0.5f)) / s.csize;
        ok = is_visible(s, st, cam, X, N, image, ix, iy, 0.5f) != 0 && get_edge_img(s, cam, image, X) != 0;   // patchOrganizerS.cpp:444-445
      }
    }
    const unsigned m = __ballot_sync(kFull, ok);
    const int pos = n + __popc(m & ((1u << lane) - 1u));
    if (ok && pos < vcap) {
      vimages[(size_t)p * vcap + pos] = image;
      vgrids[((size_t)p * vcap + pos) * 2] = ix;
      vgrids[((size_t)p * vcap + pos) * 2 + 1] = iy;
    }
    n += __popc(m);
  }
  if (lane == 0) nv[p] = n < vcap ? n : vcap;
}

// setVImagesVGrids for a batch of patches outside the table (candidates inside postProcess, or the table itself in
// additive mode): images[] and the existing vimages[] count as used, new visible target images are appended.
__global__ void k_set_vimages_batch(SceneDev s, StoreDev st, int P, int stride, const float* __restrict__ coords,
                                    const float* __restrict__ normals, const int32_t* __restrict__ images,
                                    const int32_t* __restrict__ nimages, int vstride, int32_t* __restrict__ vimages,
                                    int32_t* __restrict__ nv, int32_t* __restrict__ vgrids) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P) return;
  const int p = warp;
  const float4 c4 = __ldg(reinterpret_cast<const float4*>(coords) + p);
  const float4 n4 = __ldg(reinterpret_cast<const float4*>(normals) + p);
  const float X[4] = {c4.x, c4.y, c4.z, c4.w}, N[4] = {n4.x, n4.y, n4.z, n4.w};
  const int ni = min(nimages[p], stride);
  const int nv0 = min(nv[p], vstride);
  int n = nv0;
  for (int base = 0; base < s.tnum; base += 32) {
    const int image = base + lane;
    bool ok = false;
    int ix = 0, iy = 0;
    if (image < s.tnum) {
      bool used = false;
      for (int e = 0; e < ni; ++e) used |= (images[(size_t)p * stride + e] == image);
      for (int e = 0; e < nv0; ++e) used |= (vimages[(size_t)p * vstride + e] == image);
      if (!used) {
        CamDev cam;
        load_cam(s, image, cam);
        float ic[3];
        project(cam, X, ic);
        ix = ((int)floorf(ic[0] + 0.5f)) / s.csize;
        iy = ((int)floorf(ic[1] + 0.5f)) / s.csize;
        ok = is_visible(s, st, cam, X, N, image, ix, iy, 0.5f) != 0 && get_edge_img(s, cam, image, X) != 0;   // patchOrganizerS.cpp:444-445
      }
    }
    const unsigned m = __ballot_sync(kFull, ok);
    const int pos = n + __popc(m & ((1u << lane) - 1u));
    if (ok && pos < vstride) {
      vimages[(size_t)p * vstride + pos] = image;
      vgrids[((size_t)p * vstride + pos) * 2] = ix;
      vgrids[((size_t)p * vstride + pos) * 2 + 1] = iy;
    }
    n += __popc(m);
    __syncwarp();
  }
  if (lane == 0) nv[p] = n < vstride ? n : vstride;
}

// writePLY's vertex colour (patchOrganizerS.cpp:713-727): one thread per patch
__global__ void k_patch_colors(SceneDev s, int P, int stride, const float* __restrict__ coords, const int32_t* __restrict__ images,
                               const int32_t* __restrict__ nimages, uint8_t* __restrict__ rgb) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float4 c4 = __ldg(reinterpret_cast<const float4*>(coords) + p);
  const float X[4] = {c4.x, c4.y, c4.z, c4.w};
  const int n = min(nimages[p], stride);
  float acc[3] = {0.f, 0.f, 0.f};
  int denom = 0;
  for (int i = 0; i < n; ++i) {
    const int image = images[(size_t)p * stride + i];
    CamDev cam;
    load_cam(s, image, cam);
    float ic[3];
    project(cam, X, ic);
    const LevelDev lv = s.levels[image * s.nlevels + s.level];
    // the reference samples without a bounds test; stay inside the image instead of reading out of it
    const float x = smin(smax(ic[0], 0.0f), (float)(lv.w - 2)), y = smin(smax(ic[1], 0.0f), (float)(lv.h - 2));
    float c[3];
    get_color(lv, x, y, c);
    acc[0] += c[0]; acc[1] += c[1]; acc[2] += c[2];
    ++denom;
  }
  if (denom == 0) denom = 1;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int v = (int)floorf(acc[k] / (float)denom + 0.5f);
    rgb[3 * p + k] = (uint8_t)(v < 255 ? (v < 0 ? 0 : v) : 255);
  }
}

// CFilter::filterExactThread's test (filter.cpp:315-343): the patch stays in (image, x, y) if it is visible there
// or in one of the four neighbouring cells.  One thread per image entry.  strict = _neighborThreshold1 = 1.0.
__global__ void k_filter_exact(SceneDev s, StoreDev st, int nentries, uint8_t* __restrict__ safe) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nentries) return;
  const int image = st.images[e];
  if (image >= s.tnum) { safe[e] = 1; return; }   // non-target images are kept as they are (filter.cpp:261-266)
  const int p = st.entry_patch[e];
  const int x = st.grids[2 * e], y = st.grids[2 * e + 1];
  const float4 c4 = __ldg(reinterpret_cast<const float4*>(st.coords) + p);
  const float4 n4 = __ldg(reinterpret_cast<const float4*>(st.normals) + p);
  const float X[4] = {c4.x, c4.y, c4.z, c4.w}, N[4] = {n4.x, n4.y, n4.z, n4.w};
  CamDev cam;
  load_cam(s, image, cam);
  const int w = st.gw[image], h = st.gh[image];
  int ok = is_visible(s, st, cam, X, N, image, x, y, 1.0f);
  if (!ok && 0 < x) ok = is_visible(s, st, cam, X, N, image, x - 1, y, 1.0f);
  if (!ok && x < w - 1) ok = is_visible(s, st, cam, X, N, image, x + 1, y, 1.0f);
  if (!ok && 0 < y) ok = is_visible(s, st, cam, X, N, image, x, y - 1, 1.0f);
  if (!ok && y < h - 1) ok = is_visible(s, st, cam, X, N, image, x, y + 1, 1.0f);
  safe[e] = (uint8_t)ok;
}

// CFindMatch::isNeighbor (findMatch.cpp:120-149)
__device__ __forceinline__ int is_neighbor(const SceneDev& s, const StoreDev& st, int a, const float* Xa, const float* Na, float ua,
                                           int b, float thr) {
  const float4 xb = __ldg(reinterpret_cast<const float4*>(st.coords) + b);
  const float4 nb = __ldg(reinterpret_cast<const float4*>(st.normals) + b);
  const float Xb[4] = {xb.x, xb.y, xb.z, xb.w}, Nb[4] = {nb.x, nb.y, nb.z, nb.w};
  CamDev cam;
  load_cam(s, st.images[st.img_off[b]], cam);
  const float ub = get_unit(cam, s.level, Xb);
  const float hunit = (float)((double)(ua + ub) / 2.0 * (double)s.csize);
  if (dot4(Na, Nb) < st.cos120_f) return 0;
  const float diff[4] = {Xb[0] - Xa[0], Xb[1] - Xa[1], Xb[2] - Xa[2], Xb[3] - Xa[3]};
  const float vunit = st.dscale[a] + st.dscale[b];
  const float f0 = dot4(Na, diff);
  const float f1 = dot4(Nb, diff);
  float ftmp = (fabsf(f0) + fabsf(f1)) / 2.0f;
  ftmp /= vunit;
  float t[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) t[k] = diff[k] * 2.0f - Na[k] * f0 - Nb[k] * f1;
  const float hsize = (float)((double)sqrtf(dot4(t, t)) / 2.0 / (double)hunit);
  if (1.0f < hsize) ftmp /= smin(2.0f, hsize);
  return ftmp < thr ? 1 : 0;
}

// K6: CFilter::filterOutsideThread / computeGain (filter.cpp:88-201).  One thread per patch.
__global__ void k_gains(SceneDev s, StoreDev st, float* __restrict__ gains) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= st.P) return;
  const float4 c4 = __ldg(reinterpret_cast<const float4*>(st.coords) + p);
  const float4 n4 = __ldg(reinterpret_cast<const float4*>(st.normals) + p);
  const float X[4] = {c4.x, c4.y, c4.z, c4.w}, N[4] = {n4.x, n4.y, n4.z, n4.w};
  const float thr = st.ncc_threshold;
  float gain = smax(0.0f, st.ncc[p] - thr) * (float)st.timages[p];   // score2 (include/pmvs/patch.hpp:48-50)
  if (st.img_off[p + 1] == st.img_off[p]) { gains[p] = gain; return; }
  CamDev cam;
  load_cam(s, st.images[st.img_off[p]], cam);
  const float ua = get_unit(cam, s.level, X);
  for (int e = st.img_off[p]; e < st.img_off[p + 1]; ++e) {
    const int index = st.images[e];
    if (s.tnum <= index) continue;
    const int cell = st.cell_base[index] + st.grids[2 * e + 1] * st.gw[index] + st.grids[2 * e];
    float maxpressure = 0.0f;
    for (int j = st.cell_off[cell]; j < st.cell_off[cell + 1]; ++j) {
      const int q = st.cell_patch[j];
      if (!is_neighbor(s, st, p, X, N, ua, q, 1.0f)) maxpressure = smax(maxpressure, st.ncc[q] - thr);
    }
    gain -= maxpressure;
  }
  for (int e = st.vimg_off[p]; e < st.vimg_off[p + 1]; ++e) {
    const int index = st.vimages[e];
    if (s.tnum <= index) continue;
    load_cam(s, index, cam);
    const float pdepth = dot4(cam.oaxis, X);   // CCamera::computeDepth (source/image/camera.cpp:445-452), perspective
    const int cell = st.cell_base[index] + st.vgrids[2 * e + 1] * st.gw[index] + st.vgrids[2 * e];
    float maxpressure = 0.0f;
    for (int j = st.cell_off[cell]; j < st.cell_off[cell + 1]; ++j) {
      const int q = st.cell_patch[j];
      const float4 xq = __ldg(reinterpret_cast<const float4*>(st.coords) + q);
      const float Xq[4] = {xq.x, xq.y, xq.z, xq.w};
      const float bdepth = dot4(cam.oaxis, Xq);
      if (pdepth < bdepth && !is_neighbor(s, st, p, X, N, ua, q, 1.0f)) maxpressure = smax(maxpressure, st.ncc[q] - thr);
    }
    gain -= maxpressure;
  }
  gains[p] = gain;
}

}  // namespace pmvsb
