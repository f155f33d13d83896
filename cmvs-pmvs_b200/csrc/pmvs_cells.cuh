// cmvs-pmvs_b200/csrc/pmvs_cells.cuh
//
// Cell bookkeeping of the resident patch table on the device, and the two neighbour-search consumers:
//   * _pgrids / _vpgrids as CSR over the flattened cells of all target images, built by count -> scan -> fill ->
//     per-cell sort (CPatchOrganizerS::addPatch, source/pmvs/patchOrganizerS.cpp:312-349);
//   * CPatchOrganizerS::setVImagesVGrids for the whole table, written straight into the table's own lists
//     (CFilter::setDepthMapsVGridsVPGridsAddPatchV, source/pmvs/filter.cpp:734-783);
//   * CPatchOrganizerS::findNeighbors (patchOrganizerS.cpp:528-651) feeding CExpand::findEmptyBlocks
//     (source/pmvs/expand.cpp:108-180) and CFilter::filterNeighborThread / filterQuad (filter.cpp:357-462).
// One warp owns one patch; lanes own (image, cell) pairs of its search window.  All paths relative to /root/reference.
#pragma once
#include "pmvs_filter.cuh"

namespace pmvsb {

// ---------------------------------------------------------------------------------------------------
// exclusive scan of int32 (cell counters -> CSR offsets): per-block sums, one-block spine, apply
// ---------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;                       // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {   // blockDim.x == kScanThreads
  __shared__ int warp_sums[kScanThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < kScanThreads / 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, w, o);
      if (lane >= o) w += t;
    }
    if (lane < kScanThreads / 32) warp_sums[lane] = w;
  }
  __syncthreads();
  const int before = warp ? warp_sums[warp - 1] : 0;
  *total = warp_sums[kScanThreads / 32 - 1];
  __syncthreads();
  return before + inc - v;
}

__global__ void k_scan_tile_sums(const int32_t* __restrict__ in, int n, int32_t* __restrict__ tile_sums) {
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) s += (base + i < n) ? in[base + i] : 0;
  int total;
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void k_scan_spine(int32_t* __restrict__ tile_sums, int tiles) {   // one block
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < tiles; base += kScanThreads) {
    const int i = base + threadIdx.x;
    const int v = i < tiles ? tile_sums[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, &total);
    if (i < tiles) tile_sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
}

// out[i] = sum of in[0..i) for i < n; in and out may alias
__global__ void k_scan_apply(const int32_t* in, int n, const int32_t* __restrict__ tile_sums, int32_t* out) {
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) { v[i] = (base + i < n) ? in[base + i] : 0; s += v[i]; }
  int total;
  int run = tile_sums[blockIdx.x] + block_exclusive_scan(s, &total);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
}

// ---------------------------------------------------------------------------------------------------
// cell lists
// ---------------------------------------------------------------------------------------------------
// COO companion of a CSR list: entry -> owning patch
__global__ void k_entry_owner(int p0, int count, const int32_t* __restrict__ off, int32_t* __restrict__ owner) {
  const int p = p0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= p0 + count) return;
  for (int e = off[p]; e < off[p + 1]; ++e) owner[e] = p;
}

__global__ void k_cells_count(int tnum, int nentries, const int32_t* __restrict__ images, const int32_t* __restrict__ grids,
                              const int32_t* __restrict__ cell_base, const int32_t* __restrict__ gw, int32_t* __restrict__ count) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nentries) return;
  const int im = images[e];
  if (im >= tnum) return;   // only target images have grids (patchOrganizerS.cpp:318)
  atomicAdd(count + cell_base[im] + grids[2 * e + 1] * gw[im] + grids[2 * e], 1);
}

__global__ void k_cells_fill(int tnum, int nentries, const int32_t* __restrict__ images, const int32_t* __restrict__ grids,
                             const int32_t* __restrict__ owner, const int32_t* __restrict__ cell_base, const int32_t* __restrict__ gw,
                             const int32_t* __restrict__ cell_off, int32_t* __restrict__ cursor, int32_t* __restrict__ cell_patch) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nentries) return;
  const int im = images[e];
  if (im >= tnum) return;
  const int cell = cell_base[im] + grids[2 * e + 1] * gw[im] + grids[2 * e];
  cell_patch[cell_off[cell] + atomicAdd(cursor + cell, 1)] = owner[e];
}

// the fill order depends on the atomics: sort every (short) list so the table is the one a sequential addPatch
// loop in table order builds
__global__ void k_cells_sort(int cells, const int32_t* __restrict__ cell_off, int32_t* __restrict__ cell_patch) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cells) return;
  const int b = cell_off[c], e = cell_off[c + 1];
  for (int i = b + 1; i < e; ++i) {
    const int v = cell_patch[i];
    int j = i - 1;
    while (j >= b && cell_patch[j] > v) { cell_patch[j + 1] = cell_patch[j]; --j; }
    cell_patch[j + 1] = v;
  }
}

// ---------------------------------------------------------------------------------------------------
// setVImagesVGrids for the whole table, in place (two passes around a scan).  additive = keep the lists the table
// has and append the newly visible images (filter.cpp:764-772); otherwise start from empty lists (752-762).
// ---------------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void k_store_vimages(SceneDev s, StoreDev st, int additive, int32_t* __restrict__ count,
                                const int32_t* __restrict__ new_off, int32_t* __restrict__ new_vimages, int32_t* __restrict__ new_vgrids) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= st.P) return;
  const int p = warp;
  const float4 c4 = __ldg(reinterpret_cast<const float4*>(st.coords) + p);
  const float4 n4 = __ldg(reinterpret_cast<const float4*>(st.normals) + p);
  const float X[4] = {c4.x, c4.y, c4.z, c4.w}, N[4] = {n4.x, n4.y, n4.z, n4.w};
  const int e0 = st.img_off[p], e1 = st.img_off[p + 1];
  const int v0 = st.vimg_off[p], v1 = additive ? st.vimg_off[p + 1] : v0;
  int n = v1 - v0;
  const int out0 = FILL ? new_off[p] : 0;
  if (FILL)
    for (int e = v0 + lane; e < v1; e += 32) {
      new_vimages[out0 + e - v0] = st.vimages[e];
      new_vgrids[2 * (out0 + e - v0)] = st.vgrids[2 * e];
      new_vgrids[2 * (out0 + e - v0) + 1] = st.vgrids[2 * e + 1];
    }
  for (int base = 0; base < s.tnum; base += 32) {
    const int image = base + lane;
    bool ok = false;
    int ix = 0, iy = 0;
    if (image < s.tnum) {
      bool used = false;
      for (int e = e0; e < e1; ++e) used |= (st.images[e] == image);
      for (int e = v0; e < v1; ++e) used |= (st.vimages[e] == image);
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
    if (FILL && ok) {
      const int pos = out0 + n + __popc(m & ((1u << lane) - 1u));
      new_vimages[pos] = image;
      new_vgrids[2 * pos] = ix;
      new_vgrids[2 * pos + 1] = iy;
    }
    n += __popc(m);
  }
  if (!FILL && lane == 0) count[p] = n;
}

// ---------------------------------------------------------------------------------------------------
// neighbour search
// ---------------------------------------------------------------------------------------------------
struct PatchLists {      // _images/_grids and _vimages/_vgrids of the query patch: a table patch or a candidate outside the table
  const int32_t* images; const int32_t* grids; int nimg;
  const int32_t* vimages; const int32_t* vgrids; int nv;
};
__device__ __forceinline__ PatchLists table_lists(const StoreDev& st, int p) {
  PatchLists l;
  const int e0 = st.img_off[p], v0 = st.vimg_off[p];
  l.images = st.images + e0; l.grids = st.grids + 2 * e0; l.nimg = st.img_off[p + 1] - e0;
  l.vimages = st.vimages + v0; l.vgrids = st.vgrids + 2 * v0; l.nv = st.vimg_off[p + 1] - v0;
  return l;
}

struct NeighborQuery {   // what findNeighbors derives from the patch before it looks at any cell
  float X[4], N[4];
  float unit;     // mean getUnit over the patch's images x csize
  float radius;   // 1.5 x margin x computeRadius
  float base_radius;   // computeRadius (expand.cpp:182-198)
  float dscale;
  int nimg;
};

// getUnit summed in list order and COptim::computeUnits' second-smallest value (optim.cpp:446-471), lanes over images
__device__ __forceinline__ void neighbor_query(const SceneDev& s, const PatchLists& pl, float4 c4, float4 n4, float dscale, int margin, int lane,
                                               NeighborQuery& q) {
  q.X[0] = c4.x; q.X[1] = c4.y; q.X[2] = c4.z; q.X[3] = c4.w;
  q.N[0] = n4.x; q.N[1] = n4.y; q.N[2] = n4.z; q.N[3] = n4.w;
  q.dscale = dscale;
  const int e0 = 0, e1 = pl.nimg;
  q.nimg = e1 - e0;
  float usum = 0.0f, min1 = INFINITY, min2 = INFINITY;
  for (int base = e0; base < e1; base += 32) {
    float u = 0.0f, uu = 0.0f;
    if (base + lane < e1) {
      CamDev cam;
      load_cam(s, pl.images[base + lane], cam);
      u = get_unit(cam, s.level, q.X);
      float ray[4] = {cam.centre[0] - q.X[0], cam.centre[1] - q.X[1], cam.centre[2] - q.X[2], cam.centre[3] - q.X[3]};
      unitize4(ray);
      const float denom = dot4(ray, q.N);
      uu = 0.0f < denom ? fdiv(u, denom) : 1073741824.0f;   // (float)(INT_MAX / 2)
    }
    const int cnt = min(32, e1 - base);
    for (int i = 0; i < cnt; ++i) {   // list order: the reference's float sum
      usum += __shfl_sync(kFull, u, i);
      const float v = __shfl_sync(kFull, uu, i);
      if (v < min1) { min2 = min1; min1 = v; } else if (v < min2) min2 = v;
    }
  }
  q.unit = q.nimg > 0 ? fdiv(usum, (float)q.nimg) * (float)s.csize : 0.0f;
  q.base_radius = q.nimg >= 2 ? min2 * (float)s.csize : (q.nimg == 1 ? min1 * (float)s.csize : 0.0f);
  q.radius = 1.5f * (float)margin * q.base_radius;
}
__device__ __forceinline__ void neighbor_query(const SceneDev& s, const StoreDev& st, int p, int margin, int lane, NeighborQuery& q) {
  neighbor_query(s, table_lists(st, p), __ldg(reinterpret_cast<const float4*>(st.coords) + p), __ldg(reinterpret_cast<const float4*>(st.normals) + p),
                 st.dscale[p], margin, lane, q);
}

// CFindMatch::isNeighborRadius (findMatch.cpp:151-185)
__device__ __forceinline__ bool is_neighbor_radius(const StoreDev& st, const NeighborQuery& a, int b, float thr, float* Xb_out) {
  const float4 xb = __ldg(reinterpret_cast<const float4*>(st.coords) + b);
  const float4 nb = __ldg(reinterpret_cast<const float4*>(st.normals) + b);
  const float Xb[4] = {xb.x, xb.y, xb.z, xb.w}, Nb[4] = {nb.x, nb.y, nb.z, nb.w};
  Xb_out[0] = Xb[0]; Xb_out[1] = Xb[1]; Xb_out[2] = Xb[2]; Xb_out[3] = Xb[3];
  if (dot4(a.N, Nb) < st.cos120_f) return false;
  const float diff[4] = {Xb[0] - a.X[0], Xb[1] - a.X[1], Xb[2] - a.X[2], Xb[3] - a.X[3]};
  const float vunit = a.dscale + st.dscale[b];
  const float f0 = dot4(a.N, diff);
  const float f1 = dot4(Nb, diff);
  float ftmp = (fabsf(f0) + fabsf(f1)) / 2.0f;
  ftmp = fdiv(ftmp, vunit);
  float t[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) t[k] = diff[k] * 2.0f - a.N[k] * f0 - Nb[k] * f1;
  const float hsize = (float)((double)fsqrt(dot4(t, t)) / 2.0 / (double)a.unit);
  if (fdiv(a.radius, a.unit) < hsize) return false;
  if (1.0f < hsize) ftmp = fdiv(ftmp, smin(2.0f, hsize));
  return ftmp < thr;
}

// Walks the (2 margin + 1)^2 windows around the patch's cells in its target images (and, unless skipvis, in its
// visible-only images), hands every patch found in _pgrids / _vpgrids that passes isNeighborRadius to `hit`.
// A neighbour seen through several images is reported several times, as in the reference before its sort + unique.
template <typename Hit>
__device__ __forceinline__ void for_each_neighbor(const SceneDev& s, const StoreDev& st, const PatchLists& pl, const NeighborQuery& q, int margin,
                                                  bool skipvis, float thr, int lane, Hit hit) {
  const int ni = pl.nimg, nv = skipvis ? 0 : pl.nv;
  const int side = 2 * margin + 1, win = side * side;
  const int items = (ni + nv) * win;
  for (int it = lane; it < items; it += 32) {
    const int le = it / win, w = it - le * win;
    int image, ix, iy;
    if (le < ni) { image = pl.images[le]; ix = pl.grids[2 * le]; iy = pl.grids[2 * le + 1]; }
    else { const int v = le - ni; image = pl.vimages[v]; ix = pl.vgrids[2 * v]; iy = pl.vgrids[2 * v + 1]; }
    if (image >= s.tnum) continue;
    const int x = ix + (w % side) - margin, y = iy + (w / side) - margin;
    const int gw = st.gw[image], gh = st.gh[image];
    if (x < 0 || gw <= x || y < 0 || gh <= y) continue;
    const int cell = st.cell_base[image] + y * gw + x;
    float Xb[4];
    for (int j = st.cell_off[cell]; j < st.cell_off[cell + 1]; ++j) {
      const int b = st.cell_patch[j];
      if (is_neighbor_radius(st, q, b, thr, Xb)) hit(b, Xb);
    }
    for (int j = st.vcell_off[cell]; j < st.vcell_off[cell + 1]; ++j) {
      const int b = st.vcell_patch[j];
      if (is_neighbor_radius(st, q, b, thr, Xb)) hit(b, Xb);
    }
  }
}

// Vec4f ortho (include/numeric/vec4.hpp:303-322)
__device__ __forceinline__ void ortho4(const float* z, float* x, float* y) {
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

// CExpand::findEmptyBlocks (expand.cpp:108-180) for n table patches: bit i of mask = direction i already has a
// neighbour (fill[i] > 0); radius = computeRadius.  fill[] only ever receives non-negative terms, so "fill > 0" is an
// OR over the neighbours and needs neither the unique step nor an ordered sum.
__global__ void k_find_empty_blocks(SceneDev s, StoreDev st, int n, const int32_t* __restrict__ ids, uint8_t* __restrict__ mask,
                                    float* __restrict__ radius) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const int p = ids[warp];
  NeighborQuery q;
  neighbor_query(s, st, p, 1, lane, q);
  float xdir[4], ydir[4];
  ortho4(q.N, xdir, ydir);
  const float rlow = fdiv(q.base_radius, 6.0f), rhigh = q.base_radius * 2.5f;
  unsigned bits = 0;
  for_each_neighbor(s, st, table_lists(st, p), q, 1, false, 0.5f * 4.0f, lane, [&](int, const float* Xb) {
    const float diff[4] = {Xb[0] - q.X[0], Xb[1] - q.X[1], Xb[2] - q.X[2], Xb[3] - q.X[3]};
    float fx = dot4(diff, xdir), fy = dot4(diff, ydir);
    const float len = fsqrt(fx * fx + fy * fy);
    if (len < rlow || rhigh < len) return;
    fx = fdiv(fx, len); fy = fdiv(fy, len);
    // the reference's unqualified atan2 binds to the DOUBLE libm entry (nm -u expand.o: atan2), rounded to float on
    // assignment; a child that refinement left where findEmptyBlocks put it sits exactly on a sector boundary of its
    // parent, so the last ulp decides the sector.  CUDA's double atan2 (<= 2 ulp in double) rounds to the same float.
    float angle = (float)atan2((double)fy, (double)fx);
    if (angle < 0.0f) angle = (float)((double)angle + 2.0 * 3.14159265358979323846);
    const float findex = (float)((double)angle / (2.0 * 3.14159265358979323846 / 6.0));
    const int lindex = (int)floorf(findex), hindex = lindex + 1;
    bits |= 1u << (lindex % 6);                                   // hindex - findex > 0 always
    if ((float)lindex < findex) bits |= 1u << (hindex % 6);       // findex - lindex
  });
  bits = __reduce_or_sync(kFull, bits);
  if (lane == 0) { mask[warp] = (uint8_t)bits; radius[warp] = q.base_radius; }
}

// CFilter::filterNeighborThread + filterQuad (filter.cpp:357-462) for every table patch.
// Per warp: a hash set in shared memory makes the neighbours unique, a bitonic sort puts them in table order, the
// 5x5 normal equations of the quadric z = a x^2 + b y^2 + c xy + d x + e y are accumulated in double over the lanes
// and solved by lane 0 (the reference calls Eigen's jacobiSvd on the same system -- not available, parity unpinned).
constexpr int kNbSlots = 1024;        // hash slots per warp; the search window holds a few dozen unique neighbours
constexpr int kNbMax = 512;           // unique neighbours kept (beyond that the patch is accepted unfitted and counted)
constexpr int kNbWarps = 4;

struct NeighborSetSmem {              // per block: one hash set + one list per warp
  int32_t slots[kNbWarps][kNbSlots];
  int32_t list[kNbWarps][kNbMax];
  int32_t list_n[kNbWarps];
};

// findNeighbors(scale 4, `margin`, skipvis) of one patch into the warp's shared list, unique and in table order.
// Returns the number of unique neighbours (may exceed kNbMax: then *lost is set and the list is truncated).
__device__ __forceinline__ int collect_neighbors(const SceneDev& s, const StoreDev& st, const PatchLists& pl, const NeighborQuery& q, int margin,
                                                 bool skipvis, NeighborSetSmem& sm, int wib, int lane, bool* lost_out) {
  int32_t* tab = sm.slots[wib];
  int32_t* lst = sm.list[wib];
  for (int i = lane; i < kNbSlots; i += 32) tab[i] = -1;
  if (lane == 0) sm.list_n[wib] = 0;
  __syncwarp();
  bool lost = false;
  for_each_neighbor(s, st, pl, q, margin, skipvis, 0.5f * 4.0f, lane, [&](int b, const float*) {
    unsigned h = ((unsigned)b * 2654435761u) >> 22;   // 10 bits
    for (int probe = 0; probe < kNbSlots; ++probe) {
      const int old = atomicCAS(tab + h, -1, b);
      if (old == -1) {
        const int pos = atomicAdd(&sm.list_n[wib], 1);
        if (pos < kNbMax) lst[pos] = b; else lost = true;
        return;
      }
      if (old == b) return;
      h = (h + 1) & (kNbSlots - 1);
    }
    lost = true;
  });
  __syncwarp();
  *lost_out = __any_sync(kFull, lost);
  const int total = sm.list_n[wib];
  const int n = min(total, kNbMax);
  // table order (the reference sorts by address; any fixed order will do, this one is reproducible): bitonic sort
  int m = 1;
  while (m < n) m <<= 1;
  for (int i = n + lane; i < m; i += 32) lst[i] = INT32_MAX;
  __syncwarp();
  for (int k = 2; k <= m; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < m; i += 32) {
        const int l = i ^ j;
        if (l > i) {
          const int a = lst[i], b = lst[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { lst[i] = b; lst[l] = a; }
        }
      }
      __syncwarp();
    }
  return total;
}

// CFilter::filterQuad (filter.cpp:394-462) over the n neighbours in lst: residual of the quadric
// z = a x^2 + b y^2 + c xy + d x + e y fitted in the patch's tangent frame, in units of the patch's mean getUnit over
// its first tau images.  The 5x5 normal equations are accumulated in double over the lanes and solved by lane 0 (the
// reference calls Eigen's jacobiSvd on the same system -- not available, parity unpinned); h and the residual are
// summed in list order like the reference's float loops.
__device__ __forceinline__ float quad_residual(const SceneDev& s, const StoreDev& st, const PatchLists& pl, const NeighborQuery& q,
                                               const int32_t* lst, int n, int tau, int lane) {
  float xdir[4], ydir[4];
  ortho4(q.N, xdir, ydir);
  float hsum = 0.0f;
  for (int base = 0; base < n; base += 32) {
    float d = 0.0f;
    if (base + lane < n) {
      const float4 xb = __ldg(reinterpret_cast<const float4*>(st.coords) + lst[base + lane]);
      const float diff[4] = {xb.x - q.X[0], xb.y - q.X[1], xb.z - q.X[2], xb.w - q.X[3]};
      d = fsqrt(dot4(diff, diff));
    }
    const int cnt = min(32, n - base);
    for (int i = 0; i < cnt; ++i) hsum += __shfl_sync(kFull, d, i);
  }
  const float h = fdiv(hsum, (float)n);
  double acc[20];   // 15 distinct entries of A^T A and 5 of A^T b
#pragma unroll
  for (int k = 0; k < 20; ++k) acc[k] = 0.0;
  for (int i = lane; i < n; i += 32) {
    const float4 xb = __ldg(reinterpret_cast<const float4*>(st.coords) + lst[i]);
    const float diff[4] = {xb.x - q.X[0], xb.y - q.X[1], xb.z - q.X[2], xb.w - q.X[3]};
    const float fx = fdiv(dot4(diff, xdir), h), fy = fdiv(dot4(diff, ydir), h), fz = dot4(diff, q.N);
    const double row[5] = {(double)(fx * fx), (double)(fy * fy), (double)(fx * fy), (double)fx, (double)fy};
    int k = 0;
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
      for (int b = a; b < 5; ++b) acc[k++] += row[a] * row[b];
#pragma unroll
    for (int a = 0; a < 5; ++a) acc[15 + a] += row[a] * (double)fz;
  }
#pragma unroll
  for (int k = 0; k < 20; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(kFull, acc[k], o);
  float xs[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (lane == 0) {
    double M[5][6];
    int k = 0;
    for (int a = 0; a < 5; ++a)
      for (int b = a; b < 5; ++b) { M[a][b] = acc[k]; M[b][a] = acc[k]; ++k; }
    for (int a = 0; a < 5; ++a) M[a][5] = acc[15 + a];
    double x[5] = {0, 0, 0, 0, 0};
    bool singular = false;
    for (int c = 0; c < 5 && !singular; ++c) {
      int piv = c;
      for (int r = c + 1; r < 5; ++r) if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
      if (fabs(M[piv][c]) < 1e-300) { singular = true; break; }
      if (piv != c) for (int kk = 0; kk < 6; ++kk) { const double t = M[c][kk]; M[c][kk] = M[piv][kk]; M[piv][kk] = t; }
      for (int r = c + 1; r < 5; ++r) {
        const double f = M[r][c] / M[c][c];
        for (int kk = c; kk < 6; ++kk) M[r][kk] -= f * M[c][kk];
      }
    }
    if (!singular)
      for (int r = 4; r >= 0; --r) {
        double sacc = M[r][5];
        for (int kk = r + 1; kk < 5; ++kk) sacc -= M[r][kk] * x[kk];
        x[r] = sacc / M[r][r];
      }
    for (int a = 0; a < 5; ++a) xs[a] = (float)x[a];
  }
#pragma unroll
  for (int a = 0; a < 5; ++a) xs[a] = __shfl_sync(kFull, xs[a], 0);
  const int inum = min(tau, q.nimg);
  float unit = 0.0f;
  {
    float u = 0.0f;
    if (lane < inum) {
      CamDev cam;
      load_cam(s, pl.images[lane], cam);
      u = get_unit(cam, s.level, q.X);
    }
    for (int i = 0; i < inum; ++i) unit += __shfl_sync(kFull, u, i);
    unit = fdiv(unit, (float)inum);
  }
  float residual = 0.0f;
  for (int base = 0; base < n; base += 32) {
    float term = 0.0f;
    if (base + lane < n) {
      const float4 xb = __ldg(reinterpret_cast<const float4*>(st.coords) + lst[base + lane]);
      const float diff[4] = {xb.x - q.X[0], xb.y - q.X[1], xb.z - q.X[2], xb.w - q.X[3]};
      const float fx = fdiv(dot4(diff, xdir), h), fy = fdiv(dot4(diff, ydir), h), fz = dot4(diff, q.N);
      const float res = xs[0] * (fx * fx) + xs[1] * (fy * fy) + xs[2] * (fx * fy) + xs[3] * fx + xs[4] * fy - fz;
      term = fdiv(fabsf(res), unit);
    }
    const int cnt = min(32, n - base);
    for (int i = 0; i < cnt; ++i) residual += __shfl_sync(kFull, term, i);
  }
  return fdiv(residual, (float)(n - 5));
}

// CFilter::filterNeighborThread (filter.cpp:357-392) for every table patch: fewer than 6 neighbours from
// findNeighbors(scale 4, margin 2, skipvis 1), or a quadric residual >= quad, rejects the patch.
__global__ void __launch_bounds__(kNbWarps * 32)
k_filter_neighbor(SceneDev s, StoreDev st, float quad, int tau, uint8_t* __restrict__ reject, float* __restrict__ residual_out,
                  int32_t* __restrict__ ncount, int32_t* __restrict__ overflow) {
  __shared__ NeighborSetSmem sm;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x * kNbWarps + wib;
  if (p >= st.P) return;
  const PatchLists pl = table_lists(st, p);
  NeighborQuery q;
  neighbor_query(s, st, p, 2, lane, q);
  bool lost;
  const int total = collect_neighbors(s, st, pl, q, 2, true, sm, wib, lane, &lost);
  if (lane == 0) ncount[p] = total;
  if (lost) {
    if (lane == 0) { atomicAdd(overflow, 1); reject[p] = 0; residual_out[p] = -2.0f; }
    return;
  }
  if (total < 6) {
    if (lane == 0) { reject[p] = 1; residual_out[p] = -1.0f; }
    return;
  }
  const float residual = quad_residual(s, st, pl, q, sm.list[wib], total, tau, lane);
  if (lane == 0) { reject[p] = residual < quad ? 0 : 1; residual_out[p] = residual; }
}

// CFindMatch::isNeighbor (findMatch.cpp:120-149) between a patch outside the table (a) and table patch b
__device__ __forceinline__ int is_neighbor_ext(const SceneDev& s, const StoreDev& st, const float* Xa, const float* Na, float ua, float dscale_a,
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
  const float vunit = dscale_a + st.dscale[b];
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

// COptim::check (source/pmvs/optim.cpp:363-383) for a batch of candidates that are not in the table (postProcess calls
// it at _depth >= 2): gain = computeGain (filter.cpp:88-146) against the table's cells; reject when gain < 0, or when
// findNeighbors(scale 4, margin 2) finds more than 6 neighbours and the quadric residual is >= quad.
// One warp per candidate; lanes take one image entry each for the gain, then share the neighbour search.
__global__ void __launch_bounds__(kNbWarps * 32)
k_check_batch(SceneDev s, StoreDev st, int A, int stride, int vstride, const float* __restrict__ coords, const float* __restrict__ normals,
              const float* __restrict__ ncc, const float* __restrict__ dscale, const int32_t* __restrict__ timages,
              const int32_t* __restrict__ images, const int32_t* __restrict__ nimages, const int32_t* __restrict__ grids,
              const int32_t* __restrict__ vimages, const int32_t* __restrict__ nv, const int32_t* __restrict__ vgrids, float quad, int tau,
              float* __restrict__ gain_out, uint8_t* __restrict__ reject, int32_t* __restrict__ overflow) {
  __shared__ NeighborSetSmem sm;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * kNbWarps + wib;
  if (c >= A) return;
  PatchLists pl;
  pl.images = images + (size_t)c * stride; pl.grids = grids + (size_t)2 * c * stride; pl.nimg = min(nimages[c], stride);
  pl.vimages = vimages + (size_t)c * vstride; pl.vgrids = vgrids + (size_t)2 * c * vstride; pl.nv = min(nv[c], vstride);
  const float4 c4 = __ldg(reinterpret_cast<const float4*>(coords) + c);
  const float4 n4 = __ldg(reinterpret_cast<const float4*>(normals) + c);
  const float X[4] = {c4.x, c4.y, c4.z, c4.w}, N[4] = {n4.x, n4.y, n4.z, n4.w};
  const float thr = st.ncc_threshold, ds = dscale[c];
  // ---- gain
  float gain = smax(0.0f, ncc[c] - thr) * (float)timages[c];   // score2
  if (pl.nimg > 0) {
    CamDev cam;
    load_cam(s, pl.images[0], cam);
    const float ua = get_unit(cam, s.level, X);
    const int entries = pl.nimg + pl.nv;
    for (int base = 0; base < entries; base += 32) {
      const int e = base + lane;
      float maxpressure = 0.0f;
      if (e < entries) {
        const bool vis = e >= pl.nimg;
        const int index = vis ? pl.vimages[e - pl.nimg] : pl.images[e];
        if (index < s.tnum) {
          const int gx = vis ? pl.vgrids[2 * (e - pl.nimg)] : pl.grids[2 * e], gy = vis ? pl.vgrids[2 * (e - pl.nimg) + 1] : pl.grids[2 * e + 1];
          const int gwd = st.gw[index], ghd = st.gh[index];
          if (gx >= 0 && gx < gwd && gy >= 0 && gy < ghd) {
            const int cell = st.cell_base[index] + gy * gwd + gx;
            float pdepth = 0.0f;
            if (vis) { load_cam(s, index, cam); pdepth = dot4(cam.oaxis, X); }
            for (int j = st.cell_off[cell]; j < st.cell_off[cell + 1]; ++j) {
              const int qd = st.cell_patch[j];
              if (vis) {
                const float4 xq = __ldg(reinterpret_cast<const float4*>(st.coords) + qd);
                const float Xq[4] = {xq.x, xq.y, xq.z, xq.w};
                if (!(pdepth < dot4(cam.oaxis, Xq))) continue;
              }
              if (!is_neighbor_ext(s, st, X, N, ua, ds, qd, 1.0f)) maxpressure = smax(maxpressure, st.ncc[qd] - thr);
            }
          }
        }
      }
      const int cnt = min(32, entries - base);
      for (int i = 0; i < cnt; ++i) gain -= __shfl_sync(kFull, maxpressure, i);   // list order: the reference's float sum
    }
  }
  if (lane == 0) gain_out[c] = gain;
  if (gain < 0.0f) { if (lane == 0) reject[c] = 1; return; }
  // ---- neighbours and quadric fit
  NeighborQuery q;
  neighbor_query(s, pl, c4, n4, ds, 2, lane, q);
  bool lost;
  const int total = collect_neighbors(s, st, pl, q, 2, false, sm, wib, lane, &lost);
  if (lost) { if (lane == 0) { atomicAdd(overflow, 1); reject[c] = 0; } return; }
  int rej = 0;
  if (6 < total) rej = quad_residual(s, st, pl, q, sm.list[wib], total, tau, lane) < quad ? 0 : 1;
  if (lane == 0) reject[c] = (uint8_t)rej;
}

}  // namespace pmvsb
