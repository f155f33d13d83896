// cmvs-pmvs_b200/csrc/pmvs_table.cuh
//
// The resident patch table reorganised ON THE DEVICE, so that a filter round never marshals patches through the host:
//   * removal + renumbering in CPatchOrganizerS::collectPatches order (source/pmvs/patchOrganizerS.cpp:207-236): the
//     reference walks image by image, cell by cell, and numbers a patch at its FIRST appearance in _pgrids; a cell's
//     vector is in insertion (= creation) order, removals keep that order.  So the new number of a patch is its rank
//     under the key (first cell of the patch in the flattened cell arrays, creation sequence number): a counting sort
//     by first cell (count -> scan -> fill) and a per-bucket sort by sequence number;
//   * CFilter::filterExact's list surgery (source/pmvs/filter.cpp:240-280): image lists pruned by the visibility
//     re-test, surviving target images in ascending image order, then the non-target ones in their old order;
//   * the neighbour tests of CFilter::filterSmallGroups' labelling walk (filter.cpp:602-665) as adjacency lists.
// All paths relative to /root/reference.
#pragma once
#include "pmvs_cells.cuh"

namespace pmvsb {

// first _pgrids cell of every patch that stays (keep == null: all stay); a patch without a target image is in no cell and
// therefore not collected by the reference either
__global__ void k_tab_first_cell(StoreDev st, int tnum, const uint8_t* __restrict__ keep, int32_t* __restrict__ first_cell,
                                 int32_t* __restrict__ bucket_count) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= st.P) return;
  int best = 0x7fffffff;
  if (!keep || keep[p]) {
    for (int e = st.img_off[p]; e < st.img_off[p + 1]; ++e) {
      const int im = st.images[e];
      if (im >= tnum) continue;
      const int c = st.cell_base[im] + st.grids[2 * e + 1] * st.gw[im] + st.grids[2 * e];
      best = c < best ? c : best;
    }
  }
  if (best == 0x7fffffff) { first_cell[p] = -1; return; }
  first_cell[p] = best;
  atomicAdd(bucket_count + best, 1);
}

__global__ void k_tab_bucket_fill(int P, const int32_t* __restrict__ first_cell, const int32_t* __restrict__ bucket_off,
                                  int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int c = first_cell[p];
  if (c < 0) return;
  perm[bucket_off[c] + atomicAdd(cursor + c, 1)] = p;
}

// buckets are short (patches whose first cell is this cell): insertion sort by creation sequence number
__global__ void k_tab_bucket_sort(int cells, const int32_t* __restrict__ bucket_off, int32_t* __restrict__ perm,
                                  const int32_t* __restrict__ seq) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cells) return;
  const int b = bucket_off[c], e = bucket_off[c + 1];
  for (int i = b + 1; i < e; ++i) {
    const int v = perm[i], key = seq[v];
    int j = i - 1;
    while (j >= b && seq[perm[j]] > key) { perm[j + 1] = perm[j]; --j; }
    perm[j + 1] = v;
  }
}

// per-patch scalar fields through the permutation (new index k <- old index perm[k])
__global__ void k_tab_gather_fields(int newP, const int32_t* __restrict__ perm, const float4* __restrict__ coords, const float4* __restrict__ normals,
                                    const float* __restrict__ ncc, const float* __restrict__ dscale, const int32_t* __restrict__ timages,
                                    const int32_t* __restrict__ seq, float4* __restrict__ ncoords, float4* __restrict__ nnormals,
                                    float* __restrict__ nncc, float* __restrict__ ndscale, int32_t* __restrict__ ntimages, int32_t* __restrict__ nseq) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= newP) return;
  const int p = perm[k];
  ncoords[k] = coords[p]; nnormals[k] = normals[p]; nncc[k] = ncc[p]; ndscale[k] = dscale[p]; ntimages[k] = timages[p]; nseq[k] = seq[p];
}

// list lengths of the permuted table (slot newP = 0, the scan turns the array into offsets)
__global__ void k_tab_gather_len(int newP, const int32_t* __restrict__ perm, const int32_t* __restrict__ off, int32_t* __restrict__ len) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > newP) return;
  len[k] = k < newP ? off[perm[k] + 1] - off[perm[k]] : 0;
}

__global__ void k_tab_gather_lists(int newP, const int32_t* __restrict__ perm, const int32_t* __restrict__ off, const int32_t* __restrict__ items,
                                   const int32_t* __restrict__ cells, const int32_t* __restrict__ noff, int32_t* __restrict__ nitems,
                                   int32_t* __restrict__ ncells) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= newP) return;
  const int src = off[perm[warp]], dst = noff[warp], n = noff[warp + 1] - dst;
  for (int i = lane; i < n; i += 32) {
    nitems[dst + i] = items[src + i];
    ncells[2 * (dst + i)] = cells[2 * (src + i)];
    ncells[2 * (dst + i) + 1] = cells[2 * (src + i) + 1];
  }
}

// CFilter::filterExact (filter.cpp:240-272) for every table patch: target images whose visibility re-test passed, in
// ascending image order, then the non-target images in their old order; _timages = the number of the former.  A patch
// left with fewer than min_image_num images leaves the table (rows[p] gets length 0).  One thread per patch.
__global__ void k_tab_prune(StoreDev st, int tnum, int min_image_num, const uint8_t* __restrict__ safe, int stride, int32_t* __restrict__ rows,
                            int32_t* __restrict__ rows_n, int32_t* __restrict__ timages) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= st.P) return;
  int32_t* row = rows + (size_t)p * stride;
  int n = 0;
  for (int e = st.img_off[p]; e < st.img_off[p + 1]; ++e) {
    const int im = st.images[e];
    if (im >= tnum || !safe[e] || n >= stride) continue;
    int j = n - 1;
    while (j >= 0 && row[j] > im) { row[j + 1] = row[j]; --j; }
    row[j + 1] = im;
    ++n;
  }
  const int t = n;
  for (int e = st.img_off[p]; e < st.img_off[p + 1]; ++e) {
    const int im = st.images[e];
    if (im >= tnum && n < stride) row[n++] = im;
  }
  timages[p] = t;
  rows_n[p] = n < min_image_num ? 0 : n;
}

// stride-padded rows back into CSR lists
__global__ void k_tab_rows_to_lists(int P, int stride, const int32_t* __restrict__ rows, const int32_t* __restrict__ row_cells,
                                    const int32_t* __restrict__ off, int32_t* __restrict__ items, int32_t* __restrict__ cells) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int dst = off[p], n = off[p + 1] - dst;
  for (int i = 0; i < n; ++i) {
    items[dst + i] = rows[(size_t)p * stride + i];
    cells[2 * (dst + i)] = row_cells[((size_t)p * stride + i) * 2];
    cells[2 * (dst + i) + 1] = row_cells[((size_t)p * stride + i) * 2 + 1];
  }
}

__global__ void k_tab_flags_from_len(int P, const int32_t* __restrict__ len, uint8_t* __restrict__ keep) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P) keep[p] = len[p] > 0 ? 1 : 0;
}

// CFilter::filterSmallGroupsSub (filter.cpp:602-665) without the labels: for table patch p, the patches q of the 3x3 cells
// around p's cell in its reference image -- _pgrids list, then _vpgrids list of each cell, the reference's scan order -- with
// isNeighbor(p, q, thr).  FILL = false counts (count[p]), FILL = true writes adj[off[p] ..).  One thread per patch.
template <bool FILL>
__global__ void k_tab_group_edges(SceneDev s, StoreDev st, float thr, int32_t* __restrict__ count, const int32_t* __restrict__ off,
                                  int32_t* __restrict__ adj) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > st.P) return;
  if (p == st.P) { if (!FILL) count[p] = 0; return; }
  int n = 0;
  const int e0 = st.img_off[p];
  if (e0 < st.img_off[p + 1] && st.images[e0] < s.tnum) {
    const int index = st.images[e0], ix = st.grids[2 * e0], iy = st.grids[2 * e0 + 1];
    const int gw = st.gw[index], gh = st.gh[index];
    const float4 c4 = __ldg(reinterpret_cast<const float4*>(st.coords) + p);
    const float4 n4 = __ldg(reinterpret_cast<const float4*>(st.normals) + p);
    const float X[4] = {c4.x, c4.y, c4.z, c4.w}, N[4] = {n4.x, n4.y, n4.z, n4.w};
    CamDev cam;
    load_cam(s, index, cam);
    const float ua = get_unit(cam, s.level, X);
    int32_t* out = FILL ? adj + off[p] : nullptr;
    for (int y = -1; y <= 1; ++y) {
      const int yy = iy + y;
      if (yy < 0 || gh <= yy) continue;
      for (int x = -1; x <= 1; ++x) {
        const int xx = ix + x;
        if (xx < 0 || gw <= xx) continue;
        const int cell = st.cell_base[index] + yy * gw + xx;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          const int32_t* loff = pass == 0 ? st.cell_off : st.vcell_off;
          const int32_t* lst = pass == 0 ? st.cell_patch : st.vcell_patch;
          for (int j = loff[cell]; j < loff[cell + 1]; ++j) {
            const int q = lst[j];
            if (!is_neighbor(s, st, p, X, N, ua, q, thr)) continue;
            if (FILL) out[n] = q;
            ++n;
          }
        }
      }
    }
  }
  if (!FILL) count[p] = n;
}

// ---------------------------------------------------------------------------------------------------
// plumbing of the fused candidate evaluation (pmvsb_evaluate_batch): CSR <-> stride-padded rows, compaction
// ---------------------------------------------------------------------------------------------------
__global__ void k_rows_from_csr(int P, int stride, const int32_t* __restrict__ off, const int32_t* __restrict__ items, int32_t* __restrict__ rows,
                                int32_t* __restrict__ n) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int b = off[p], len = min(off[p + 1] - b, stride);
  for (int i = 0; i < len; ++i) rows[(size_t)p * stride + i] = items[b + i];
  n[p] = len;
}
// flag[i] = (verdict[i] == 0), slot P = 0 (the scan turns the array into positions, the last slot into the count)
__global__ void k_flag_zero(int P, const int32_t* __restrict__ verdict, int32_t* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > P) return;
  flag[i] = (i < P && verdict[i] == 0) ? 1 : 0;
}
// survivors of a stage move to the front: per-patch scalars and the stride-padded image rows
__global__ void k_compact_patches(int P, int stride, const int32_t* __restrict__ verdict, const int32_t* __restrict__ pos, const int32_t* __restrict__ src_index,
                                  const float4* __restrict__ coords, const float4* __restrict__ normals, const int32_t* __restrict__ rows,
                                  const int32_t* __restrict__ n, const float* __restrict__ a, const float* __restrict__ b, int32_t* __restrict__ out_index,
                                  float4* __restrict__ ocoords, float4* __restrict__ onormals, int32_t* __restrict__ orows, int32_t* __restrict__ on,
                                  float* __restrict__ oa, float* __restrict__ ob) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P || verdict[warp] != 0) return;
  const int j = pos[warp];
  if (lane == 0) {
    out_index[j] = src_index ? src_index[warp] : warp;
    ocoords[j] = coords[warp]; onormals[j] = normals[warp]; on[j] = n[warp];
    if (a) oa[j] = a[warp];
    if (b) ob[j] = b[warp];
  }
  const int len = n[warp];
  for (int i = lane; i < len; i += 32) orows[(size_t)j * stride + i] = rows[(size_t)warp * stride + i];
}
// same for the companions of an accepted candidate: cell rows (2 ints per image) and per-patch scalars
__global__ void k_compact_extras(int P, int stride, const int32_t* __restrict__ verdict, const int32_t* __restrict__ pos, const int32_t* __restrict__ n,
                                 const int32_t* __restrict__ cells, const float* __restrict__ ncc, const uint8_t* __restrict__ ok,
                                 const int32_t* __restrict__ timages, const float* __restrict__ tmp, int32_t* __restrict__ ocells, float* __restrict__ oncc,
                                 int32_t* __restrict__ otimages, float* __restrict__ otmp) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P || verdict[warp] != 0) return;
  const int j = pos[warp];
  if (lane == 0) { oncc[j] = ok[warp] ? ncc[warp] : -1.0f; otimages[j] = timages[warp]; otmp[j] = tmp[warp]; }   // a failed optimiser leaves _ncc untouched (optim.cpp:649-655)
  const int len = n[warp];
  for (int i = lane; i < 2 * len; i += 32) ocells[(size_t)j * 2 * stride + i] = cells[(size_t)warp * 2 * stride + i];
}
// verdicts of a later stage scattered back to the candidates they belong to (index[j] = candidate of compacted slot j)
__global__ void k_scatter_verdict(int n, const int32_t* __restrict__ index, const int32_t* __restrict__ v, const uint8_t* __restrict__ rej, int value,
                                  int32_t* __restrict__ verdict) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const bool bad = v ? v[j] != 0 : rej[j] != 0;
  if (bad) verdict[index[j]] = value;
}
__global__ void k_copy_len(int n, const int32_t* __restrict__ len, int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  out[i] = i < n ? len[i] : 0;
}
__global__ void k_merge_reject(int n, const uint8_t* __restrict__ rej, const float* __restrict__ gain, int32_t* __restrict__ verdict, float* __restrict__ tmp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  tmp[i] = gain[i];                       // COptim::check stores the gain in _tmp (optim.cpp:366)
  verdict[i] = rej[i] ? 1 : 0;
}

}  // namespace pmvsb
