// cmvs-pmvs_b200/host/pipeline.cpp -- wave-based seed / expand / filter rounds over the C ABI.
//
// Reference behaviour followed (paths relative to /root/reference/source/pmvs):
//   CFindMatch::run        findMatch.cpp:187-220   seed, then 3 x (expand, filter, thresholds -= 0.05)
//   CSeed                  seed.cpp:40-384         epipolar candidates, <= 2 successes per feature, 1 patch per cell
//   CExpand                expand.cpp:17-406       6-direction empty-block search, cell counters
//   CFilter                filter.cpp:13-665       outside / exact / neighbour (quadric) / small groups
//   CPatchOrganizerS       patchOrganizerS.cpp     cell lists, neighbours, writers
// Differences by construction: candidates of one wave are generated from one snapshot of the grids and are
// evaluated together; they are committed in priority order with the cell rules re-checked at commit time.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <list>
#include <numeric>
#include <queue>
#include <random>
#include <sstream>

#include "pmvs_host.hpp"

namespace pmvs {

namespace {
inline float dot4(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]; }
inline float norm4(const float* a) { return std::sqrt(dot4(a, a)); }
inline void unitize4(float* v) {
  const float l = dot4(v, v);
  if (l != 1.0f && l != 0.0f) { const float s = std::sqrt(l); v[0] /= s; v[1] /= s; v[2] /= s; v[3] /= s; }
}
// include/numeric/vec4.hpp:303-322
void ortho(const float* z, float* x, float* y) {
  x[0] = x[1] = x[2] = x[3] = 0.0f;
  y[0] = y[1] = y[2] = y[3] = 0.0f;
  if (std::fabs(z[0]) > 0.5f) { x[0] = z[1]; x[1] = -z[0]; x[2] = 0; }
  else if (std::fabs(z[1]) > 0.5f) { x[1] = z[2]; x[2] = -z[1]; x[0] = 0; }
  else { x[2] = z[0]; x[0] = -z[2]; x[1] = 0; }
  unitize4(x);
  y[0] = z[1] * x[2] - z[2] * x[1];
  y[1] = z[2] * x[0] - z[0] * x[2];
  y[2] = z[0] * x[1] - z[1] * x[0];
}
}  // namespace

// ---------------------------------------------------------------------------------------------- geometry
void Pipeline::project(int image, const float* X, float* o) const {   // include/image/camera.hpp:89-108
  const Camera& c = cams_[image];
  for (int i = 0; i < 3; ++i) o[i] = dot4(c.P[i], X);
  if (o[2] <= 0.0f) { o[0] = -65535.0f; o[1] = -65535.0f; o[2] = -1.0f; return; }
  const float z = o[2];
  o[0] /= z; o[1] /= z; o[2] /= z;
  o[0] = std::max((float)(INT_MIN + 3.0f), std::min((float)(INT_MAX - 3.0f), o[0]));
  o[1] = std::max((float)(INT_MIN + 3.0f), std::min((float)(INT_MAX - 3.0f), o[1]));
}

float Pipeline::get_unit(int image, const float* X) const {          // optim.cpp:1116-1124
  const Camera& c = cams_[image];
  const float d[4] = {X[0] - c.centre[0], X[1] - c.centre[1], X[2] - c.centre[2], X[3] - c.centre[3]};
  const float fz = norm4(d);
  if (c.ipscale == 0.0f) return 1.0f;
  return (float)(2.0 * fz * (1 << opt_.level) / c.ipscale);
}

int Pipeline::get_mask(int image, int ix, int iy) const {   // include/image/image.hpp:553-565
  const std::vector<unsigned char>& m = masks_[image];
  if (m.empty()) return 1;
  if (ix < 0 || lw_[image] <= ix || iy < 0 || lh_[image] <= iy) return 1;
  return m[(size_t)iy * lw_[image] + ix];
}

int Pipeline::get_mask(int image, const float* X) const {   // include/image/photo.hpp:44-49, image.hpp:540-551
  if (masks_[image].empty()) return 1;
  float ic[3];
  project(image, X, ic);
  return get_mask(image, (int)std::floor(ic[0] + 0.5f), (int)std::floor(ic[1] + 0.5f));
}

int Pipeline::get_edge(int image, const float* X) const {   // photo.hpp:51-59, image.hpp:567-592
  const std::vector<unsigned char>& m = edges_[image];
  if (m.empty()) return 1;
  float ic[3];
  project(image, X, ic);
  if (ic[0] < 0 || lw_[image] - 1 <= ic[0] || ic[1] < 0 || lh_[image] - 1 <= ic[1]) return 0;
  const int ix = (int)std::floor(ic[0] + 0.5f), iy = (int)std::floor(ic[1] + 0.5f);
  if (ix < 0 || lw_[image] <= ix || iy < 0 || lh_[image] <= iy) return 1;
  return m[(size_t)iy * lw_[image] + ix];
}

// `_pss.getMask(coord, _level) == 0 || insideBimages(coord) == 0` (expand.cpp:212, seed.cpp:314; include/image/photoSetS.hpp:109-116,
// findMatch.cpp:109-118): true = the point passes
bool Pipeline::mask_gate(const float* X) const {
  if (any_mask_)
    for (int i = 0; i < num_; ++i)
      if (get_mask(i, X) == 0) return false;
  for (int index : opt_.bindexes) {
    float ic[3];
    project(index, X, ic);
    if (ic[0] < 0.0 || lw_[index] - 1 < ic[0] || ic[1] < 0.0 || lh_[index] - 1 < ic[1]) return false;
  }
  return true;
}

bool Pipeline::is_neighbor(const Patch& l, const Patch& r, float hunit, float thr, float radius) const {   // findMatch.cpp:125-185
  static const double cos120 = std::cos(120.0 * M_PI / 180.0);   // same double as the reference's per-call expression
  if (dot4(l.normal, r.normal) < cos120) return false;
  float diff[4];
  for (int k = 0; k < 4; ++k) diff[k] = r.coord[k] - l.coord[k];
  const float vunit = l.dscale + r.dscale;
  const float f0 = dot4(l.normal, diff), f1 = dot4(r.normal, diff);
  float ftmp = (std::fabs(f0) + std::fabs(f1)) / 2.0f;
  ftmp /= vunit;
  float t[4];
  for (int k = 0; k < 4; ++k) t[k] = diff[k] * 2 - l.normal[k] * f0 - r.normal[k] * f1;
  const float hsize = (float)(norm4(t) / 2.0 / hunit);
  if (radius >= 0.0f && radius / hunit < hsize) return false;
  if (1.0f < hsize) ftmp /= std::min(2.0f, hsize);
  return ftmp < thr;
}

bool Pipeline::is_neighbor(const Patch& l, const Patch& r, float thr) const {   // findMatch.cpp:120-123
  const float hunit = (float)((get_unit(l.images[0], l.coord) + get_unit(r.images[0], r.coord)) / 2.0 * opt_.csize);
  return is_neighbor(l, r, hunit, thr, -1.0f);
}

// ---------------------------------------------------------------------------------------------- bookkeeping
int Pipeline::add_patch(Patch&& moved) {   // patchOrganizerS.cpp:312-349
  const int id = (int)patches_.size();
  patches_.push_back(std::move(moved));   // the four lists change owner instead of being copied
  const Patch& p = patches_.back();
  patches_.back().alive = true;
  // the host keeps the OCCUPANCY of _pgrids (what the cell rules read); the lists themselves (_pgrids, _vpgrids) live
  // on the device (pmvsb_store_append / pmvsb_store_rebuild)
  for (size_t i = 0; i < p.images.size(); ++i) {
    const int im = p.images[i];
    if (tnum_ <= im) continue;
    ++grids_[im].occ[(size_t)p.grids[i][1] * grids_[im].gw + p.grids[i][0]];
  }
  return id;
}

// ids of the live patches in creation order
std::vector<int> Pipeline::live_patches() const {
  std::vector<int> ids;
  for (int id = 0; id < (int)patches_.size(); ++id)
    if (patches_[id].alive) ids.push_back(id);
  return ids;
}

// CPatch fields of a list of patches as the arrays pmvsb_store_upload / pmvsb_store_append take
struct Pipeline::TableArrays {
  std::vector<float> coords, normals, ncc, dsc;
  std::vector<int32_t> ioff, voff, images, grids, vimages, vgrids, timages;
};

void Pipeline::marshal(const std::vector<int>& ids, TableArrays& t) const {
  const int P = (int)ids.size();
  t.coords.resize((size_t)4 * P); t.normals.resize((size_t)4 * P); t.ncc.resize(P); t.dsc.resize(P); t.timages.resize(P);
  t.ioff.assign(P + 1, 0); t.voff.assign(P + 1, 0);
  for (int k = 0; k < P; ++k) {
    t.ioff[k + 1] = t.ioff[k] + (int32_t)patches_[ids[k]].images.size();
    t.voff[k + 1] = t.voff[k] + (int32_t)patches_[ids[k]].vimages.size();
  }
  t.images.resize(std::max(1, t.ioff[P])); t.grids.resize((size_t)2 * std::max(1, t.ioff[P]));
  t.vimages.resize(std::max(1, t.voff[P])); t.vgrids.resize((size_t)2 * std::max(1, t.voff[P]));
  parallel_for(P, threads_, [&](int k) {
    const Patch& p = patches_[ids[k]];
    for (int c = 0; c < 4; ++c) { t.coords[4 * k + c] = p.coord[c]; t.normals[4 * k + c] = p.normal[c]; }
    t.ncc[k] = p.ncc; t.dsc[k] = p.dscale; t.timages[k] = p.timages;
    int e = t.ioff[k];
    for (size_t i = 0; i < p.images.size(); ++i, ++e) { t.images[e] = p.images[i]; t.grids[2 * e] = p.grids[i][0]; t.grids[2 * e + 1] = p.grids[i][1]; }
    e = t.voff[k];
    for (size_t i = 0; i < p.vimages.size(); ++i, ++e) { t.vimages[e] = p.vimages[i]; t.vgrids[2 * e] = p.vgrids[i][0]; t.vgrids[2 * e + 1] = p.vgrids[i][1]; }
  }, 1024);
}

// The live patches go to the GPU once (after the seed round); from then on the table lives there: expansion appends to it
// (append_table), the filter rounds reorganise it in place (device_rebuild) and the host mirrors what it needs (sync_table).
void Pipeline::upload_table(const std::vector<int>& ids) {
  Tick tk(this, "gpu.upload_table+depth_maps");
  TableArrays t;
  marshal(ids, t);
  if (pmvsb_set_thresholds(gpu_, ncc_threshold_, ncc_threshold_before_)) die("set_thresholds");
  if (pmvsb_set_depth(gpu_, depth_)) die("set_depth");
  if (pmvsb_store_upload(gpu_, (int)ids.size(), t.coords.data(), t.normals.data(), t.ncc.data(), t.dsc.data(), t.ioff.data(), t.images.data(),
                         t.grids.data(), t.voff.data(), t.vimages.data(), t.vgrids.data(), t.timages.data())) die("store_upload");
  if (pmvsb_store_set_seq(gpu_, 0, (int)ids.size(), ids.data())) die("store_set_seq");   // creation order = patch id
  table_ids_ = ids;
  device_rebuild(2, nullptr);   // collectPatches numbering + depth maps; _vimages as they are
}

// CPatchOrganizerS::removePatch for the table patches with keep == 0 (null: none), collectPatches renumbering and
// CFilter::setDepthMapsVGridsVPGridsAddPatchV(additive) (filter.cpp:734-783), all on the device; mode 2 = renumbering and
// depth maps only.  The host follows with the permutation.
void Pipeline::device_rebuild(int mode, const std::vector<uint8_t>* keep) {
  Tick tk(this, "gpu.table_rebuild");
  if (pmvsb_set_thresholds(gpu_, ncc_threshold_, ncc_threshold_before_)) die("set_thresholds");
  if (pmvsb_set_depth(gpu_, depth_)) die("set_depth");
  std::vector<int32_t> perm(std::max<size_t>(table_ids_.size(), 1));
  int32_t n = 0;
  if (pmvsb_store_rebuild(gpu_, keep ? keep->data() : nullptr, mode, &n, perm.data())) die("store_rebuild");
  std::vector<int> ids(n);
  for (int k = 0; k < n; ++k) ids[k] = table_ids_[perm[k]];
  table_ids_.swap(ids);
  table_index_.assign(patches_.size(), -1);
  for (int k = 0; k < n; ++k) table_index_[table_ids_[k]] = k;
}

// the table's lists back into the host patches (what the next expansion round and the writers read), the patches that left
// the table marked dead, the occupancy of _pgrids recounted
void Pipeline::sync_table(bool full) {
  Tick tk(this, "host.sync_table");
  int32_t P = 0, E = 0, VE = 0;
  if (pmvsb_store_counts(gpu_, &P, &E, &VE)) die("store_counts");
  if (P != (int)table_ids_.size()) { std::cerr << "sync_table: table size mismatch" << std::endl; std::exit(1); }
  // staging reused from call to call (a fresh vector of this size is zero-filled and page-faulted in every time)
  static std::vector<int32_t> seq, ti, off, im, gr, voff, vim, vgr;
  auto room = [](std::vector<int32_t>& v, size_t n) { if (v.size() < n) v.resize(n + n / 4); };
  room(seq, (size_t)std::max(P, 1)); room(ti, (size_t)std::max(P, 1)); room(off, (size_t)P + 1); room(im, (size_t)std::max(E, 1)); room(gr, (size_t)2 * std::max(E, 1));
  { Tick tk2(this, "host.sync_table.download");
  if (pmvsb_store_download_lists(gpu_, seq.data(), ti.data(), off.data(), im.data(), gr.data())) die("store_download_lists");
  // between rounds the host only needs the image lists (parents of the next expansion) and the occupancy of _pgrids; the
  // per-patch cells and visible-image lists are read by the writers only (full = true, once, from write())
  if (full) {
    room(voff, (size_t)P + 1); room(vim, (size_t)std::max(VE, 1)); room(vgr, (size_t)2 * std::max(VE, 1));
    if (pmvsb_store_download_vimages(gpu_, voff.data(), vim.data(), vgr.data())) die("store_download_vimages");
  } }
  { Tick tk2(this, "host.sync_table.lists");
  for (Patch& p : patches_) p.alive = false;
  parallel_for(P, threads_, [&](int k) {
    if (seq[k] != table_ids_[k]) { std::cerr << "sync_table: the device and host numberings disagree" << std::endl; std::exit(1); }
    Patch& p = patches_[table_ids_[k]];
    p.alive = true;
    p.timages = ti[k];
    p.images.assign(im.begin() + off[k], im.begin() + off[k + 1]);
    if (!full) return;
    const int n = off[k + 1] - off[k], nv = voff[k + 1] - voff[k];
    p.grids.resize(n);
    for (int i = 0; i < n; ++i) p.grids[i] = {gr[(size_t)2 * (off[k] + i)], gr[(size_t)2 * (off[k] + i) + 1]};
    p.vimages.assign(vim.begin() + voff[k], vim.begin() + voff[k + 1]);
    p.vgrids.resize(nv);
    for (int i = 0; i < nv; ++i) p.vgrids[i] = {vgr[(size_t)2 * (voff[k] + i)], vgr[(size_t)2 * (voff[k] + i) + 1]};
  }, 1024);
  }
  Tick tk3(this, "host.sync_table.occupancy");
  parallel_for(tnum_, threads_, [&](int t) { std::fill(grids_[t].occ.begin(), grids_[t].occ.end(), 0); }, 1);
  // one relaxed atomic increment per (patch, image) entry: entries of one cell are far apart in the list, so the threads rarely meet
  parallel_for(E, threads_, [&](int e) {
    if (im[e] < tnum_) __atomic_fetch_add(&grids_[im[e]].occ[(size_t)gr[(size_t)2 * e + 1] * grids_[im[e]].gw + gr[(size_t)2 * e]], 1, __ATOMIC_RELAXED);
  }, 1 << 16);
}

// CPatchOrganizerS::addPatch + updateDepthMaps for patches committed since the table was uploaded
void Pipeline::append_table(const std::vector<int>& ids) {
  if (ids.empty()) return;
  Tick tk(this, "gpu.append_table");
  table_index_.resize(patches_.size(), -1);
  const int first = (int)table_ids_.size();
  for (int id : ids) { table_index_[id] = (int)table_ids_.size(); table_ids_.push_back(id); }
  static TableArrays t;   // reused from wave to wave (marshal sizes every array)
  { Tick tk2(this, "gpu.append_table.marshal"); marshal(ids, t); }
  { Tick tk2(this, "gpu.append_table.store_append");
  if (pmvsb_store_append(gpu_, (int)ids.size(), t.coords.data(), t.normals.data(), t.ncc.data(), t.dsc.data(), t.ioff.data(), t.images.data(),
                         t.grids.data(), t.voff.data(), t.vimages.data(), t.vgrids.data(), t.timages.data())) die("store_append"); }
  if (pmvsb_store_set_seq(gpu_, first, (int)ids.size(), ids.data())) die("store_set_seq");
}

// ---------------------------------------------------------------------------------------------- evaluate a wave
void Pipeline::evaluate_range(std::vector<Candidate>& all, std::vector<int>& all_verdict, int lo, int hi, bool gather) {
  const int P = hi - lo;
  if (P <= 0 && !gather) return;
  Candidate* cands = all.data() + lo;      // this rank's shard of the wave
  int* verdict = all_verdict.data() + lo;
  // the reference's per-candidate contract -- preProcess, refinePatch, postProcess (seed.cpp:397-409, expand.cpp:225-237) -- as
  // ONE library call for the whole shard: the stages chain on the device, only the accepted candidates' records come back
  // scratch reused across waves (the pipeline thread is the only caller; the worker threads of parallel_for see the same objects)
  static std::vector<float> coords, normals, acoords, anormals, ascal;
  static std::vector<int32_t> ioff, images, v, aindex, ati, aoff, aim, agr, avoff, avim, avgr;
  coords.resize((size_t)4 * P); normals.resize((size_t)4 * P); ioff.resize((size_t)P + 1); v.resize(P);
  ioff[0] = 0;
  for (int k = 0; k < P; ++k) ioff[k + 1] = ioff[k] + (int32_t)cands[k].patch.images.size();
  images.resize(std::max(1, ioff[P]));
  parallel_for(P, threads_, [&](int k) {
    const Patch& p = cands[k].patch;
    for (int c = 0; c < 4; ++c) { coords[4 * k + c] = p.coord[c]; normals[4 * k + c] = p.normal[c]; }
    std::copy(p.images.begin(), p.images.end(), images.begin() + ioff[k]);
  }, 1024);
  if (pmvsb_set_thresholds(gpu_, ncc_threshold_, ncc_threshold_before_)) die("set_thresholds");
  if (pmvsb_set_depth(gpu_, depth_)) die("set_depth");
  int32_t A = 0, E = 0, VE = 0, refined = 0;
  { Tick tk(this, "gpu.evaluate");
    if (pmvsb_evaluate_batch(gpu_, P, coords.data(), normals.data(), ioff.data(), images.data(), opt_.quad, &A, &E, &VE, &refined)) die("evaluate_batch"); }
  int nv = P;   // verdicts that come back
  if (gather) {
    // multi-GPU: every rank's accepted records all-gathered device to device (NCCL); from here on the whole wave is local
    Tick tk(this, "gpu.allgather_wave");
    int rc = pmvsb_evaluate_allgather(gpu_, lo, (int)all.size());
    if (rc == PMVSB_EGROW) {
      // the wave's message is larger than the mailbox slots (every rank sees the same sizes and arrives here): all ranks unmap,
      // meet, re-export larger mailboxes, map again and repeat the exchange
      const size_t need = pmvsb_peer_needed(gpu_);
      if (pmvsb_peer_close(gpu_)) die("peer_close");
      char token = 0;
      std::vector<char> tokens(dist_.world);
      dist_.allgather(&token, 1, tokens.data());
      if (!peer_bringup(need + need / 2)) die("peer mailboxes (re-export)");
      if (is_root()) std::cerr << "peer mailboxes re-exported with " << need + need / 2 << " bytes per slot" << std::endl;
      rc = pmvsb_evaluate_allgather(gpu_, lo, (int)all.size());
    }
    if (rc) die("evaluate_allgather");
    if (pmvsb_evaluate_counts(gpu_, &nv, &A, &E, &VE)) die("evaluate_counts");
    cands = all.data(); verdict = all_verdict.data();
    v.resize(nv);
  }
  aindex.resize(std::max(A, 1)); acoords.resize((size_t)4 * std::max(A, 1)); anormals.resize((size_t)4 * std::max(A, 1)); ascal.resize((size_t)4 * std::max(A, 1));
  ati.resize(std::max(A, 1)); aoff.resize((size_t)A + 1); aim.resize(std::max(E, 1)); agr.resize((size_t)2 * std::max(E, 1)); avoff.resize((size_t)A + 1);
  avim.resize(std::max(VE, 1)); avgr.resize((size_t)2 * std::max(VE, 1));
  if (pmvsb_evaluate_fetch(gpu_, v.data(), aindex.data(), acoords.data(), anormals.data(), ascal.data(), ati.data(), aoff.data(), aim.data(), agr.data(),
                           avoff.data(), avim.data(), avgr.data())) die("evaluate_fetch");
  for (int k = 0; k < nv; ++k) verdict[k] = v[k];
  parallel_for(A, threads_, [&](int j) {
    Patch& p = cands[aindex[j]].patch;
    for (int c = 0; c < 4; ++c) { p.coord[c] = acoords[4 * j + c]; p.normal[c] = anormals[4 * j + c]; }
    p.ncc = ascal[4 * j]; p.dscale = ascal[4 * j + 1]; p.ascale = ascal[4 * j + 2]; p.tmp = ascal[4 * j + 3];
    p.timages = ati[j];
    const int n = aoff[j + 1] - aoff[j], nv = avoff[j + 1] - avoff[j];
    p.images.assign(aim.begin() + aoff[j], aim.begin() + aoff[j + 1]);
    p.grids.resize(n);
    for (int i = 0; i < n; ++i) p.grids[i] = {agr[(size_t)2 * (aoff[j] + i)], agr[(size_t)2 * (aoff[j] + i) + 1]};
    p.vimages.assign(avim.begin() + avoff[j], avim.begin() + avoff[j + 1]);
    p.vgrids.resize(nv);
    for (int i = 0; i < nv; ++i) p.vgrids[i] = {avgr[(size_t)2 * (avoff[j] + i)], avgr[(size_t)2 * (avoff[j] + i) + 1]};
  }, 512);
}

// pre -> refine -> post (+ vimages at depth >= 1) for a wave.  With several GPUs every rank evaluates a contiguous shard
// and the results are all-gathered, so that all ranks commit the same wave.
void Pipeline::evaluate(std::vector<Candidate>& cands, std::vector<int>& verdict) {
  Tick tk_all(this, "evaluate.total");
  const int P = (int)cands.size();
  verdict.assign(P, 1);
  if (P == 0) return;
  // a small wave is evaluated whole on every rank: the evaluation is deterministic, so all ranks hold the same results without
  // any exchange, and an exchange would cost more than the split saves (PMVSB_SHARD_MIN)
  if (dist_.world == 1 || P < dist_.shard_min) { evaluate_range(cands, verdict, 0, P, false); return; }
  int lo = 0, hi = P;
  Dist::shard(P, dist_.world, dist_.rank, lo, hi);
  const bool device = !dist_.tcp_exchange;   // peer mailboxes or NCCL: the exchange happens between the GPUs' memories
  evaluate_range(cands, verdict, lo, hi, device);
  if (!device) exchange_results(cands, verdict);   // PMVSB_EXCHANGE=tcp: host-side exchange over the rendezvous sockets
}

// One message per rank and wave: the verdicts of its shard, then ONLY the accepted candidates' records (header, the
// CPatch scalars post-processing produced, image and visible-image lists at their real lengths).  The sizes travel over
// the rendezvous sockets first (that is also where a rank that died is noticed); the messages, padded to the longest, go
// through one NCCL all-gather (or the sockets with PMVSB_EXCHANGE=tcp).
void Pipeline::exchange_results(std::vector<Candidate>& cands, std::vector<int>& verdict) {
  Tick tk(this, "gpu.allgather_wave");
  const int P = (int)cands.size(), W = dist_.world;
  int lo = 0, hi = 0;
  Dist::shard(P, W, dist_.rank, lo, hi);
  std::vector<int32_t> msg;
  msg.reserve((size_t)(hi - lo) * 64);
  for (int k = lo; k < hi; ++k) msg.push_back(verdict[k]);
  for (int k = lo; k < hi; ++k) {
    if (verdict[k] != 0) continue;
    const Patch& p = cands[k].patch;
    const int32_t ni = (int32_t)p.images.size(), nv = (int32_t)p.vimages.size();
    msg.push_back(ni); msg.push_back(nv); msg.push_back(p.timages);
    const float f[12] = {p.coord[0], p.coord[1], p.coord[2], p.coord[3], p.normal[0], p.normal[1], p.normal[2], p.normal[3], p.ncc, p.dscale, p.ascale, p.tmp};
    const size_t at = msg.size();
    msg.resize(at + 12);
    std::memcpy(msg.data() + at, f, sizeof(f));
    for (int i = 0; i < ni; ++i) { msg.push_back(p.images[i]); msg.push_back(p.grids[i][0]); msg.push_back(p.grids[i][1]); }
    for (int i = 0; i < nv; ++i) { msg.push_back(p.vimages[i]); msg.push_back(p.vgrids[i][0]); msg.push_back(p.vgrids[i][1]); }
  }
  std::vector<int64_t> sizes(W, 0);
  const int64_t mine = (int64_t)msg.size();
  dist_.allgather(&mine, sizeof(mine), sizes.data());
  const size_t longest = (size_t)*std::max_element(sizes.begin(), sizes.end());
  if (longest == 0) return;
  msg.resize(longest, 0);
  std::vector<int32_t> all(longest * W);
  if (dist_.tcp_exchange) dist_.allgather(msg.data(), longest * sizeof(int32_t), all.data());
  else if (pmvsb_allgather(gpu_, msg.data(), longest * sizeof(int32_t), all.data())) die("allgather");
  exchanged_bytes_ += (double)longest * sizeof(int32_t) * W;
  for (int rk = 0; rk < W; ++rk) {
    if (rk == dist_.rank) continue;
    int rlo = 0, rhi = 0;
    Dist::shard(P, W, rk, rlo, rhi);
    const int32_t* r = all.data() + (size_t)rk * longest;
    const int32_t* q = r + (rhi - rlo);
    for (int k = rlo; k < rhi; ++k) {
      verdict[k] = r[k - rlo];
      if (verdict[k] != 0) continue;
      Patch& p = cands[k].patch;
      const int ni = q[0], nv = q[1];
      p.timages = q[2];
      float f[12];
      std::memcpy(f, q + 3, sizeof(f));
      for (int c = 0; c < 4; ++c) { p.coord[c] = f[c]; p.normal[c] = f[4 + c]; }
      p.ncc = f[8]; p.dscale = f[9]; p.ascale = f[10]; p.tmp = f[11];
      q += 15;
      p.images.resize(ni); p.grids.resize(ni);
      for (int i = 0; i < ni; ++i, q += 3) { p.images[i] = q[0]; p.grids[i] = {q[1], q[2]}; }
      p.vimages.resize(nv); p.vgrids.resize(nv);
      for (int i = 0; i < nv; ++i, q += 3) { p.vimages[i] = q[0]; p.vgrids[i] = {q[1], q[2]}; }
    }
  }
}

// ---------------------------------------------------------------------------------------------- seed round
namespace {
// fundamental matrix between two 3x4 projections (rows as double[4]); F[a][b] = det of the two rows of P0 other than a
// stacked on the two rows of P1 other than b (include/image/camera.hpp:129-151)
double det4(const double* a, const double* b, const double* c, const double* d) {
  const double m[4][4] = {{a[0], a[1], a[2], a[3]}, {b[0], b[1], b[2], b[3]}, {c[0], c[1], c[2], c[3]}, {d[0], d[1], d[2], d[3]}};
  double det = 0.0;
  for (int j = 0; j < 4; ++j) {
    double sub[3][3];
    for (int r = 1; r < 4; ++r) { int cc = 0; for (int k = 0; k < 4; ++k) if (k != j) sub[r - 1][cc++] = m[r][k]; }
    const double d3 = sub[0][0] * (sub[1][1] * sub[2][2] - sub[1][2] * sub[2][1]) - sub[0][1] * (sub[1][0] * sub[2][2] - sub[1][2] * sub[2][0]) +
                      sub[0][2] * (sub[1][0] * sub[2][1] - sub[1][1] * sub[2][0]);
    det += ((j % 2) ? -1.0 : 1.0) * m[0][j] * d3;
  }
  return det;
}
void fundamental(const std::vector<double>& P0, const std::vector<double>& P1, double F[3][3]) {
  const double* p0[3] = {&P0[0], &P0[4], &P0[8]};
  const double* p1[3] = {&P1[0], &P1[4], &P1[8]};
  const int o[3][2] = {{1, 2}, {2, 0}, {0, 1}};
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) F[a][b] = det4(p0[o[a][0]], p0[o[a][1]], p1[o[b][0]], p1[o[b][1]]);
}
}  // namespace

void Pipeline::seed_round() {
  Tick tk(this, "round.seed");
  Stats st;
  std::vector<int> order(tnum_);
  std::iota(order.begin(), order.end(), 0);
  std::mt19937 gen(42);   // seed.cpp:38
  std::shuffle(order.begin(), order.end(), gen);
  std::cerr << "adding seeds " << std::endl;
  for (auto& g : grids_) std::fill(g.counts.begin(), g.counts.end(), 0);
  for (int im = 0; im < tnum_; ++im)
    for (size_t c = 0; c < grids_[im].occ.size(); ++c)
      if (grids_[im].occ[c] != 0) grids_[im].counts[c] = (unsigned char)count_threshold2_;
  auto can_add = [&](int image, int x, int y) {   // seed.cpp:325-338
    if (!get_mask(image, opt_.csize * x, opt_.csize * y)) return false;
    if (tnum_ <= image) return true;
    const size_t c = (size_t)y * grids_[image].gw + x;
    if (grids_[image].occ[c] != 0) return false;
    return !(count_threshold2_ <= grids_[image].counts[c]);
  };
  // CSeed::canAdd of every cell of every image, kept current as the round commits (the candidate kernel reads it)
  std::vector<int> cell_base(num_ + 1, 0);
  for (int i = 0; i < num_; ++i) cell_base[i + 1] = cell_base[i] + grids_[i].gw * grids_[i].gh;
  std::vector<uint8_t> blocked((size_t)cell_base[num_]);
  parallel_for(num_, threads_, [&](int i) {
    for (int y = 0; y < grids_[i].gh; ++y)
      for (int x = 0; x < grids_[i].gw; ++x) blocked[(size_t)cell_base[i] + (size_t)y * grids_[i].gw + x] = can_add(i, x, y) ? 0 : 1;
  }, 1);
  auto refresh = [&](int image, int cell) { blocked[(size_t)cell_base[image] + cell] = can_add(image, cell % grids_[image].gw, cell / grids_[image].gw) ? 0 : 1; };
  std::vector<int32_t> sc_ref, sc_other(1 << 18), sc_other_feat(1 << 18);
  std::vector<float> sc_coord((size_t)4 << 18), sc_resp(1 << 18);
  const double cos_a0 = std::cos(60.0f * M_PI / 180.0f);
  for (int index : order) {
    // COptim::collectImages (optim.cpp:66-93)
    std::vector<std::pair<float, int>> cand_im;
    for (int j : opt_.visdata2[index]) {
      if (opt_.sequence != -1 && opt_.sequence < std::abs(index - j)) continue;
      const float d = cams_[index].oaxis[0] * cams_[j].oaxis[0] + cams_[index].oaxis[1] * cams_[j].oaxis[1] + cams_[index].oaxis[2] * cams_[j].oaxis[2];
      if (d < cos_a0) continue;
      cand_im.push_back({distances_[index][j], j});
    }
    std::sort(cand_im.begin(), cand_im.end());
    std::vector<int> indexes;
    for (int i = 0; i < std::min(tau_, (int)cand_im.size()); ++i) indexes.push_back(cand_im[i].second);
    if (indexes.empty()) continue;
    // ---- enumerate the wave on the GPU: every (cell, feature, epipolar candidate) of this image from the current snapshot
    // (CSeed::collectCells / collectCandidates / unproject, seed.cpp:207-384: pmvsb_seed_candidates)
    std::vector<Candidate> wave;
    const ImageGrid& g = grids_[index];
    { Tick tk2(this, "gpu.seed.candidates");
      int32_t nref = 0, total = 0;
      if ((int)sc_ref.size() < 4 * 65536) sc_ref.resize(4 * 65536);
      for (int attempt = 0; attempt < 3; ++attempt) {
        const int cap_ref = (int)sc_ref.size() / 4, cap = (int)sc_other.size();
        Tick tk3(this, "gpu.seed.candidates.call");
        if (pmvsb_seed_candidates(gpu_, index, (int)indexes.size(), indexes.data(), blocked.data(), cap_ref, &nref, sc_ref.data(), sc_ref.data() + cap_ref,
                                  sc_ref.data() + 2 * cap_ref, sc_ref.data() + 3 * cap_ref, cap, &total, sc_coord.data(), sc_other.data(),
                                  sc_other_feat.data(), sc_resp.data())) die("seed_candidates");
        if (nref <= cap_ref && total <= cap) break;
        if (nref > cap_ref) sc_ref.resize((size_t)4 * nref);
        if (total > cap) { sc_coord.resize((size_t)4 * total); sc_other.resize(total); sc_other_feat.resize(total); sc_resp.resize(total); }
      }
      const int cap_ref = (int)sc_ref.size() / 4;
      const int32_t *rfeat = sc_ref.data(), *rcell = rfeat + cap_ref, *rstart = rcell + cap_ref, *rcount = rstart + cap_ref;
      size_t n = 0;
      for (int r = 0; r < nref; ++r) n += (size_t)rcount[r];
      wave.resize(n);
      std::vector<size_t> at((size_t)nref + 1, 0);
      for (int r = 0; r < nref; ++r) at[r + 1] = at[r] + (size_t)rcount[r];
      parallel_for(nref, threads_, [&](int r) {
        for (int j = 0; j < rcount[r]; ++j) {
          const int h = rstart[r] + j;
          Candidate& c = wave[at[r] + j];
          const int other = sc_other[h];
          for (int k4 = 0; k4 < 4; ++k4) { c.patch.coord[k4] = sc_coord[(size_t)4 * h + k4]; c.patch.normal[k4] = cams_[index].centre[k4] - c.patch.coord[k4]; }
          unitize4(c.patch.normal);
          c.patch.normal[3] = 0.0f;
          c.patch.images = {index, other};
          c.cell = rcell[r]; c.feature = rfeat[r]; c.order = j;
          c.parent = other;   // other image (for the counters)
          const Feature& p1 = features_[other][sc_other_feat[h]];
          c.dir = (((int)std::floor(p1.y + 0.5f)) / opt_.csize) * grids_[other].gw + ((int)std::floor(p1.x + 0.5f)) / opt_.csize;
        }
      }, 16);
    }
    std::vector<int> verdict;
    evaluate(wave, verdict);
    // ---- commit: replay the reference's sequential walk (seed.cpp:140-199) over the evaluated wave.  The cells are visited
    // in the reference's order and canAdd is asked again on the LIVE grids -- for the reference cell before its features are
    // tried, for the other image's cell before a candidate counts as a trial -- because a patch committed earlier in this wave
    // may have taken the cell since the snapshot; such cells / candidates are skipped exactly as the sequential walk would
    // (no trial counted).  The trial counters are unsigned chars that wrap, as in the reference.
    int total = 0;
    size_t k = 0;
    std::vector<char> open_now;
    while (k < wave.size()) {
      const int cell = wave[k].cell;
      size_t cell_end = k;
      while (cell_end < wave.size() && wave[cell_end].cell == cell) ++cell_end;
      bool placed = false;
      size_t f0 = k;
      if (!can_add(index, cell % g.gw, cell / g.gw)) f0 = cell_end;
      while (f0 < cell_end && !placed) {
        size_t f1 = f0;
        while (f1 < cell_end && wave[f1].feature == wave[f0].feature) ++f1;
        int count = 0;
        const Patch* best = nullptr;
        float best_score = 0.0f;
        // collectCandidates asks canAdd for the whole list BEFORE the first trial of this feature (seed.cpp:153 -> 287)
        open_now.resize(f1 - f0);
        for (size_t c = f0; c < f1; ++c) open_now[c - f0] = can_add(wave[c].parent, wave[c].dir % grids_[wave[c].parent].gw, wave[c].dir / grids_[wave[c].parent].gw);
        for (size_t c = f0; c < f1; ++c) {
          const int other = wave[c].parent, ocell = wave[c].dir;
          if (!open_now[c - f0]) continue;
          // trial counters of both cells (seed.cpp:175-181)
          ++grids_[index].counts[cell];
          if (other < tnum_) { ++grids_[other].counts[ocell]; refresh(other, ocell); }
          ++st.trial;
          if (verdict[c] == 1) { ++st.fail0; continue; }
          if (verdict[c] == 2) { ++st.fail1; continue; }
          ++st.pass;
          ++count;
          const Patch& p = wave[c].patch;
          const float score = std::max(0.0f, p.ncc - ncc_threshold_) * (int)p.images.size();   // CPatch::score
          if (!best || best_score < score) { best = &p; best_score = score; }
          if (count_threshold0_ <= count) break;
        }
        if (count != 0 && best) {
          const int id = add_patch(Patch(*best));
          for (size_t i = 0; i < patches_[id].images.size(); ++i)
            if (patches_[id].images[i] < tnum_) refresh(patches_[id].images[i], patches_[id].grids[i][1] * grids_[patches_[id].images[i]].gw + patches_[id].grids[i][0]);
          ++total;
          placed = true;
        }
        f0 = f1;
      }
      refresh(index, cell);
      k = cell_end;
    }
    std::cerr << '(' << index << ',' << total << ')' << std::flush;
  }
  std::cerr << "done" << std::endl;
  std::cerr << "Total pass fail0 fail1 refinepatch: " << st.trial << ' ' << st.pass << ' ' << st.fail0 << ' ' << st.fail1 << ' ' << st.pass + st.fail1 << std::endl;
}

// ---------------------------------------------------------------------------------------------- expansion
bool Pipeline::check_counts(const Patch& p) const {   // expand.cpp:258-323; true = reject (host/cell_rules.hpp)
  return pmvs::check_counts(p.images.data(), p.grids.empty() ? nullptr : &p.grids[0][0], (int)p.images.size(), tnum_, cell_views_.data(), count_threshold1_,
                            opt_.minImageNum, depth_);
}

bool Pipeline::update_counts(const Patch& p) {   // expand.cpp:325-406; true = the new patch joins the queue
  return pmvs::update_counts(p.images.data(), p.grids.empty() ? nullptr : &p.grids[0][0], (int)p.images.size(), p.vimages.data(),
                             p.vgrids.empty() ? nullptr : &p.vgrids[0][0], (int)p.vimages.size(), tnum_, cell_views_.data(), count_threshold1_);
}

void Pipeline::expand_round() {
  Tick tk(this, "round.expand");
  Stats st;
  for (auto& g : grids_) std::fill(g.counts.begin(), g.counts.end(), 0);
  for (Patch& p : patches_) p.flag = 0;
  // the queue is ordered by _tmp (patchOrganizerS.hpp:10-15); a wave takes the whole frontier, best first
  // the table is on the device in collectPatches order (uploaded after the seed round, reorganised by the filter rounds)
  if (!table_ready_) { upload_table(live_patches()); table_ready_ = true; }
  else device_rebuild(2, nullptr);   // CExpand::run starts from collectPatches + fresh depth maps (expand.cpp:42-47)
  std::vector<int> frontier = table_ids_;
  for (int id : frontier) patches_[id].flag = 1;
  std::cerr << "Expanding patches..." << std::flush;
  const double two_pi = 2 * M_PI;
  int wave_no = 0;
  size_t last_table_size = patches_.size();
  while (!frontier.empty()) {
    { Tick tk2(this, "host.expand.sort_frontier");
      // (key, id) pairs: the comparison reads contiguous memory instead of two Patch records per call
      std::vector<std::pair<float, int>> keyed(frontier.size());
      for (size_t k = 0; k < frontier.size(); ++k) keyed[k] = {patches_[frontier[k]].tmp, frontier[k]};
      std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first > b.first; });
      for (size_t k = 0; k < frontier.size(); ++k) frontier[k] = keyed[k].second; }
    std::vector<Candidate> wave;
    // findEmptyBlocks (expand.cpp:108-180): the neighbour search of every frontier patch is one kernel over the resident
    // table; it returns the directions that already have a neighbour and computeRadius
    frontier.erase(std::remove_if(frontier.begin(), frontier.end(), [&](int id) { return !patches_[id].alive; }), frontier.end());
    const int F = (int)frontier.size();
    std::vector<int32_t> fidx(F);
    std::vector<uint8_t> fmask(F);
    std::vector<float> fradius(F);
    for (int fi = 0; fi < F; ++fi) {
      fidx[fi] = table_index_[frontier[fi]];
      if (fidx[fi] < 0) { std::cerr << "expand: frontier patch is not in the GPU table" << std::endl; std::exit(1); }
    }
    { Tick tk2(this, "gpu.find_empty_blocks");
    if (pmvsb_find_empty_blocks_store(gpu_, F, fidx.data(), fmask.data(), fradius.data())) die("find_empty_blocks_store"); }
    std::vector<std::vector<Candidate>> per_parent(F);
    { Tick tk2(this, "host.expand.candidates");
    parallel_for(F, threads_, [&](int fi) {
      const int id = frontier[fi];
      std::vector<Candidate>& mine = per_parent[fi];   // this parent's candidates; concatenated in frontier order below
      const Patch& pp = patches_[id];
      float xdir[4], ydir[4];
      ortho(pp.normal, xdir, ydir);
      const int dnum = 6;
      const float radius = fradius[fi];
      for (int i = 0; i < dnum; ++i) {
        if (fmask[fi] & (1 << i)) continue;
        if (pp.dflag & (1 << i)) continue;
        const float angle = (float)(two_pi * i / dnum);
        Candidate c;
        const float cs = (float)(std::cos(angle) * radius), sn = (float)(std::sin(angle) * radius);
        for (int k = 0; k < 4; ++k) { c.patch.coord[k] = pp.coord[k] + cs * xdir[k] + sn * ydir[k]; c.patch.normal[k] = pp.normal[k]; }
        c.patch.flag = 1;
        c.parent = id; c.dir = i;
        // setGridsImages (patchOrganizerS.cpp:383-398): parent's images that see the candidate inside their grid
        for (int im : pp.images) {
          float ic[3];
          project(im, c.patch.coord, ic);
          const int ix = ((int)std::floor(ic[0] + 0.5f)) / opt_.csize, iy = ((int)std::floor(ic[1] + 0.5f)) / opt_.csize;
          if (0 <= ix && ix < grids_[im].gw && 0 <= iy && iy < grids_[im].gh) { c.patch.images.push_back(im); c.patch.grids.push_back({ix, iy}); }
        }
        if (c.patch.images.empty()) { patches_[id].dflag |= (unsigned char)(1 << i); continue; }
        if (!mask_gate(c.patch.coord)) { patches_[id].dflag |= (unsigned char)(1 << i); continue; }   // expand.cpp:212
        if (check_counts(c.patch)) { patches_[id].dflag |= (unsigned char)(1 << i); continue; }
        if (any_edge_) {   // COptim::removeImagesEdge (optim.cpp:385-396, expand.cpp:219)
          size_t out = 0;
          for (size_t k = 0; k < c.patch.images.size(); ++k)
            if (get_edge(c.patch.images[k], c.patch.coord)) { c.patch.images[out] = c.patch.images[k]; c.patch.grids[out] = c.patch.grids[k]; ++out; }
          c.patch.images.resize(out); c.patch.grids.resize(out);
          if (out == 0) { patches_[id].dflag |= (unsigned char)(1 << i); continue; }
        }
        mine.push_back(c);
      }
    }, 256);
    }
    { Tick tk2(this, "host.expand.concat");
      size_t total = 0;
      for (auto& v : per_parent) total += v.size();
      wave.reserve(total);
      for (auto& v : per_parent) for (auto& c : v) wave.push_back(std::move(c));
      per_parent.clear(); }
    std::vector<int> verdict;
    evaluate(wave, verdict);
    // commit in parent-priority order; cells may have been taken by an earlier commit of this wave
    std::vector<int> next;
    { Tick tk_commit(this, "host.expand.commit");
    for (size_t k = 0; k < wave.size(); ++k) {
      Candidate& c = wave[k];
      ++st.trial;
      bool fail = false;
      if (verdict[k] == 1) { ++st.fail0; fail = true; }
      else if (verdict[k] == 2) { ++st.fail1; fail = true; }
      else {
        Patch& p = c.patch;
        // the cell rules are re-checked against the grids as they are NOW (what a sequential run would have seen)
        if (check_counts(p)) { ++st.fail0; fail = true; }
        if (!fail) {
          ++st.pass;
          const bool requeue = update_counts(p);
          p.flag = 1;
          const int nid = add_patch(std::move(p));   // the wave's copy is not read again (only c.parent / c.dir)
          if (requeue) next.push_back(nid);
        }
      }
      if (fail) patches_[c.parent].dflag |= (unsigned char)(1 << c.dir);
    }
    }
    // CPatchOrganizerS::addPatch + updateDepthMaps for the patches committed by this wave: extend the resident table
    {
      std::vector<int> fresh;
      for (size_t id = last_table_size; id < patches_.size(); ++id) fresh.push_back((int)id);
      append_table(fresh);
      last_table_size = patches_.size();
    }
    frontier.swap(next);
    ++wave_no;
    std::cerr << '[' << wave_no << ':' << wave.size() << "->" << frontier.size() << ']' << std::flush;
  }
  std::cerr << std::endl << "Total pass fail0 fail1 refinepatch: " << st.trial << ' ' << st.pass << ' ' << st.fail0 << ' ' << st.fail1 << ' ' << st.pass + st.fail1 << std::endl;
}

// ---------------------------------------------------------------------------------------------- filters
// CFilter::run (filter.cpp:13-27).  The table stays on the device for the whole round: every filter is a kernel over it that
// answers with one keep flag per patch, removal + renumbering + depth maps + _vimages/_vpgrids are pmvsb_store_rebuild, and the
// host reads the lists back once, at the end of the round (sync_table).
void Pipeline::apply_keep(const char* name, const std::vector<uint8_t>& keep) {
  const int P = (int)keep.size();
  int count = 0;
  for (int k = 0; k < P; ++k) count += keep[k] ? 0 : 1;
  std::cerr << P << " -> " << P - count << " (" << 100.0f * (P - count) / std::max(P, 1) << "%)" << std::endl;
  (void)name;
  device_rebuild(1, &keep);   // setDepthMapsVGridsVPGridsAddPatchV(1) after every filter
}

void Pipeline::filter_outside() {   // filter.cpp:29-86
  Tick tk(this, "filter.outside");
  std::cerr << "FilterOutside" << std::endl;
  const int P = (int)table_ids_.size();
  if (P == 0) return;
  std::vector<float> gains(P);
  if (pmvsb_compute_gains_store(gpu_, gains.data())) die("compute_gains_store");
  std::vector<uint8_t> keep(P);
  double ave = 0.0, ave2 = 0.0;
  for (int k = 0; k < P; ++k) {
    ave += gains[k]; ave2 += (double)gains[k] * gains[k];
    keep[k] = gains[k] < 0.0f ? 0 : 1;
  }
  ave /= P; ave2 /= P;
  std::cerr << "Gain (ave/var): " << ave << ' ' << std::sqrt(std::max(0.0, ave2 - ave * ave)) << std::endl;
  apply_keep("outside", keep);
}

void Pipeline::filter_exact() {   // filter.cpp:203-355
  Tick tk(this, "filter.exact");
  std::cerr << "Filter Exact: " << std::flush;
  const int P = (int)table_ids_.size();
  if (P == 0) return;
  std::vector<uint8_t> keep(P);
  if (pmvsb_filter_exact_apply_store(gpu_, keep.data())) die("filter_exact_apply_store");
  std::cerr << std::endl;
  apply_keep("exact", keep);
}

void Pipeline::filter_neighbor() {   // filter.cpp:357-392, 464-519 (times = 1)
  Tick tk(this, "filter.neighbor");
  std::cerr << "FilterNeighbor:\t" << std::flush;
  const int P = (int)table_ids_.size();
  if (P == 0) return;
  std::vector<uint8_t> keep(P, 0);
  int32_t overflow = 0;
  if (pmvsb_filter_neighbor_store(gpu_, opt_.quad, keep.data(), nullptr, nullptr, &overflow)) die("filter_neighbor_store");
  if (overflow) std::cerr << "(" << overflow << " patches with more neighbours than the kernel keeps: not fitted) ";
  for (int k = 0; k < P; ++k) keep[k] = keep[k] ? 0 : 1;   // the kernel answers with reject flags
  apply_keep("neighbor", keep);
}

void Pipeline::filter_small_groups() {   // filter.cpp:524-665
  Tick tk(this, "filter.small_groups");
  std::cerr << "FilterGroups:\t" << std::flush;
  const int P = (int)table_ids_.size();
  if (P == 0) return;
  std::vector<uint8_t> keep(P);
  int32_t threshold = 0;
  if (pmvsb_filter_small_groups_store(gpu_, neighbor_threshold2_, keep.data(), &threshold)) die("filter_small_groups_store");
  std::cerr << threshold << std::endl;
  apply_keep("groups", keep);
}

void Pipeline::filter_round() {   // filter.cpp:13-27
  Tick tk(this, "round.filter");
  device_rebuild(0, nullptr);
  filter_outside();
  filter_exact();
  filter_neighbor();
  filter_small_groups();
  sync_table(false);
}

void Pipeline::run() {   // findMatch.cpp:187-220
  patches_.reserve((size_t)1 << 21);   // the commit loops append patch records one by one: no re-allocation of a growing table
  seed_round();
  ++depth_;
  for (int t = 0; t < 3; ++t) {
    expand_round();
    filter_round();
    ncc_threshold_ -= 0.05f;           // updateThreshold (findMatch.cpp:23-28)
    ncc_threshold_before_ -= 0.05f;
    count_threshold1_ = 2;
    ++depth_;
  }
}

// ---------------------------------------------------------------------------------------------- writers
namespace {
// formats records [0, P) on the CPU threads (one text block per slice, same iostream formatting as a single stream)
// and writes the blocks in order
template <typename F>
void write_records(const std::string& path, const std::string& header, int P, int threads, int precision, F record) {
  const int slices = std::max(1, std::min(threads * 4, (P + 4095) / 4096));
  std::vector<std::string> text(slices);
  pmvs::parallel_for(slices, threads, [&](int sidx) {
    const int b = (int)((long long)P * sidx / slices), e = (int)((long long)P * (sidx + 1) / slices);
    std::ostringstream o;
    if (precision > 0) o << std::setprecision(precision);
    for (int k = b; k < e; ++k) record(o, k);
    text[sidx] = o.str();
  }, 1);
  std::ofstream o(path.c_str(), std::ios::binary);
  o << header;
  for (const std::string& t : text) o.write(t.data(), (std::streamsize)t.size());
}
}  // namespace

void Pipeline::write(const std::string& base, bool ply, bool patch, bool pset) {   // patchOrganizerS.cpp:89-132, 687-779
  if (!is_root()) return;   // every rank holds the same patches; one of them writes
  if (table_ready_) sync_table(true);   // the writers read _vimages (.patch); the rounds did not need them on the host
  {
  Tick tk(this, "write.total");
  const std::vector<int> ids = table_ready_ ? table_ids_ : live_patches();
  const int P = (int)ids.size();
  if (ply) {
    std::vector<uint8_t> rgb((size_t)3 * std::max(P, 1), 0);
    if (P > 0) {
      int stride = 1;
      for (int id : ids) stride = std::max(stride, (int)patches_[id].images.size());
      std::vector<float> coords((size_t)4 * P);
      std::vector<int32_t> images((size_t)stride * P, 0), nimages(P);
      for (int k = 0; k < P; ++k) {
        const Patch& p = patches_[ids[k]];
        for (int c = 0; c < 4; ++c) coords[4 * k + c] = p.coord[c];
        nimages[k] = (int)p.images.size();
        for (size_t i = 0; i < p.images.size(); ++i) images[(size_t)k * stride + i] = p.images[i];
      }
      if (pmvsb_patch_colors_batch(gpu_, P, stride, coords.data(), images.data(), nimages.data(), rgb.data())) die("patch_colors_batch");
    }
    std::ostringstream h;
    h << "ply\nformat ascii 1.0\nelement vertex " << P << "\nproperty float x\nproperty float y\nproperty float z\nproperty float nx\nproperty float ny\n"
      << "property float nz\nproperty uchar diffuse_red\nproperty uchar diffuse_green\nproperty uchar diffuse_blue\nproperty float quality\nend_header\n";
    write_records(base + ".ply", h.str(), P, threads_, std::numeric_limits<double>::max_digits10, [&](std::ostream& o, int k) {
      const Patch& p = patches_[ids[k]];
      o << p.coord[0] << ' ' << p.coord[1] << ' ' << p.coord[2] << ' ' << p.normal[0] << ' ' << p.normal[1] << ' ' << p.normal[2] << ' '
        << (int)rgb[3 * k] << ' ' << (int)rgb[3 * k + 1] << ' ' << (int)rgb[3 * k + 2] << ' ' << p.ncc << '\n';
    });
  }
  if (patch) {
    // same text as the reference's `ofstr << patch` (source/pmvs/patch.cpp:30-48); newlines instead of std::endl flushes
    std::ostringstream h;
    h << "PATCHES" << '\n' << P << '\n';
    write_records(base + ".patch", h.str(), P, threads_, std::numeric_limits<double>::max_digits10, [&](std::ostream& o, int k) {
      const Patch& p = patches_[ids[k]];
      o << "PATCHS" << '\n'
        << p.coord[0] << ' ' << p.coord[1] << ' ' << p.coord[2] << ' ' << p.coord[3] << '\n'
        << p.normal[0] << ' ' << p.normal[1] << ' ' << p.normal[2] << ' ' << p.normal[3] << '\n'
        << p.ncc << ' ' << p.dscale << ' ' << p.ascale << '\n'
        << (int)p.images.size() << '\n';
      for (int im : p.images) o << image_ids_[im] << ' ';
      o << '\n' << (int)p.vimages.size() << '\n';
      for (int im : p.vimages) o << image_ids_[im] << ' ';
      o << '\n' << "\n";
    });
  }
  if (pset)
    write_records(base + ".pset", "", P, threads_, 0, [&](std::ostream& o, int k) {
      const Patch& p = patches_[ids[k]];
      o << p.coord[0] << ' ' << p.coord[1] << ' ' << p.coord[2] << ' ' << p.normal[0] << ' ' << p.normal[1] << ' ' << p.normal[2] << "\n";
    });
  std::cerr << "wrote " << P << " patches to " << base << ".*" << std::endl;
  }
  for (const auto& kv : seconds_) std::cerr << "time " << kv.first << ' ' << kv.second << " s" << std::endl;
  if (dist_.world > 1 && !dist_.tcp_exchange) exchanged_bytes_ = pmvsb_exchanged_bytes(gpu_);
  if (dist_.world > 1) std::cerr << "exchange " << dist_.world << " ranks, " << exchanged_bytes_ / 1.0e6 << " MB all-gathered over "
                                 << (dist_.exchange == Dist::kTcp ? "tcp" : dist_.exchange == Dist::kPeer ? "peer memory (CUDA IPC mailboxes)" : "nccl")
                                 << ", waves below " << dist_.shard_min << " candidates evaluated whole on every rank" << std::endl;
}

}  // namespace pmvs
