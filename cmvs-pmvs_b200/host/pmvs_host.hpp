// cmvs-pmvs_b200/host/pmvs_host.hpp -- host side of the pmvs2 drop-in binary (C++17).
//
// What stays on the host (BASELINE.json north_star): option / camera / image file handling, seed candidate
// enumeration, the per-image cell bookkeeping and commit rules (CPatchOrganizerS's role), the small-group filter and
// the output writers.  Everything photometric or neighbour-searching (features, pre/postProcess incl. check,
// refinePatch, depth maps, visibility, gains, findEmptyBlocks, filterNeighbor) is a batched call into
// libpmvs_b200.so (include/pmvs_b200.h).
//
// The reference works patch by patch from worker threads; this driver works in WAVES: all candidates that can be
// generated from the current grid snapshot are evaluated by one launch per stage and then committed in the
// reference's priority order, re-checking the cell rules at commit time.
#pragma once
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <map>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pmvs_b200.h"
#include "cell_rules.hpp"

namespace pmvs {

// The host bookkeeping loops that only READ the grids (neighbour searches, marshalling) run on the option file's
// `CPU` threads, as the reference's do; body(i) must touch only item i's own data.
template <typename F>
void parallel_for(int n, int threads, F body, int chunk = 64) {
  threads = std::max(1, std::min(threads, (n + chunk - 1) / chunk));
  if (threads == 1) { for (int i = 0; i < n; ++i) body(i); return; }
  std::atomic<int> next(0);
  auto work = [&]() {
    for (;;) {
      const int b = next.fetch_add(chunk);
      if (b >= n) break;
      const int e = std::min(n, b + chunk);
      for (int i = b; i < e; ++i) body(i);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < threads; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
}

struct Options {   // source/pmvs/option.cpp:10-28, 47-109
  int level = 1, csize = 2, wsize = 7, minImageNum = 3, CPU = 4, useBound = 0, useVisData = 0, sequence = -1;
  float threshold = 0.7f, setEdge = 0.0f, maxAngleDeg = 10.0f, quad = 2.5f;
  int tflag = -10, oflag = -10;
  std::vector<int> timages, oimages;
  std::vector<std::vector<int>> visdata2;   // indexes into timages ++ oimages
  std::vector<int> bindexes;
  std::string prefix, option;
};
Options parse_options(const std::string& prefix, const std::string& option);

// one process per GPU (distributed.cpp); world = 1 is the ordinary single-GPU run
struct Dist {
  int rank = 0, world = 1, local_rank = 0, port = 0;
  // how a wave's results travel between the ranks (PMVSB_EXCHANGE): peer = stores into each other's GPU memory over NVLink
  // (CUDA IPC mailboxes, the default), nccl = one ncclAllGather per wave, tcp = over the rendezvous sockets below
  enum Exchange { kPeer, kNccl, kTcp };
  Exchange exchange = kPeer;
  bool tcp_exchange = false;      // exchange == kTcp
  int shard_min = 4096;           // PMVSB_SHARD_MIN: waves with fewer candidates are evaluated whole on every rank (no exchange)
  std::string master_addr;
  std::vector<int> fds;           // rank 0: one socket per peer (index = rank); peers: fds[0] = the socket to rank 0
  static Dist from_env();
  void connect();                 // persistent star through rank 0 (no-op when world == 1)
  void broadcast_from_root(void* buf, size_t n) const;
  void allgather(const void* send, size_t n, void* recv) const;   // n bytes per rank, rank order
  // contiguous balanced shard of n items for rank r
  static void shard(int n, int world, int r, int& lo, int& hi) { lo = (int)((long long)n * r / world); hi = (int)((long long)n * (r + 1) / world); }
};

struct Camera {
  float P[3][4];       // at the working level
  float centre[4], oaxis[4], xaxis[3], yaxis[3], zaxis[3], ipscale;
};

struct Feature {       // PMVS3::CPoint
  float x, y, response;
  int type;            // 0 Harris, 1 DoG
};

struct Patch {         // Patch::CPatch
  float coord[4] = {0, 0, 0, 1}, normal[4] = {0, 0, 0, 0};
  float ncc = -1.0f, dscale = 0.0f, ascale = 0.0f, tmp = 0.0f;
  int timages = 0, flag = 0;
  unsigned char dflag = 0;
  bool alive = true;
  std::vector<int> images, vimages;
  std::vector<std::array<int, 2>> grids, vgrids;
};

struct ImageGrid {     // per image: CPatchOrganizerS::_pgrids / _vpgrids / _counts / _dpgrids of that image
  int gw = 0, gh = 0;
  std::vector<int> occ;                    // number of patches in each cell of _pgrids (target images only); the lists are on the device
  std::vector<unsigned char> counts;
};

struct Stats { long trial = 0, pass = 0, fail0 = 0, fail1 = 0; };

class Pipeline {
 public:
  explicit Pipeline(const Options& o, const Dist& dist = Dist());
  bool is_root() const { return dist_.rank == 0; }
  ~Pipeline();
  void load();                 // images + cameras -> GPU context, features
  void run();                  // seed, 3 x (expand, filter)
  void write(const std::string& base, bool ply, bool patch, bool pset);

 private:
  // ---- set-up
  void detect_features();
  // ---- geometry helpers (f32, the reference's formulas)
  void project(int image, const float* X, float* out3) const;
  float get_unit(int image, const float* X) const;
  bool is_neighbor(const Patch& l, const Patch& r, float hunit, float thr, float radius) const;   // radius < 0: no radius test
  bool is_neighbor(const Patch& l, const Patch& r, float thr) const;
  // masks / edges / bounding images (CImage::getMask, CPhoto::getMask / getEdge, CFindMatch::insideBimages)
  int get_mask(int image, int ix, int iy) const;          // include/image/image.hpp:553-565 at the working level
  int get_mask(int image, const float* X) const;          // include/image/photo.hpp:44-49
  int get_edge(int image, const float* X) const;          // photo.hpp:51-59
  bool mask_gate(const float* X) const;                   // getMask(coord, level) != 0 && insideBimages(coord) != 0
  // ---- bookkeeping
  int add_patch(Patch&& p);                 // CPatchOrganizerS::addPatch (takes the lists over)
  std::vector<int> live_patches() const;    // ids of the live patches in creation order
  void device_rebuild(int mode, const std::vector<uint8_t>* keep);
  void apply_keep(const char* name, const std::vector<uint8_t>& keep);
  void sync_table(bool full);
  struct TableArrays;
  void marshal(const std::vector<int>& ids, TableArrays& t) const;
  void upload_table(const std::vector<int>& ids);
  void append_table(const std::vector<int>& ids);
  // ---- rounds
  void seed_round();
  void expand_round();
  void filter_round();
  void filter_outside();
  void filter_exact();
  void filter_neighbor();
  void filter_small_groups();
  bool check_counts(const Patch& p) const;
  bool update_counts(const Patch& p);
  struct Candidate { Patch patch; int parent = -1; int dir = -1; int cell = -1; int feature = -1; int order = 0; };
  // pre -> refine -> post (+ vimages at depth >= 1) for a batch; verdict[i] = 0 accepted, 1 failed in preProcess, 2 in postProcess
  void evaluate(std::vector<Candidate>& cands, std::vector<int>& verdict);
  void evaluate_range(std::vector<Candidate>& cands, std::vector<int>& verdict, int lo, int hi, bool gather);
  void exchange_results(std::vector<Candidate>& cands, std::vector<int>& verdict);
  bool peer_bringup(size_t slot_bytes);   // maps the ranks' mailboxes into each other (setup.cpp); the same answer on every rank
  void die(const std::string& where) const;

  Options opt_;
  Dist dist_;
  pmvsb_ctx* gpu_ = nullptr;
  int num_ = 0, tnum_ = 0, tau_ = 0, depth_ = 0, threads_ = 1;
  float ncc_threshold_ = 0.7f, ncc_threshold_before_ = 0.4f;
  int count_threshold0_ = 2, count_threshold1_ = 4, count_threshold2_ = 2;
  float neighbor_threshold_ = 0.5f, neighbor_threshold1_ = 1.0f, neighbor_threshold2_ = 1.0f;
  std::vector<int> image_ids_;              // file numbers, targets first
  std::vector<Camera> cams_;
  std::vector<int> lw_, lh_;                // image size at the working level
  std::vector<std::vector<unsigned char>> masks_, edges_;   // working-level maps as the GPU built them (empty = none)
  bool any_mask_ = false, any_edge_ = false;
  std::vector<std::vector<double>> P0_;     // level-`level` projection in double for the epipolar search
  std::vector<std::vector<float>> distances_;
  std::vector<std::vector<Feature>> features_;
  std::vector<ImageGrid> grids_;
  std::vector<CellGridView> cell_views_;   // grids_ as the cell rules take them (host/cell_rules.hpp)
  std::vector<Patch> patches_;
  std::vector<int> table_ids_;              // table index -> patch id, for the table resident on the GPU
  std::vector<int> table_index_;            // patch id -> table index (-1: not in the table)
  double exchanged_bytes_ = 0.0;            // wave results received over the all-gather (multi-GPU runs)
  bool table_ready_ = false;                // the table has been uploaded (after the seed round)
  std::map<std::string, double> seconds_;   // wall time per phase (printed by write())
 public:
  struct Tick {
    Pipeline* p; std::string k; std::chrono::steady_clock::time_point t0;
    Tick(Pipeline* pp, const std::string& kk) : p(pp), k(kk), t0(std::chrono::steady_clock::now()) {}
    ~Tick() { p->seconds_[k] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
  };
};

}  // namespace pmvs
