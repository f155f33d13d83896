// cmvs-pmvs_b200/host/setup.cpp -- option file, cameras, images, feature detection (host side).
//
// File contracts follow the reference so that existing scene directories work unchanged:
//   option file   /root/reference/source/pmvs/option.cpp:30-145   (keys, '#' comments, fatal unknown key)
//   vis.dat       option.cpp:160-299
//   cameras       <prefix>txt/%08d.txt, "CONTOUR" + 3x4 matrix    (source/image/camera.cpp:13-54, 257-270)
//   images        <prefix>visualize/%08d.ppm (binary P6)           (source/image/photoSetS.cpp:29-72)
// Feature detection = Harris (sigma 4) + DoG (scales 1..3), 4 strongest per 32x32 block of the working level
// (source/pmvs/detectFeatures.cpp:78-81, harris.cpp, dog.cpp, detector.hpp); SURVEY section 8f row 1 keeps it on the host.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <set>
#include <sstream>
#include <thread>

#include "pmvs_host.hpp"

namespace pmvs {

static void fatal(const std::string& msg) {
  std::cerr << msg << std::endl;
  std::exit(1);
}

// ---------------------------------------------------------------------------------------------- options
Options parse_options(const std::string& prefix, const std::string& option) {
  Options o;
  o.prefix = prefix;
  o.option = option;
  std::ifstream in((prefix + option).c_str());
  if (!in.is_open()) fatal("Cannot open option file: " + prefix + option);
  std::string name;
  while (in >> name) {
    if (name[0] == '#') { std::string rest; std::getline(in, rest); continue; }
    if (name == "level") in >> o.level;
    else if (name == "csize") in >> o.csize;
    else if (name == "threshold") in >> o.threshold;
    else if (name == "wsize") in >> o.wsize;
    else if (name == "minImageNum") in >> o.minImageNum;
    else if (name == "CPU") in >> o.CPU;
    else if (name == "setEdge") in >> o.setEdge;
    else if (name == "useBound") in >> o.useBound;
    else if (name == "useVisData") in >> o.useVisData;
    else if (name == "sequence") in >> o.sequence;
    else if (name == "quad") in >> o.quad;
    else if (name == "maxAngle") in >> o.maxAngleDeg;
    else if (name == "timages") {
      in >> o.tflag;
      if (o.tflag == -1) {
        int a, b;
        in >> a >> b;
        for (int i = a; i < b; ++i) o.timages.push_back(i);
      } else if (0 < o.tflag) {
        for (int i = 0; i < o.tflag; ++i) { int v; in >> v; o.timages.push_back(v); }
      } else {
        fatal("tflag is not valid: " + std::to_string(o.tflag));
      }
    } else if (name == "oimages") {
      in >> o.oflag;
      if (o.oflag == -1) {
        int a, b;
        in >> a >> b;
        for (int i = a; i < b; ++i) o.oimages.push_back(i);
      } else if (0 <= o.oflag) {
        for (int i = 0; i < o.oflag; ++i) { int v; in >> v; o.oimages.push_back(v); }
      } else if (o.oflag != -2 && o.oflag != -3) {
        fatal("oflag is not valid: " + std::to_string(o.oflag));
      }
    } else {
      fatal("Unrecognizable option: " + name);
    }
  }
  if (o.tflag == -10 || o.oflag == -10) fatal("_tflag and _oflag not specified: " + std::to_string(o.tflag) + " " + std::to_string(o.oflag));

  std::map<int, int> tdict;
  for (int i = 0; i < (int)o.timages.size(); ++i) tdict[o.timages[i]] = i;

  // oimages from vis.dat: every image co-visible with a target image that is not itself a target
  if (o.oflag == -2) {
    std::ifstream vis((prefix + "vis.dat").c_str());
    if (!vis.is_open()) fatal("No vis.dat although specified to initOimages: \n" + prefix + "vis.dat");
    std::string header;
    int n;
    vis >> header >> n;
    o.oimages.clear();
    for (int c = 0; c < n; ++c) {
      int id, k;
      vis >> id >> k;
      const bool is_target = tdict.count(c) != 0;
      for (int i = 0; i < k; ++i) {
        int other;
        vis >> other;
        if (is_target && !tdict.count(other)) o.oimages.push_back(other);
      }
    }
    std::sort(o.oimages.begin(), o.oimages.end());
    o.oimages.erase(std::unique(o.oimages.begin(), o.oimages.end()), o.oimages.end());
  }

  std::vector<int> images = o.timages;
  images.insert(images.end(), o.oimages.begin(), o.oimages.end());
  const int num = (int)images.size();
  o.visdata2.assign(num, {});
  if (o.useVisData == 0) {
    for (int y = 0; y < num; ++y)
      for (int x = 0; x < num; ++x)
        if (x != y) o.visdata2[y].push_back(x);
  } else {
    std::map<int, int> dict;
    for (int i = 0; i < num; ++i) dict[images[i]] = i;
    std::ifstream vis((prefix + "vis.dat").c_str());
    if (!vis.is_open()) fatal("No vis.dat although specified to initVisdata2: \n" + prefix + "vis.dat");
    std::string header;
    int n;
    vis >> header >> n;
    std::vector<std::vector<char>> m(num, std::vector<char>(num, 0));
    for (int c = 0; c < n; ++c) {
      int id, k;
      vis >> id >> k;
      const auto it0 = dict.find(c);
      for (int i = 0; i < k; ++i) {
        int other;
        vis >> other;
        const auto it1 = dict.find(other);
        if (it0 != dict.end() && it1 != dict.end()) { o.visdata2[it0->second].push_back(it1->second); m[it0->second][it1->second] = 1; }
      }
    }
    (void)m;   // the symmetrised matrix (_visdata) is not read anywhere on the path; the lists are what addImages uses
  }
  if (o.useBound) {
    std::ifstream b((prefix + "bimages.dat").c_str());
    if (!b.is_open()) fatal("File not found: " + prefix + "bimages.dat");
    int n;
    b >> n;
    for (int i = 0; i < n; ++i) {
      int v;
      b >> v;
      if (tdict.count(v)) o.bindexes.push_back(tdict[v]);
    }
  }
  std::cerr << "--------------------------------------------------" << std::endl
            << "--- Summary of specified options ---" << std::endl
            << "# of timages: " << o.timages.size() << (o.tflag == -1 ? " (range specification)" : " (enumeration)") << std::endl
            << "# of oimages: " << o.oimages.size() << std::endl
            << "level: " << o.level << "  csize: " << o.csize << std::endl
            << "threshold: " << o.threshold << "  wsize: " << o.wsize << std::endl
            << "minImageNum: " << o.minImageNum << "  CPU: " << o.CPU << std::endl
            << "useVisData: " << o.useVisData << "  sequence: " << o.sequence << std::endl
            << "--------------------------------------------------" << std::endl;
  return o;
}

// ---------------------------------------------------------------------------------------------- files
static bool read_camera(const std::string& path, float* P) {
  std::ifstream in(path.c_str());
  if (!in.is_open()) return false;
  std::string header;
  in >> header;
  if (header != "CONTOUR") {
    if (header == "CONTOUR2" || header == "CONTOUR3") fatal("Camera format " + header + " is not supported by pmvs-b200 (use CONTOUR 3x4 matrices): " + path);
    fatal("Unrecognizable txt format");
  }
  for (int i = 0; i < 12; ++i) in >> P[i];
  return !in.fail();
}

static bool read_ppm(const std::string& path, std::vector<unsigned char>& rgb, int& w, int& h) {
  FILE* fp = std::fopen(path.c_str(), "rb");
  if (!fp) return false;
  auto next_int = [&](int& v) {
    for (;;) {
      int c = std::fgetc(fp);
      if (c == EOF) return false;
      if (c == '#') { while (c != '\n' && c != EOF) c = std::fgetc(fp); continue; }
      if (c == ' ' || c == '\t' || c == '\n' || c == '\r') continue;
      std::ungetc(c, fp);
      break;
    }
    return std::fscanf(fp, "%d", &v) == 1;
  };
  char magic[3] = {0, 0, 0};
  int maxv = 0;
  const bool ok = std::fscanf(fp, "%2s", magic) == 1 && std::string(magic) == "P6" && next_int(w) && next_int(h) && next_int(maxv) && maxv == 255 && w > 0 && h > 0;
  if (!ok) { std::fclose(fp); return false; }
  std::fgetc(fp);
  rgb.resize((size_t)w * h * 3);
  const size_t got = std::fread(rgb.data(), 1, rgb.size(), fp);
  std::fclose(fp);
  return got == rgb.size();
}

// masks/%08d.{pgm,pbm}, edges/%08d.{pgm,pbm} (CImage::completeName + readPGMImage / readPBMImage,
// source/image/image.cpp:38-75, 507-567, 619-668): binary P5, or binary P4 read as ONE continuous bit stream (the reference does
// not skip the row padding) with a set bit = outside (0) and a clear bit = inside (255).
static bool read_gray_map(const std::string& base, std::vector<unsigned char>& out, int& w, int& h) {
  for (const char* ext : {".pgm", ".pbm"}) {
    std::ifstream in((base + ext).c_str(), std::ios::binary);
    if (!in.is_open()) continue;
    const bool pgm = ext[2] == 'g';
    std::string header;
    in >> header;
    in.get();
    if (header != (pgm ? "P5" : "P4")) { std::cerr << "Only accept binary " << (pgm ? "pgm" : "pbm") << " format: " << base << ext << std::endl; return false; }
    while (in.peek() == '#') { std::string line; std::getline(in, line); }
    int maxv = 255;
    in >> w >> h;
    if (pgm) in >> maxv;
    in.get();
    if (in.fail() || w < 1 || h < 1) return false;
    out.assign((size_t)w * h, 0);
    if (pgm) {
      in.read(reinterpret_cast<char*>(out.data()), (std::streamsize)out.size());
      if ((size_t)in.gcount() != out.size()) fatal("Truncated map file: " + base + ext);
    } else {
      size_t count = 0;
      unsigned char byte = 0;
      while (count < out.size()) {
        if (!in.read(reinterpret_cast<char*>(&byte), 1)) fatal("Truncated map file: " + base + ext);
        for (int j = 0; j < 8 && count < out.size(); ++j, byte = (unsigned char)(byte << 1)) out[count++] = (byte >> 7) ? 0 : 255;
      }
    }
    return true;
  }
  return false;
}

void Pipeline::die(const std::string& where) const {
  std::cerr << where << ": " << (gpu_ ? pmvsb_last_error(gpu_) : "no GPU context") << std::endl;
  std::exit(1);
}

Pipeline::Pipeline(const Options& o, const Dist& dist) : opt_(o), dist_(dist) {
  image_ids_ = o.timages;
  image_ids_.insert(image_ids_.end(), o.oimages.begin(), o.oimages.end());
  tnum_ = (int)o.timages.size();
  num_ = (int)image_ids_.size();
  tau_ = std::min(o.minImageNum * 2, num_);                 // findMatch.cpp:56
  ncc_threshold_ = o.threshold;
  ncc_threshold_before_ = o.threshold - 0.3f;               // findMatch.cpp:104
  threads_ = std::max(1, std::min(o.CPU, (int)std::thread::hardware_concurrency() / std::max(1, dist.world)));   // the ranks share the host cores
  if (const char* t = std::getenv("PMVSB_HOST_THREADS")) { const int v = std::atoi(t); if (v > 0) threads_ = std::min(threads_, v); }   // pmvs2_clusters: clusters side by side
}

Pipeline::~Pipeline() {
  if (gpu_) pmvsb_destroy(gpu_);
}

void Pipeline::load() {
  Tick tk(this, "load.total");
  if (num_ == 0 || tnum_ == 0) fatal("No target images");
  if (num_ > PMVSB_MAX_VIEWS)
    std::cerr << "pmvs-b200: " << num_ << " images in one option file; a patch may list at most " << PMVSB_MAX_VIEWS
              << " visible images (the run stops with an error if one needs more -- cluster the scene with CMVS / genOption)" << std::endl;
  // the files are read by the CPU threads while this thread brings the CUDA context up
  struct Loaded {
    float P[12]; std::vector<unsigned char> rgb; int w = 0, h = 0; std::string error;
    std::vector<unsigned char> map[2]; int mw[2] = {0, 0}, mh[2] = {0, 0};   // masks/ and edges/ files, when present
  };
  std::vector<Loaded> loaded(num_);
  std::thread reader([&]() {
    parallel_for(num_, threads_, [&](int i) {
      Loaded& L = loaded[i];
      char name[1024];
      std::snprintf(name, sizeof(name), "%stxt/%08d.txt", opt_.prefix.c_str(), image_ids_[i]);
      if (!read_camera(name, L.P)) {
        std::snprintf(name, sizeof(name), "%stxt/%04d.txt", opt_.prefix.c_str(), image_ids_[i]);
        if (!read_camera(name, L.P)) { L.error = std::string("Cannot read camera: ") + name; return; }
      }
      bool four_digits = false;
      std::snprintf(name, sizeof(name), "%svisualize/%08d.ppm", opt_.prefix.c_str(), image_ids_[i]);
      if (!read_ppm(name, L.rgb, L.w, L.h)) {
        std::snprintf(name, sizeof(name), "%svisualize/%04d.ppm", opt_.prefix.c_str(), image_ids_[i]);
        if (!read_ppm(name, L.rgb, L.w, L.h)) {
          std::snprintf(name, sizeof(name), "%svisualize/%08d.jpg", opt_.prefix.c_str(), image_ids_[i]);
          if (std::ifstream(name)) L.error = std::string("JPEG input is not supported by pmvs-b200 (convert to binary PPM): ") + name;
          else L.error = "Unsupported iamge format found. Stop allocation: " + std::string(name);
          return;
        }
        four_digits = true;
      }
      for (int which = 0; which < 2; ++which) {   // photoSetS.cpp:43-49, 60-62: same digit count as the image name
        std::snprintf(name, sizeof(name), four_digits ? "%s%s/%04d" : "%s%s/%08d", opt_.prefix.c_str(), which == 0 ? "masks" : "edges", image_ids_[i]);
        if (read_gray_map(name, L.map[which], L.mw[which], L.mh[which]) && (L.mw[which] != L.w || L.mh[which] != L.h))
          L.error = std::string(which == 0 ? "Mask" : "Edge map") + " and image differ in size: " + name;
      }
    }, 1);
  });
  int created;
  { Tick tk2(this, "load.create_gpu_context");
    int ndev = pmvsb_device_count();
    created = pmvsb_create(&gpu_, ndev > 0 ? dist_.local_rank % ndev : dist_.local_rank, num_, tnum_, opt_.level, opt_.csize, opt_.wsize, opt_.minImageNum, opt_.threshold, opt_.maxAngleDeg); }
  { Tick tk2(this, "load.wait_for_files"); reader.join(); }
  if (created != 0)
    fatal("pmvs-b200: cannot create a GPU context (CUDA device required; there is no CPU fallback) or bad options");
  // The exchange between the GPUs of this run is brought up on its own thread while the images go to the GPU; joined at the end
  // of load().  Peer mailboxes (default): export mine, gather every rank's 64-byte handle over the rendezvous sockets, map them
  // -- tens of milliseconds.  NCCL (PMVSB_EXCHANGE=nccl, or when a rank cannot map a peer): ncclCommInitRank, 0.5 - 1.5 s.
  std::thread nccl_init;
  double nccl_seconds = 0.0;   // written by the thread, read after the join
  const auto t_nccl = std::chrono::steady_clock::now();
  if (dist_.world > 1 && dist_.exchange != Dist::kTcp)
    nccl_init = std::thread([&]() {
      if (dist_.exchange == Dist::kPeer) {
        const char* mb = std::getenv("PMVSB_PEER_SLOT_MB");
        const size_t slot = std::max<size_t>(4096, (size_t)((mb && *mb ? std::atof(mb) : 64.0) * 1048576.0));   // grows on demand (PMVSB_EGROW)
        if (!peer_bringup(slot)) {
          if (is_root()) std::cerr << "pmvs2: the GPUs of this run cannot map each other's memory (" << pmvsb_last_error(gpu_) << "); falling back to NCCL" << std::endl;
          dist_.exchange = Dist::kNccl;
        }
      }
      if (dist_.exchange == Dist::kNccl) {
        uint8_t id[128] = {0};
        if (dist_.rank == 0 && pmvsb_comm_unique_id(gpu_, id)) die("comm_unique_id");
        dist_.broadcast_from_root(id, sizeof(id));
        if (pmvsb_comm_init(gpu_, dist_.rank, dist_.world, id)) die("comm_init");
      }
      nccl_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_nccl).count();
    });
  cams_.resize(num_);
  lw_.resize(num_); lh_.resize(num_);
  std::cerr << "Reading images: " << std::flush;
  for (int i = 0; i < num_; ++i) {
    Loaded& L = loaded[i];
    if (!L.error.empty()) fatal(L.error);
    if (pmvsb_upload_camera(gpu_, i, L.P)) die("upload_camera");
    if (pmvsb_upload_image(gpu_, i, L.w, L.h, L.rgb.data())) die("upload_image");
    for (int which = 0; which < 2; ++which)
      if (!L.map[which].empty()) {
        std::cerr << (which == 0 ? "Read mask: " : "Read edge: ") << image_ids_[i] << std::endl;
        if (pmvsb_upload_mask(gpu_, i, which, L.mw[which], L.mh[which], L.map[which].data())) die("upload_mask");
        std::vector<unsigned char>().swap(L.map[which]);
      }
    if (pmvsb_set_visdata2(gpu_, i, opt_.visdata2[i].data(), (int)opt_.visdata2[i].size())) die("set_visdata2");
    std::vector<unsigned char>().swap(L.rgb);
    std::cerr << '*' << std::flush;
  }
  std::cerr << std::endl;
  if (opt_.setEdge != 0.0f && pmvsb_set_edge(gpu_, opt_.setEdge)) die("set_edge");   // findMatch.cpp:74-76
  if (!opt_.bindexes.empty() && pmvsb_set_bimages(gpu_, opt_.bindexes.data(), (int)opt_.bindexes.size())) die("set_bimages");
  if (pmvsb_finalize_scene(gpu_)) die("finalize_scene");
  masks_.assign(num_, {}); edges_.assign(num_, {});
  for (int i = 0; i < num_; ++i)
    for (int which = 0; which < 2; ++which) {
      int present = 0, w = 0, h = 0;
      if (pmvsb_download_mask(gpu_, i, which, nullptr, &present)) die("download_mask");
      if (!present) continue;
      if (pmvsb_image_dims(gpu_, i, opt_.level, &w, &h)) die("image_dims");
      std::vector<unsigned char>& m = which == 0 ? masks_[i] : edges_[i];
      m.resize((size_t)w * h);
      if (pmvsb_download_mask(gpu_, i, which, m.data(), &present)) die("download_mask");
      (which == 0 ? any_mask_ : any_edge_) = true;
    }
  grids_.resize(num_);
  P0_.resize(num_);
  for (int i = 0; i < num_; ++i) {
    float P[12];
    Camera& c = cams_[i];
    if (pmvsb_get_camera(gpu_, i, opt_.level, P, c.centre, c.oaxis, c.xaxis, c.yaxis, c.zaxis, &c.ipscale)) die("get_camera");
    P0_[i].resize(12);
    for (int k = 0; k < 12; ++k) { c.P[k / 4][k % 4] = P[k]; P0_[i][k] = P[k]; }
    if (pmvsb_image_dims(gpu_, i, opt_.level, &lw_[i], &lh_[i])) die("image_dims");
    ImageGrid& g = grids_[i];
    if (pmvsb_grid_dims(gpu_, i, &g.gw, &g.gh)) die("grid_dims");
    if (i < tnum_) {
      g.occ.assign((size_t)g.gw * g.gh, 0);
      g.counts.assign((size_t)g.gw * g.gh, 0);
    }
  }
  cell_views_.resize(tnum_);
  for (int i = 0; i < tnum_; ++i) cell_views_[i] = {grids_[i].gw, grids_[i].gh, grids_[i].occ.data(), grids_[i].counts.data()};
  // CPhotoSetS::setDistances (source/image/photoSetS.cpp:195-235): baseline / mean baseline + axis divergence beyond 10 degrees
  distances_.assign(num_, std::vector<float>(num_, 0.0f));
  float avedis = 0.0f;
  int denom = 0;
  for (int i = 0; i < num_; ++i)
    for (int j = 0; j < num_; ++j) {
      if (i == j) continue;
      float d[4];
      for (int k = 0; k < 4; ++k) d[k] = cams_[i].centre[k] - cams_[j].centre[k];
      const float len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3]);
      distances_[i][j] = len;
      avedis += len;
      ++denom;
    }
  if (denom) {
    avedis /= denom;
    if (avedis == 0.0f) fatal("All the optical centers are identical..?");
    const float margin = std::cos(10.0f * M_PI / 180.0f);
    for (int i = 0; i < num_; ++i)
      for (int j = 0; j < num_; ++j) {
        distances_[i][j] /= avedis;
        const float dot = cams_[i].oaxis[0] * cams_[j].oaxis[0] + cams_[i].oaxis[1] * cams_[j].oaxis[1] + cams_[i].oaxis[2] * cams_[j].oaxis[2];
        distances_[i][j] += std::max(0.0f, 1.0f - dot - margin);
      }
  }
  detect_features();
  if (nccl_init.joinable()) {
    { Tick tk2(this, "load.wait_for_exchange"); nccl_init.join(); }
    seconds_[dist_.exchange == Dist::kPeer ? "load.peer_mailboxes(background)" : "load.nccl_init(background)"] += nccl_seconds;
  }
}

// Every rank (re)allocates its mailbox with `slot_bytes` per rank, the 64-byte IPC handles go round the rendezvous sockets (which is
// also the barrier: every mailbox exists before anybody maps it), every rank maps the others'.  All ranks return the same
// answer: true only when every rank mapped every peer (a rank that failed still takes part in the two exchanges).
bool Pipeline::peer_bringup(size_t slot_bytes) {
  const int W = dist_.world;
  uint8_t mine[PMVSB_PEER_HANDLE_BYTES] = {0};
  int32_t ok = pmvsb_peer_export(gpu_, dist_.rank, W, slot_bytes, mine) == 0 ? 1 : 0;
  std::vector<uint8_t> all((size_t)PMVSB_PEER_HANDLE_BYTES * W);
  dist_.allgather(mine, sizeof(mine), all.data());
  if (ok) ok = pmvsb_peer_open(gpu_, all.data()) == 0 ? 1 : 0;
  std::vector<int32_t> oks(W, 0);
  dist_.allgather(&ok, sizeof(ok), oks.data());
  bool every = true;
  for (int k = 0; k < W; ++k) every = every && oks[k] == 1;
  if (!every) pmvsb_peer_close(gpu_);
  return every;
}

// ---------------------------------------------------------------------------------------------- features
// CDetectFeatures::run (source/pmvs/detectFeatures.cpp:13-125): Harris + DoG per image, on the device pyramid
void Pipeline::detect_features() {
  Tick tk(this, "gpu.detect_features");
  features_.assign(num_, {});
  const int fcsize = 16;   // findMatch.cpp:81
  size_t total = 0;
  std::vector<float> xy, resp;
  std::vector<int32_t> type;
  for (int i = 0; i < num_; ++i) {
    const int blocks = ((lw_[i] + 2 * fcsize - 1) / (2 * fcsize)) * ((lh_[i] + 2 * fcsize - 1) / (2 * fcsize));
    const int cap = 8 * blocks;   // at most 4 points per block and detector
    xy.resize((size_t)2 * cap); resp.resize(cap); type.resize(cap);
    int32_t n = 0;
    if (pmvsb_detect_features(gpu_, i, fcsize, cap, xy.data(), resp.data(), type.data(), &n)) die("detect_features");
    features_[i].resize(n);
    for (int k = 0; k < n; ++k) features_[i][k] = {xy[2 * k], xy[2 * k + 1], resp[k], type[k]};
    if (pmvsb_set_features(gpu_, i, n, xy.data(), type.data())) die("set_features");   // CSeed::readPoints: binned on the device side
    total += (size_t)n;
  }
  std::cerr << "features: " << total << " in " << num_ << " images" << std::endl;
}

}  // namespace pmvs
