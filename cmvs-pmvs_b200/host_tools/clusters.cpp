// cmvs-pmvs_b200/host_tools/clusters.cpp -- `pmvs2_clusters prefix [--gpus N] [--merge NAME] [PATCH] [PSET]`
//
// Runs the per-cluster option files that genOption writes (option-0000, option-0001, ...; the commands of the
// generated pmvs.sh, /root/reference/source/genOption.cpp:66-74) as independent `pmvs2` processes, one cluster per
// GPU at a time: cluster c of a batch of N goes to GPU (slot) via CUDA_VISIBLE_DEVICES.  Clusters share nothing
// ("replicas only", no collective); when all are done the per-cluster models are concatenated into
// models/<NAME>.ply / .patch / .pset (default NAME = option-all), which is the merge the reference leaves to the user
// (each of its cluster runs writes models/option-%04d.* and nothing combines them).
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <algorithm>
#include <vector>

namespace {

bool exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0; }

std::string sibling(const char* argv0, const char* name) {   // bin/pmvs2 next to this binary
  char buf[4096];
  const ssize_t n = readlink("/proc/self/exe", buf, sizeof(buf) - 1);
  std::string self = n > 0 ? std::string(buf, (size_t)n) : std::string(argv0);
  const size_t slash = self.rfind('/');
  return (slash == std::string::npos ? std::string("./") : self.substr(0, slash + 1)) + name;
}

int gpu_count_from_env() {
  // CUDA_VISIBLE_DEVICES of the parent lists the devices the clusters may use; otherwise /proc/driver/nvidia/gpus
  if (const char* v = getenv("CUDA_VISIBLE_DEVICES")) {
    if (!*v) return 0;
    int n = 1;
    for (const char* p = v; *p; ++p) n += *p == ',';
    return n;
  }
  int n = 0;
  for (int i = 0; i < 64; ++i) {
    char path[64];
    snprintf(path, sizeof(path), "/dev/nvidia%d", i);
    if (exists(path)) ++n;
  }
  return n;
}

std::vector<std::string> visible_devices(int n) {
  std::vector<std::string> ids;
  if (const char* v = getenv("CUDA_VISIBLE_DEVICES")) {
    std::stringstream ss(v);
    std::string tok;
    while (std::getline(ss, tok, ',')) if (!tok.empty()) ids.push_back(tok);
  }
  for (int i = (int)ids.size(); i < n; ++i) ids.push_back(std::to_string(i));
  ids.resize((size_t)n);
  return ids;
}

// element count + byte offset of the body: "element vertex N" ... "end_header\n" for .ply, "PATCHES\nN\n" for .patch.
// The file is read once into `all`; the body is all[*body_at ..) (cluster models run to hundreds of megabytes: no copies).
bool split_model(const std::string& path, const char* kind, long long* count, std::string* all, size_t* body_at) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  struct stat st;
  if (fstat(fileno(f), &st) != 0) { fclose(f); return false; }
  all->resize((size_t)st.st_size);
  const size_t got = st.st_size ? fread(&(*all)[0], 1, (size_t)st.st_size, f) : 0;
  fclose(f);
  if (got != (size_t)st.st_size) return false;
  if (!strcmp(kind, "pset")) {
    long long n = 0;
    for (const char* p = all->data(), *e = p + all->size(); (p = (const char*)memchr(p, '\n', (size_t)(e - p))) != nullptr; ++p) ++n;
    *count = n; *body_at = 0;
    return true;
  }
  if (!strcmp(kind, "ply")) {
    const size_t ev = all->find("element vertex ");
    const size_t eh = all->find("end_header\n");
    if (ev == std::string::npos || eh == std::string::npos) return false;
    *count = atoll(all->c_str() + ev + 15);
    *body_at = eh + 11;
    return true;
  }
  if (all->compare(0, 8, "PATCHES\n") != 0) return false;
  const size_t nl = all->find('\n', 8);
  if (nl == std::string::npos) return false;
  *count = atoll(all->c_str() + 8);
  *body_at = nl + 1;
  return true;
}

}  // namespace

int main(int argc, char* argv[]) {
  if (argc < 2) {
    std::cerr << "Usage: " << argv[0] << " prefix [--gpus N] [--merge NAME] [--no-merge] [PATCH] [PSET]" << std::endl
              << "runs <prefix>option-0000, option-0001, ... (genOption's files) with pmvs2, one cluster per GPU at a time," << std::endl
              << "then concatenates the cluster models into <prefix>models/NAME.{ply,patch,pset} (NAME defaults to option-all)" << std::endl;
    return 1;
  }
  const std::string prefix(argv[1]);
  int gpus = -1;
  bool merge = true, patch = false, pset = false;
  std::string merged = "option-all";
  for (int i = 2; i < argc; ++i) {
    const std::string a(argv[i]);
    if (a == "--gpus" && i + 1 < argc) gpus = atoi(argv[++i]);
    else if (a == "--merge" && i + 1 < argc) merged = argv[++i];
    else if (a == "--no-merge") merge = false;
    else if (a == "PATCH") patch = true;
    else if (a == "PSET") pset = true;
    else { std::cerr << "pmvs2_clusters: unknown argument " << a << std::endl; return 1; }
  }
  if (gpus < 0) gpus = gpu_count_from_env();
  if (gpus < 1) { std::cerr << "pmvs2_clusters: no GPU visible (there is no CPU path)" << std::endl; return 1; }
  std::vector<std::string> options;
  for (int c = 0;; ++c) {
    char name[64];
    snprintf(name, sizeof(name), "option-%04d", c);
    if (!exists(prefix + name)) break;
    options.push_back(name);
  }
  if (options.empty()) { std::cerr << "pmvs2_clusters: no " << prefix << "option-0000 (run genOption first)" << std::endl; return 1; }
  const std::string pmvs2 = sibling(argv[0], "pmvs2");
  if (!exists(pmvs2)) { std::cerr << "pmvs2_clusters: " << pmvs2 << " not found" << std::endl; return 1; }
  mkdir((prefix + "models").c_str(), 0755);
  const std::vector<std::string> devices = visible_devices(gpus);

  // cluster -> GPU: a free slot takes the next cluster (clusters differ in size, so no fixed c mod N schedule)
  const auto t0 = std::chrono::steady_clock::now();
  std::map<pid_t, std::pair<int, int>> running;   // pid -> (cluster, slot)
  std::vector<int> free_slots;
  for (int s = gpus - 1; s >= 0; --s) free_slots.push_back(s);
  size_t next = 0;
  int failures = 0;
  while (next < options.size() || !running.empty()) {
    while (next < options.size() && !free_slots.empty()) {
      const int slot = free_slots.back();
      free_slots.pop_back();
      const int c = (int)next++;
      const pid_t pid = fork();
      if (pid < 0) { perror("fork"); return 1; }
      if (pid == 0) {
        setenv("CUDA_VISIBLE_DEVICES", devices[(size_t)slot].c_str(), 1);
        // the clusters that run side by side share the host cores (file parsing, bookkeeping, writers): an equal share each
        setenv("PMVSB_HOST_THREADS", std::to_string(std::max(1, (int)sysconf(_SC_NPROCESSORS_ONLN) / std::max(1, std::min(gpus, (int)options.size())))).c_str(), 0);
        unsetenv("WORLD_SIZE"); unsetenv("RANK"); unsetenv("LOCAL_RANK");   // a cluster is one single-GPU run
        const std::string log = prefix + "models/" + options[(size_t)c] + ".log";
        if (FILE* f = freopen(log.c_str(), "w", stdout)) { (void)f; dup2(fileno(stdout), fileno(stderr)); }
        std::vector<const char*> av = {pmvs2.c_str(), prefix.c_str(), options[(size_t)c].c_str()};
        if (patch) av.push_back("PATCH");
        if (pset) av.push_back("PSET");
        av.push_back(nullptr);
        execv(pmvs2.c_str(), const_cast<char* const*>(av.data()));
        perror("execv pmvs2");
        _exit(127);
      }
      running[pid] = std::make_pair(c, slot);
      std::cerr << "cluster " << c << " -> GPU " << devices[(size_t)slot] << " (pid " << pid << ")" << std::endl;
    }
    int status = 0;
    const pid_t done = wait(&status);
    if (done < 0) break;
    const auto it = running.find(done);
    if (it == running.end()) continue;
    const int c = it->second.first;
    free_slots.push_back(it->second.second);
    running.erase(it);
    const bool ok = WIFEXITED(status) && WEXITSTATUS(status) == 0;
    if (!ok) { ++failures; std::cerr << "cluster " << c << " FAILED, see " << prefix << "models/" << options[(size_t)c] << ".log" << std::endl; }
    else std::cerr << "cluster " << c << " done at " << std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() << " s" << std::endl;
  }
  if (failures) return 1;

  if (merge) {
    struct Kind { const char* ext; bool on; };
    const Kind kinds[3] = {{"ply", true}, {"patch", patch}, {"pset", pset}};
    for (const Kind& k : kinds) {
      if (!k.on) continue;
      long long total = 0;
      std::vector<std::string> files(options.size());
      std::vector<size_t> body_at(options.size(), 0);
      for (size_t c = 0; c < options.size(); ++c) {
        long long n = 0;
        const std::string path = prefix + "models/" + options[c] + "." + k.ext;
        if (!split_model(path, k.ext, &n, &files[c], &body_at[c])) { std::cerr << "pmvs2_clusters: cannot parse " << path << std::endl; return 1; }
        total += n;
      }
      std::ofstream out((prefix + "models/" + merged + "." + k.ext).c_str(), std::ios::binary);
      if (!strcmp(k.ext, "ply"))   // header of CPatchOrganizerS::writePLY (patchOrganizerS.cpp:693-706)
        out << "ply\nformat ascii 1.0\nelement vertex " << total << "\nproperty float x\nproperty float y\nproperty float z\nproperty float nx\n"
            << "property float ny\nproperty float nz\nproperty uchar diffuse_red\nproperty uchar diffuse_green\nproperty uchar diffuse_blue\n"
            << "property float quality\nend_header\n";
      else if (!strcmp(k.ext, "patch"))
        out << "PATCHES\n" << total << '\n';
      for (size_t c = 0; c < files.size(); ++c) out.write(files[c].data() + body_at[c], (std::streamsize)(files[c].size() - body_at[c]));
      std::cerr << "merged " << total << " patches of " << options.size() << " clusters into " << prefix << "models/" << merged << '.' << k.ext << std::endl;
    }
  }
  std::cerr << "time clusters.total " << std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() << " s on " << gpus << " GPU(s)" << std::endl;
  return 0;
}
