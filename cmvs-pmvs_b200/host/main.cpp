// cmvs-pmvs_b200/host/main.cpp -- `pmvs2 prefix option_file [PATCH] [PSET]`, the reference binary's command line
// (/root/reference/source/pmvs.cpp:7-63; genOption's scripts call it pmvs2, this fork's CMake target is pmvs3).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include <unistd.h>

#include "pmvs_host.hpp"

int main(int argc, char* argv[]) {
  if (argc < 3) {
    std::cerr << "Usage: " << argv[0] << " prefix option_file [Optional export]" << std::endl << std::endl
              << "--------------------------------------------------" << std::endl
              << "level       1    csize    2" << std::endl
              << "threshold   0.7  wsize    7" << std::endl
              << "minImageNum 3    CPU      4" << std::endl
              << "useVisData  0    sequence -1" << std::endl
              << "quad        2.5  maxAngle 10.0" << std::endl
              << "--------------------------------------------------" << std::endl
              << "2 ways to specify targetting images" << std::endl
              << "timages  5  1 3 5 7 9 (enumeration)" << std::endl
              << "        -1  0 24 (range specification)" << std::endl
              << "--------------------------------------------------" << std::endl
              << "4 ways to specify other images" << std::endl
              << "oimages  5  0 2 4 6 8 (enumeration)" << std::endl
              << "        -1  24 48 (range specification)" << std::endl << std::endl
              << "[Optional export] PATCH PSET" << std::endl;
    return 1;
  }
  const auto t_start = std::chrono::steady_clock::now();
  for (int i = 0; i < argc; ++i) std::cout << std::endl << argv[i];
  std::cout << std::endl;
  const pmvs::Options opt = pmvs::parse_options(argv[1], argv[2]);
  pmvs::Dist dist = pmvs::Dist::from_env();   // WORLD_SIZE > 1: one process per GPU (distributed.cpp)
  // rank 0 reports; the other ranks' streams go to a sink that is never destroyed (the iostream library flushes cerr / cout
  // at exit, after function-local statics are gone)
  if (dist.rank != 0) { std::ofstream* quiet = new std::ofstream("/dev/null"); std::cerr.rdbuf(quiet->rdbuf()); std::cout.rdbuf(quiet->rdbuf()); }
  dist.connect();
  pmvs::Pipeline pipe(opt, dist);
  pipe.load();
  pipe.run();
  bool patch = false, pset = false;
  for (int i = 3; i < argc; ++i) {
    const std::string a(argv[i]);
    if (a == "PATCH") patch = true;
    if (a == "PSET") pset = true;
  }
  pipe.write(std::string(argv[1]) + "models/" + argv[2], true, patch, pset);   // rank 0 writes
  std::cerr << "time main.total " << std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() << " s" << std::endl;
  // The models are on disk and every stream is flushed.  Leave without unwinding: freeing a few thousand device allocations one
  // by one and tearing the CUDA context down from user space costs ~0.5 s of a ~3 s run, and the driver releases everything
  // when the process goes away anyway.
  std::cout.flush();
  std::cerr.flush();
  std::fflush(nullptr);
  _exit(0);
}
