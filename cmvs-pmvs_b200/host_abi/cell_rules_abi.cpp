// cmvs-pmvs_b200/host_abi/cell_rules_abi.cpp -- C hooks around the host-side cell rules (lib/libpmvs_host.so, no CUDA), so that
// the tests can hold the code pmvs2 runs (host/cell_rules.hpp) to the reference's CExpand::checkCounts / updateCounts.
// Grids of the target images are passed flattened: cell_base[i] = first cell of image i (cell_base[tnum] = total).
#include <vector>

#include "../host/cell_rules.hpp"

namespace {
std::vector<pmvs::CellGridView> views(int tnum, const int* gw, const int* gh, const int* cell_base, const int* occ, unsigned char* counts) {
  std::vector<pmvs::CellGridView> v(tnum);
  for (int i = 0; i < tnum; ++i) v[i] = {gw[i], gh[i], occ + cell_base[i], counts + cell_base[i]};
  return v;
}
}  // namespace

extern "C" {

// P candidates with CSR lists (off[P+1], images, grids = 2 per entry); verdict[k] = 1 when CExpand::checkCounts rejects
int pmvsh_check_counts_batch(int tnum, const int* gw, const int* gh, const int* cell_base, const int* occ, const unsigned char* counts, int P, const int* off,
                             const int* images, const int* grids, int count_threshold1, int min_image_num, int depth, unsigned char* verdict) {
  const auto v = views(tnum, gw, gh, cell_base, occ, const_cast<unsigned char*>(counts));
  for (int k = 0; k < P; ++k)
    verdict[k] = pmvs::check_counts(images + off[k], grids + 2 * off[k], off[k + 1] - off[k], tnum, v.data(), count_threshold1, min_image_num, depth) ? 1 : 0;
  return 0;
}

// CExpand::updateCounts for P patches IN ORDER (the counters change as it goes); requeue[k] = its return value
int pmvsh_update_counts_batch(int tnum, const int* gw, const int* gh, const int* cell_base, const int* occ, unsigned char* counts, int P, const int* off,
                              const int* images, const int* grids, const int* voff, const int* vimages, const int* vgrids, int count_threshold1,
                              unsigned char* requeue) {
  const auto v = views(tnum, gw, gh, cell_base, occ, counts);
  for (int k = 0; k < P; ++k)
    requeue[k] = pmvs::update_counts(images + off[k], grids + 2 * off[k], off[k + 1] - off[k], vimages + voff[k], vgrids + 2 * voff[k], voff[k + 1] - voff[k], tnum,
                                     v.data(), count_threshold1) ? 1 : 0;
  return 0;
}

}  // extern "C"
