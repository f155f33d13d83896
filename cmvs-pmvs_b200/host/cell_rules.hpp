// cmvs-pmvs_b200/host/cell_rules.hpp -- the per-cell trial rules of the expansion round, on plain arrays.
//
// CExpand::checkCounts and updateCounts (/root/reference/source/pmvs/expand.cpp:258-323, 325-406) read and bump
// CPatchOrganizerS::_counts and ask whether a cell of _pgrids holds a patch.  The host keeps exactly that per image: the
// occupancy of _pgrids and the unsigned-char trial counters (north_star: "the Pgrid/cell bookkeeping stays in C++ on the host").
// Free functions so that the pipeline and the test hooks (cell_rules_abi.cpp -> lib/libpmvs_host.so) run the same code.
#pragma once

namespace pmvs {

struct CellGridView {        // one image's grid
  int gw, gh;
  const int* occ;            // patches per cell of _pgrids (target images)
  unsigned char* counts;     // CPatchOrganizerS::_counts
};

// CExpand::checkCounts (expand.cpp:258-323): true = the candidate is rejected.  grids = (x, y) per image entry.
inline bool check_counts(const int* images, const int* grids, int n, int tnum, const CellGridView* g, int count_threshold1, int min_image_num, int depth) {
  int full = 0, empty = 0;
  for (int i = 0; i < n; ++i) {
    const int im = images[i];
    if (tnum <= im) continue;
    const int ix = grids[2 * i], iy = grids[2 * i + 1];
    if (ix < 0 || g[im].gw <= ix || iy < 0 || g[im].gh <= iy) continue;
    const int c = iy * g[im].gw + ix;
    if (g[im].occ[c] != 0) { ++full; continue; }
    if (count_threshold1 <= g[im].counts[c]) ++full; else ++empty;
  }
  if (depth <= 1) return empty < min_image_num && full != 0;     // the first expansion is expensive: stricter
  return empty < min_image_num - 1 && full != 0;
}

// CExpand::updateCounts (expand.cpp:325-406): bumps the counters of the patch's cells in _images (target images) and _vimages;
// true = the new patch joins the expansion queue (some cell was still below the threshold).
inline bool update_counts(const int* images, const int* grids, int n, const int* vimages, const int* vgrids, int nv, int tnum, const CellGridView* g,
                          int count_threshold1) {
  int empty = 0;
  auto visit = [&](int im, int ix, int iy) {
    if (ix < 0 || g[im].gw <= ix || iy < 0 || g[im].gh <= iy) return;
    const int c = iy * g[im].gw + ix;
    if (!(count_threshold1 <= g[im].counts[c])) ++empty;
    ++g[im].counts[c];       // unsigned char: wraps like the reference's
  };
  for (int i = 0; i < n; ++i)
    if (images[i] < tnum) visit(images[i], grids[2 * i], grids[2 * i + 1]);
  for (int i = 0; i < nv; ++i) visit(vimages[i], vgrids[2 * i], vgrids[2 * i + 1]);
  return empty != 0;
}

}  // namespace pmvs
