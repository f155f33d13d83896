// oracle/shim/CImg.h -- TEST INFRASTRUCTURE.  Stand-in for CImg 2.7.5 covering the calls made at
// /root/reference/source/image/image.cpp:476-498,676-701,750-784: load / load_jpeg / is_empty /
// width / height / spectrum / size / operator()(x,y,z,c) and CImgException.
// Only binary PNM (P6 colour, P5 grey) is decoded -- the synthetic scenes are PPM, i.e. lossless.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace cimg_library {
struct CImgException {
  std::string msg;
  explicit CImgException(const std::string& m) : msg(m) {}
  const char* what() const { return msg.c_str(); }
};

template <typename T>
class CImg {
 public:
  CImg() : _w(0), _h(0), _c(0) {}
  CImg& load(const char* file) { return load_pnm(file); }
  CImg& load_jpeg(const char* file) { throw CImgException(std::string("no jpeg in shim: ") + file); }
  CImg& load_pnm(const char* file) {
    FILE* fp = std::fopen(file, "rb");
    if (!fp) throw CImgException(std::string("cannot open ") + file);
    char magic[3] = {0, 0, 0};
    int w = 0, h = 0, maxv = 0;
    bool ok = (std::fscanf(fp, "%2s", magic) == 1) && next_int(fp, w) && next_int(fp, h) && next_int(fp, maxv);
    int ch = (ok && std::strcmp(magic, "P6") == 0) ? 3 : ((ok && std::strcmp(magic, "P5") == 0) ? 1 : 0);
    if (!ok || ch == 0 || maxv != 255 || w <= 0 || h <= 0) {
      std::fclose(fp);
      throw CImgException(std::string("not a binary 8-bit PNM: ") + file);
    }
    std::fgetc(fp);  // single whitespace after maxval
    std::vector<unsigned char> raw((size_t)w * h * ch);
    size_t got = std::fread(raw.data(), 1, raw.size(), fp);
    std::fclose(fp);
    if (got != raw.size()) throw CImgException(std::string("short read: ") + file);
    _w = w; _h = h; _c = ch;
    _data.resize(raw.size());
    // CImg stores planar (x fastest, then y, then channel)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x)
        for (int c = 0; c < ch; ++c)
          _data[((size_t)c * h + y) * w + x] = (T)raw[((size_t)y * w + x) * ch + c];
    return *this;
  }
  bool is_empty() const { return _data.empty(); }
  int width() const { return _w; }
  int height() const { return _h; }
  int spectrum() const { return _c; }
  size_t size() const { return _data.size(); }
  T operator()(int x, int y, int /*z*/, int c) const { return _data[((size_t)c * _h + y) * _w + x]; }

 private:
  static bool next_int(FILE* fp, int& v) {
    for (;;) {
      int ch = std::fgetc(fp);
      if (ch == EOF) return false;
      if (ch == '#') { while (ch != '\n' && ch != EOF) ch = std::fgetc(fp); continue; }
      if (ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r') continue;
      std::ungetc(ch, fp);
      break;
    }
    return std::fscanf(fp, "%d", &v) == 1;
  }
  int _w, _h, _c;
  std::vector<T> _data;
};
}  // namespace cimg_library
