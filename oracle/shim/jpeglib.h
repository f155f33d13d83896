/* oracle/shim/jpeglib.h -- TEST INFRASTRUCTURE.  Declaration-only stand-in for libjpeg-turbo's
 * header so that /root/reference/source/image/image.cpp (writeJpegImage, lines 786-830, never
 * called by pmvs) compiles.  The functions abort if ever reached. */
#ifndef PMVS_ORACLE_JPEGLIB_SHIM_H
#define PMVS_ORACLE_JPEGLIB_SHIM_H
#include <stdio.h>
#include <stdlib.h>
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#define METHODDEF(type) static type
typedef unsigned char JSAMPLE;
typedef JSAMPLE* JSAMPROW;
typedef JSAMPROW* JSAMPARRAY;
typedef unsigned int JDIMENSION;
typedef int boolean;
typedef enum { JCS_UNKNOWN, JCS_GRAYSCALE, JCS_RGB } J_COLOR_SPACE;
struct jpeg_common_struct;
typedef struct jpeg_common_struct* j_common_ptr;
struct jpeg_error_mgr {
  void (*error_exit)(j_common_ptr);
  void (*output_message)(j_common_ptr);
};
struct jpeg_common_struct { struct jpeg_error_mgr* err; };
struct jpeg_compress_struct {
  struct jpeg_error_mgr* err;
  JDIMENSION image_width, image_height, next_scanline;
  int input_components;
  J_COLOR_SPACE in_color_space;
};
typedef struct jpeg_compress_struct* j_compress_ptr;
static inline void pmvs_jpeg_shim_die(void) { fprintf(stderr, "jpeg shim reached\n"); abort(); }
static inline struct jpeg_error_mgr* jpeg_std_error(struct jpeg_error_mgr* e) { return e; }
static inline void jpeg_create_compress(j_compress_ptr) { pmvs_jpeg_shim_die(); }
static inline void jpeg_stdio_dest(j_compress_ptr, FILE*) { pmvs_jpeg_shim_die(); }
static inline void jpeg_set_defaults(j_compress_ptr) { pmvs_jpeg_shim_die(); }
static inline void jpeg_set_quality(j_compress_ptr, int, boolean) { pmvs_jpeg_shim_die(); }
static inline void jpeg_start_compress(j_compress_ptr, boolean) { pmvs_jpeg_shim_die(); }
static inline JDIMENSION jpeg_write_scanlines(j_compress_ptr, JSAMPARRAY, JDIMENSION) { pmvs_jpeg_shim_die(); return 0; }
static inline void jpeg_finish_compress(j_compress_ptr) { pmvs_jpeg_shim_die(); }
static inline void jpeg_destroy_compress(j_compress_ptr) { pmvs_jpeg_shim_die(); }
#endif
