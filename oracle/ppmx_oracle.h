/*
 * ppmx_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the per-pixel operators of the reference program
 * (/root/reference/ppmx-edward.c, cited below as ref:LINE).  It exists so the
 * CUDA path can be checked bit for bit; nothing in the product
 * (imageprocessingtools_b200/, include/) may include, link or call it.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity status: PINNED.  Every function here is compared byte for byte with
 * the reference itself, compiled from its own source into oracle/_ref/ by
 * oracle/Makefile (tests/test_oracle_vs_ref.py), and with the golden digests
 * that tests/golden/make_golden.py recorded from that compiled reference.
 * The reference ships no tests or vectors of its own.
 *
 * Extension operators (orx_*: k x k convolution, histogram) have NO reference
 * counterpart -- "parity unpinned", self-oracle only.
 *
 * All images are flat, packed, row-major: 3 bytes per pixel (r,g,b) unless a
 * parameter says "plane" (1 byte per pixel, the .r member of the reference's
 * pixel struct).
 */
#ifndef PPMX_ORACLE_H
#define PPMX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_OK 0
#define ORC_ERR (-1)

#define ORC_FT_PPM 0 /* ref:22 */
#define ORC_FT_PGM 1 /* ref:23 */
#define ORC_FT_PBM 2 /* ref:24 */

/* ref:986-1003.  out_r_plane[w*h] = (r+g+b)/3 (the .r member; .g/.b stay 0). */
int orc_gray(const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out_r_plane);

/* ref:949-971.  out_r_plane[w*h] in {0,1}. */
int orc_mono(const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out_r_plane);

/* ref:268-284.  Packs arbitrary .r bytes the way the P4 writer does; returns the
 * number of bytes written ( h * ceil(w/8) ). */
size_t orc_pack_pbm(const uint8_t *r_plane, uint32_t w, uint32_t h, uint8_t *out);

/* ref:888-913.  In place.  dir 1 = vertical, 0 = horizontal. */
int orc_flip(uint8_t *rgb, uint32_t w, uint32_t h, int dir);

/* ref:477-489, 495-500 */
double orc_cubic(double x);
int orc_mod(int a, int b);

/* ref:649-656 (raw formula) and ref:687-691 (angle folding + call). */
void orc_calc_rot_size(double angle, uint32_t w, uint32_t h, uint32_t *nw, uint32_t *nh);
void orc_rotate_size(double angle_deg, uint32_t w, uint32_t h, uint32_t *nw, uint32_t *nh);

/* ref:673-789.  out must hold nw*nh*3 bytes with (nw,nh) from orc_rotate_size;
 * angle 0 copies the input (the reference aliases the buffer, ref:701-705). */
int orc_rotate(const uint8_t *rgb, uint32_t w, uint32_t h, double angle_deg, uint8_t *out);

/* ref:516-641.  Flat tables: weights[out_size*taps], indices[out_size*taps],
 * malloc'd here, released with orc_free. */
int orc_calc_contributions(int in_size, int out_size, double scale, double k_width,
                           int *taps, double **weights, int **indices);
void orc_free(void *p);

/* ref:808-872.  dim 0: height pass (out is out_size x w), dim 1: width pass
 * (out is h x out_size). */
int orc_imresize(const uint8_t *rgb, uint32_t w, uint32_t h, int out_size, int dim,
                 const double *weights, const int *indices, int taps, uint8_t *out);

/* ref:1096-1120: the "-wN" driver section.  *out is malloc'd. */
int orc_resize(const uint8_t *rgb, uint32_t w, uint32_t h, uint32_t new_w,
               uint8_t **out, uint32_t *out_w, uint32_t *out_h);

/* Whole pipeline on a decoded raster, ref:1084-1155 + raster part of ref:263-291.
 * flags are 0/1; resize_w and angle as given to -w / -r.  *out is malloc'd and
 * holds exactly the bytes the reference writes after the header. */
typedef struct orc_flags {
    int resize_enable, rotate_enable, flipv_enable, fliph_enable, gray_enable, mono_enable;
    uint32_t resize_w;
    int angle;
} orc_flags;
int orc_process(const uint8_t *rgb, uint32_t w, uint32_t h, const orc_flags *f,
                uint8_t **out, size_t *out_bytes, uint32_t *out_w, uint32_t *out_h, int *file_type);

/* ref:239-261: header text for the output file; returns its length. */
int orc_header(char *dst, size_t cap, int file_type, uint32_t w, uint32_t h, uint32_t maxval);

/* ---- extension operators: NO reference counterpart, parity unpinned ---- */

/* k x k integer convolution on each channel. border: symmetric mirror (the aux table
 * idiom of ref:551-555,589); result = clamp(floor(acc/div + 0.5) + bias) done in
 * integers as floor_div(2*acc + div, 2*div) + bias; clamp <0 -> 0, >255 -> 255.
 * coef is k*k int32 row-major, div > 0.  y0/full_h let a row band be computed:
 * rgb points at the full image, output rows [y0, y0+band_h) are produced. */
int orx_conv(const uint8_t *rgb, uint32_t w, uint32_t h, int k, const int32_t *coef,
             int32_t div, int32_t bias, uint8_t *out);

/* 256-bin histogram of (r+g+b)/3 (grey of ref:1000). bins[256] u64, overwritten. */
int orx_hist_gray(const uint8_t *rgb, uint32_t w, uint32_t h, uint64_t *bins);

/* levels: out[i] = lut[in[i]] for every byte.  orx_levels_lut_linear: lut[v] = 0 for v <= lo, 255 for v >= hi,
 * else round((v - lo) * 255.0 / (hi - lo)) with round(x) = floor(x + 0.5) (ref:27), here written with doubles the
 * way the reference writes its roundings (the product library evaluates the same value in integers). */
int orx_levels(const uint8_t *src, size_t nbytes, const uint8_t *lut, uint8_t *out);
int orx_levels_lut_linear(int lo, int hi, uint8_t *lut);

/* synthetic input generator shared by tests and bench:
 * s = s*1664525 + 1013904223; r = s>>24, g = s>>16, b = s>>8. */
void orc_lcg_fill(uint8_t *rgb, size_t npix, uint32_t seed);

#ifdef __cplusplus
}
#endif
#endif
