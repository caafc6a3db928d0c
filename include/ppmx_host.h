/*
 * ppmx_host.h -- the host side, in C like the reference: the same operator functions, the
 * same handler fields, the same PPM in / P6-P5-P4 out, with every pixel loop replaced by a
 * call into the CUDA ABI of ppmx_gpu.h.
 *
 * "ref:N" = /root/reference/ppmx-edward.c line N.  Function names are the reference's with
 * a ppmx_ prefix; argument meaning, return values (0 / -1) and the one-line stdout messages
 * follow the reference so its callers and tests read the same.  Rasters live in HBM: the
 * handler's buff / new_buff are device rasters, not row-pointer arrays.
 *
 * Everything that needs libm stays here, on the host, so tables and sizes are bit-identical
 * to the reference's glibc results: cubic (ref:477), mod (ref:495), calc_contributions
 * (ref:516), calc_rot_size (ref:649).
 */
#ifndef PPMX_HOST_H
#define PPMX_HOST_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#include "ppmx_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ref:59-66 */
typedef struct ppmx_args_flag {
    char resize_enable, rotate_enable, flipv_enable, fliph_enable, gray_enable, mono_enable;
} ppmx_args_flag;

/* ref:46-56, with device rasters in place of pixel** */
typedef struct ppmx_img_info {
    unsigned int height, width, new_height, new_width, max_color, size, file_type;
    ppmx_gpu_image *buff;
    ppmx_gpu_image *new_buff; /* may alias buff (flip, ref:896; rotate 0, ref:703) */
} ppmx_img_info;

/* ref:76-88 */
typedef struct ppmx_image_handler {
    ppmx_img_info imginfo;
    ppmx_args_flag arg_flag;
    ppmx_gpu_ctx *ctx;          /* new: the device context every operator runs on */
    unsigned char *file_buffer; /* whole input file, in pinned memory */
    const char *filename;
    size_t filesize;
    size_t index_buffer;        /* offset of the raster in file_buffer after the header */
    unsigned int output_width_size;
    double angle;
    char norotate;
    int conv_preset;            /* extension flags -blur/-blur7/-sharpen/-edge/-gauss5/-gauss7/-gauss9/-sharpen7 (0 = none); not in the reference */
    int levels_enable, levels_lo, levels_hi; /* extension flag -levelsLO-HI (black/white point); not in the reference */
} ppmx_image_handler;

/* ref:92-96, flat: weights[out_size][weights_sz], indices likewise */
typedef struct ppmx_contributions {
    double *weights;
    int32_t *indices;
    int weights_sz;
    int out_size;
} ppmx_contributions;

/* ---- host-only mathematics ----------------------------------------------------------- */
double ppmx_cubic(double x);   /* ref:477-489 */
int ppmx_mod(int a, int b);    /* ref:495-500 */
int ppmx_calc_contributions(int in_size, int out_size, double scale, double k_width,
                            ppmx_contributions *contrib);                 /* ref:516-641 */
void ppmx_contributions_free(ppmx_contributions *contrib);
void ppmx_calc_rot_size(double angle, unsigned int old_width, unsigned int old_height,
                        unsigned int *new_width, unsigned int *new_height); /* ref:649-656 */

/* ---- the operators: int op(handler), reading imginfo.buff, producing imginfo.new_buff -- */
int ppmx_gray(ppmx_image_handler *handler);                          /* ref:986-1003 */
int ppmx_mono(ppmx_image_handler *handler);                          /* ref:949-971  */
int ppmx_flip(ppmx_image_handler *handler, unsigned char flip_direction); /* ref:888-913 */
int ppmx_rotate(ppmx_image_handler *handler);                        /* ref:673-789  */
int ppmx_imresize(ppmx_image_handler *handler, int out_size, int dim, const double *weights,
                  const int32_t *indices, int weights_sz);           /* ref:808-872  */
void ppmx_renewBuffer(ppmx_image_handler *handler);                  /* ref:1019-1026 */

/* ---- the pipeline --------------------------------------------------------------------- */

/* A ready-to-run op chain for ppmx_gpu_apply: the flag order and renewBuffer conditions of
 * ref:1084-1155, with rotate geometry and resize tables already computed on the host. */
#define PPMX_PLAN_MAX_OPS 12 /* resize 2 + rotate + conv + levels + gray/mono + one flip = 7 at most today */
typedef struct ppmx_plan {
    ppmx_op ops[PPMX_PLAN_MAX_OPS];
    int nops;
    ppmx_contributions contrib[2]; /* owned tables of the two resize passes */
    unsigned char levels_lut[256]; /* table of the extension levels stage, if any */
} ppmx_plan;
int ppmx_plan_chain(const ppmx_args_flag *flags, unsigned int output_width_size, double angle,
                    unsigned int width, unsigned int height, ppmx_plan *plan);
/* The same with one EXTENSION stage (no reference counterpart): a preset k x k convolution inserted after
 * resize/rotate and before gray/mono/flip.  Presets: 1 blur 3x3 (1 2 1; 2 4 2; 1 2 1)/16, 2 box blur 7x7 /49,
 * 3 sharpen (0 -1 0; -1 5 -1; 0 -1 0), 4 edge (-1 ... 8 ... -1).  0 = none = ppmx_plan_chain. */
#define PPMX_CONV_NONE 0
#define PPMX_CONV_BLUR3 1
#define PPMX_CONV_BLUR7 2
#define PPMX_CONV_SHARPEN 3
#define PPMX_CONV_EDGE 4
#define PPMX_CONV_GAUSS5 5   /* -gauss5: binomial (1 4 6 4 1)^2 / 256 */
#define PPMX_CONV_GAUSS7 6   /* -gauss7: binomial (1 6 15 20 15 6 1)^2 / 4096 */
#define PPMX_CONV_SHARPEN7 7 /* -sharpen7: unsharp mask 2 I - gauss7 */
#define PPMX_CONV_GAUSS9 8   /* -gauss9: binomial (1 8 28 56 70 56 28 8 1)^2 / 65536 */
int ppmx_plan_chain_ext(const ppmx_args_flag *flags, unsigned int output_width_size, double angle,
                        unsigned int width, unsigned int height, int conv_preset, ppmx_plan *plan);
/* The same plus a second EXTENSION stage after the convolution: levels with black point lo and white point hi
 * (0 <= lo < hi <= 255; lo < 0 = no levels stage), table = ppmx_levels_lut_linear(lo, hi). */
int ppmx_plan_chain_ext2(const ppmx_args_flag *flags, unsigned int output_width_size, double angle,
                         unsigned int width, unsigned int height, int conv_preset, int levels_lo, int levels_hi,
                         ppmx_plan *plan);
/* EXTENSION helpers for PPMX_OP_LEVELS (no reference counterpart).  lut[v] = 0 below lo, 255 above hi and
 * round((v - lo) * 255 / (hi - lo)) in between, with the reference's round(x) = floor(x + 0.5) (ref:27) evaluated
 * in integers.  Returns -1 unless 0 <= lo < hi <= 255. */
int ppmx_levels_lut_linear(int lo, int hi, unsigned char lut[256]);
/* Black and white points from a 256-bin histogram (PPMX_OP_HIST_GRAY): the smallest lo / largest hi such that at
 * most clip_permille / 1000 of the pixels lie below lo / above hi.  Returns -1 for an empty or flat histogram. */
int ppmx_levels_points_from_hist(const unsigned long long hist[256], unsigned int clip_permille, int *lo, int *hi);
void ppmx_plan_free(ppmx_plan *plan);

/* Header tokenizer of ref:333-456 on an in-memory file: width, height, maxval and the offset
 * of the raster.  Same acceptance rules and messages (P6 only, comments, exact size). */
int ppmx_parse_header(const unsigned char *file, size_t filesize, unsigned int *width,
                      unsigned int *height, unsigned int *max_color, size_t *raster_offset);
/* EXTENSION (the reference rejects both, ref:386, 453): ASCII P3 and 16-bit samples.  probe reads the header of a
 * P3 or P6 file of any maxval 1..65535; decode writes width * height * 3 bytes to dst -- samples as they are for
 * maxval <= 255 (header maxval kept, like the reference does for P6, ref:259), scaled to 0..255 with
 * round(v * 255 / maxval) (ref:27 rounding, in integers) for larger maxvals (*out_max_color = 255). */
#define PPMX_PNM_P6_8 0
#define PPMX_PNM_P6_16 1
#define PPMX_PNM_P3 2
int ppmx_probe_pnm(const unsigned char *file, size_t filesize, unsigned int *width, unsigned int *height,
                   unsigned int *max_color, size_t *raster_offset, int *format);
int ppmx_decode_pnm(const unsigned char *file, size_t filesize, size_t raster_offset, int format, unsigned int width,
                    unsigned int height, unsigned int max_color, unsigned char *dst, unsigned int *out_max_color);

/* Header text of ref:239-261 ("# generated by ppmx_edward"); returns its length. */
int ppmx_format_header(char *dst, size_t cap, int file_type, unsigned int width,
                       unsigned int height, unsigned int max_color);

/* Row-band partition of a raster of full_h rows over nranks GPUs: contiguous bands
 * whose starts are multiples of `align` rows (4 keeps the Bayer phase of mono on band boundaries,
 * 1 otherwise), sizes differing by at most one unit.  Rank `rank` gets rows [*y0, *y0 + *rows);
 * trailing ranks may get 0 rows when full_h is small.  No reference counterpart. */
int ppmx_band_plan(unsigned int full_h, int nranks, int rank, unsigned int align, unsigned int *y0,
                   unsigned int *rows);

/* Synthetic raster for benchmarks and tests (no reference counterpart): the LCG
 * s = s * 1664525 + 1013904223 (mod 2^32), one step per pixel, r = s >> 24, g = s >> 16, b = s >> 8.
 * Writes pixels [first, first + npix) of the sequence started from `seed` (a rank fills only its rows). */
void ppmx_synth_lcg(unsigned char *rgb, size_t first, size_t npix, uint32_t seed);

int ppmx_getImageInfo(ppmx_image_handler *handler);   /* ref:409-456: parse + upload        */
int ppmx_putImageToFile(ppmx_image_handler *handler); /* ref:221-301: download + one fwrite */
int ppmx_doProcessPPM(ppmx_image_handler *handler);   /* ref:1053-1172                      */
/* EXTENSION (the reference takes exactly one file, ref:180; the command line offers this as -batch): the handler's
 * flags applied to every file, `<name>.out` beside each, on ONE device context, with file n+1 being read into pinned
 * memory and result n-1 being written while the device works on file n.  Returns -1 if any file failed. */
int ppmx_doProcessBatch(ppmx_image_handler *handler, const char *const *filenames, int nfiles);
void ppmx_usage(void);                                /* ref:194-205 */
int ppmx_main(int argc, char *argv[]);                /* ref:117-191 */

#ifdef __cplusplus
}
#endif
#endif /* PPMX_HOST_H */
