/*
 * ref_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Flat-buffer entry points around the UNMODIFIED reference translation unit.
 * The reference source is not copied into this repository: it is #included from
 * where it lies (-DREF_SOURCE="\"/root/reference/ppmx-edward.c\"", see oracle/Makefile)
 * and the result goes to oracle/_ref/libppmx_ref.so only.  Its main() is renamed so
 * the object can live in a shared library.
 *
 * Every ref_* function converts a packed raster to the reference's own
 * row-pointer layout (with the reference's allocator), calls the reference
 * function, flattens the answer and optionally reports the seconds spent inside
 * the reference call alone (op_seconds), which is what bench.py quotes as the
 * CPU baseline ("kind": "reference").
 */
#ifndef REF_SOURCE
#error "build with -DREF_SOURCE=\"/path/to/ppmx-edward.c\""
#endif

#define main ppmx_ref_cli_main
#include REF_SOURCE
#undef main
#undef round /* the reference's macro (its line 27) must not leak into the code below */

#include <stdint.h>
#include <time.h>

static double now_s(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static int load_rows(ppm_image_handler *hd, const uint8_t *rgb, uint32_t w, uint32_t h)
{
    uint32_t y;
    memset(hd, 0, sizeof(*hd));
    if (image_buff_alloc(&hd->imginfo.buff, h, w) != PPM_NOERROR) return -1;
    for (y = 0; y < h; y++) memcpy(hd->imginfo.buff[y], rgb + (size_t)y * w * 3, (size_t)w * 3);
    hd->imginfo.width = w;
    hd->imginfo.height = h;
    hd->imginfo.max_color = 255;
    hd->imginfo.file_type = FILETYPE_PPM;
    return 0;
}

static void store_rows(pixel **rows, uint32_t w, uint32_t h, uint8_t *out_rgb)
{
    uint32_t y;
    for (y = 0; y < h; y++) memcpy(out_rgb + (size_t)y * w * 3, rows[y], (size_t)w * 3);
}

static void drop(ppm_image_handler *hd)
{
    if (hd->imginfo.new_buff && hd->imginfo.new_buff != hd->imginfo.buff)
        releaseBuffer(&hd->imginfo.new_buff, hd->imginfo.new_height);
    if (hd->imginfo.buff) releaseBuffer(&hd->imginfo.buff, hd->imginfo.height);
}

/* out_rgb receives the full pixel structs (r,g,b) of new_buff: w*h*3 bytes. */
int ref_gray(const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out_rgb, int *file_type, double *op_seconds)
{
    ppm_image_handler hd;
    double t0;
    int rc;
    if (load_rows(&hd, rgb, w, h)) return -1;
    t0 = now_s();
    rc = gray(&hd);
    if (op_seconds) *op_seconds = now_s() - t0;
    if (rc == PPM_NOERROR) store_rows(hd.imginfo.new_buff, hd.imginfo.new_width, hd.imginfo.new_height, out_rgb);
    if (file_type) *file_type = (int)hd.imginfo.file_type;
    drop(&hd);
    return rc;
}

int ref_mono(const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out_rgb, int *file_type, double *op_seconds)
{
    ppm_image_handler hd;
    double t0;
    int rc;
    if (load_rows(&hd, rgb, w, h)) return -1;
    t0 = now_s();
    rc = mono(&hd);
    if (op_seconds) *op_seconds = now_s() - t0;
    if (rc == PPM_NOERROR) store_rows(hd.imginfo.new_buff, hd.imginfo.new_width, hd.imginfo.new_height, out_rgb);
    if (file_type) *file_type = (int)hd.imginfo.file_type;
    drop(&hd);
    return rc;
}

int ref_flip(const uint8_t *rgb, uint32_t w, uint32_t h, int dir, uint8_t *out_rgb, double *op_seconds)
{
    ppm_image_handler hd;
    double t0;
    int rc;
    if (load_rows(&hd, rgb, w, h)) return -1;
    t0 = now_s();
    rc = flip(&hd, (unsigned char)dir);
    if (op_seconds) *op_seconds = now_s() - t0;
    if (rc == PPM_NOERROR) store_rows(hd.imginfo.new_buff, hd.imginfo.new_width, hd.imginfo.new_height, out_rgb);
    drop(&hd);
    return rc;
}

void ref_rotate_size(double angle_deg, uint32_t w, uint32_t h, uint32_t *nw, uint32_t *nh)
{
    double a = angle_deg; /* same folding the reference applies before its size call */
    if (a >= 270) a = 360 - a;
    else if (a > 180) a = a - 180;
    else if (a > 90) a = 180 - a;
    calc_rot_size(a, w, h, nw, nh);
}

/* out_rgb must hold ref_rotate_size() pixels. */
int ref_rotate(const uint8_t *rgb, uint32_t w, uint32_t h, double angle_deg, uint8_t *out_rgb,
               uint32_t *nw, uint32_t *nh, double *op_seconds)
{
    ppm_image_handler hd;
    double t0;
    int rc;
    if (load_rows(&hd, rgb, w, h)) return -1;
    hd.angle = angle_deg;
    t0 = now_s();
    rc = rotate(&hd);
    if (op_seconds) *op_seconds = now_s() - t0;
    if (rc == PPM_NOERROR) {
        if (hd.norotate) { hd.imginfo.new_width = w; hd.imginfo.new_height = h; }
        store_rows(hd.imginfo.new_buff, hd.imginfo.new_width, hd.imginfo.new_height, out_rgb);
        if (nw) *nw = hd.imginfo.new_width;
        if (nh) *nh = hd.imginfo.new_height;
    }
    drop(&hd);
    return rc;
}

double ref_cubic(double x) { return cubic(x); }
int ref_mod(int a, int b) { return mod(a, b); }

/* flat copies of the reference's contribution tables; caller frees with ref_free */
int ref_calc_contributions(int in_size, int out_size, double scale, double k_width,
                           int *taps, double **weights, int **indices)
{
    contributions c;
    int y, z;
    double *fw;
    int *fi;
    memset(&c, 0, sizeof(c));
    if (calc_contributions(in_size, out_size, scale, k_width, &c) != PPM_NOERROR) return -1;
    fw = (double *)malloc(sizeof(double) * (size_t)out_size * (size_t)(c.weights_sz ? c.weights_sz : 1));
    fi = (int *)malloc(sizeof(int) * (size_t)out_size * (size_t)(c.weights_sz ? c.weights_sz : 1));
    for (y = 0; y < out_size; y++) {
        for (z = 0; z < c.weights_sz; z++) {
            fw[(size_t)y * c.weights_sz + z] = c.weights[y][z];
            fi[(size_t)y * c.weights_sz + z] = c.indices[y][z];
        }
        free(c.weights[y]);
        free(c.indices[y]);
    }
    free(c.weights);
    free(c.indices);
    *taps = c.weights_sz;
    *weights = fw;
    *indices = fi;
    return 0;
}

void ref_free(void *p) { free(p); }

int ref_imresize(const uint8_t *rgb, uint32_t w, uint32_t h, int out_size, int dim,
                 const double *weights, const int *indices, int taps, uint8_t *out_rgb, double *op_seconds)
{
    ppm_image_handler hd;
    double **wr;
    int **ir;
    int y, rc;
    double t0;
    if (load_rows(&hd, rgb, w, h)) return -1;
    wr = (double **)malloc(sizeof(double *) * (size_t)out_size);
    ir = (int **)malloc(sizeof(int *) * (size_t)out_size);
    for (y = 0; y < out_size; y++) {
        wr[y] = (double *)weights + (size_t)y * taps;
        ir[y] = (int *)indices + (size_t)y * taps;
    }
    t0 = now_s();
    rc = imresize(&hd, out_size, dim, wr, ir, taps);
    if (op_seconds) *op_seconds = now_s() - t0;
    if (rc == PPM_NOERROR) store_rows(hd.imginfo.new_buff, hd.imginfo.new_width, hd.imginfo.new_height, out_rgb);
    free(wr);
    free(ir);
    drop(&hd);
    return rc;
}

int ref_sizeof_pixel(void) { return (int)sizeof(pixel); }
