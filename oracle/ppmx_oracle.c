/*
 * ppmx_oracle.c -- TEST INFRASTRUCTURE ONLY (see ppmx_oracle.h).
 *
 * Plain-C restatement of the reference operators on flat packed rasters.
 * "ref:N" = /root/reference/ppmx-edward.c line N.  Build WITHOUT FMA contraction
 * (gcc -O2 -ffp-contract=off, no -march=native): the bicubic paths are
 * order- and rounding-sensitive.
 *
 * Parity: pinned against the compiled reference (oracle/_ref) by
 * tests/test_oracle_vs_ref.py and tests/golden/.
 */
#include "ppmx_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* the reference overrides round() with this (ref:27) */
static double rnd_half_up(double v) { return floor(v + 0.5); }

static const double ORC_PI = 3.14159265358979323846; /* ref:12 */

#define PX(buf, width, y, x) ((buf) + ((size_t)(y) * (size_t)(width) + (size_t)(x)) * 3u)

/* ------------------------------------------------------------------ gray */
int orc_gray(const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out_r_plane)
{
    size_t n = (size_t)w * h, i;
    for (i = 0; i < n; i++) {
        int sum = rgb[3 * i] + rgb[3 * i + 1] + rgb[3 * i + 2]; /* ref:1000, int arithmetic */
        out_r_plane[i] = (uint8_t)(sum / 3);
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ mono */
int orc_mono(const uint8_t *rgb, uint32_t w, uint32_t h, uint8_t *out_r_plane)
{
    /* ref:954 -- 4x4 Bayer thresholds, indexed x-major (ref:967) */
    static const double bayer[16] = {0.1250, 1.0000, 0.1875, 0.8125, 0.6250, 0.3750, 0.6875, 0.4375,
                                     0.2500, 0.8750, 0.0625, 0.9375, 0.7500, 0.5000, 0.5625, 0.3125};
    uint32_t x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++) {
            const uint8_t *p = PX(rgb, w, y, x);
            unsigned char grey = (unsigned char)((p[0] + p[1] + p[2]) / 3); /* ref:966 */
            double thr = bayer[(x % 4) * 4 + (y % 4)] * 255;               /* ref:967 */
            out_r_plane[(size_t)y * w + x] = (grey >= thr) ? 0 : 1;
        }
    return ORC_OK;
}

/* ------------------------------------------------------------------ P4 packer */
size_t orc_pack_pbm(const uint8_t *r_plane, uint32_t w, uint32_t h, uint8_t *out)
{
    size_t n = 0;
    uint32_t x, y;
    for (y = 0; y < h; y++) {
        int acc = 0;           /* ref:270, an int whose low byte is written */
        unsigned char slot = 1; /* ref:272 "tmp": 1-based bit slot in the current byte */
        for (x = 0; x < w; x++, slot++) {
            acc |= r_plane[(size_t)y * w + x] << (8 - slot); /* ref:273 */
            if (slot % 8 == 0) {                               /* ref:274-278 */
                out[n++] = (uint8_t)(acc & 0xff);
                slot = 0;
                acc = 0;
            }
        }
        if (slot - 1 != 0) out[n++] = (uint8_t)(acc & 0xff); /* ref:280-282, row padding */
    }
    return n;
}

/* ------------------------------------------------------------------ flip */
int orc_flip(uint8_t *rgb, uint32_t w, uint32_t h, int dir)
{
    uint32_t x, y;
    uint8_t t[3];
    if (dir) { /* vertical, ref:899-904 */
        for (y = 0; y < h / 2; y++)
            for (x = 0; x < w; x++) {
                uint8_t *a = PX(rgb, w, y, x), *b = PX(rgb, w, h - 1 - y, x);
                memcpy(t, a, 3); memcpy(a, b, 3); memcpy(b, t, 3);
            }
    } else { /* horizontal, ref:906-911 */
        for (y = 0; y < h; y++)
            for (x = 0; x < w / 2; x++) {
                uint8_t *a = PX(rgb, w, y, x), *b = PX(rgb, w, y, w - 1 - x);
                memcpy(t, a, 3); memcpy(a, b, 3); memcpy(b, t, 3);
            }
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ helpers */
double orc_cubic(double x)
{
    /* Keys cubic convolution, a = -0.5, ref:477-489.  Same association as the source. */
    double a1 = fabs(x), a2 = a1 * a1, a3 = a2 * a1, r = 0;
    if (a1 <= 1) r = (1.5 * a3) - (2.5 * a2) + 1;
    if ((1 < a1) && (a1 <= 2)) r = r + ((-0.5 * a3) + (2.5 * a2) - (4 * a1) + 2);
    return r;
}

int orc_mod(int a, int b)
{
    int r = 0; /* ref:495-500 */
    if (b != 0) r = a % b;
    return r < 0 ? r + b : r;
}

void orc_calc_rot_size(double angle, uint32_t w, uint32_t h, uint32_t *nw, uint32_t *nh)
{
    double t = (angle * ORC_PI) / 180.0; /* ref:653 */
    *nw = (uint32_t)rnd_half_up((w * cos(t)) + (h * sin(t))); /* ref:654 */
    *nh = (uint32_t)rnd_half_up((w * sin(t)) + (h * cos(t))); /* ref:655 */
}

void orc_rotate_size(double angle_deg, uint32_t w, uint32_t h, uint32_t *nw, uint32_t *nh)
{
    double a = angle_deg; /* fold into [0,90], ref:687-689 */
    if (a >= 270) a = 360 - a;
    else if (a > 180) a = a - 180;
    else if (a > 90) a = 180 - a;
    orc_calc_rot_size(a, w, h, nw, nh);
}

/* ------------------------------------------------------------------ rotate */
int orc_rotate(const uint8_t *rgb, uint32_t w, uint32_t h, double angle_deg, uint8_t *out)
{
    uint32_t nw, nh;
    int x, y;
    orc_rotate_size(angle_deg, w, h, &nw, &nh);

    if (angle_deg == 0) { /* ref:701-705: output aliases the input */
        memcpy(out, rgb, (size_t)w * h * 3);
        return ORC_OK;
    }
    memset(out, 0, (size_t)nw * nh * 3); /* image_buff_alloc zero-fills, ref:930 */

    if (angle_deg == 90) { /* ref:714-717 */
        for (y = 0; y < (int)h; y++)
            for (x = 0; x < (int)w; x++) memcpy(PX(out, nw, x, nw - y - 1), PX(rgb, w, y, x), 3);
    } else if (angle_deg == 180) { /* ref:718-721 */
        for (y = 0; y < (int)h; y++)
            for (x = 0; x < (int)w; x++) memcpy(PX(out, nw, nh - y - 1, nw - x - 1), PX(rgb, w, y, x), 3);
    } else if (angle_deg == 270) { /* ref:722-725 */
        for (y = 0; y < (int)nh; y++)
            for (x = 0; x < (int)nw; x++) memcpy(PX(out, nw, nh - y - 1, x), PX(rgb, w, x, y), 3);
    } else { /* ref:726-786 */
        double theta = (angle_deg * ORC_PI) / 180.0;                       /* ref:692 */
        int xc = (int)floor(w / 2), yc = (int)floor(h / 2);                /* ref:694-695 */
        int xo = (int)(floor(nw / 2) - floor(w / 2));                      /* ref:697 */
        int yo = (int)(floor(nh / 2) - floor(h / 2));                      /* ref:698 */
        for (y = 0; y < (int)nh; y++)
            for (x = 0; x < (int)nw; x++) {
                int x0 = (x - xo) - xc, y0 = (y - yo) - yc;                /* ref:731-735 */
                double nX = ((cos(theta) * (double)x0) + (sin(theta) * (double)y0) + xc);  /* ref:741 */
                double nY = (-(sin(theta) * (double)x0) + (cos(theta) * (double)y0) + yc); /* ref:742 */
                double rx = rnd_half_up(nX), ry = rnd_half_up(nY);
                uint8_t *dst = PX(out, nw, y, x);
                if (!((rx < w) && (ry < h) && (ry >= 0) && (rx >= 0))) continue; /* ref:744 */
                /* ref:752; w-2 / h-2 are unsigned in the source */
                if (rx > 1 && ry > 1 && rx < (double)(uint32_t)(w - 2) && ry < (double)(uint32_t)(h - 2)) {
                    double q[3] = {0.0, 0.0, 0.0};
                    int i, j, c;
                    for (j = 0; j < 4; j++) {
                        double p[3] = {0.0, 0.0, 0.0};
                        int v = (int)(floor(nY) - 1 + j); /* ref:758 */
                        for (i = 0; i < 4; i++) {
                            int u = (int)(floor(nX) - 1 + i); /* ref:761 */
                            const uint8_t *s = PX(rgb, w, v, u);
                            for (c = 0; c < 3; c++) p[c] += (s[c] * orc_cubic(nX - u)); /* ref:762-764 */
                        }
                        for (c = 0; c < 3; c++) q[c] += p[c] * orc_cubic(nY - v); /* ref:766-768 */
                    }
                    for (c = 0; c < 3; c++) {
                        if (q[c] < 0) q[c] = 0.0;      /* ref:771-773 */
                        if (q[c] >= 256) q[c] = 255.0; /* ref:775-777 */
                        dst[c] = (uint8_t)(int)q[c];   /* truncation, ref:779-781 */
                    }
                } else {
                    memcpy(dst, PX(rgb, w, (int)ry, (int)rx), 3); /* nearest, ref:783 */
                }
            }
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ resize */
void orc_free(void *p) { free(p); }

int orc_calc_contributions(int in_size, int out_size, double scale, double k_width,
                           int *taps, double **weights, int **indices)
{
    int P, x, y, kept = 0, n2 = in_size * 2;
    int *mirror = NULL, *idx = NULL, *oi = NULL;
    double *wt = NULL, *ow = NULL;
    unsigned char *keep = NULL;

    if (in_size < 1 || out_size < 1 || !(scale > 0)) return ORC_ERR; /* undefined in the reference */

    if (scale < 1.0) k_width = k_width / scale; /* ref:533 */
    P = (int)ceil(k_width) + 2;                 /* ref:535 */

    mirror = (int *)malloc((size_t)n2 * sizeof(int));
    idx = (int *)malloc((size_t)out_size * P * sizeof(int));
    wt = (double *)malloc((size_t)out_size * P * sizeof(double));
    keep = (unsigned char *)calloc((size_t)P, 1);
    if (!mirror || !idx || !wt || !keep) goto fail;

    /* symmetric extension table 0..n-1,n-1..0, ref:551-555 */
    for (x = 0; x < in_size; x++) { mirror[x] = x; mirror[n2 - 1 - x] = x; }

    for (y = 0; y < out_size; y++) {
        double u = ((y + 1) / scale) + (0.5 * (1 - (1 / scale))); /* ref:562 */
        double sum = 0.0;
        for (x = 0; x < P; x++) {
            int id = (int)(floor(u - (k_width / 2)) + (x - 1)); /* ref:563 */
            double wv;
            if (scale < 1.0) wv = scale * orc_cubic((u - (double)id - 1) * scale); /* ref:571-572 */
            else wv = orc_cubic(u - (double)id - 1);                               /* ref:578 */
            idx[(size_t)y * P + x] = id;
            wt[(size_t)y * P + x] = wv;
        }
        for (x = 0; x < P; x++) sum += wt[(size_t)y * P + x];  /* ref:583 */
        for (x = 0; x < P; x++) wt[(size_t)y * P + x] /= sum;  /* ref:584 */
        for (x = 0; x < P; x++)                                /* ref:589 */
            idx[(size_t)y * P + x] = mirror[orc_mod(idx[(size_t)y * P + x], n2)];
    }

    /* only ROW 0 decides which tap columns survive, ref:591-602 */
    for (x = 0; x < P; x++)
        if (wt[x] != 0.0f) { keep[x] = 1; kept++; }

    ow = (double *)calloc((size_t)out_size * (kept ? kept : 1), sizeof(double));
    oi = (int *)calloc((size_t)out_size * (kept ? kept : 1), sizeof(int));
    if (!ow || !oi) goto fail;
    for (y = 0; y < out_size; y++) { /* ref:616-624 */
        int k = 0;
        for (x = 0; x < P; x++)
            if (keep[x]) {
                oi[(size_t)y * kept + k] = idx[(size_t)y * P + x];
                ow[(size_t)y * kept + k] = wt[(size_t)y * P + x];
                k++;
            }
    }
    free(mirror); free(idx); free(wt); free(keep);
    *taps = kept; *weights = ow; *indices = oi;
    return ORC_OK;
fail:
    free(mirror); free(idx); free(wt); free(keep); free(ow); free(oi);
    return ORC_ERR;
}

static uint8_t quantise(double s)
{
    s = rnd_half_up(s); /* ref:831-837 */
    return (s < 0.0f) ? 0 : (s >= 256) ? 255 : (uint8_t)(int)s;
}

int orc_imresize(const uint8_t *rgb, uint32_t w, uint32_t h, int out_size, int dim,
                 const double *weights, const int *indices, int taps, uint8_t *out)
{
    int x, y, z, c;
    if (dim == 0) { /* height pass, ref:814-839 */
        for (y = 0; y < out_size; y++)
            for (x = 0; x < (int)w; x++) {
                double s[3] = {0.0, 0.0, 0.0};
                for (z = 0; z < taps; z++) {
                    const uint8_t *p = PX(rgb, w, indices[(size_t)y * taps + z], x);
                    for (c = 0; c < 3; c++) s[c] = s[c] + (p[c] * weights[(size_t)y * taps + z]);
                }
                for (c = 0; c < 3; c++) PX(out, w, y, x)[c] = quantise(s[c]);
            }
    } else { /* width pass, ref:840-869 */
        for (y = 0; y < (int)h; y++)
            for (x = 0; x < out_size; x++) {
                double s[3] = {0.0, 0.0, 0.0};
                for (z = 0; z < taps; z++) {
                    const uint8_t *p = PX(rgb, w, y, indices[(size_t)x * taps + z]);
                    for (c = 0; c < 3; c++) s[c] = s[c] + p[c] * weights[(size_t)x * taps + z];
                }
                for (c = 0; c < 3; c++) PX(out, out_size, y, x)[c] = quantise(s[c]);
            }
    }
    return ORC_OK;
}

int orc_resize(const uint8_t *rgb, uint32_t w, uint32_t h, uint32_t new_w,
               uint8_t **out, uint32_t *out_w, uint32_t *out_h)
{
    double scale[2];
    uint32_t new_h;
    int taps[2] = {0, 0}, first, second, rc = ORC_ERR;
    double *wt[2] = {NULL, NULL};
    int *ix[2] = {NULL, NULL};
    uint8_t *mid = NULL, *fin = NULL;

    if ((int)new_w < 1 || w < 1 || h < 1) return ORC_ERR;  /* ref:1096 */
    scale[1] = (double)((double)new_w / w);                 /* ref:1098 */
    new_h = (uint32_t)((double)h * scale[1]);               /* ref:1099, truncation */
    if (new_h < 1) return ORC_ERR;                          /* reference divides by zero here */
    scale[0] = (double)((double)new_h / h);                 /* ref:1100 */
    if (scale[0] < scale[1]) { first = 0; second = 1; }     /* ref:1102-1103 */
    else { first = 1; second = 0; }

    if (orc_calc_contributions((int)h, (int)new_h, scale[0], 4.0, &taps[0], &wt[0], &ix[0])) goto done;
    if (orc_calc_contributions((int)w, (int)new_w, scale[1], 4.0, &taps[1], &wt[1], &ix[1])) goto done;

    if (first == 0) { /* height then width; the intermediate is 8-bit (ref:1118) */
        mid = (uint8_t *)malloc((size_t)w * new_h * 3 + 1);
        fin = (uint8_t *)malloc((size_t)new_w * new_h * 3 + 1);
        if (!mid || !fin) goto done;
        orc_imresize(rgb, w, h, (int)new_h, 0, wt[0], ix[0], taps[0], mid);
        orc_imresize(mid, w, new_h, (int)new_w, 1, wt[1], ix[1], taps[1], fin);
    } else {
        mid = (uint8_t *)malloc((size_t)new_w * h * 3 + 1);
        fin = (uint8_t *)malloc((size_t)new_w * new_h * 3 + 1);
        if (!mid || !fin) goto done;
        orc_imresize(rgb, w, h, (int)new_w, 1, wt[1], ix[1], taps[1], mid);
        orc_imresize(mid, new_w, h, (int)new_h, 0, wt[0], ix[0], taps[0], fin);
    }
    (void)second;
    *out = fin; fin = NULL;
    *out_w = new_w; *out_h = new_h;
    rc = ORC_OK;
done:
    free(mid); free(fin);
    free(wt[0]); free(wt[1]); free(ix[0]); free(ix[1]);
    return rc;
}

/* ------------------------------------------------------------------ pipeline */
typedef struct {
    uint8_t *px; /* packed rgb */
    uint32_t w, h;
} orc_img;

static uint8_t *plane_to_rgb(const uint8_t *plane, size_t n)
{
    /* gray/mono fill only .r of a zeroed pixel buffer (ref:996-1000, 961-968) */
    uint8_t *o = (uint8_t *)calloc(n * 3 + 1, 1);
    size_t i;
    if (o) for (i = 0; i < n; i++) o[3 * i] = plane[i];
    return o;
}

int orc_process(const uint8_t *rgb, uint32_t w, uint32_t h, const orc_flags *f,
                uint8_t **out, size_t *out_bytes, uint32_t *out_w, uint32_t *out_h, int *file_type)
{
    /* buff / new_buff hand-over exactly as ref:1084-1155: renewBuffer before
     * rotate iff -w, before gray/mono/flip iff -w or -r.  new_buff may alias buff. */
    orc_img cur = {NULL, w, h}, nxt = {NULL, 0, 0};
    int ft = ORC_FT_PPM, rc = ORC_ERR;
    int renew = f->resize_enable || f->rotate_enable;
    size_t n;

    cur.px = (uint8_t *)malloc((size_t)w * h * 3 + 1);
    if (!cur.px) return ORC_ERR;
    memcpy(cur.px, rgb, (size_t)w * h * 3);

#define RENEW() do { if (nxt.px != cur.px) free(cur.px); cur = nxt; nxt.px = NULL; } while (0)

    if (f->resize_enable) {
        if (orc_resize(cur.px, cur.w, cur.h, f->resize_w, &nxt.px, &nxt.w, &nxt.h)) goto done;
    }
    if (f->rotate_enable) {
        if (f->resize_enable) RENEW();
        orc_rotate_size((double)f->angle, cur.w, cur.h, &nxt.w, &nxt.h);
        if (f->angle == 0) { nxt.px = cur.px; nxt.w = cur.w; nxt.h = cur.h; }
        else {
            nxt.px = (uint8_t *)malloc((size_t)nxt.w * nxt.h * 3 + 1);
            if (!nxt.px) goto done;
            orc_rotate(cur.px, cur.w, cur.h, (double)f->angle, nxt.px);
        }
    }
    if (f->gray_enable || f->mono_enable) {
        uint8_t *plane;
        if (renew) RENEW();
        n = (size_t)cur.w * cur.h;
        plane = (uint8_t *)malloc(n + 1);
        if (!plane) goto done;
        if (f->gray_enable) { orc_gray(cur.px, cur.w, cur.h, plane); ft = ORC_FT_PGM; }
        else { orc_mono(cur.px, cur.w, cur.h, plane); ft = ORC_FT_PBM; }
        if (nxt.px && nxt.px != cur.px) free(nxt.px);
        nxt.px = plane_to_rgb(plane, n);
        nxt.w = cur.w; nxt.h = cur.h;
        free(plane);
        if (!nxt.px) goto done;
    }
    if (f->flipv_enable || f->fliph_enable) {
        if (renew) RENEW();
        /* flip works on buff in place and aliases new_buff to it (ref:896); without
         * -w/-r a preceding gray/mono result is simply dropped (leaked in the source) */
        if (nxt.px && nxt.px != cur.px) free(nxt.px);
        orc_flip(cur.px, cur.w, cur.h, f->flipv_enable ? 1 : 0);
        nxt = cur;
    }
    if (!nxt.px) goto done; /* "no data to write", ref:235 */

    n = (size_t)nxt.w * nxt.h;
    if (ft == ORC_FT_PGM) { /* ref:263-267: only .r is written */
        size_t i;
        uint8_t *o = (uint8_t *)malloc(n + 1);
        if (!o) goto done;
        for (i = 0; i < n; i++) o[i] = nxt.px[3 * i];
        *out = o; *out_bytes = n;
    } else if (ft == ORC_FT_PBM) { /* ref:268-284 */
        size_t i;
        uint8_t *plane = (uint8_t *)malloc(n + 1), *o = (uint8_t *)malloc((size_t)nxt.h * ((nxt.w + 7) / 8) + 1);
        if (!plane || !o) { free(plane); free(o); goto done; }
        for (i = 0; i < n; i++) plane[i] = nxt.px[3 * i];
        *out_bytes = orc_pack_pbm(plane, nxt.w, nxt.h, o);
        *out = o;
        free(plane);
    } else { /* ref:285-291 */
        uint8_t *o = (uint8_t *)malloc(n * 3 + 1);
        if (!o) goto done;
        memcpy(o, nxt.px, n * 3);
        *out = o; *out_bytes = n * 3;
    }
    *out_w = nxt.w; *out_h = nxt.h; *file_type = ft;
    rc = ORC_OK;
done:
    if (nxt.px && nxt.px != cur.px) free(nxt.px);
    free(cur.px);
    return rc;
#undef RENEW
}

int orc_header(char *dst, size_t cap, int file_type, uint32_t w, uint32_t h, uint32_t maxval)
{
    const char *magic = file_type == ORC_FT_PGM ? "P5" : file_type == ORC_FT_PBM ? "P4" : "P6"; /* ref:239-247 */
    if (file_type == ORC_FT_PBM) /* no maxval line for P4, ref:258 */
        return snprintf(dst, cap, "%s\n# generated by ppmx_edward\n%u %u\n", magic, w, h);
    return snprintf(dst, cap, "%s\n# generated by ppmx_edward\n%u %u\n%u\n", magic, w, h, maxval);
}

/* ------------------------------------------------------------------ extensions (unpinned) */
static int mirror_index(int i, int n)
{
    /* same symmetric extension as the resize tables (ref:551-555,589) */
    int m = orc_mod(i, 2 * n);
    return m < n ? m : 2 * n - 1 - m;
}

static int64_t floor_div(int64_t a, int64_t b) /* b > 0 */
{
    int64_t q = a / b;
    if ((a % b != 0) && (a < 0)) q--;
    return q;
}

int orx_conv(const uint8_t *rgb, uint32_t w, uint32_t h, int k, const int32_t *coef,
             int32_t div, int32_t bias, uint8_t *out)
{
    int r = k / 2, x, y, dx, dy, c;
    if (k < 1 || !(k & 1) || div < 1) return ORC_ERR;
    for (y = 0; y < (int)h; y++)
        for (x = 0; x < (int)w; x++) {
            int64_t acc[3] = {0, 0, 0};
            for (dy = 0; dy < k; dy++) {
                int sy = mirror_index(y + dy - r, (int)h);
                for (dx = 0; dx < k; dx++) {
                    const uint8_t *p = PX(rgb, w, sy, mirror_index(x + dx - r, (int)w));
                    int32_t cf = coef[dy * k + dx];
                    for (c = 0; c < 3; c++) acc[c] += (int64_t)cf * p[c];
                }
            }
            for (c = 0; c < 3; c++) {
                /* floor(acc/div + 0.5) in integers, then bias, then the ref:835 clamp */
                int64_t v = floor_div(2 * acc[c] + div, 2 * (int64_t)div) + bias;
                PX(out, w, y, x)[c] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
            }
        }
    return ORC_OK;
}

int orx_hist_gray(const uint8_t *rgb, uint32_t w, uint32_t h, uint64_t *bins)
{
    size_t n = (size_t)w * h, i;
    memset(bins, 0, 256 * sizeof(uint64_t));
    for (i = 0; i < n; i++) bins[(rgb[3 * i] + rgb[3 * i + 1] + rgb[3 * i + 2]) / 3]++;
    return ORC_OK;
}

int orx_levels(const uint8_t *src, size_t nbytes, const uint8_t *lut, uint8_t *out)
{
    size_t i;
    for (i = 0; i < nbytes; i++) out[i] = lut[src[i]];
    return ORC_OK;
}

int orx_levels_lut_linear(int lo, int hi, uint8_t *lut)
{
    int v;
    if (lo < 0 || hi > 255 || lo >= hi) return ORC_ERR;
    for (v = 0; v < 256; v++) {
        double x = floor((double)(v - lo) * 255.0 / (double)(hi - lo) + 0.5); /* the round() of ref:27 */
        lut[v] = (uint8_t)(v <= lo ? 0 : v >= hi ? 255 : (int)x);
    }
    return ORC_OK;
}

void orc_lcg_fill(uint8_t *rgb, size_t npix, uint32_t seed)
{
    uint32_t s = seed;
    size_t i;
    for (i = 0; i < npix; i++) {
        s = s * 1664525u + 1013904223u;
        rgb[3 * i] = (uint8_t)(s >> 24);
        rgb[3 * i + 1] = (uint8_t)(s >> 16);
        rgb[3 * i + 2] = (uint8_t)(s >> 8);
    }
}
