/*
 * ppmx_host.c -- host side of ppmx-b200, in C like the reference (see include/ppmx_host.h).
 *
 * "ref:N" = /root/reference/ppmx-edward.c line N.  This file holds what the reference does
 * on the CPU that is NOT a pixel loop: flag handling, the P6 tokenizer, header writing, the
 * op ordering of doProcessPPM, and the libm-dependent tables (contributions, rotation size).
 * Every pixel loop is a call into libppmx_gpu.so.  Build with -O2 -ffp-contract=off: the
 * contribution weights must be bit-identical to the reference's (no FMA contraction).
 */
#include "../../include/ppmx_host.h"

#include <ctype.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define PPMX_PI 3.14159265358979323846 /* the reference's own literal, ref:12 */

/* one-line message on stdout, then -1: the reference's CHECK_ERROR convention (ref:31-36) */
#define BAIL(msg)            \
    do {                     \
        printf("%s", msg);   \
        return PPMX_ERROR;   \
    } while (0)

static double half_up(double v) { return floor(v + 0.5); } /* the reference's round(), ref:27 */

/* ------------------------------------------------------------------ host mathematics */

double ppmx_cubic(double x)
{
    /* Keys kernel, a = -0.5; the grouping of every product and sum is the source's (ref:484-486) */
    const double t = fabs(x), t2 = t * t, t3 = t2 * t;
    double k = 0;
    if (t <= 1) k = (1.5 * t3) - (2.5 * t2) + 1;
    if ((1 < t) && (t <= 2)) k = k + ((-0.5 * t3) + (2.5 * t2) - (4 * t) + 2);
    return k;
}

int ppmx_mod(int a, int b)
{
    int m = (b != 0) ? a % b : 0;
    return m < 0 ? m + b : m;
}

void ppmx_calc_rot_size(double angle, unsigned int old_width, unsigned int old_height,
                        unsigned int *new_width, unsigned int *new_height)
{
    const double th = (angle * PPMX_PI) / 180.0;
    *new_width = (unsigned int)half_up((old_width * cos(th)) + (old_height * sin(th)));
    *new_height = (unsigned int)half_up((old_width * sin(th)) + (old_height * cos(th)));
}

/* the size rotate() asks for: the angle is folded into [0,90] first (ref:687-691) */
static void rotated_size(double angle, unsigned int w, unsigned int h, unsigned int *nw, unsigned int *nh)
{
    double a = angle;
    if (a >= 270) a = 360 - a;
    else if (a > 180) a = a - 180;
    else if (a > 90) a = 180 - a;
    ppmx_calc_rot_size(a, w, h, nw, nh);
}

/* One row of the contribution table before pruning: P candidate taps for output index y.
 * ref:558-589 does this in five separate sweeps over the whole table; per row the values are
 * the same because no step looks at another row. */
static void contribution_row(int y, int in_size, double scale, double kw, int P, double *w, int *id)
{
    const double u = ((y + 1) / scale) + (0.5 * (1 - (1 / scale))); /* ref:562 */
    const double left = floor(u - (kw / 2));
    double sum = 0.0;
    int x;
    for (x = 0; x < P; x++) {
        id[x] = (int)(left + (x - 1)); /* ref:563 */
        if (scale < 1.0) w[x] = scale * ppmx_cubic((u - (double)id[x] - 1) * scale); /* ref:571-572 */
        else w[x] = ppmx_cubic(u - (double)id[x] - 1);                               /* ref:578 */
    }
    for (x = 0; x < P; x++) sum += w[x]; /* ref:583 */
    for (x = 0; x < P; x++) w[x] /= sum; /* ref:584 */
    for (x = 0; x < P; x++) {            /* symmetric extension, ref:551-555 + 589 */
        int m = ppmx_mod(id[x], 2 * in_size);
        id[x] = (m < in_size) ? m : 2 * in_size - 1 - m;
    }
}

int ppmx_calc_contributions(int in_size, int out_size, double scale, double k_width, ppmx_contributions *c)
{
    double kw = k_width, *row_w = NULL;
    int P, x, y, kept = 0, *row_i = NULL;
    unsigned char *keep = NULL;

    memset(c, 0, sizeof(*c));
    if (in_size < 1 || out_size < 1 || !(scale > 0)) BAIL("error: allocating ind2store\n"); /* what ref:595 ends in */
    if (scale < 1.0) kw = kw / scale; /* ref:533 */
    P = (int)ceil(kw) + 2;            /* ref:535 */

    row_w = (double *)malloc(sizeof(double) * (size_t)P);
    row_i = (int *)malloc(sizeof(int) * (size_t)P);
    keep = (unsigned char *)calloc((size_t)P, 1);
    if (!row_w || !row_i || !keep) goto nomem;

    /* which tap columns survive is decided by output index 0 alone (ref:591-602) */
    contribution_row(0, in_size, scale, kw, P, row_w, row_i);
    for (x = 0; x < P; x++)
        if (row_w[x] != 0.0f) { keep[x] = 1; kept++; }

    c->weights = (double *)calloc((size_t)out_size * (size_t)(kept ? kept : 1), sizeof(double));
    c->indices = (int32_t *)calloc((size_t)out_size * (size_t)(kept ? kept : 1), sizeof(int32_t));
    if (!c->weights || !c->indices) goto nomem;
    for (y = 0; y < out_size; y++) { /* ref:616-624 */
        int k = 0;
        contribution_row(y, in_size, scale, kw, P, row_w, row_i);
        for (x = 0; x < P; x++)
            if (keep[x]) {
                c->weights[(size_t)y * kept + k] = row_w[x];
                c->indices[(size_t)y * kept + k] = row_i[x];
                k++;
            }
    }
    c->weights_sz = kept;
    c->out_size = out_size;
    free(row_w); free(row_i); free(keep);
    return PPMX_OK;
nomem:
    free(row_w); free(row_i); free(keep);
    ppmx_contributions_free(c);
    BAIL("error. allocating memory for weights and indices\n");
}

void ppmx_contributions_free(ppmx_contributions *c)
{
    if (!c) return;
    free(c->weights);
    free(c->indices);
    memset(c, 0, sizeof(*c));
}

/* ------------------------------------------------------------------ operators on device rasters */

static void set_new(ppmx_image_handler *h, ppmx_gpu_image *img)
{
    unsigned int w = 0, hh = 0;
    /* a superseded result (the reference leaks it, e.g. gray's output in "-gray -fh") goes back to the pool */
    if (h->imginfo.new_buff && h->imginfo.new_buff != h->imginfo.buff) ppmx_gpu_image_free(h->ctx, h->imginfo.new_buff);
    h->imginfo.new_buff = img;
    ppmx_gpu_image_info(img, &w, &hh, NULL, NULL, NULL);
    h->imginfo.new_width = w;
    h->imginfo.new_height = hh;
}

static int run_simple(ppmx_image_handler *h, int kind, int flip_dir)
{
    ppmx_op op;
    ppmx_gpu_image *out = NULL;
    memset(&op, 0, sizeof(op));
    op.kind = kind;
    op.flip_direction = flip_dir;
    if (!h || !h->ctx || !h->imginfo.buff) BAIL("Error: no image loaded\n");
    if (ppmx_gpu_op(h->ctx, &op, h->imginfo.buff, &out, NULL) != PPMX_OK) return PPMX_ERROR;
    set_new(h, out);
    return PPMX_OK;
}

int ppmx_gray(ppmx_image_handler *h)
{
    if (run_simple(h, PPMX_OP_GRAY, 0) != PPMX_OK) return PPMX_ERROR;
    h->imginfo.file_type = PPMX_FILETYPE_PGM; /* ref:991 */
    return PPMX_OK;
}

int ppmx_mono(ppmx_image_handler *h)
{
    if (run_simple(h, PPMX_OP_MONO, 0) != PPMX_OK) return PPMX_ERROR;
    h->imginfo.file_type = PPMX_FILETYPE_PBM; /* ref:956 */
    return PPMX_OK;
}

int ppmx_flip(ppmx_image_handler *h, unsigned char flip_direction)
{
    /* the reference swaps inside buff and points new_buff at it (ref:896); the device kernel
     * writes a fresh raster which then takes buff's place */
    if (run_simple(h, PPMX_OP_FLIP, flip_direction ? 1 : 0) != PPMX_OK) return PPMX_ERROR;
    ppmx_gpu_image_free(h->ctx, h->imginfo.buff);
    h->imginfo.buff = h->imginfo.new_buff;
    return PPMX_OK;
}

static void fill_rotate_op(ppmx_op *op, double angle, unsigned int w, unsigned int h)
{
    const double th = (angle * PPMX_PI) / 180.0; /* ref:692 */
    memset(op, 0, sizeof(*op));
    op->kind = PPMX_OP_ROTATE;
    op->angle_deg = (int32_t)angle;
    op->cos_t = cos(th); /* the two loop-invariant libm values of ref:741-742 */
    op->sin_t = sin(th);
    rotated_size(angle, w, h, &op->new_width, &op->new_height);
}

int ppmx_rotate(ppmx_image_handler *h)
{
    ppmx_op op;
    ppmx_gpu_image *out = NULL;
    if (!h || !h->ctx || !h->imginfo.buff) BAIL("Error: no image loaded\n");
    fill_rotate_op(&op, h->angle, h->imginfo.width, h->imginfo.height);
    h->imginfo.new_width = op.new_width;
    h->imginfo.new_height = op.new_height;
    if (h->angle == 0) { /* ref:701-705 */
        h->norotate = 1;
        h->imginfo.new_buff = h->imginfo.buff;
        return PPMX_OK;
    }
    if (ppmx_gpu_op(h->ctx, &op, h->imginfo.buff, &out, NULL) != PPMX_OK) return PPMX_ERROR;
    set_new(h, out);
    return PPMX_OK;
}

int ppmx_imresize(ppmx_image_handler *h, int out_size, int dim, const double *weights, const int32_t *indices,
                  int weights_sz)
{
    ppmx_op op;
    ppmx_gpu_image *out = NULL;
    if (!h || !h->ctx || !h->imginfo.buff) BAIL("Error: no image loaded\n");
    memset(&op, 0, sizeof(op));
    op.kind = PPMX_OP_IMRESIZE;
    op.dim = dim;
    op.out_size = out_size;
    op.weights_sz = weights_sz;
    op.weights = weights;
    op.indices = indices;
    if (ppmx_gpu_op(h->ctx, &op, h->imginfo.buff, &out, NULL) != PPMX_OK) return PPMX_ERROR;
    set_new(h, out);
    return PPMX_OK;
}

void ppmx_renewBuffer(ppmx_image_handler *h)
{
    if (h->imginfo.buff && h->imginfo.buff != h->imginfo.new_buff) ppmx_gpu_image_free(h->ctx, h->imginfo.buff);
    h->imginfo.buff = h->imginfo.new_buff;
    h->imginfo.height = h->imginfo.new_height;
    h->imginfo.width = h->imginfo.new_width;
    h->imginfo.new_buff = NULL;
}

/* ------------------------------------------------------------------ the op chain as data */

/* extension presets (no reference counterpart) */
static const int32_t k_blur3[9] = {1, 2, 1, 2, 4, 2, 1, 2, 1};
static const int32_t k_sharpen[9] = {0, -1, 0, -1, 5, -1, 0, -1, 0};
static const int32_t k_edge[9] = {-1, -1, -1, -1, 8, -1, -1, -1, -1};
/* binomial blurs and the 7x7 unsharp mask, filled on first use (outer products of a row of Pascal's triangle) */
static int32_t k_gauss5[25], k_gauss7[49], k_sharpen7[49], k_gauss9[81];
static void fill_binomial_presets(void)
{
    static const int32_t b5[5] = {1, 4, 6, 4, 1}, b7[7] = {1, 6, 15, 20, 15, 6, 1}, b9[9] = {1, 8, 28, 56, 70, 56, 28, 8, 1};
    int i, j;
    if (k_gauss5[0]) return; /* (idempotent: a racing second caller writes the same values) */
    for (i = 0; i < 7; i++)
        for (j = 0; j < 7; j++) {
            k_gauss7[i * 7 + j] = b7[i] * b7[j];
            k_sharpen7[i * 7 + j] = -b7[i] * b7[j] + (i == 3 && j == 3 ? 2 * 4096 : 0);
        }
    for (i = 0; i < 9; i++)
        for (j = 0; j < 9; j++) k_gauss9[i * 9 + j] = b9[i] * b9[j];
    for (i = 4; i >= 0; i--)
        for (j = 4; j >= 0; j--) k_gauss5[i * 5 + j] = b5[i] * b5[j]; /* (element 0 last: it is the "filled" flag) */
}
static const int32_t k_box7[49] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                   1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};

int ppmx_plan_chain(const ppmx_args_flag *f, unsigned int output_width_size, double angle, unsigned int width,
                    unsigned int height, ppmx_plan *plan)
{
    return ppmx_plan_chain_ext(f, output_width_size, angle, width, height, PPMX_CONV_NONE, plan);
}

int ppmx_plan_chain_ext(const ppmx_args_flag *f, unsigned int output_width_size, double angle, unsigned int width,
                        unsigned int height, int conv_preset, ppmx_plan *plan)
{
    return ppmx_plan_chain_ext2(f, output_width_size, angle, width, height, conv_preset, -1, -1, plan);
}

/* EXTENSION: levels table, round(x) = floor(x + 0.5) as ref:27, in integers */
int ppmx_levels_lut_linear(int lo, int hi, unsigned char lut[256])
{
    int v;
    if (lo < 0 || hi > 255 || lo >= hi) return PPMX_ERROR;
    for (v = 0; v < 256; v++) {
        if (v <= lo) lut[v] = 0;
        else if (v >= hi) lut[v] = 255;
        else lut[v] = (unsigned char)((2 * (v - lo) * 255 + (hi - lo)) / (2 * (hi - lo)));
    }
    return PPMX_OK;
}

int ppmx_levels_points_from_hist(const unsigned long long hist[256], unsigned int clip_permille, int *lo, int *hi)
{
    unsigned long long total = 0, allow, acc;
    int v, a, b;
    for (v = 0; v < 256; v++) total += hist[v];
    if (!total || clip_permille > 499) return PPMX_ERROR;
    allow = total / 1000 * clip_permille + total % 1000 * clip_permille / 1000;
    for (a = 0, acc = 0; a < 255 && acc + hist[a] <= allow; a++) acc += hist[a];
    for (b = 255, acc = 0; b > 0 && acc + hist[b] <= allow; b--) acc += hist[b];
    if (a >= b) return PPMX_ERROR;
    *lo = a;
    *hi = b;
    return PPMX_OK;
}

int ppmx_plan_chain_ext2(const ppmx_args_flag *f, unsigned int output_width_size, double angle, unsigned int width,
                         unsigned int height, int conv_preset, int levels_lo, int levels_hi, ppmx_plan *plan)
{
    unsigned int w = width, h = height;
    /* the condition of ref:1138,1143,1148,1153; an extension stage counts like -w / -r */
    int renew = f->resize_enable || f->rotate_enable;
    int n = 0;
    memset(plan, 0, sizeof(*plan));
    /* the reference's command line refuses these pairs (ref:130, 135, 166, 171); a caller of the C planner gets
     * the same answer instead of a chain the reference can not express (and the op array stays within its bound) */
    if ((f->gray_enable && f->mono_enable) || (f->flipv_enable && f->fliph_enable))
        BAIL("Error: Conflicting options not allowed\n");

    if (f->resize_enable) { /* ref:1084-1130 */
        double scale[2];
        unsigned int new_w = output_width_size, new_h;
        int first, second, pass;
        if ((int)new_w < 1) BAIL("invalid option for new width\n"); /* ref:1096 */
        scale[1] = (double)((double)new_w / w);                      /* ref:1098 */
        new_h = (unsigned int)((double)h * scale[1]);                /* ref:1099 */
        scale[0] = (double)((double)new_h / h);                      /* ref:1100 */
        if (scale[0] < scale[1]) { first = 0; second = 1; }          /* ref:1102-1103 */
        else { first = 1; second = 0; }
        if (ppmx_calc_contributions((int)h, (int)new_h, scale[0], 4.0, &plan->contrib[0]) != PPMX_OK) goto bad;
        if (ppmx_calc_contributions((int)w, (int)new_w, scale[1], 4.0, &plan->contrib[1]) != PPMX_OK) goto bad;
        for (pass = 0; pass < 2; pass++) { /* ref:1115-1120 */
            int dim = pass ? second : first;
            ppmx_op *op = &plan->ops[n++];
            op->kind = PPMX_OP_IMRESIZE;
            op->renew_before = pass; /* renewBuffer between the passes, ref:1118 */
            op->dim = dim;
            op->out_size = dim ? (int)new_w : (int)new_h;
            op->weights_sz = plan->contrib[dim].weights_sz;
            op->weights = plan->contrib[dim].weights;
            op->indices = plan->contrib[dim].indices;
        }
        w = new_w;
        h = new_h;
    }
    if (f->rotate_enable) { /* ref:1132-1135 */
        ppmx_op *op = &plan->ops[n++];
        fill_rotate_op(op, angle, w, h);
        op->renew_before = f->resize_enable ? 1 : 0;
        if (angle != 0) { w = op->new_width; h = op->new_height; }
    }
    if (conv_preset != PPMX_CONV_NONE) { /* extension stage, handed over like the reference's own stages */
        ppmx_op *op = &plan->ops[n++];
        op->kind = PPMX_OP_CONV;
        op->renew_before = renew;
        op->conv_bias = 0;
        switch (conv_preset) {
        case PPMX_CONV_BLUR3: op->conv_k = 3; op->conv_div = 16; op->conv_coef = k_blur3; break;
        case PPMX_CONV_BLUR7: op->conv_k = 7; op->conv_div = 49; op->conv_coef = k_box7; break;
        case PPMX_CONV_SHARPEN: op->conv_k = 3; op->conv_div = 1; op->conv_coef = k_sharpen; break;
        case PPMX_CONV_EDGE: op->conv_k = 3; op->conv_div = 1; op->conv_coef = k_edge; break;
        case PPMX_CONV_GAUSS5: fill_binomial_presets(); op->conv_k = 5; op->conv_div = 256; op->conv_coef = k_gauss5; break;
        case PPMX_CONV_GAUSS7: fill_binomial_presets(); op->conv_k = 7; op->conv_div = 4096; op->conv_coef = k_gauss7; break;
        case PPMX_CONV_SHARPEN7: fill_binomial_presets(); op->conv_k = 7; op->conv_div = 4096; op->conv_coef = k_sharpen7; break;
        case PPMX_CONV_GAUSS9: fill_binomial_presets(); op->conv_k = 9; op->conv_div = 65536; op->conv_coef = k_gauss9; break;
        default: printf("Error: unknown convolution preset\n"); goto bad;
        }
        renew = 1;
    }
    if (levels_lo >= 0) { /* second extension stage */
        ppmx_op *op = &plan->ops[n++];
        if (ppmx_levels_lut_linear(levels_lo, levels_hi, plan->levels_lut) != PPMX_OK) {
            printf("Error: invalid levels (need 0 <= lo < hi <= 255)\n");
            goto bad;
        }
        op->kind = PPMX_OP_LEVELS;
        op->renew_before = renew;
        op->levels_lut = plan->levels_lut;
        renew = 1;
    }
    if (f->gray_enable) { /* ref:1137-1140 */
        plan->ops[n].kind = PPMX_OP_GRAY;
        plan->ops[n++].renew_before = renew;
    }
    if (f->mono_enable) { /* ref:1142-1145 */
        plan->ops[n].kind = PPMX_OP_MONO;
        plan->ops[n++].renew_before = renew;
    }
    if (f->flipv_enable) { /* ref:1147-1150 */
        plan->ops[n].kind = PPMX_OP_FLIP;
        plan->ops[n].flip_direction = 1;
        plan->ops[n++].renew_before = renew;
    }
    if (f->fliph_enable) { /* ref:1152-1155 */
        plan->ops[n].kind = PPMX_OP_FLIP;
        plan->ops[n].flip_direction = 0;
        plan->ops[n++].renew_before = renew;
    }
    if (n > PPMX_PLAN_MAX_OPS) { /* can not happen with the stages above (7 at most); guards future additions */
        printf("Error: op chain too long\n");
        goto bad;
    }
    plan->nops = n;
    return PPMX_OK;
bad:
    ppmx_plan_free(plan);
    return PPMX_ERROR;
}

void ppmx_plan_free(ppmx_plan *plan)
{
    if (!plan) return;
    ppmx_contributions_free(&plan->contrib[0]);
    ppmx_contributions_free(&plan->contrib[1]);
    plan->nops = 0;
}

/* ------------------------------------------------------------------ row bands over GPUs */

int ppmx_band_plan(unsigned int full_h, int nranks, int rank, unsigned int align, unsigned int *y0,
                   unsigned int *rows)
{
    unsigned int units, base, extra, u0, u1, a, b;
    if (nranks < 1 || rank < 0 || rank >= nranks || align < 1 || !y0 || !rows) return PPMX_ERROR;
    units = (full_h + align - 1) / align;          /* the last unit may be short */
    base = units / (unsigned int)nranks;
    extra = units % (unsigned int)nranks;          /* the first `extra` ranks get one unit more */
    u0 = (unsigned int)rank * base + ((unsigned int)rank < extra ? (unsigned int)rank : extra);
    u1 = u0 + base + ((unsigned int)rank < extra ? 1u : 0u);
    a = u0 * align; b = u1 * align;
    if (a > full_h) a = full_h;
    if (b > full_h) b = full_h;
    *y0 = a;
    *rows = b - a;
    return PPMX_OK;
}

/* ------------------------------------------------------------------ synthetic rasters */

/* The benchmark's input generator: s = s * 1664525 + 1013904223 mod 2^32, one step per pixel,
 * r = s >> 24, g = s >> 16, b = s >> 8.  Fills pixels [first, first + npix) of the sequence that starts from `seed`,
 * so every rank of a multi-process job can produce its own rows of ONE raster: the state after `first` steps comes
 * from the affine map of a jump, composed by squaring (a k-step jump is s -> A s + C mod 2^32). */
void ppmx_synth_lcg(unsigned char *rgb, size_t first, size_t npix, uint32_t seed)
{
    uint32_t A = 1664525u, Cc = 1013904223u, ja = 1u, jc = 0u, s = seed;
    size_t k = first, i;
    while (k) { /* (ja, jc) := jump by the bits of `first` */
        if (k & 1u) { ja = ja * A; jc = jc * A + Cc; }
        Cc = Cc * A + Cc; /* doubling: s -> A (A s + C) + C */
        A = A * A;
        k >>= 1;
    }
    s = ja * s + jc;
    for (i = 0; i < npix; i++) {
        s = s * 1664525u + 1013904223u;
        rgb[3 * i] = (unsigned char)(s >> 24);
        rgb[3 * i + 1] = (unsigned char)(s >> 16);
        rgb[3 * i + 2] = (unsigned char)(s >> 8);
    }
}

/* ------------------------------------------------------------------ P6 in, P6/P5/P4 out */

/* cursor over the in-memory file with the reference's lookahead rules (ref:333-347) */
typedef struct {
    const unsigned char *p;
    size_t n, i;
    int cur; /* current character or EOF */
} hdr_cursor;

static void hdr_next(hdr_cursor *c)
{
    if (c->cur == EOF) return;
    c->cur = (c->i < c->n) ? c->p[c->i++] : EOF;
    if (c->cur == '#') { /* a comment reads as one newline, ref:341-346 */
        while (c->i < c->n && c->p[c->i] != '\n') c->i++;
        if (c->i < c->n) c->i++;
        c->cur = '\n';
    }
}

/* returns 0 for a number (value in *v), 1 for the magic word P6, -1 otherwise (ref:360-394) */
static int hdr_token(hdr_cursor *c, unsigned int *v)
{
    char word[16];
    int k = 0;
    while (c->cur != EOF && isspace(c->cur)) hdr_next(c);
    if (c->cur != EOF && isdigit(c->cur)) {
        do {
            if (k < (int)sizeof(word) - 1) word[k++] = (char)c->cur;
            hdr_next(c);
        } while (c->cur != EOF && isdigit(c->cur));
        word[k] = 0;
        *v = (unsigned int)atoi(word);
        return 0;
    }
    if (c->cur != EOF && isalpha(c->cur)) {
        do {
            if (k < (int)sizeof(word) - 1) word[k++] = (char)c->cur;
            hdr_next(c);
        } while (c->cur != EOF && isalnum(c->cur));
        word[k] = 0;
        hdr_next(c); /* the reference steps once more after a word, ref:388 */
        return strcmp(word, "P6") == 0 ? 1 : -2;
    }
    return -1;
}

int ppmx_parse_header(const unsigned char *file, size_t filesize, unsigned int *width, unsigned int *height,
                      unsigned int *max_color, size_t *raster_offset)
{
    hdr_cursor c;
    unsigned int v = 0;
    size_t need;
    int t;
    c.p = file; c.n = filesize; c.i = 0; c.cur = '\n'; /* ref:1072-1073 */

    t = hdr_token(&c, &v);
    if (t == -1) BAIL("error in getting next token. wrong format.\n"); /* ref:416 */
    if (t != 1) BAIL("error. invalid file format.\n");                  /* ref:417: only P6 */
    t = hdr_token(&c, width);
    if (t == -1) BAIL("error in getting next token. wrong format.\n");
    if (t != 0) BAIL("error. invalid file format. unable to parse width from input file.\n");
    t = hdr_token(&c, height);
    if (t == -1) BAIL("error in getting next token. wrong format.\n");
    if (t != 0) BAIL("error. invalid file format. unable to parse height from input file.\n");
    t = hdr_token(&c, max_color);
    if (t == -1) BAIL("error in getting next token. wrong format.\n");
    if (t != 0) BAIL("error. invalid file format. unable to parse maximum color from input file.\n");

    /* one byte per channel always (ref:316-318); the file must end exactly with the raster */
    need = (size_t)(*width) * (size_t)(*height) * 3;
    if (need >= 3 && c.i + need - 3 > filesize) BAIL("Error: unexpected end of file.\n"); /* ref:315 */
    if (c.i + need != filesize) BAIL("file format error\n");                              /* ref:453 */
    *raster_offset = c.i;
    return PPMX_OK;
}

/* ------------------------------------------------------------------ EXTENSION: P3 and 16-bit input
 * The reference reads binary P6 with one byte per sample only (ref:386 accepts the word "P6" alone, ref:453 demands
 * width * height * 3 bytes after the header).  The two other PPM encodings are decoded here, on the host, into the
 * same packed 8-bit raster; no reference counterpart, no oracle beyond tests/test_host_logic.py.
 *   P3           ASCII decimal samples separated by white space, '#' comments run to the end of the line
 *   maxval > 255 two bytes per sample, most significant first (P6), or larger decimals (P3)
 * Samples above maxval are clamped to it.  maxval <= 255: bytes are taken as they are and the header's maxval is
 * written back unchanged, exactly like the reference treats P6 (ref:259).  maxval > 255: samples are scaled to
 * 0..255 with the reference's round() (floor(x + 0.5), ref:27) in integers and the result is written as maxval 255. */

static int pnm_skip_space(const unsigned char *f, size_t n, size_t *i)
{
    while (*i < n) {
        if (f[*i] == '#') { while (*i < n && f[*i] != '\n') (*i)++; }
        else if (isspace(f[*i])) (*i)++;
        else return 1;
    }
    return 0;
}

static int pnm_number(const unsigned char *f, size_t n, size_t *i, unsigned int *v)
{
    unsigned long long t = 0;
    int digits = 0;
    if (!pnm_skip_space(f, n, i)) return 0;
    while (*i < n && isdigit(f[*i])) {
        t = t * 10 + (unsigned)(f[*i] - '0');
        if (t > 0xFFFFFFFFull) return 0;
        (*i)++;
        digits++;
    }
    *v = (unsigned int)t;
    return digits > 0;
}

int ppmx_probe_pnm(const unsigned char *file, size_t filesize, unsigned int *width, unsigned int *height,
                   unsigned int *max_color, size_t *raster_offset, int *format)
{
    size_t i = 2, need;
    int p3;
    if (filesize < 2 || file[0] != 'P' || (file[1] != '3' && file[1] != '6')) BAIL("error. invalid file format.\n"); /* ref:417 */
    p3 = file[1] == '3';
    if (!pnm_number(file, filesize, &i, width)) BAIL("error. invalid file format. unable to parse width from input file.\n");
    if (!pnm_number(file, filesize, &i, height)) BAIL("error. invalid file format. unable to parse height from input file.\n");
    if (!pnm_number(file, filesize, &i, max_color))
        BAIL("error. invalid file format. unable to parse maximum color from input file.\n");
    if (*max_color < 1 || *max_color > 65535) BAIL("error. invalid file format. unable to parse maximum color from input file.\n");
    if (p3) {
        *format = PPMX_PNM_P3;
        *raster_offset = i;
        return PPMX_OK;
    }
    if (i >= filesize || !isspace(file[i])) BAIL("file format error\n"); /* one white-space byte ends the header */
    i++;
    need = (size_t)(*width) * (size_t)(*height) * 3 * (*max_color > 255 ? 2 : 1);
    if (i + need != filesize) BAIL("file format error\n"); /* ref:453 */
    *format = *max_color > 255 ? PPMX_PNM_P6_16 : PPMX_PNM_P6_8;
    *raster_offset = i;
    return PPMX_OK;
}

int ppmx_decode_pnm(const unsigned char *file, size_t filesize, size_t raster_offset, int format, unsigned int width,
                    unsigned int height, unsigned int max_color, unsigned char *dst, unsigned int *out_max_color)
{
    const size_t n = (size_t)width * height * 3;
    const unsigned int scale = max_color > 255;
    size_t i = raster_offset, k;
    *out_max_color = scale ? 255u : max_color;
    for (k = 0; k < n; k++) {
        unsigned int v;
        if (format == PPMX_PNM_P3) {
            if (!pnm_number(file, filesize, &i, &v)) BAIL("Error: unexpected end of file.\n"); /* ref:315 */
        } else if (format == PPMX_PNM_P6_16) {
            if (i + 2 > filesize) BAIL("Error: unexpected end of file.\n");
            v = ((unsigned int)file[i] << 8) | file[i + 1];
            i += 2;
        } else {
            if (i + 1 > filesize) BAIL("Error: unexpected end of file.\n");
            v = file[i++];
        }
        if (v > max_color) v = max_color;
        /* round(v * 255 / maxval) with round(x) = floor(x + 0.5) (ref:27), in integers */
        dst[k] = (unsigned char)(scale ? (2ull * v * 255u + max_color) / (2ull * max_color) : v);
    }
    if (format == PPMX_PNM_P3) { /* nothing but white space and comments may follow the last sample (cf. ref:453) */
        if (pnm_skip_space(file, filesize, &i)) BAIL("file format error\n");
    }
    return PPMX_OK;
}

int ppmx_format_header(char *dst, size_t cap, int file_type, unsigned int width, unsigned int height,
                       unsigned int max_color)
{
    const char *magic = (file_type == PPMX_FILETYPE_PGM) ? "P5" : (file_type == PPMX_FILETYPE_PBM) ? "P4" : "P6";
    if (file_type == PPMX_FILETYPE_PBM) /* P4 carries no maxval line, ref:258 */
        return snprintf(dst, cap, "%s\n# generated by ppmx_edward\n%u %u\n", magic, width, height);
    return snprintf(dst, cap, "%s\n# generated by ppmx_edward\n%u %u\n%u\n", magic, width, height, max_color);
}

int ppmx_getImageInfo(ppmx_image_handler *h)
{
    unsigned int w = 0, hh = 0, mx = 0;
    size_t off = 0;
    if (ppmx_parse_header(h->file_buffer, h->filesize, &w, &hh, &mx, &off) != PPMX_OK) return PPMX_ERROR;
    h->imginfo.file_type = PPMX_FILETYPE_PPM; /* ref:418 */
    h->imginfo.width = w;
    h->imginfo.height = hh;
    h->imginfo.max_color = mx;
    h->imginfo.size = hh * w;
    h->imginfo.new_width = h->imginfo.new_height = 0;
    h->index_buffer = off;
    if (ppmx_gpu_upload(h->ctx, h->file_buffer + off, w, hh, PPMX_LAYOUT_RGB8, &h->imginfo.buff) != PPMX_OK)
        BAIL("error. can not allocate memory\n");
    return PPMX_OK;
}

static int write_output(const char *in_name, int file_type, unsigned int w, unsigned int h, unsigned int maxval,
                        const unsigned char *raster, size_t nbytes)
{
    char header[96];
    size_t len = strlen(in_name);
    char *name = (char *)malloc(len + 5);
    FILE *fp;
    int hl;
    if (!name) BAIL("error. can not allocate memory\n");
    memcpy(name, in_name, len);
    memcpy(name + len, ".out", 5); /* ref:229-233 */
    fp = fopen(name, "wb");
    free(name);
    if (!fp) BAIL("Error: unable to open file for writing\n"); /* ref:237 */
    hl = ppmx_format_header(header, sizeof(header), file_type, w, h, maxval);
    if (fwrite(header, 1, (size_t)hl, fp) != (size_t)hl || (nbytes && fwrite(raster, 1, nbytes, fp) != nbytes)) {
        fclose(fp);
        BAIL("Error: failed in writing to file\n");
    }
    fclose(fp);
    return PPMX_OK;
}

int ppmx_putImageToFile(ppmx_image_handler *h)
{
    size_t cap, n = 0;
    unsigned char *out;
    int rc;
    if (h->imginfo.new_buff == NULL) BAIL("Error: no data to write\n"); /* ref:235 */
    cap = (size_t)h->imginfo.new_width * h->imginfo.new_height * 3 + 16;
    out = (unsigned char *)ppmx_gpu_host_alloc(h->ctx, cap);
    if (!out) BAIL("error. can not allocate memory\n");
    rc = ppmx_gpu_download(h->ctx, h->imginfo.new_buff, (int)h->imginfo.file_type, out, cap, &n);
    if (rc == PPMX_OK)
        rc = write_output(h->filename, (int)h->imginfo.file_type, h->imginfo.new_width, h->imginfo.new_height,
                          h->imginfo.max_color, out, n);
    ppmx_gpu_host_free(h->ctx, out);
    /* the writer releases new_buff (ref:293-298); buff goes with it when it is the same raster */
    if (h->imginfo.new_buff == h->imginfo.buff) h->imginfo.buff = NULL;
    ppmx_gpu_image_free(h->ctx, h->imginfo.new_buff);
    h->imginfo.new_buff = NULL;
    return rc;
}

/* "P6", then a maxval above 255: two bytes per sample (the reference fails its size check on such a file, ref:453) */
static int pnm_is_16bit(const unsigned char *f, size_t n)
{
    size_t i = 2;
    unsigned int w, hh, mx;
    if (n < 2 || f[0] != 'P' || f[1] != '6') return 0;
    return pnm_number(f, n, &i, &w) && pnm_number(f, n, &i, &hh) && pnm_number(f, n, &i, &mx) && mx > 255 && mx <= 65535;
}

/* ---- one file = one job: read -> plan + device chain -> write ---------------------------------------- */

typedef struct ppmx_job {
    const char *filename;
    unsigned char *file_buffer; /* the whole input file, pinned */
    size_t filesize, file_cap;
    unsigned char *decoded;     /* pinned 8-bit raster of a P3 / 16-bit input (extension), else NULL */
    size_t decoded_cap;
    const unsigned char *raster;
    unsigned int w, h, maxval;
    unsigned char *out;         /* what follows the output header, pinned */
    size_t out_cap, out_bytes;
    unsigned int ow, oh;
    int ft, rc;
} ppmx_job;

/* pinned buffers are expensive to make (the pages are locked one by one): a batch keeps a few and reuses them */
typedef struct ppmx_pin_cache {
    void *p[8];
    size_t cap[8];
    pthread_mutex_t mu;
    ppmx_gpu_ctx *ctx;
} ppmx_pin_cache;

static void *pin_get(ppmx_pin_cache *pc, size_t need, size_t *cap)
{
    int i, best = -1;
    void *p;
    pthread_mutex_lock(&pc->mu);
    for (i = 0; i < 8; i++)
        if (pc->p[i] && pc->cap[i] >= need && (best < 0 || pc->cap[i] < pc->cap[best])) best = i;
    if (best >= 0) {
        p = pc->p[best];
        *cap = pc->cap[best];
        pc->p[best] = NULL;
        pthread_mutex_unlock(&pc->mu);
        return p;
    }
    pthread_mutex_unlock(&pc->mu);
    *cap = need;
    return ppmx_gpu_host_alloc(pc->ctx, need);
}

static void pin_put(ppmx_pin_cache *pc, void *p, size_t cap)
{
    int i, slot = -1;
    void *drop = p;
    if (!p) return;
    pthread_mutex_lock(&pc->mu);
    for (i = 0; i < 8 && slot < 0; i++)
        if (!pc->p[i]) slot = i;
    if (slot < 0) /* full: keep the larger buffers */
        for (i = 0; i < 8; i++)
            if (pc->cap[i] < cap && (slot < 0 || pc->cap[i] < pc->cap[slot])) slot = i;
    if (slot >= 0) {
        drop = pc->p[slot];
        pc->p[slot] = p;
        pc->cap[slot] = cap;
    }
    pthread_mutex_unlock(&pc->mu);
    if (drop) ppmx_gpu_host_free(pc->ctx, drop);
}

static void pin_cache_free(ppmx_pin_cache *pc)
{
    int i;
    for (i = 0; i < 8; i++)
        if (pc->p[i]) { ppmx_gpu_host_free(pc->ctx, pc->p[i]); pc->p[i] = NULL; }
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ref:1058-1069 + 409-456: the whole file into pinned memory; after the header it IS the packed raster (ref:316-318) */
static int job_read(ppmx_pin_cache *pc, ppmx_job *j)
{
    FILE *fp = fopen(j->filename, "rb");
    long sz;
    size_t off = 0;
    if (!fp) BAIL("error. can not open file\n"); /* ref:1059 */
    if (fseek(fp, 0, SEEK_END) < 0) { fclose(fp); BAIL("error. can not set file position in fseek.\n"); }
    sz = ftell(fp);
    rewind(fp);
    j->filesize = (size_t)(sz < 0 ? 0 : sz);
    j->file_buffer = (unsigned char *)pin_get(pc, j->filesize + 1, &j->file_cap);
    if (!j->file_buffer) { fclose(fp); BAIL("error. can not allocate memory\n"); }
    if (fread(j->file_buffer, 1, j->filesize, fp) != j->filesize) {
        fclose(fp);
        BAIL("error in reading input file.\n"); /* ref:1069 */
    }
    fclose(fp);
    if (j->filesize >= 2 && j->file_buffer[0] == 'P' && (j->file_buffer[1] == '3' || pnm_is_16bit(j->file_buffer, j->filesize))) {
        /* EXTENSION: an encoding the reference rejects (ref:386, 453) is decoded on the host into the same raster */
        int fmt = 0;
        if (ppmx_probe_pnm(j->file_buffer, j->filesize, &j->w, &j->h, &j->maxval, &off, &fmt) != PPMX_OK) return PPMX_ERROR;
        j->decoded = (unsigned char *)pin_get(pc, (size_t)j->w * j->h * 3 + 1, &j->decoded_cap);
        if (!j->decoded) BAIL("error. can not allocate memory\n");
        if (ppmx_decode_pnm(j->file_buffer, j->filesize, off, fmt, j->w, j->h, j->maxval, j->decoded, &j->maxval) != PPMX_OK)
            return PPMX_ERROR;
        j->raster = j->decoded;
        return PPMX_OK;
    }
    if (ppmx_parse_header(j->file_buffer, j->filesize, &j->w, &j->h, &j->maxval, &off) != PPMX_OK) return PPMX_ERROR;
    j->raster = j->file_buffer + off;
    return PPMX_OK;
}

/* ref:1084-1155: the op chain, as one ppmx_gpu_apply call */
static int job_run(const ppmx_image_handler *h, ppmx_gpu_ctx *ctx, ppmx_pin_cache *pc, ppmx_job *j)
{
    ppmx_plan plan;
    unsigned int ow = j->w, oh = j->h;
    int i, rc;
    memset(&plan, 0, sizeof(plan));
    if (ppmx_plan_chain_ext2(&h->arg_flag, h->output_width_size, h->angle, j->w, j->h, h->conv_preset,
                             h->levels_enable ? h->levels_lo : -1, h->levels_hi, &plan) != PPMX_OK) return PPMX_ERROR;
    if (plan.nops == 0) { ppmx_plan_free(&plan); BAIL("Error: no data to write\n"); } /* ref:235 */
    /* the largest raster any stage can hand to the writer is RGB at the final size */
    for (i = 0; i < plan.nops; i++) {
        int lay = PPMX_LAYOUT_RGB8;
        if (plan.ops[i].kind == PPMX_OP_IMRESIZE || plan.ops[i].kind == PPMX_OP_ROTATE)
            ppmx_gpu_op_output(&plan.ops[i], ow, oh, PPMX_LAYOUT_RGB8, &ow, &oh, &lay);
    }
    j->out = (unsigned char *)pin_get(pc, (size_t)ow * oh * 3 + 16, &j->out_cap);
    if (!j->out) { ppmx_plan_free(&plan); BAIL("error. can not allocate memory\n"); }
    rc = ppmx_gpu_apply(ctx, plan.ops, plan.nops, j->raster, j->w, j->h, j->out, j->out_cap, &j->out_bytes, &j->ow, &j->oh, &j->ft);
    ppmx_plan_free(&plan);
    return rc;
}

static int job_write(ppmx_job *j) /* ref:221-301: header + ONE fwrite of the raster */
{
    return write_output(j->filename, j->ft, j->ow, j->oh, j->maxval, j->out, j->out_bytes);
}

static void job_release(ppmx_pin_cache *pc, ppmx_job *j)
{
    pin_put(pc, j->out, j->out_cap);
    pin_put(pc, j->decoded, j->decoded_cap);
    pin_put(pc, j->file_buffer, j->file_cap);
    j->out = j->decoded = j->file_buffer = NULL;
}

static int open_ctx(ppmx_image_handler *h, int *own)
{
    *own = 0;
    if (!h->ctx) {
        /* PPMX_DEVICE = n: that GPU; PPMX_DEVICE = all: every visible GPU (one large raster is cut into row bands,
         * a batch of files is dealt round-robin) */
        const char *dev = getenv("PPMX_DEVICE");
        int rc = (dev && strcmp(dev, "all") == 0) ? ppmx_gpu_init_multi(&h->ctx, NULL, 0) : ppmx_gpu_init(&h->ctx, dev ? atoi(dev) : 0);
        if (rc != PPMX_OK) return PPMX_ERROR;
        *own = 1;
    }
    return PPMX_OK;
}

int ppmx_doProcessPPM(ppmx_image_handler *h)
{
    ppmx_job j;
    ppmx_pin_cache pc;
    int own_ctx = 0, rc = PPMX_ERROR;
    const int trace = getenv("PPMX_TRACE") != NULL;
    double t0 = now_s(), t1, t2, t3, t4;

    memset(&j, 0, sizeof(j));
    memset(&pc, 0, sizeof(pc));
    pthread_mutex_init(&pc.mu, NULL);
    if (open_ctx(h, &own_ctx) != PPMX_OK) return PPMX_ERROR;
    pc.ctx = h->ctx;
    t1 = now_s();
    j.filename = h->filename;
    t2 = t3 = t1;
    if (job_read(&pc, &j) == PPMX_OK) {
        h->imginfo.width = j.w; h->imginfo.height = j.h; h->imginfo.max_color = j.maxval;
        t2 = now_s();
        if (job_run(h, h->ctx, &pc, &j) == PPMX_OK) {
            h->imginfo.new_width = j.ow; h->imginfo.new_height = j.oh; h->imginfo.file_type = (unsigned int)j.ft;
            t3 = now_s();
            rc = job_write(&j);
        }
    }
    t4 = now_s();
    if (trace)
        fprintf(stderr, "ppmx trace: device context %.1f ms, read+parse %.1f ms, chain (upload, kernels, download) %.1f ms, write %.1f ms\n",
                1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3));
    job_release(&pc, &j);
    pin_cache_free(&pc);
    pthread_mutex_destroy(&pc.mu);
    if (own_ctx) { ppmx_gpu_free(h->ctx); h->ctx = NULL; }
    return rc;
}

/* ---- EXTENSION: several files in one run (the reference takes exactly one, ref:180) -----------------------
 * One device context for all of them, and a three-stage pipeline: a reader thread brings file n+1 into pinned memory
 * and a writer thread puts result n-1 on disk while the calling thread runs the device chain of file n. */

typedef struct ppmx_queue {
    ppmx_job *slot[4];
    int head, count, closed;
    pthread_mutex_t mu;
    pthread_cond_t cv;
} ppmx_queue;

static void q_init(ppmx_queue *q) { memset(q, 0, sizeof(*q)); pthread_mutex_init(&q->mu, NULL); pthread_cond_init(&q->cv, NULL); }
static void q_push(ppmx_queue *q, ppmx_job *j)
{
    pthread_mutex_lock(&q->mu);
    while (q->count == 2) pthread_cond_wait(&q->cv, &q->mu); /* at most two jobs wait between stages */
    q->slot[(q->head + q->count) % 4] = j;
    q->count++;
    pthread_cond_broadcast(&q->cv);
    pthread_mutex_unlock(&q->mu);
}
static ppmx_job *q_pop(ppmx_queue *q)
{
    ppmx_job *j = NULL;
    pthread_mutex_lock(&q->mu);
    while (q->count == 0 && !q->closed) pthread_cond_wait(&q->cv, &q->mu);
    if (q->count) {
        j = q->slot[q->head];
        q->head = (q->head + 1) % 4;
        q->count--;
        pthread_cond_broadcast(&q->cv);
    }
    pthread_mutex_unlock(&q->mu);
    return j;
}
static void q_close(ppmx_queue *q)
{
    pthread_mutex_lock(&q->mu);
    q->closed = 1;
    pthread_cond_broadcast(&q->cv);
    pthread_mutex_unlock(&q->mu);
}

typedef struct ppmx_batch {
    ppmx_job *jobs;
    int njobs;
    ppmx_pin_cache pc;
    ppmx_queue read_q, write_q;
} ppmx_batch;

static void *reader_main(void *arg)
{
    ppmx_batch *b = (ppmx_batch *)arg;
    int i;
    for (i = 0; i < b->njobs; i++) {
        b->jobs[i].rc = job_read(&b->pc, &b->jobs[i]);
        q_push(&b->read_q, &b->jobs[i]);
    }
    q_close(&b->read_q);
    return NULL;
}

static void *writer_main(void *arg)
{
    ppmx_batch *b = (ppmx_batch *)arg;
    ppmx_job *j;
    while ((j = q_pop(&b->write_q)) != NULL) {
        if (j->rc == PPMX_OK) j->rc = job_write(j);
        job_release(&b->pc, j);
    }
    return NULL;
}

int ppmx_doProcessBatch(ppmx_image_handler *h, const char *const *filenames, int nfiles)
{
    ppmx_batch b;
    pthread_t rd, wr;
    ppmx_job *j;
    int own_ctx = 0, i, failed = 0;
    const int trace = getenv("PPMX_TRACE") != NULL;
    double t0 = now_s();
    if (nfiles < 1 || !filenames) BAIL("Error: invalid options\n");
    if (open_ctx(h, &own_ctx) != PPMX_OK) return PPMX_ERROR;
    memset(&b, 0, sizeof(b));
    b.jobs = (ppmx_job *)calloc((size_t)nfiles, sizeof(ppmx_job));
    if (!b.jobs) { if (own_ctx) { ppmx_gpu_free(h->ctx); h->ctx = NULL; } BAIL("error. can not allocate memory\n"); }
    b.njobs = nfiles;
    for (i = 0; i < nfiles; i++) b.jobs[i].filename = filenames[i];
    pthread_mutex_init(&b.pc.mu, NULL);
    b.pc.ctx = h->ctx;
    q_init(&b.read_q);
    q_init(&b.write_q);
    pthread_create(&rd, NULL, reader_main, &b);
    pthread_create(&wr, NULL, writer_main, &b);
    while ((j = q_pop(&b.read_q)) != NULL) {
        if (j->rc == PPMX_OK) j->rc = job_run(h, h->ctx, &b.pc, j);
        q_push(&b.write_q, j);
    }
    q_close(&b.write_q);
    pthread_join(rd, NULL);
    pthread_join(wr, NULL);
    for (i = 0; i < nfiles; i++) failed += b.jobs[i].rc != PPMX_OK;
    if (trace) fprintf(stderr, "ppmx trace: %d file(s), %d failed, %.1f ms in all\n", nfiles, failed, 1e3 * (now_s() - t0));
    pin_cache_free(&b.pc);
    pthread_mutex_destroy(&b.pc.mu);
    free(b.jobs);
    if (own_ctx) { ppmx_gpu_free(h->ctx); h->ctx = NULL; }
    return failed ? PPMX_ERROR : PPMX_OK;
}

/* ------------------------------------------------------------------ command line */

void ppmx_usage(void)
{
    printf("ppmx-edward [options] (input filename)\n");
    printf("Options -fv  Flip vertically\n");
    printf("        -fh  Flip horizontally\n");
    printf("        -w(new width) Scale to the new width\n");
    printf("        -w100 means new width is 100\n");
    printf("        -r(angle)  Rotate (CW)\n");
    printf("        -r30 means rotate 30 degree CW.\n");
    printf("        -mono Convert to bilevel (.pbm) format\n");
    printf("        -gray  Convert to grayscale (.pgm) format\n");
}

static int all_digits(const char *s)
{
    for (; *s; s++)
        if (!isdigit((unsigned char)*s)) return 0;
    return 1;
}

int ppmx_main(int argc, char *argv[])
{
    ppmx_image_handler hd;
    int i, have_file = 0, batch = 0, nfiles = 0, rc;
    const char **files = (const char **)calloc((size_t)(argc > 0 ? argc : 1), sizeof(char *));
    if (!files) BAIL("error. can not allocate memory\n");
    memset(&hd, 0, sizeof(hd));
    for (i = 1; i < argc; i++) /* EXTENSION flag -batch: any number of input files in one run (the reference takes one, ref:180) */
        if (strcmp(argv[i], "-batch") == 0) batch = 1;

    for (i = 1; i < argc; i++) { /* same options, checks and messages as ref:125-183 */
        const char *a = argv[i];
        if (a[0] != '-') {
            if (have_file && !batch) { free(files); BAIL("Error: invalid options\n"); } /* ref:180 */
            hd.filename = a;
            files[nfiles++] = a;
            have_file = 1;
        } else if (strcmp(a, "-batch") == 0) {
            continue;
        } else if (a[1] == 'f') {
            if (a[2] == 'h') {
                if (hd.arg_flag.fliph_enable) BAIL("Error: Duplicate options not allowed\n");
                if (hd.arg_flag.flipv_enable) BAIL("Error: Conflicting options not allowed\n");
                hd.arg_flag.fliph_enable = 1;
            } else if (a[2] == 'v') {
                if (hd.arg_flag.flipv_enable) BAIL("Error: Duplicate options not allowed\n");
                if (hd.arg_flag.fliph_enable) BAIL("Error: Conflicting options not allowed\n");
                hd.arg_flag.flipv_enable = 1;
            } else {
                BAIL("Error: invalid option for flip.\nallowed options are -fh -fv only.\n");
            }
        } else if (a[1] == 'w') {
            if (!all_digits(a + 2)) BAIL("Error: invalid option for scaling.\n");
            if (hd.arg_flag.resize_enable) BAIL("Error: Duplicate options not allowed\n");
            hd.arg_flag.resize_enable = 1;
            hd.output_width_size = (unsigned int)atoi(a + 2);
        } else if (a[1] == 'r') {
            if (a[2] == 0) BAIL("Error: invalid option for rotate\n");
            if (hd.arg_flag.rotate_enable) BAIL("Error: Duplicate options not allowed\n");
            hd.arg_flag.rotate_enable = 1;
            if (!all_digits(a + 2)) BAIL("Error: invalid option for rotate.\n");
            hd.angle = (double)atoi(a + 2);
            if (hd.angle < 0 || hd.angle >= 360) BAIL("Error: invalid option for rotate.\n");
        } else if (strcmp(a + 1, "blur") == 0 || strcmp(a + 1, "blur7") == 0 || strcmp(a + 1, "sharpen") == 0 ||
                   strcmp(a + 1, "edge") == 0 || strcmp(a + 1, "gauss5") == 0 || strcmp(a + 1, "gauss7") == 0 ||
                   strcmp(a + 1, "gauss9") == 0 || strcmp(a + 1, "sharpen7") == 0) { /* EXTENSION flags: the reference rejects them (ref:175) */
            if (hd.conv_preset) BAIL("Error: Duplicate options not allowed\n");
            hd.conv_preset = strcmp(a + 1, "blur") == 0 ? PPMX_CONV_BLUR3 : strcmp(a + 1, "blur7") == 0 ? PPMX_CONV_BLUR7
                             : strcmp(a + 1, "sharpen") == 0 ? PPMX_CONV_SHARPEN : strcmp(a + 1, "edge") == 0 ? PPMX_CONV_EDGE
                             : strcmp(a + 1, "gauss5") == 0 ? PPMX_CONV_GAUSS5 : strcmp(a + 1, "gauss7") == 0 ? PPMX_CONV_GAUSS7
                             : strcmp(a + 1, "gauss9") == 0 ? PPMX_CONV_GAUSS9 : PPMX_CONV_SHARPEN7;
        } else if (strncmp(a + 1, "levels", 6) == 0) { /* EXTENSION flag -levelsLO-HI, e.g. -levels16-235 */
            int lo = -1, hi = -1, used = 0;
            if (hd.levels_enable) BAIL("Error: Duplicate options not allowed\n");
            if (sscanf(a + 7, "%d-%d%n", &lo, &hi, &used) != 2 || a[7 + used] != 0 || lo < 0 || hi > 255 || lo >= hi)
                BAIL("Error: invalid option for levels.\n");
            hd.levels_enable = 1;
            hd.levels_lo = lo;
            hd.levels_hi = hi;
        } else if (strcmp(a + 1, "gray") == 0) {
            if (hd.arg_flag.gray_enable) BAIL("Error: Duplicate options not allowed\n");
            if (hd.arg_flag.mono_enable) BAIL("Error: Conflicting options not allowed\n");
            hd.arg_flag.gray_enable = 1;
        } else if (strcmp(a + 1, "mono") == 0) {
            if (hd.arg_flag.mono_enable) BAIL("Error: Duplicate options not allowed\n");
            if (hd.arg_flag.gray_enable) BAIL("Error: Conflicting options not allowed\n");
            hd.arg_flag.mono_enable = 1;
        } else {
            printf("Error: invalid option: %s\n", a + 1);
            ppmx_usage();
            return PPMX_ERROR;
        }
    }
    if (!have_file) {
        free(files);
        ppmx_usage();
        return PPMX_ERROR;
    }
    rc = batch ? ppmx_doProcessBatch(&hd, files, nfiles) : ppmx_doProcessPPM(&hd);
    free(files);
    return rc != 0 ? PPMX_ERROR : 0;
}
