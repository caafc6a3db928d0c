/* tests/c/stepwise.c -- the operator-by-operator binding of INTEGRATION.md section 3, as a program.
 *
 * It walks the control flow of the reference's doProcessPPM (ref:1084-1155 of
 * /root/reference/ppmx-edward.c: resize -> rotate -> gray -> mono -> flipv -> fliph with renewBuffer
 * before a stage iff -w or -r was given) but calls the ppmx_* host functions of include/ppmx_host.h one
 * at a time -- ppmx_getImageInfo, ppmx_calc_contributions, ppmx_imresize, ppmx_renewBuffer, ppmx_rotate,
 * ppmx_gray, ppmx_mono, ppmx_flip, ppmx_putImageToFile -- instead of the one-shot ppmx_gpu_apply.
 * tests/test_gpu_parity.py compares its output files with those of ppmx-b200 and of the reference.
 *
 * usage: stepwise [-fv] [-fh] [-wN] [-rN] [-gray] [-mono] file.ppm      (writes file.ppm.out)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ppmx_host.h"

static int read_file(ppmx_image_handler *h)
{
    FILE *fp = fopen(h->filename, "rb");
    long sz;
    if (!fp) return -1;
    fseek(fp, 0, SEEK_END);
    sz = ftell(fp);
    rewind(fp);
    h->filesize = (size_t)sz;
    h->file_buffer = (unsigned char *)ppmx_gpu_host_alloc(h->ctx, h->filesize + 1);
    if (!h->file_buffer || fread(h->file_buffer, 1, h->filesize, fp) != h->filesize) { fclose(fp); return -1; }
    fclose(fp);
    return 0;
}

int main(int argc, char *argv[])
{
    ppmx_image_handler hd;
    int i, rc = 1;
    memset(&hd, 0, sizeof(hd));
    for (i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (!strcmp(a, "-fv")) hd.arg_flag.flipv_enable = 1;
        else if (!strcmp(a, "-fh")) hd.arg_flag.fliph_enable = 1;
        else if (!strcmp(a, "-gray")) hd.arg_flag.gray_enable = 1;
        else if (!strcmp(a, "-mono")) hd.arg_flag.mono_enable = 1;
        else if (a[0] == '-' && a[1] == 'w') { hd.arg_flag.resize_enable = 1; hd.output_width_size = (unsigned)atoi(a + 2); }
        else if (a[0] == '-' && a[1] == 'r') { hd.arg_flag.rotate_enable = 1; hd.angle = (double)atoi(a + 2); }
        else hd.filename = a;
    }
    if (!hd.filename || ppmx_gpu_init(&hd.ctx, 0) != PPMX_OK) return 1;
    if (read_file(&hd) != 0 || ppmx_getImageInfo(&hd) != PPMX_OK) goto done;
    hd.imginfo.new_buff = NULL;

    if (hd.arg_flag.resize_enable) { /* ref:1084-1130 */
        ppmx_contributions c[2];
        double scale[2];
        int order[2] = {0, 0}, im_sz[2], dim;
        unsigned int out_width;
        hd.imginfo.new_width = hd.output_width_size;
        scale[1] = (double)((double)hd.imginfo.new_width / hd.imginfo.width);
        hd.imginfo.new_height = (unsigned int)((double)hd.imginfo.height * scale[1]);
        scale[0] = (double)((double)hd.imginfo.new_height / hd.imginfo.height);
        if (scale[0] < scale[1]) order[1] = 1; else order[0] = 1;
        if (ppmx_calc_contributions((int)hd.imginfo.height, (int)hd.imginfo.new_height, scale[0], 4.0, &c[0]) != PPMX_OK) goto done;
        if (ppmx_calc_contributions((int)hd.imginfo.width, (int)hd.imginfo.new_width, scale[1], 4.0, &c[1]) != PPMX_OK) goto done;
        im_sz[0] = (int)hd.imginfo.new_height;
        im_sz[1] = (int)hd.imginfo.new_width;
        out_width = hd.imginfo.new_width;
        dim = order[0];
        if (ppmx_imresize(&hd, im_sz[dim], dim, c[dim].weights, c[dim].indices, c[dim].weights_sz) != PPMX_OK) goto done;
        hd.imginfo.new_width = out_width; /* ref:1117 */
        ppmx_renewBuffer(&hd);
        /* the reference leaves imginfo.width at the FINAL width here (ref:1117-1118) although the buffer still
         * has the old one; the device raster knows its own size, so restore the true width for the next pass */
        ppmx_gpu_image_info(hd.imginfo.buff, &hd.imginfo.width, &hd.imginfo.height, NULL, NULL, NULL);
        dim = order[1];
        if (ppmx_imresize(&hd, im_sz[dim], dim, c[dim].weights, c[dim].indices, c[dim].weights_sz) != PPMX_OK) goto done;
        ppmx_contributions_free(&c[0]);
        ppmx_contributions_free(&c[1]);
    }
    if (hd.arg_flag.rotate_enable) { /* ref:1132-1135 */
        if (hd.arg_flag.resize_enable) ppmx_renewBuffer(&hd);
        ppmx_rotate(&hd);
    }
    if (hd.arg_flag.gray_enable) { /* ref:1137-1140 */
        if (hd.arg_flag.resize_enable || hd.arg_flag.rotate_enable) ppmx_renewBuffer(&hd);
        ppmx_gray(&hd);
    }
    if (hd.arg_flag.mono_enable) { /* ref:1142-1145 */
        if (hd.arg_flag.resize_enable || hd.arg_flag.rotate_enable) ppmx_renewBuffer(&hd);
        ppmx_mono(&hd);
    }
    if (hd.arg_flag.flipv_enable) { /* ref:1147-1150 */
        if (hd.arg_flag.resize_enable || hd.arg_flag.rotate_enable) ppmx_renewBuffer(&hd);
        ppmx_flip(&hd, 1);
    }
    if (hd.arg_flag.fliph_enable) { /* ref:1152-1155 */
        if (hd.arg_flag.resize_enable || hd.arg_flag.rotate_enable) ppmx_renewBuffer(&hd);
        ppmx_flip(&hd, 0);
    }
    rc = ppmx_putImageToFile(&hd) == PPMX_OK ? 0 : 255;
done:
    if (hd.imginfo.buff) ppmx_gpu_image_free(hd.ctx, hd.imginfo.buff);
    if (hd.file_buffer) ppmx_gpu_host_free(hd.ctx, hd.file_buffer);
    ppmx_gpu_free(hd.ctx);
    return rc;
}
