/*
 * ppmx_gpu.h -- the drop-in boundary: a C ABI (plain pointers and sizes, no CUDA or
 * torch types) behind which the per-pixel loops of ppmx-edward.c run on a B200.
 *
 * "ref:N" below = /root/reference/ppmx-edward.c line N.  The reference has no plugin or
 * FFI layer; its operator interface is `int op(ppm_image_handler *h [, args])` -- gray
 * (ref:986), mono (ref:949), flip (ref:888), rotate (ref:673), imresize (ref:808) -- each
 * reading imginfo.buff and producing imginfo.new_buff, returning PPM_NOERROR 0 /
 * PPM_ERROR -1 (ref:16-17).  The entry points here replace the LOOPS inside those
 * functions and nothing else; everything that touches libm (cos, sin, output sizes,
 * contribution tables) stays on the host (see ppmx_host.h) and arrives here as numbers.
 *
 * Conventions kept from the reference: return 0 / -1; one line on stdout describing a
 * failure (CHECK_ERROR, ref:31-36); a context is used by one host thread at a time.
 * There is no CPU fallback: without a usable CUDA device every call fails with -1.
 */
#ifndef PPMX_GPU_H
#define PPMX_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPMX_OK 0     /* PPM_NOERROR, ref:17 */
#define PPMX_ERROR (-1) /* PPM_ERROR,   ref:16 */

/* output file kinds, ref:22-24 */
#define PPMX_FILETYPE_PPM 0
#define PPMX_FILETYPE_PGM 1
#define PPMX_FILETYPE_PBM 2

/* How a device raster is stored.  RGB8 is the reference's `pixel` array (ref:39-43), rows
 * contiguous.  R8 keeps only the .r member of each pixel (what gray/mono produce: .g/.b
 * stay 0 there, ref:996-1000) -- 1 byte per pixel.  BITS is the P4 raster the writer
 * emits (ref:268-284): MSB first, every row padded to a whole byte. */
#define PPMX_LAYOUT_RGB8 0
#define PPMX_LAYOUT_R8 1
#define PPMX_LAYOUT_BITS 2

/* operator kinds; one per reference loop nest */
#define PPMX_OP_GRAY 0      /* ref:998-1000            RGB8 -> R8, file type PGM            */
#define PPMX_OP_MONO 1      /* ref:964-969             RGB8 -> R8 of 0/1, file type PBM     */
#define PPMX_OP_FLIP 2      /* ref:898-911             any byte layout, same layout out     */
#define PPMX_OP_ROTATE 3    /* ref:714-786             RGB8 -> RGB8                         */
#define PPMX_OP_IMRESIZE 4  /* ref:820-838 / 846-868   RGB8 -> RGB8, one separable pass     */
#define PPMX_OP_MONO_BITS 5 /* ref:964-969 fused with the P4 packer ref:268-284: RGB8 -> BITS */
#define PPMX_OP_PACK_PBM 6  /* ref:268-284 alone       R8 or RGB8 (.r) -> BITS              */
#define PPMX_OP_EXTRACT_R 7 /* ref:263-267             RGB8 -> R8 (.r of every pixel)       */
/* extensions: no counterpart in the reference ("parity unpinned", self-oracle only) */
#define PPMX_OP_CONV 16     /* k x k integer convolution, mirror border, RGB8 -> RGB8       */
#define PPMX_OP_HIST_GRAY 17 /* 256-bin histogram of (r+g+b)/3, RGB8 -> 256 x u64 (no image) */
#define PPMX_OP_GRAY_HIST 18 /* gray and its histogram in one pass over the raster          */
#define PPMX_OP_LEVELS 19   /* levels: every byte through a 256-entry table, RGB8 or R8, same layout out */

typedef struct ppmx_op {
    int32_t kind;            /* PPMX_OP_*                                                   */
    int32_t renew_before;    /* chain only: hand new_buff over to buff first (renewBuffer,
                                ref:1019-1026; the callers' conditions are ref:1133-1153)   */
    /* flip, ref:888 */
    int32_t flip_direction;  /* 1 vertical, 0 horizontal (ref:882-883)                      */
    /* rotate, ref:673: the host evaluates libm once (ref:653-655, 741-742)                 */
    int32_t angle_deg;       /* 0..359 as given to -r (ref:161-162)                         */
    double cos_t, sin_t;     /* cos/sin(angle * M_PI / 180) from the host libm              */
    uint32_t new_width, new_height; /* calc_rot_size on the folded angle (ref:687-691)      */
    /* imresize, ref:808: one pass; flat host tables [out_size][weights_sz]                 */
    int32_t dim;             /* 0 = height pass, 1 = width pass                             */
    int32_t out_size;
    int32_t weights_sz;
    const double *weights;
    const int32_t *indices;
    /* extension: convolution */
    int32_t conv_k;          /* odd, 1..15                                                  */
    int32_t conv_div;        /* > 0                                                         */
    int32_t conv_bias;
    const int32_t *conv_coef; /* k*k, row major                                             */
    /* extension: histogram.  GRAY_HIST inside a chain writes 256 bins per raster here (host
     * memory, raster i of a batch at hist_out + 256*i); ppmx_gpu_op takes its own argument.   */
    uint64_t *hist_out;
    /* extension: levels.  256 bytes on the host; ppmx_levels_lut_linear (ppmx_host.h) builds the usual
     * black-point / white-point stretch, and a histogram from HIST_GRAY picks the points.           */
    const uint8_t *levels_lut;
} ppmx_op;

typedef struct ppmx_gpu_ctx ppmx_gpu_ctx;     /* one CUDA device (or several), streams, a private buffer pool */
typedef struct ppmx_gpu_image ppmx_gpu_image; /* a raster resident in HBM                   */
typedef struct ppmx_gpu_graph ppmx_gpu_graph; /* a recorded sequence of raw launches        */
typedef struct ppmx_gpu_chain ppmx_gpu_chain; /* an op chain prepared for device-resident rasters */

/* ---- lifetime ------------------------------------------------------------------------ */

/* Binds a context to CUDA device `device`, creates its streams and a stream-ordered pool of its
 * own (the device's default pool is not touched).  Replaces nothing in the reference (it has no
 * state beyond the handler on main's stack, ref:119). */
int ppmx_gpu_init(ppmx_gpu_ctx **ctx, int device);
/* One context over several devices of this process (devices == NULL or ndev == 0: every visible
 * device).  ppmx_gpu_apply on such a context cuts ONE large raster into row bands, one per device
 * (BASELINE config 4: halo rows are uploaded with each band straight from the host raster);
 * ppmx_gpu_apply_batch deals whole rasters round-robin (config 5, image-parallel).  Everything
 * else (upload, op, download, raw memory) runs on the first device. */
int ppmx_gpu_init_multi(ppmx_gpu_ctx **ctx, const int *devices, int ndev);
int ppmx_gpu_device_count(const ppmx_gpu_ctx *ctx);
void ppmx_gpu_free(ppmx_gpu_ctx *ctx);

/* Pinned host memory for rasters, so upload/download run at PCIe speed (portable: usable from
 * every device of a multi-device context).  Replaces the per-row mallocs of getImageInfo /
 * image_buff_alloc (ref:440-449, 922-934).  host_register pins memory the caller already owns
 * (e.g. a shared mapping that several ranks of a job read their bands from). */
void *ppmx_gpu_host_alloc(ppmx_gpu_ctx *ctx, size_t bytes);
void ppmx_gpu_host_free(ppmx_gpu_ctx *ctx, void *p);
int ppmx_gpu_host_register(ppmx_gpu_ctx *ctx, void *p, size_t bytes);
int ppmx_gpu_host_unregister(ppmx_gpu_ctx *ctx, void *p);

/* ---- the reference-facing call: one op chain, host raster in, host raster out ------- */

/* Runs ops[0..nops) on the packed RGB raster `src_rgb` (w*h*3 bytes, exactly the bytes that
 * follow a P6 header, ref:316-318) with the reference's buff/new_buff hand-over rules
 * (ref:1084-1155, driven by ops[i].renew_before) and writes to `dst` exactly the bytes
 * putImageToFile emits after its header (ref:263-291): RGB triples, or .r bytes for PGM, or
 * packed bits for PBM.  Operators whose result the reference computes but never writes (the
 * leaked grey raster of "-gray -fh") are not computed; the bytes are the same.
 * A large raster is cut into row parts that go round-robin over the context's streams, so the
 * upload of one part, the kernels of the previous and the download of the one before overlap;
 * on a multi-device context the parts are spread over the devices as row bands. */
int ppmx_gpu_apply(ppmx_gpu_ctx *ctx, const ppmx_op *ops, int nops,
                   const uint8_t *src_rgb, uint32_t w, uint32_t h,
                   uint8_t *dst, size_t dst_cap, size_t *dst_bytes,
                   uint32_t *out_w, uint32_t *out_h, int *out_file_type);

/* Same chain over `count` equally sized rasters laid back to back in src/dst (config 5,
 * image-parallel): uploads, kernels and downloads of consecutive rasters overlap. dst_stride
 * is the distance between outputs in dst; every output has the same size. */
int ppmx_gpu_apply_batch(ppmx_gpu_ctx *ctx, const ppmx_op *ops, int nops,
                         const uint8_t *src_rgb, uint32_t w, uint32_t h, int count,
                         uint8_t *dst, size_t dst_stride, size_t *dst_bytes_each,
                         uint32_t *out_w, uint32_t *out_h, int *out_file_type);

/* One rank of a multi-process job (one process per GPU): the same as ppmx_gpu_apply, but only the
 * output rows of row band `band` of `nbands` are produced (cuts as ppmx_band_plan with align 4,
 * returned in band_y0 / band_rows).  src is the WHOLE w x h raster on the host -- only the rows the
 * band needs (its own plus halo rows: k/2 for a convolution, the table's reach for a resize height
 * pass, the mirrored band for a vertical flip) are read -- and dst the WHOLE output, of which only
 * the band's rows are written; typically both are one shared pinned mapping.  Chains holding a
 * 90 / 270 degree or free rotation can not be cut (nbands must be 1 for them). */
int ppmx_gpu_apply_band(ppmx_gpu_ctx *ctx, const ppmx_op *ops, int nops,
                        const uint8_t *src_rgb, uint32_t w, uint32_t h, int band, int nbands,
                        uint8_t *dst, size_t dst_cap, size_t *dst_bytes,
                        uint32_t *out_w, uint32_t *out_h, int *out_file_type,
                        uint32_t *band_y0, uint32_t *band_rows);

/* The rows ppmx_gpu_apply_band touches for band `band` of `nbands`: output rows [out_y0, out_y0 +
 * out_rows) are written, source rows [src_y0, src_y0 + src_rows) are read.  No device work. */
int ppmx_gpu_band_rows(const ppmx_op *ops, int nops, uint32_t w, uint32_t h, int band, int nbands,
                       uint32_t *out_y0, uint32_t *out_rows, uint32_t *src_y0, uint32_t *src_rows);

/* What a chain will produce for a w x h raster, without running it: geometry, writer's file type,
 * bytes, whether it can be cut into row bands, and the number of kernels one part launches. */
int ppmx_gpu_chain_info(const ppmx_op *ops, int nops, uint32_t w, uint32_t h, uint32_t *out_w,
                        uint32_t *out_h, int *out_file_type, size_t *out_bytes, int *splittable,
                        int *kernels);

/* The same chain for rasters that are ALREADY in HBM (a device-resident batch): prepare linearises and fuses the
 * chain for w x h rasters, uploads the resize tables and allocates the intermediate rasters once; run then only
 * launches kernels (d_src -> d_dst, caller-owned device memory, any stream; recordable with ppmx_gpu_graph_*).
 * One prepared chain serves one stream at a time (the intermediates are its own).  Histogram stages are not taken. */
int ppmx_gpu_chain_prepare(ppmx_gpu_ctx *ctx, const ppmx_op *ops, int nops, uint32_t w, uint32_t h,
                           ppmx_gpu_chain **chain);
int ppmx_gpu_chain_run(ppmx_gpu_chain *chain, const void *d_src, void *d_dst, void *stream);
/* geometry, writer's file type and bytes of the result; kernels per run; bytes the stages read + write per run */
int ppmx_gpu_chain_info2(const ppmx_gpu_chain *chain, uint32_t *out_w, uint32_t *out_h, int *out_file_type,
                         size_t *out_bytes, int *kernels, size_t *bytes_moved);
void ppmx_gpu_chain_free(ppmx_gpu_chain *chain);

/* ---- device-resident rasters: one operator per call ---------------------------------- */

/* getImageInfo's raster copy (ref:444-449) becomes one asynchronous H2D copy. */
int ppmx_gpu_upload(ppmx_gpu_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, int layout,
                    ppmx_gpu_image **img);
/* Uninitialised raster in HBM (image_buff_alloc, ref:922-934, without the memset). */
int ppmx_gpu_image_alloc(ppmx_gpu_ctx *ctx, uint32_t w, uint32_t h, int layout, ppmx_gpu_image **img);
void ppmx_gpu_image_free(ppmx_gpu_ctx *ctx, ppmx_gpu_image *img); /* releaseBuffer, ref:464-471 */
int ppmx_gpu_image_info(const ppmx_gpu_image *img, uint32_t *w, uint32_t *h, int *layout,
                        size_t *bytes, void **device_ptr);

/* One reference operator: reads `src`, allocates and returns `*dst` (the callee-allocates
 * rule of ref:996/961/706/818).  FLIP returns a new raster too; the in-place aliasing of
 * ref:896 is reproduced by the chain logic, not here.  HIST_GRAY returns no image
 * (*dst = NULL) and writes hist_out[256] on the host. */
int ppmx_gpu_op(ppmx_gpu_ctx *ctx, const ppmx_op *op, const ppmx_gpu_image *src,
                ppmx_gpu_image **dst, uint64_t *hist_out);

/* putImageToFile's raster loop (ref:263-291) for a device raster: converts to the byte
 * stream of `file_type` if needed and copies it to the host; blocks until it has landed. */
int ppmx_gpu_download(ppmx_gpu_ctx *ctx, const ppmx_gpu_image *img, int file_type,
                      uint8_t *dst, size_t dst_cap, size_t *dst_bytes);

int ppmx_gpu_sync(ppmx_gpu_ctx *ctx);

/* ---- raw launches on caller-owned device memory ---------------------------------------
 * For callers that manage HBM and streams themselves (bench.py measures device-resident
 * throughput this way).  d_src/d_dst are device pointers; `stream` is a CUstream /
 * cudaStream_t passed as void* (NULL = default stream); d_hist is 256 x u64 on the device,
 * accumulated into (HIST ops only).  d_top / d_bottom: optional device pointers (possibly
 * PEER memory of a neighbouring GPU) to the halo rows above / below a row band; see
 * ppmx_band below.  Nothing is allocated, copied or synchronised.
 * Memory contract: rasters whose rows are not 16-byte aligned are read as whole aligned 16-byte vectors (and flips of
 * such rasters as aligned 4-byte words), so up to 15 bytes before the first and after the last byte of a raster -- never
 * beyond the 16-byte granule that holds them -- may be READ (nothing outside a raster is ever written).  Memory from
 * cudaMalloc / a torch allocator (256-/512-byte granules) always satisfies this; a raster carved out of a larger buffer
 * must not start or end inside the last 16 bytes of a mapping.  Halo rows are never read beyond `halo` rows. */
typedef struct ppmx_band {
    uint32_t full_h;   /* height of the whole raster this band belongs to (0 = not a band)  */
    uint32_t y0;       /* first row of the band in the whole raster                          */
    const void *d_top; /* rows [y0-halo, y0) as a packed raster, or NULL at the image top    */
    const void *d_bottom; /* rows [y0+h, y0+h+halo), or NULL at the image bottom             */
    uint32_t halo;     /* rows available behind d_top / d_bottom                             */
    uint32_t out_y0;   /* imresize height pass only: this band produces output rows          */
    uint32_t out_rows; /*   [out_y0, out_y0 + out_rows) of the op's out_size rows            */
} ppmx_band;

int ppmx_gpu_launch(const ppmx_op *op, const void *d_src, uint32_t w, uint32_t h, int src_layout,
                    void *d_dst, const ppmx_band *band, void *d_hist, void *d_tables, void *stream);

/* A sequence of ppmx_gpu_launch calls recorded into a CUDA graph and replayed with one driver call:
 * begin puts `stream` into capture mode, the launches issued on it (and on streams forked from it
 * with events) are recorded instead of run, end instantiates the graph (*kernel_nodes = kernels in
 * it).  graph_launch enqueues the whole recording on a stream. */
int ppmx_gpu_graph_begin(void *stream);
int ppmx_gpu_graph_end(void *stream, ppmx_gpu_graph **graph, uint64_t *kernel_nodes);
int ppmx_gpu_graph_launch(ppmx_gpu_graph *graph, void *stream);
void ppmx_gpu_graph_free(ppmx_gpu_graph *graph);

/* Device-side copy of an imresize table pair for ppmx_gpu_launch (d_tables): returns a
 * device allocation holding weights then indices; release with ppmx_gpu_tables_free. */
int ppmx_gpu_tables_upload(const ppmx_op *op, void **d_tables);
void ppmx_gpu_tables_free(void *d_tables);

/* size in bytes of a raster */
size_t ppmx_gpu_layout_bytes(uint32_t w, uint32_t h, int layout);

/* Output geometry of one operator applied to a w x h raster of `layout`. */
int ppmx_gpu_op_output(const ppmx_op *op, uint32_t w, uint32_t h, int layout,
                       uint32_t *out_w, uint32_t *out_h, int *out_layout);

/* ---- multi-GPU (one process per GPU): peer access to a neighbour's band ---------------
 * Exports a raster's allocation as a 64-byte CUDA IPC handle, and maps a peer's handle
 * into this process; halo rows are then read over NVLink by the kernels themselves. */
int ppmx_gpu_device_alloc(ppmx_gpu_ctx *ctx, size_t bytes, void **device_ptr); /* exportable HBM */
void ppmx_gpu_device_free(ppmx_gpu_ctx *ctx, void *device_ptr);
/* plain copies on the context's first stream, blocking: kind 0 = host->device, 1 = device->host,
 * 2 = device->device (also between a local and a peer-mapped pointer) */
int ppmx_gpu_copy(ppmx_gpu_ctx *ctx, void *dst, const void *src, size_t bytes, int kind);
int ppmx_gpu_ipc_export(ppmx_gpu_ctx *ctx, const void *device_ptr, uint8_t handle[64]);
int ppmx_gpu_ipc_open(ppmx_gpu_ctx *ctx, const uint8_t handle[64], void **device_ptr);
int ppmx_gpu_ipc_close(ppmx_gpu_ctx *ctx, void *device_ptr);

/* Switches for benchmarking (not needed for correct results): key "pdl" turns programmatic
 * dependent launch on (1, default) or off (0), process-wide.  Key "variant" selects an alternative
 * kernel implementation (0 = default) in the TUNING build only (libppmx_gpu_tuning.so, the same
 * sources with -DPPMX_TUNING, loaded by tools/sweep.py and the variant tests); the release library
 * carries no alternatives and answers -1 to any variant but 0.  Returns 0, or -1. */
int ppmx_gpu_set_tuning(const char *key, int value);

/* number of kernel launches issued through this library since load (bench: gpu_launches) */
uint64_t ppmx_gpu_launch_count(void);
const char *ppmx_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PPMX_GPU_H */
