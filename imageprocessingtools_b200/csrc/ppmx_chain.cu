// ppmx_chain.cu -- the op chain of doProcessPPM (ref:1084-1155 = /root/reference/ppmx-edward.c) on the device:
// host raster in, the writer's raster (ref:263-291) out.
//
// 1. The chain is LINEARISED.  The reference hands buffers over through buff / new_buff with renewBuffer
//    (ref:1019-1026) under per-flag conditions (ref:1133-1153); an operator always reads `buff`.  Simulating those
//    hand-overs on node numbers instead of rasters gives the one path of operators that actually feeds the writer
//    (e.g. "-gray -fh" without -w/-r: flip of the ORIGINAL raster, then the writer's .r extraction -- the grey raster
//    is computed and leaked by the reference and simply not computed here; same bytes out).
// 2. Every stage of that path maps a range of its output rows to the range of input rows it needs (pointwise: the
//    same rows; k x k convolution: k/2 rows more on either side; vertical flip / 180 degrees: the mirrored range;
//    resize height pass: the rows its contribution table names).  A "part" = a range of OUTPUT rows; walking the
//    stages backwards gives the rows of the SOURCE raster to upload for it -- halo rows come straight from the host
//    raster, no device-to-device exchange is needed when the source is on the host.
// 3. Parts are spread over the context's streams (upload of part n+1, kernels of part n and download of part n-1
//    overlap, also for ONE raster) and, for a multi-device context or one rank of a multi-process job, over GPUs
//    as row bands (BASELINE config 4).
// There is no CPU fallback: every byte of every output row is produced by a kernel launch.
#include "ppmx_ctx.h"

#include <cstdlib>
#include <cstring>
#include <new>

using namespace ppmx;

namespace {

constexpr int kFusedKind = 1000;  // internal stage kind: geometry + pointwise tail in one kernel (ppmx_fused.cu)

struct Stage {
    ppmx_op op;
    GeomOp go = {};     // kFusedKind only
    int op_index = -1;  // position in the caller's op list (device tables); -1: the writer's conversion
    uint32_t in_w = 0, in_h = 0, out_w = 0, out_h = 0;
    int in_layout = PPMX_LAYOUT_RGB8, out_layout = PPMX_LAYOUT_RGB8;
    bool passthrough = false;  // produces no raster (histogram of a superseded grey raster)
};

struct Pipeline {
    std::vector<Stage> st;
    uint32_t w = 0, h = 0, out_w = 0, out_h = 0;
    int out_layout = PPMX_LAYOUT_RGB8, file_type = PPMX_FILETYPE_PPM;
    bool splittable = true;
    int hist_stage = -1;
    uint64_t *hist_out = nullptr;
    size_t in_pitch() const { return row_bytes(w, PPMX_LAYOUT_RGB8); }
    size_t out_pitch() const { return row_bytes(out_w, out_layout); }
    size_t out_bytes() const { return out_pitch() * out_h; }
};

bool is_hist(int kind) { return kind == PPMX_OP_GRAY_HIST || kind == PPMX_OP_HIST_GRAY; }

// can this stage produce a range of output rows from a range of input rows?
bool stage_splits(const Stage &s)
{
    switch (s.op.kind) {
    case PPMX_OP_ROTATE: return s.op.angle_deg == 180;  // 90 / 270: a transpose; other angles: a slanted source strip
    case kFusedKind: return !s.go.transpose;
    default: return true;
    }
}

// one output row <-> one input row (no neighbours): later stages of this kind keep a histogram's rows a partition
bool stage_row_exact(const Stage &s)
{
    switch (s.op.kind) {
    case PPMX_OP_CONV: return s.op.conv_k <= 1;
    case PPMX_OP_IMRESIZE: return s.op.dim != 0;
    case PPMX_OP_ROTATE: return s.op.angle_deg == 180;
    case kFusedKind: return !s.go.transpose;
    default: return true;
    }
}

// ---- fusion: runs of flips / right-angle rotations / grey / mono / the writer's conversions become ONE stage ----

// where a pixel of the group's source raster sits after the geometric stages seen so far:
// cx = sx * (swap ? y : x) + tx,  cy = sy * (swap ? x : y) + ty   (a signed permutation: the 8 orientations)
struct Aff {
    bool swap = false;
    int sx = 1, sy = 1;
    long tx = 0, ty = 0;
    uint32_t cw = 0, ch = 0;
    void fliph() { sx = -sx; tx = (long)cw - 1 - tx; }                 // ref:906-911
    void flipv() { sy = -sy; ty = (long)ch - 1 - ty; }                 // ref:899-904
    void rot90()                                                        // new[x][newW-1-y] = old[y][x], ref:717
    {
        const int nsx = -sy, nsy = sx;
        const long ntx = (long)ch - 1 - ty, nty = tx;
        swap = !swap; sx = nsx; sy = nsy; tx = ntx; ty = nty;
        const uint32_t t = cw; cw = ch; ch = t;
    }
    void rot270()                                                       // new[newH-1-y][x] = old[x][y], ref:725
    {
        const int nsx = sy, nsy = -sx;
        const long ntx = ty, nty = (long)cw - 1 - tx;
        swap = !swap; sx = nsx; sy = nsy; tx = ntx; ty = nty;
        const uint32_t t = cw; cw = ch; ch = t;
    }
};

bool fusable_kind(const Stage &s)
{
    switch (s.op.kind) {
    case PPMX_OP_FLIP:
    case PPMX_OP_GRAY:
    case PPMX_OP_MONO:
    case PPMX_OP_MONO_BITS:
    case PPMX_OP_EXTRACT_R:
    case PPMX_OP_PACK_PBM: return !s.passthrough;
    case PPMX_OP_ROTATE: return s.op.angle_deg == 90 || s.op.angle_deg == 180 || s.op.angle_deg == 270;
    default: return false;
    }
}

// tries to replace stages [i0, i1) by one fused stage; false = leave them as they are
bool fuse_run(const std::vector<Stage> &st, size_t i0, size_t i1, Stage *out)
{
    if (st[i0].in_layout != PPMX_LAYOUT_RGB8) return false;
    Aff cur, at_mono;
    cur.cw = st[i0].in_w;
    cur.ch = st[i0].in_h;
    int point = 0;  // 0 rgb, 1 grey, 2 red, 3 mono (bits pending), 4 mono packed
    for (size_t i = i0; i < i1; i++) {
        const ppmx_op &op = st[i].op;
        switch (op.kind) {
        case PPMX_OP_FLIP:
            if (op.flip_direction) cur.flipv();
            else cur.fliph();
            break;
        case PPMX_OP_ROTATE:
            if (op.angle_deg == 90) cur.rot90();
            else if (op.angle_deg == 270) cur.rot270();
            else { cur.fliph(); cur.flipv(); }  // ref:721
            break;
        case PPMX_OP_GRAY:
            if (point) return false;
            point = 1;
            break;
        case PPMX_OP_EXTRACT_R:
            if (point) return false;
            point = 2;
            break;
        case PPMX_OP_MONO:
        case PPMX_OP_MONO_BITS:
            if (point) return false;
            point = op.kind == PPMX_OP_MONO ? 3 : 4;
            at_mono = cur;
            break;
        case PPMX_OP_PACK_PBM:
            if (point != 3) return false;  // the packer on raw .r bytes (the "-mono -fh" quirk) keeps its own kernel
            point = 4;
            break;
        default: return false;
        }
    }
    if (point == 3) return false;  // a 0/1 plane that is not packed inside the run
    Stage f = st[i0];
    memset(&f.op, 0, sizeof(f.op));
    f.op.kind = kFusedKind;
    f.op_index = -1;
    f.out_w = st[i1 - 1].out_w;
    f.out_h = st[i1 - 1].out_h;
    f.out_layout = st[i1 - 1].out_layout;
    GeomOp &g = f.go;
    g.transpose = cur.swap ? 1 : 0;
    g.rev_x = cur.sx < 0;
    g.rev_y = cur.sy < 0;
    g.point = point == 4 ? 3 : point;
    if (point == 4) {
        g.mx_from_y = at_mono.swap ? 1 : 0;
        g.mx_neg = at_mono.sx < 0;
        g.my_neg = at_mono.sy < 0;
        g.mx_add = (int)(((at_mono.tx % 4) + 4) % 4);
        g.my_add = (int)(((at_mono.ty % 4) + 4) % 4);
    }
    if (!geom_point_supported(f.in_w, f.in_h, g)) return false;
    *out = f;
    return true;
}

void fuse_pipeline(std::vector<Stage> &st)
{
    if (getenv("PPMX_NO_FUSE")) return;  // (tests compare the fused chain with the stage-by-stage one)
    std::vector<Stage> out;
    for (size_t i = 0; i < st.size();) {
        size_t j = i;
        while (j < st.size() && fusable_kind(st[j])) j++;
        bool done = false;
        // the longest fusable run that starts here; a single stage keeps its own (faster) kernel unless it is a
        // 90 / 270 degree rotation of a raster the bulk-copy transposer can not take (sides not multiples of 16)
        for (size_t e = j; e > i && !done; e--) {
            const bool lone = e - i == 1;
            if (lone) {
                const ppmx_op &op = st[i].op;
                const bool odd_transpose = op.kind == PPMX_OP_ROTATE && op.angle_deg != 180 && ((st[i].in_w | st[i].in_h) & 15u);
                if (!odd_transpose) break;
            }
            Stage f;
            if (fuse_run(st, i, e, &f)) {
                out.push_back(f);
                i = e;
                done = true;
            }
        }
        if (!done) out.push_back(st[i++]);
    }
    st.swap(out);
}

int build_pipeline(const ppmx_op *ops, int nops, uint32_t w, uint32_t h, Pipeline *P)
{
    struct Node {
        int parent, op_index;
    };
    std::vector<Node> nodes(1, Node{-1, -1});  // node 0 = the uploaded raster
    int buff = 0, newb = -1, ft = PPMX_FILETYPE_PPM;
    for (int i = 0; i < nops; i++) {
        const ppmx_op &op = ops[i];
        if (op.renew_before && newb >= 0) {  // renewBuffer, ref:1019-1026
            buff = newb;
            newb = -1;
        }
        switch (op.kind) {
        case PPMX_OP_GRAY:
        case PPMX_OP_GRAY_HIST:
        case PPMX_OP_MONO:
            nodes.push_back(Node{buff, i});
            newb = (int)nodes.size() - 1;
            ft = op.kind == PPMX_OP_MONO ? PPMX_FILETYPE_PBM : PPMX_FILETYPE_PGM;  // ref:956, 991
            break;
        case PPMX_OP_FLIP:  // ref:896: works on buff itself and aliases new_buff to it
            nodes.push_back(Node{buff, i});
            buff = newb = (int)nodes.size() - 1;
            break;
        case PPMX_OP_ROTATE:
            if (op.angle_deg == 0) {  // ref:701-705
                newb = buff;
                break;
            }
            /* fall through */
        case PPMX_OP_IMRESIZE:
        case PPMX_OP_CONV:
        case PPMX_OP_LEVELS:
            nodes.push_back(Node{buff, i});
            newb = (int)nodes.size() - 1;
            break;
        default:
            return fail("operator not allowed in a chain");
        }
    }
    if (newb < 0) return fail("Error: no data to write");  // ref:235

    std::vector<int> path;  // node numbers from the first operator to the raster the writer gets
    for (int n = newb; n > 0; n = nodes[n].parent) path.insert(path.begin(), n);

    P->st.clear();
    P->w = w;
    P->h = h;
    P->file_type = ft;
    P->hist_stage = -1;
    P->hist_out = nullptr;
    uint32_t cw = w, ch = h;
    int cl = PPMX_LAYOUT_RGB8;
    auto side_hists = [&](int parent_node) {
        // a gray+hist raster that a later operator supersedes (its histogram is still asked for): counted on its
        // input as a stage that hands its raster through untouched
        for (size_t n = 1; n < nodes.size(); n++) {
            const ppmx_op &op = ops[nodes[n].op_index];
            bool on_path = false;
            for (int q : path) on_path = on_path || q == (int)n;
            if (op.kind != PPMX_OP_GRAY_HIST || on_path || nodes[n].parent != parent_node) continue;
            Stage s;
            s.op = op;
            s.op.kind = PPMX_OP_HIST_GRAY;
            s.op_index = nodes[n].op_index;
            s.in_w = s.out_w = cw;
            s.in_h = s.out_h = ch;
            s.in_layout = s.out_layout = cl;
            if (cl != PPMX_LAYOUT_RGB8) continue;  // (a grey raster's grey histogram: not a chain the planner builds)
            s.passthrough = true;
            P->st.push_back(s);
        }
    };
    side_hists(0);
    for (int n : path) {
        Stage s;
        s.op = ops[nodes[n].op_index];
        s.op_index = nodes[n].op_index;
        s.in_w = cw;
        s.in_h = ch;
        s.in_layout = cl;
        if (ppmx_gpu_op_output(&s.op, cw, ch, cl, &s.out_w, &s.out_h, &s.out_layout) != PPMX_OK)
            return fail("operator does not accept this raster layout");
        if (s.op.kind == PPMX_OP_IMRESIZE && (s.op.weights_sz < 1 || !s.op.weights || !s.op.indices))
            return fail("imresize: bad tables");
        if (s.op.kind == PPMX_OP_CONV && !s.op.conv_coef) return fail("conv: no coefficients");
        if (s.op.kind == PPMX_OP_LEVELS && !s.op.levels_lut) return fail("levels: no table");
        P->st.push_back(s);
        cw = s.out_w;
        ch = s.out_h;
        cl = s.out_layout;
        side_hists(n);
    }
    // the writer's raster loop (ref:263-291): .r bytes for P5, packed bits for P4, triples for P6
    int conv_kind = -1;
    if (ft == PPMX_FILETYPE_PGM) {
        if (cl == PPMX_LAYOUT_RGB8) conv_kind = PPMX_OP_EXTRACT_R;
        else if (cl != PPMX_LAYOUT_R8) return fail("PGM output from a bit raster");
    } else if (ft == PPMX_FILETYPE_PBM) {
        if (cl != PPMX_LAYOUT_BITS) conv_kind = PPMX_OP_PACK_PBM;
    } else if (cl != PPMX_LAYOUT_RGB8) {
        return fail("PPM output needs an RGB raster");
    }
    if (conv_kind >= 0) {
        if (conv_kind == PPMX_OP_PACK_PBM && !P->st.empty() && P->st.back().op.kind == PPMX_OP_MONO) {
            // a bilevel result nothing else touches goes straight to packed bits (mono + ref:268-284 in one kernel)
            P->st.back().op.kind = PPMX_OP_MONO_BITS;
            P->st.back().out_layout = PPMX_LAYOUT_BITS;
        } else {
            Stage s;
            memset(&s.op, 0, sizeof(s.op));
            s.op.kind = conv_kind;
            s.in_w = s.out_w = cw;
            s.in_h = s.out_h = ch;
            s.in_layout = cl;
            s.out_layout = conv_kind == PPMX_OP_EXTRACT_R ? PPMX_LAYOUT_R8 : PPMX_LAYOUT_BITS;
            P->st.push_back(s);
        }
        cl = P->st.back().out_layout;
    }
    P->out_w = cw;
    P->out_h = ch;
    P->out_layout = cl;
    fuse_pipeline(P->st);

    P->splittable = true;
    for (size_t i = 0; i < P->st.size(); i++) {
        const Stage &s = P->st[i];
        if (!stage_splits(s)) P->splittable = false;
        if (is_hist(s.op.kind)) {
            if (P->hist_stage >= 0) return fail("one histogram per chain");
            if (!s.op.hist_out) return fail("gray+hist in a chain needs op.hist_out");
            P->hist_stage = (int)i;
            P->hist_out = s.op.hist_out;
            // every row must be counted exactly once: the rows of the parts at this stage have to be a partition
            for (size_t j = i + 1; j < P->st.size(); j++)
                if (!stage_row_exact(P->st[j])) P->splittable = false;
        }
    }
    return PPMX_OK;
}

struct Rows {
    uint32_t lo, hi;
};

// rows of the stage's INPUT needed for rows [a, b) of its output
Rows stage_in_rows(const Stage &s, uint32_t a, uint32_t b)
{
    const uint32_t H = s.in_h;
    switch (s.op.kind) {
    case kFusedKind:
        if (s.go.transpose) return Rows{0, H};
        return s.go.rev_y ? Rows{H - b, H - a} : Rows{a, b};
    case PPMX_OP_FLIP:
        if (s.op.flip_direction) return Rows{H - b, H - a};  // ref:899-904
        return Rows{a, b};
    case PPMX_OP_ROTATE:
        if (s.op.angle_deg == 180) return Rows{H - b, H - a};  // ref:718-721
        return Rows{0, H};
    case PPMX_OP_CONV: {
        const uint32_t r = (uint32_t)s.op.conv_k / 2;  // mirrored rows (ref:551-555 idiom) lie within r of the edge
        return Rows{a > r ? a - r : 0, b + r < H ? b + r : H};
    }
    case PPMX_OP_IMRESIZE:
        if (s.op.dim == 0) {  // ref:826-830: the rows named by the table
            int lo = (int)H, hi = -1;
            const int K = s.op.weights_sz;
            for (uint32_t y = a; y < b; y++)
                for (int z = 0; z < K; z++) {
                    const int v = s.op.indices[(size_t)y * K + z];
                    lo = v < lo ? v : lo;
                    hi = v > hi ? v : hi;
                }
            if (hi < lo) return Rows{0, 0};
            if (lo < 0) lo = 0;
            if (hi >= (int)H) hi = (int)H - 1;
            return Rows{(uint32_t)lo, (uint32_t)hi + 1};
        }
        return Rows{a, b};
    default:
        return Rows{a, b};
    }
}

// Launches one stage: S holds rows [in.lo, in.hi) of the stage's input, D receives rows [a, b) of its output.
int launch_stage(const Stage &s, const uint8_t *S, Rows in, uint8_t *D, uint32_t a, uint32_t b, unsigned long long *d_hist,
                 const DeviceTables *tables, cudaStream_t stream)
{
    const size_t pitch = row_bytes(s.in_w, s.in_layout);
    const bool whole = a == 0 && b == s.out_h && in.lo == 0 && in.hi == s.in_h;
    Band band;
    switch (s.op.kind) {
    case kFusedKind: {
        GeomOp go = s.go;
        cudaError_t e;
        if (go.transpose) {
            if (!whole) return fail("internal: a transposing stage needs the whole raster");
            e = geom_point(S, D, s.in_w, s.in_h, 0, go, stream);
        } else {
            const uint32_t first = go.rev_y ? s.in_h - b : a;  // first source row of this part
            // mono's Bayer phase is counted in whole-raster rows: shift it by the part's first source row
            const int shift = (int)(first & 3u);
            if (go.mx_from_y) go.mx_add = (go.mx_add + (go.mx_neg ? 4 - shift : shift)) & 3;
            else go.my_add = (go.my_add + (go.my_neg ? 4 - shift : shift)) & 3;
            e = geom_point(S + (size_t)(first - in.lo) * pitch, D, s.in_w, b - a, 0, go, stream);
        }
        if (e != cudaSuccess) return fail("fused geometry stage", e);
        return PPMX_OK;
    }
    case PPMX_OP_ROTATE:
        if (s.op.angle_deg == 180) {
            ppmx_op op = s.op;  // the mirrored rows, turned as a raster of their own
            return launch_op(&op, S + (size_t)((s.in_h - b) - in.lo) * pitch, s.in_w, b - a, s.in_layout, D, band, nullptr, nullptr,
                             stream);
        }
        if (!whole) return fail("internal: this rotation needs the whole raster");
        return launch_op(&s.op, S, s.in_w, s.in_h, s.in_layout, D, band, nullptr, nullptr, stream);
    case PPMX_OP_FLIP:
        if (s.op.flip_direction)
            return launch_op(&s.op, S + (size_t)((s.in_h - b) - in.lo) * pitch, s.in_w, b - a, s.in_layout, D, band, nullptr, nullptr,
                             stream);
        return launch_op(&s.op, S + (size_t)(a - in.lo) * pitch, s.in_w, b - a, s.in_layout, D, band, nullptr, nullptr, stream);
    case PPMX_OP_CONV: {
        if (whole) return launch_op(&s.op, S, s.in_w, s.in_h, s.in_layout, D, band, nullptr, nullptr, stream);
        const uint32_t r = (uint32_t)s.op.conv_k / 2;
        band.full_h = s.in_h;
        band.y0 = a;
        band.halo = r;
        // rows [a - r, a) and [b, b + r) lie in the same buffer (address arithmetic on integers: for a < r the
        // "row a - r" is a virtual row before the raster's first one, never dereferenced)
        const intptr_t base = (intptr_t)S - (intptr_t)in.lo * (intptr_t)pitch;
        band.top = a > 0 ? (const uint8_t *)(base + ((intptr_t)a - (intptr_t)r) * (intptr_t)pitch) : nullptr;
        band.bottom = b < s.in_h ? (const uint8_t *)(base + (intptr_t)b * (intptr_t)pitch) : nullptr;
        return launch_op(&s.op, (const uint8_t *)(base + (intptr_t)a * (intptr_t)pitch), s.in_w, b - a, s.in_layout, D, band, nullptr,
                         nullptr, stream);
    }
    case PPMX_OP_IMRESIZE:
        if (s.op.dim == 0) {
            if (!whole) {
                band.full_h = s.in_h;
                band.y0 = in.lo;
                band.out_y0 = a;
                band.out_rows = b - a;
            }
            return launch_op(&s.op, S, s.in_w, in.hi - in.lo, s.in_layout, D, band, nullptr, tables, stream);
        }
        return launch_op(&s.op, S + (size_t)(a - in.lo) * pitch, s.in_w, b - a, s.in_layout, D, band, nullptr, tables, stream);
    default:  // pointwise stages; mono takes the first row's number for the Bayer phase (ref:967)
        if (!whole) {
            band.full_h = s.in_h;
            band.y0 = a;
        }
        return launch_op(&s.op, S + (size_t)(a - in.lo) * pitch, s.in_w, b - a, s.in_layout, D, band, d_hist, tables, stream);
    }
}

// What one device does for one call: tables and histogram bins in HBM, parts in flight per lane.
struct DeviceJob {
    ppmx_gpu_ctx *c = nullptr;
    std::vector<DeviceTables> tables;  // by caller op index
    cudaEvent_t done[kLanes][kInFlight] = {};
    unsigned issued[kLanes] = {};
    unsigned long long *d_hist = nullptr;  // count x 256 bins
    int count = 0;

    int begin(ppmx_gpu_ctx *ctx, const Pipeline &P, int nops, int nrasters)
    {
        c = ctx;
        count = nrasters;
        PPMX_CK(cudaSetDevice(c->device), "cudaSetDevice");
        tables.assign(nops, DeviceTables());
        bool any = false;
        for (const Stage &s : P.st)
            if (s.op.kind == PPMX_OP_IMRESIZE && s.op_index >= 0 && !tables[s.op_index].base) {
                if (upload_tables(c, &s.op, &tables[s.op_index], c->lane[0]) != PPMX_OK) return PPMX_ERROR;
                any = true;
            }
        if (P.hist_stage >= 0) {
            const size_t nb = (size_t)count * 256 * sizeof(unsigned long long);
            PPMX_CK(pool_alloc(c, (void **)&d_hist, nb, c->lane[0]), "pool alloc (hist)");
            PPMX_CK(cudaMemsetAsync(d_hist, 0, nb, c->lane[0]), "clear hist");
            any = true;
        }
        if (any) {
            PPMX_CK(cudaEventRecord(c->tables_ready, c->lane[0]), "event record");
            for (int l = 1; l < kLanes; l++) PPMX_CK(cudaStreamWaitEvent(c->lane[l], c->tables_ready, 0), "event wait");
        }
        return PPMX_OK;
    }

    // rows [a, b) of the output of raster `index`: upload what they need, run the stages, download
    int part(const Pipeline &P, int lane, const uint8_t *src, uint8_t *dst, uint32_t a, uint32_t b, int index)
    {
        if (a >= b) return PPMX_OK;
        PPMX_CK(cudaSetDevice(c->device), "cudaSetDevice");
        cudaStream_t s = c->lane[lane];
        const int slot = (int)(issued[lane]++ % kInFlight);
        // at most kInFlight parts per lane are enqueued ahead of the GPU, so the pool holds a bounded number of
        // rasters however long the batch is
        if (done[lane][slot]) PPMX_CK(cudaEventSynchronize(done[lane][slot]), "event sync");
        else PPMX_CK(cudaEventCreateWithFlags(&done[lane][slot], cudaEventDisableTiming), "event create");

        const size_t m = P.st.size();
        std::vector<Rows> need(m + 1);
        need[m] = Rows{a, b};
        for (size_t i = m; i-- > 0;) {
            need[i] = P.st[i].passthrough ? need[i + 1] : stage_in_rows(P.st[i], need[i + 1].lo, need[i + 1].hi);
            if (need[i].hi <= need[i].lo) return fail("internal: empty source range");
        }
        uint8_t *S = nullptr;
        const size_t in_bytes = (size_t)(need[0].hi - need[0].lo) * P.in_pitch();
        PPMX_CK(pool_alloc(c, (void **)&S, in_bytes + 16, s), "can not allocate image buff in HBM");  // wording of ref:926
        int rc = PPMX_OK;
        if (cudaMemcpyAsync(S, src + (size_t)need[0].lo * P.in_pitch(), in_bytes, cudaMemcpyHostToDevice, s) != cudaSuccess)
            rc = fail("upload");
        for (size_t i = 0; i < m && rc == PPMX_OK; i++) {
            const Stage &st = P.st[i];
            unsigned long long *dh = ((int)i == P.hist_stage) ? d_hist + 256 * (size_t)index : nullptr;
            const DeviceTables *t = (st.op.kind == PPMX_OP_IMRESIZE && st.op_index >= 0) ? &tables[st.op_index] : nullptr;
            if (st.passthrough) {
                rc = launch_stage(st, S, need[i], nullptr, need[i + 1].lo, need[i + 1].hi, dh, t, s);
                continue;
            }
            uint8_t *D = nullptr;
            const size_t nb = (size_t)(need[i + 1].hi - need[i + 1].lo) * row_bytes(st.out_w, st.out_layout);
            if (pool_alloc(c, (void **)&D, nb + 16, s) != cudaSuccess) {
                rc = fail("can not allocate image buff in HBM");
                break;
            }
            rc = launch_stage(st, S, need[i], D, need[i + 1].lo, need[i + 1].hi, dh, t, s);
            cudaFreeAsync(S, s);
            S = D;
        }
        if (rc == PPMX_OK && cudaMemcpyAsync(dst + (size_t)a * P.out_pitch(), S, (size_t)(b - a) * P.out_pitch(),
                                              cudaMemcpyDeviceToHost, s) != cudaSuccess)
            rc = fail("download");
        cudaFreeAsync(S, s);
        cudaEventRecord(done[lane][slot], s);
        return rc;
    }

    // after the last part: the bins go to `host` (count x 256) once every lane has finished counting
    int hist_download(uint64_t *host)
    {
        if (!d_hist) return PPMX_OK;
        PPMX_CK(cudaSetDevice(c->device), "cudaSetDevice");
        for (int l = 1; l < kLanes; l++) {
            PPMX_CK(cudaEventRecord(c->lane_done[l], c->lane[l]), "event record");
            PPMX_CK(cudaStreamWaitEvent(c->lane[0], c->lane_done[l], 0), "event wait");
        }
        PPMX_CK(cudaMemcpyAsync(host, d_hist, (size_t)count * 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->lane[0]),
                "hist D2H");
        return PPMX_OK;
    }

    // waits for everything this job enqueued and releases what it holds; safe after any failure
    int finish()
    {
        if (!c) return PPMX_OK;
        int rc = PPMX_OK;
        cudaSetDevice(c->device);
        for (int l = 0; l < kLanes; l++) {
            cudaError_t e = cudaStreamSynchronize(c->lane[l]);
            if (e != cudaSuccess && rc == PPMX_OK) rc = fail("stream sync", e);
            for (int k = 0; k < kInFlight; k++)
                if (done[l][k]) {
                    cudaEventDestroy(done[l][k]);
                    done[l][k] = nullptr;
                }
        }
        for (auto &t : tables)
            if (t.base) cudaFreeAsync(t.base, c->lane[0]);
        tables.clear();
        if (d_hist) cudaFreeAsync(d_hist, c->lane[0]);
        d_hist = nullptr;
        c = nullptr;
        return rc;
    }
    ~DeviceJob() { finish(); }
};

// rows [*a, *b) of `total` for piece i of n; starts are multiples of `align` (the same cut as ppmx_band_plan, ppmx_host.c)
void split_rows(uint32_t total, uint32_t n, uint32_t i, uint32_t align, uint32_t *a, uint32_t *b)
{
    const uint32_t units = (total + align - 1) / align, base = units / n, extra = units % n;
    uint32_t u0 = i * base + (i < extra ? i : extra), u1 = u0 + base + (i < extra ? 1u : 0u);
    u0 *= align;
    u1 *= align;
    *a = u0 > total ? total : u0;
    *b = u1 > total ? total : u1;
}

constexpr size_t kPartBytes = 4u << 20;  // a part moves about this much over PCIe, once a raster is at least twice as large
constexpr uint32_t kMaxParts = 32;

// PPMX_PART_BYTES in the environment overrides the part size (tests cut small rasters into many parts with it)
size_t part_bytes()
{
    const char *e = getenv("PPMX_PART_BYTES");
    const long long v = e ? atoll(e) : 0;
    return v > 0 ? (size_t)v : kPartBytes;
}

uint32_t parts_for(const Pipeline &P, uint32_t rows)
{
    if (!P.splittable || rows < 8) return 1;
    const size_t bytes = (size_t)rows * (P.in_pitch() > P.out_pitch() ? P.in_pitch() : P.out_pitch());
    const size_t per = part_bytes();
    if (bytes < 2 * per) return 1;
    size_t n = (bytes + per - 1) / per;
    if (n > kMaxParts) n = kMaxParts;
    if (n > rows / 4) n = rows / 4;
    return (uint32_t)(n < 1 ? 1 : n);
}

// The common body of ppmx_gpu_apply / _batch / _band.  band_n > 0: only output rows of band `band_i` of `band_n`
// (one rank of a multi-process job); else everything, over all devices of the context.
int apply_impl(ppmx_gpu_ctx *ctx, const ppmx_op *ops, int nops, const uint8_t *src, uint32_t w, uint32_t h, int count,
               uint8_t *dst, size_t dst_stride, int band_i, int band_n, size_t *dst_bytes_each, uint32_t *out_w, uint32_t *out_h,
               int *out_file_type, uint32_t *band_y0, uint32_t *band_rows)
{
    if (!ctx || !ops || nops < 1 || !src || !dst || count < 1) return fail("ppmx_gpu_apply: bad argument");
    Pipeline P;
    if (build_pipeline(ops, nops, w, h, &P) != PPMX_OK) return PPMX_ERROR;
    if (P.out_bytes() > dst_stride) return fail("destination buffer too small");

    std::vector<ppmx_gpu_ctx *> devs;
    if (ctx->children.empty()) devs.push_back(ctx);
    else devs = ctx->children;
    if (band_n > 0) devs.resize(1);  // a rank drives one device
    if (band_n > 1 && !P.splittable) return fail("this op chain can not be cut into row bands (it holds a 90/270 degree or free rotation)");
    const uint32_t ndev = (uint32_t)devs.size();

    std::vector<DeviceJob> jobs(ndev);
    int rc = PPMX_OK;
    for (uint32_t d = 0; d < ndev && rc == PPMX_OK; d++) rc = jobs[d].begin(devs[d], P, nops, count);

    uint32_t my_a = 0, my_b = P.out_h;
    if (band_n > 0) split_rows(P.out_h, (uint32_t)band_n, (uint32_t)band_i, 4, &my_a, &my_b);
    const size_t in_each = (size_t)w * h * 3;
    // one raster over several devices as row bands (config 4) when it can be cut; a batch goes image-parallel
    const bool bands_over_devices = ndev > 1 && count == 1 && P.splittable;
    if (rc == PPMX_OK) {
        if (bands_over_devices) {
            std::vector<Rows> dev_rows(ndev);
            uint32_t nparts = 1;
            for (uint32_t d = 0; d < ndev; d++) {
                split_rows(P.out_h, ndev, d, 4, &dev_rows[d].lo, &dev_rows[d].hi);
                const uint32_t n = parts_for(P, dev_rows[d].hi - dev_rows[d].lo);
                nparts = n > nparts ? n : nparts;
            }
            for (uint32_t j = 0; j < nparts && rc == PPMX_OK; j++)      // part j of every device before part j + 1 of any:
                for (uint32_t d = 0; d < ndev && rc == PPMX_OK; d++) {  // all links start moving at once
                    uint32_t a, b;
                    split_rows(dev_rows[d].hi - dev_rows[d].lo, nparts, j, 4, &a, &b);
                    rc = jobs[d].part(P, (int)(j % kLanes), src, dst, dev_rows[d].lo + a, dev_rows[d].lo + b, 0);
                }
        } else {
            const uint32_t nparts = parts_for(P, my_b - my_a);
            unsigned seq = 0;
            for (int i = 0; i < count && rc == PPMX_OK; i++) {
                DeviceJob &job = jobs[(uint32_t)i % ndev];
                for (uint32_t j = 0; j < nparts && rc == PPMX_OK; j++, seq++) {
                    uint32_t a, b;
                    split_rows(my_b - my_a, nparts, j, 4, &a, &b);
                    rc = job.part(P, (int)((seq / ndev) % kLanes), src + (size_t)i * in_each, dst + (size_t)i * dst_stride, my_a + a,
                                  my_a + b, i);
                }
            }
        }
    }
    // histogram bins: straight into the caller's array from one device, summed over the devices otherwise
    std::vector<std::vector<uint64_t>> partial;
    if (rc == PPMX_OK && P.hist_stage >= 0) {
        if (ndev == 1) rc = jobs[0].hist_download(P.hist_out);
        else {
            partial.assign(ndev, std::vector<uint64_t>((size_t)count * 256));
            for (uint32_t d = 0; d < ndev && rc == PPMX_OK; d++) rc = jobs[d].hist_download(partial[d].data());
        }
    }
    for (uint32_t d = 0; d < ndev; d++) {  // always: nothing may still be writing into dst when this call returns
        const int r = jobs[d].finish();
        if (rc == PPMX_OK) rc = r;
    }
    if (rc != PPMX_OK) return rc;
    if (!partial.empty())
        for (size_t k = 0; k < (size_t)count * 256; k++) {
            uint64_t t = 0;
            for (uint32_t d = 0; d < ndev; d++) t += partial[d][k];
            P.hist_out[k] = t;
        }
    if (dst_bytes_each) *dst_bytes_each = P.out_bytes();
    if (out_w) *out_w = P.out_w;
    if (out_h) *out_h = P.out_h;
    if (out_file_type) *out_file_type = P.file_type;
    if (band_y0) *band_y0 = my_a;
    if (band_rows) *band_rows = my_b - my_a;
    return PPMX_OK;
}

}  // namespace

extern "C" int ppmx_gpu_apply_batch(ppmx_gpu_ctx *c, const ppmx_op *ops, int nops, const uint8_t *src, uint32_t w, uint32_t h,
                                    int count, uint8_t *dst, size_t dst_stride, size_t *dst_bytes_each, uint32_t *out_w,
                                    uint32_t *out_h, int *out_file_type)
{
    return apply_impl(c, ops, nops, src, w, h, count, dst, dst_stride, 0, 0, dst_bytes_each, out_w, out_h, out_file_type, nullptr,
                      nullptr);
}

extern "C" int ppmx_gpu_apply(ppmx_gpu_ctx *c, const ppmx_op *ops, int nops, const uint8_t *src, uint32_t w, uint32_t h,
                              uint8_t *dst, size_t dst_cap, size_t *dst_bytes, uint32_t *out_w, uint32_t *out_h,
                              int *out_file_type)
{
    return apply_impl(c, ops, nops, src, w, h, 1, dst, dst_cap, 0, 0, dst_bytes, out_w, out_h, out_file_type, nullptr, nullptr);
}

extern "C" int ppmx_gpu_apply_band(ppmx_gpu_ctx *c, const ppmx_op *ops, int nops, const uint8_t *src, uint32_t w, uint32_t h,
                                   int band, int nbands, uint8_t *dst, size_t dst_cap, size_t *dst_bytes, uint32_t *out_w,
                                   uint32_t *out_h, int *out_file_type, uint32_t *band_y0, uint32_t *band_rows)
{
    if (nbands < 1 || band < 0 || band >= nbands) return fail("ppmx_gpu_apply_band: bad band number");
    return apply_impl(c, ops, nops, src, w, h, 1, dst, dst_cap, band, nbands, dst_bytes, out_w, out_h, out_file_type, band_y0,
                      band_rows);
}

extern "C" int ppmx_gpu_chain_info(const ppmx_op *ops, int nops, uint32_t w, uint32_t h, uint32_t *out_w, uint32_t *out_h,
                                   int *out_file_type, size_t *out_bytes, int *splittable, int *kernels)
{
    if (!ops || nops < 1) return fail("ppmx_gpu_chain_info: bad argument");
    Pipeline P;
    if (build_pipeline(ops, nops, w, h, &P) != PPMX_OK) return PPMX_ERROR;
    if (out_w) *out_w = P.out_w;
    if (out_h) *out_h = P.out_h;
    if (out_file_type) *out_file_type = P.file_type;
    if (out_bytes) *out_bytes = P.out_bytes();
    if (splittable) *splittable = P.splittable ? 1 : 0;
    if (kernels) *kernels = (int)P.st.size();
    return PPMX_OK;
}

extern "C" int ppmx_gpu_band_rows(const ppmx_op *ops, int nops, uint32_t w, uint32_t h, int band, int nbands, uint32_t *out_y0,
                                  uint32_t *out_rows, uint32_t *src_y0, uint32_t *src_rows)
{
    if (!ops || nops < 1 || nbands < 1 || band < 0 || band >= nbands) return fail("ppmx_gpu_band_rows: bad argument");
    Pipeline P;
    if (build_pipeline(ops, nops, w, h, &P) != PPMX_OK) return PPMX_ERROR;
    if (nbands > 1 && !P.splittable) return fail("this op chain can not be cut into row bands (it holds a 90/270 degree or free rotation)");
    uint32_t a, b;
    split_rows(P.out_h, (uint32_t)nbands, (uint32_t)band, 4, &a, &b);
    Rows need{a, b};
    if (a < b)
        for (size_t i = P.st.size(); i-- > 0;)
            if (!P.st[i].passthrough) need = stage_in_rows(P.st[i], need.lo, need.hi);
    if (out_y0) *out_y0 = a;
    if (out_rows) *out_rows = b - a;
    if (src_y0) *src_y0 = a < b ? need.lo : 0;
    if (src_rows) *src_rows = a < b ? need.hi - need.lo : 0;
    return PPMX_OK;
}

// ---------------------------------------------------------------------------------------------
// a chain prepared once and run on rasters that are ALREADY in HBM (device-resident batches, benchmarks): the
// intermediate rasters and resize tables are allocated at prepare time, so a run is kernel launches only and can be
// recorded into a CUDA graph (ppmx_gpu_graph_*)
// ---------------------------------------------------------------------------------------------

struct ppmx_gpu_chain {
    ppmx_gpu_ctx *c = nullptr;
    Pipeline P;
    std::vector<DeviceTables> tables;
    std::vector<uint8_t *> mid;  // output raster of every stage but the last
};

extern "C" void ppmx_gpu_chain_free(ppmx_gpu_chain *ch)
{
    if (!ch) return;
    if (ch->c) {
        cudaSetDevice(ch->c->device);
        cudaStreamSynchronize(ch->c->lane[0]);
        for (uint8_t *p : ch->mid)
            if (p) cudaFreeAsync(p, ch->c->lane[0]);
        for (auto &t : ch->tables)
            if (t.base) cudaFreeAsync(t.base, ch->c->lane[0]);
    }
    delete ch;
}

extern "C" int ppmx_gpu_chain_prepare(ppmx_gpu_ctx *c, const ppmx_op *ops, int nops, uint32_t w, uint32_t h, ppmx_gpu_chain **out)
{
    if (!c || !ops || nops < 1 || !out) return fail("ppmx_gpu_chain_prepare: bad argument");
    *out = nullptr;
    c = primary(c);
    PPMX_CK(cudaSetDevice(c->device), "cudaSetDevice");
    ppmx_gpu_chain *ch = new (std::nothrow) ppmx_gpu_chain();
    if (!ch) return fail("out of host memory");
    ch->c = c;
    int rc = build_pipeline(ops, nops, w, h, &ch->P);
    if (rc == PPMX_OK && ch->P.hist_stage >= 0) rc = fail("a prepared chain can not hold a histogram stage");
    ch->tables.assign(nops, DeviceTables());
    ch->mid.assign(ch->P.st.size(), nullptr);
    size_t last = 0;
    for (size_t i = 0; i < ch->P.st.size(); i++)
        if (!ch->P.st[i].passthrough) last = i;
    for (size_t i = 0; i < ch->P.st.size() && rc == PPMX_OK; i++) {
        const Stage &st = ch->P.st[i];
        if (st.op.kind == PPMX_OP_IMRESIZE && st.op_index >= 0 && !ch->tables[st.op_index].base)
            rc = upload_tables(c, &st.op, &ch->tables[st.op_index], c->lane[0]);
        if (rc == PPMX_OK && i != last && !st.passthrough &&
            pool_alloc(c, (void **)&ch->mid[i], row_bytes(st.out_w, st.out_layout) * st.out_h + 16, c->lane[0]) != cudaSuccess)
            rc = fail("can not allocate image buff in HBM");
    }
    if (rc == PPMX_OK && cudaStreamSynchronize(c->lane[0]) != cudaSuccess) rc = fail("sync");
    if (rc != PPMX_OK) {
        ppmx_gpu_chain_free(ch);
        return rc;
    }
    *out = ch;
    return PPMX_OK;
}

extern "C" int ppmx_gpu_chain_info2(const ppmx_gpu_chain *ch, uint32_t *out_w, uint32_t *out_h, int *out_file_type, size_t *out_bytes,
                                    int *kernels, size_t *bytes_moved)
{
    if (!ch) return fail("ppmx_gpu_chain_info2: null chain");
    const Pipeline &P = ch->P;
    if (out_w) *out_w = P.out_w;
    if (out_h) *out_h = P.out_h;
    if (out_file_type) *out_file_type = P.file_type;
    if (out_bytes) *out_bytes = P.out_bytes();
    if (kernels) *kernels = (int)P.st.size();
    if (bytes_moved) {  // what the stages read and write, by their algorithmic raster sizes
        size_t t = 0;
        for (const Stage &st : P.st)
            t += row_bytes(st.in_w, st.in_layout) * st.in_h + (st.passthrough ? 0 : row_bytes(st.out_w, st.out_layout) * st.out_h);
        *bytes_moved = t;
    }
    return PPMX_OK;
}

extern "C" int ppmx_gpu_chain_run(ppmx_gpu_chain *ch, const void *d_src, void *d_dst, void *stream)
{
    if (!ch || !d_src || !d_dst) return fail("ppmx_gpu_chain_run: null argument");
    const Pipeline &P = ch->P;
    const uint8_t *S = (const uint8_t *)d_src;
    cudaStream_t s = (cudaStream_t)stream;
    if (P.st.empty()) {  // (a chain of "-r0" alone: the raster as it is)
        PPMX_CK(cudaMemcpyAsync(d_dst, d_src, P.out_bytes(), cudaMemcpyDeviceToDevice, s), "copy");
        return PPMX_OK;
    }
    for (size_t i = 0; i < P.st.size(); i++) {
        const Stage &st = P.st[i];
        const DeviceTables *t = (st.op.kind == PPMX_OP_IMRESIZE && st.op_index >= 0) ? &ch->tables[st.op_index] : nullptr;
        uint8_t *D = ch->mid[i] ? ch->mid[i] : (uint8_t *)d_dst;
        if (launch_stage(st, S, Rows{0, st.in_h}, D, 0, st.out_h, nullptr, t, s) != PPMX_OK) return PPMX_ERROR;
        S = D;
    }
    return PPMX_OK;
}
