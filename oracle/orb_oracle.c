/*
 * orb_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's ORB front end (bmhopkinson/hyslam), used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
 * checker and CPU baseline.  Nothing under hyslam_b200/ may call into this file.
 *
 * PARITY STATUS: *pinned to the reference's own code* for the extractor, Hamming distance and stereo
 * rows: oracle/_ref compiles hySLAM's ORBExtractor / ORBFinder / DescriptorDistance / FeatureDescriptor /
 * Stereomatcher / FeatureViews translation units UNMODIFIED (oracle/Makefile `ref`) against oracle/cvshim,
 * and tests/test_oracle_vs_ref.py requires this file to reproduce their output bit for bit (C1 / C2 frames,
 * other settings, strided views; tests/golden/ref_*.npz keep those results for boxes without the reference
 * tree).  The reference has no tests or golden vectors of its own on this path.  The
 * third-party arithmetic underneath (OpenCV: resize / GaussianBlur / FAST / fastAtan2 / gemm) is pinned
 * bit-for-bit against cv2 4.13 in tests/test_oracle_vs_cv2.py and tests/test_cvshim_vs_cv2.py.
 * The matcher rows (criteria, grid / window query, projection x4, triangulation gate, Fuse, SearchBySim3,
 * SearchForInitialization) are pinned the same way: oracle/_ref also compiles FeatureMatcher.cc, MatchCriteria.cpp,
 * Frame.cc, KeyFrame.cc, MapPoint.cc, Camera.cpp and LandMarkMatches.cpp unmodified, and
 * tests/test_oracle_match_vs_ref.py compares what a hySLAM caller observes after each call.
 * Still unpinned (no reference code can be executed): BoW quantisation -- DBoW2 and the ORB vocabulary are not in the
 * reference tree.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Compile with -ffp-contract=off: the reference is an x86-64 baseline build (no FMA).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <pthread.h>

#define ORC_API __attribute__((visibility("default")))

enum { ORC_OK = 0, ORC_EINVAL = -1, ORC_ECAPACITY = -2, ORC_EUNSUPPORTED = -3, ORC_ENOMEM = -4 };

#define ORC_MAX_LEVELS 16
#define EDGE_THRESHOLD 19      /* ORBExtractor.cpp:74 */
#define PATCH_SIZE 31          /* ORBExtractor.cpp:73 */
#define HALF_PATCH_SIZE 15     /* ORBFinder.cpp:14 */
#define FAST_THRESHOLD 20      /* ORBFinder.h:92 + setter bug ORBFinder.cpp:58-60: always 20 */

/* cv::KeyPoint layout (28 bytes): pt.x pt.y size angle response octave class_id */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orc_keypoint;

typedef struct {
    int32_t nfeatures;
    float scale_factor;   /* FeatureExtractorSettings::fScaleFactor is float (FeatureExtractorSettings.h:22) */
    int32_t nlevels;
    int32_t cell_px;      /* N_CELLS: really the cell edge in px (ORBExtractor.cpp:409) */
    int32_t ini_th;       /* dead in the reference (ORBFinder.cpp:58-60) */
    int32_t min_th;       /* dead in the reference */
} orc_params;

static const int8_t brief_pattern[1024] = {
#include "../include/hyorb_brief_pattern.inc"
};

/* cvRound: round-half-to-even (SSE cvtss2si / lrint under the default rounding mode). */
static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_round_d(double v) { return (int)lrint(v); }

/* ------------------------------------------------------------------------------------------ */
/* a1: scale tables and per-level quotas -- ORBExtractor.cpp:76-119                           */
/* ------------------------------------------------------------------------------------------ */
ORC_API int orc_scale_tables(const orc_params *p, float *scale, float *inv_scale, float *sigma2,
                             float *inv_sigma2, int32_t *quota)
{
    if (!p || p->nlevels < 1 || p->nlevels > ORC_MAX_LEVELS) return ORC_EINVAL;
    const double scaleFactor = (double)p->scale_factor;  /* ORBExtractor.h: double scaleFactor = float setting */
    const int n = p->nlevels;
    scale[0] = 1.0f; sigma2[0] = 1.0f;
    for (int i = 1; i < n; i++) {
        scale[i] = (float)((double)scale[i - 1] * scaleFactor);       /* :92 float*double -> double -> float */
        sigma2[i] = scale[i] * scale[i];                                /* :93 */
    }
    for (int i = 0; i < n; i++) {
        inv_scale[i] = 1.0f / scale[i];                                 /* :100 */
        inv_sigma2[i] = 1.0f / sigma2[i];                               /* :101 */
    }
    float factor = (float)(1.0 / scaleFactor);                          /* :107 1.0f/double -> double -> float */
    /* :108  nfeatures*(1 - factor) is int*float -> float; pow in double narrowed to float */
    float nDesired = (float)p->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)n));
    int sum = 0;
    for (int l = 0; l < n - 1; l++) {
        quota[l] = cv_round_f(nDesired);                                /* :113 */
        sum += quota[l];
        nDesired *= factor;                                             /* :115 */
    }
    quota[n - 1] = p->nfeatures - sum > 0 ? p->nfeatures - sum : 0;     /* :117 */
    return ORC_OK;
}

/* level size -- ORBExtractor.cpp:568-569 */
ORC_API void orc_level_size(int W, int H, float inv_scale, int32_t *w, int32_t *h)
{
    *w = cv_round_f((float)W * inv_scale);
    *h = cv_round_f((float)H * inv_scale);
}

/* ------------------------------------------------------------------------------------------ */
/* a2: cv::resize(INTER_LINEAR), 8UC1 -- called at ORBExtractor.cpp:577.                       */
/* OpenCV algorithm (imgproc/resize.cpp: resizeGeneric_ + HResizeLinear + VResizeLinear<uchar, */
/* int, short>): Q11 coefficients, (b*(h>>4))>>16 vertical combine, (+2)>>2.                   */
/* ------------------------------------------------------------------------------------------ */
/* horizontal: out-of-range taps reset the fraction (resize.cpp xofs/alpha loop); vertical: the
 * fraction is kept and the two source rows are clipped instead (yofs/beta loop + clip()). */
static void linear_coeffs(int dst_n, int src_n, int *ofs, int16_t *c0, int16_t *c1, int vertical)
{
    const double inv_scale = (double)dst_n / src_n;
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < dst_n; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (!vertical) {
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= src_n - 1) { s = src_n - 1; f = 0.f; }
        }
        ofs[d] = s;
        c0[d] = (int16_t)cv_round_f((1.f - f) * 2048.f);
        c1[d] = (int16_t)cv_round_f(f * 2048.f);
    }
}

ORC_API int orc_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride,
                                 uint8_t *dst, int dw, int dh, int dstride)
{
    if (sw < 1 || sh < 1 || dw < 1 || dh < 1) return ORC_EINVAL;
    if (sw == 2 * dw && sh == 2 * dh) {
        /* cv::resize silently switches INTER_LINEAR to the INTER_AREA 2x2 fast path */
        for (int y = 0; y < dh; y++)
            for (int x = 0; x < dw; x++) {
                const uint8_t *p = src + (size_t)(2 * y) * sstride + 2 * x;
                dst[(size_t)y * dstride + x] = (uint8_t)((p[0] + p[1] + p[sstride] + p[sstride + 1] + 2) >> 2);
            }
        return ORC_OK;
    }
    int *xo = (int *)malloc(sizeof(int) * (dw + dh));
    int16_t *c = (int16_t *)malloc(sizeof(int16_t) * 2 * (dw + dh));
    int32_t *rows = (int32_t *)malloc(sizeof(int32_t) * 2 * dw);
    if (!xo || !c || !rows) { free(xo); free(c); free(rows); return ORC_ENOMEM; }
    int *yo = xo + dw;
    int16_t *xa0 = c, *xa1 = c + dw, *yb0 = c + 2 * dw, *yb1 = c + 2 * dw + dh;
    linear_coeffs(dw, sw, xo, xa0, xa1, 0);
    linear_coeffs(dh, sh, yo, yb0, yb1, 1);
    for (int y = 0; y < dh; y++) {
        const int sy0 = yo[y] < 0 ? 0 : (yo[y] > sh - 1 ? sh - 1 : yo[y]);
        const int sy1 = yo[y] + 1 < 0 ? 0 : (yo[y] + 1 > sh - 1 ? sh - 1 : yo[y] + 1);
        const uint8_t *r0 = src + (size_t)sy0 * sstride, *r1 = src + (size_t)sy1 * sstride;
        int32_t *h0 = rows, *h1 = rows + dw;
        for (int x = 0; x < dw; x++) {
            const int s0 = xo[x];
            const int s1 = s0 + 1 < sw ? s0 + 1 : sw - 1;
            h0[x] = r0[s0] * xa0[x] + r0[s1] * xa1[x];
            h1[x] = r1[s0] * xa0[x] + r1[s1] * xa1[x];
        }
        const int b0 = yb0[y], b1 = yb1[y];
        for (int x = 0; x < dw; x++) {
            int v = (((b0 * (h0[x] >> 4)) >> 16) + ((b1 * (h1[x] >> 4)) >> 16) + 2) >> 2;
            dst[(size_t)y * dstride + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    free(xo); free(c); free(rows);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* a6: cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101), 8UC1 -- ORBExtractor.cpp:536-537.   */
/* OpenCV fixed-point path: separable integer kernel [18,34,48,56,48,34,18]/256; horizontal    */
/* pass exact Q8, vertical pass Q16, out = (v + 32768) >> 16.                                  */
/* ------------------------------------------------------------------------------------------ */
static const int gk7[7] = {18, 34, 48, 56, 48, 34, 18};

static inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

ORC_API int orc_gaussian7_u8(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride)
{
    if (w < 1 || h < 1) return ORC_EINVAL;
    uint16_t *tmp = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)w * h);
    if (!tmp) return ORC_ENOMEM;
    for (int y = 0; y < h; y++) {
        const uint8_t *r = src + (size_t)y * sstride;
        for (int x = 0; x < w; x++) {
            int acc = 0;
            if (x >= 3 && x < w - 3) for (int k = 0; k < 7; k++) acc += gk7[k] * r[x + k - 3];
            else for (int k = 0; k < 7; k++) acc += gk7[k] * r[reflect101(x + k - 3, w)];
            tmp[(size_t)y * w + x] = (uint16_t)acc;   /* <= 255*256 */
        }
    }
    for (int y = 0; y < h; y++) {
        const uint16_t *rr[7];
        for (int k = 0; k < 7; k++) rr[k] = tmp + (size_t)reflect101(y + k - 3, h) * w;
        for (int x = 0; x < w; x++) {
            uint32_t acc = 0;
            for (int k = 0; k < 7; k++) acc += (uint32_t)gk7[k] * rr[k][x];
            dst[(size_t)y * dstride + x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
    free(tmp);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* a4: cv::FAST(img, kps, 20, true) = FAST-9/16 with 3x3 NMS on an ROI -- ORBFinder.cpp:66-68. */
/* OpenCV algorithm (features2d/fast.cpp FAST_t<16>, fast_score.cpp cornerScore<16>).          */
/* Output order: row-major.  kp = (x, y, response) relative to the ROI origin.                 */
/* ------------------------------------------------------------------------------------------ */
static const int ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

static int fast_corner_score(const uint8_t *p, int stride, int threshold)
{
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 25; k++) d[k] = v - p[ring_dy[k & 15] * stride + ring_dx[k & 15]];
    int best = threshold;   /* a0 in cornerScore<16>; result = max over arcs - 1 */
    for (int k = 0; k < 16; k++) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; j++) {
            if (d[k + j] < mn) mn = d[k + j];
            if (d[k + j] > mx) mx = d[k + j];
        }
        if (mn > best) best = mn;          /* arc of 9 pixels all darker than centre by > best  */
        if (-mx > best) best = -mx;        /* arc of 9 pixels all brighter */
    }
    return best - 1;
}

static int fast_is_corner(const uint8_t *p, int stride, int threshold)
{
    const int v = p[0];
    unsigned dark = 0, bright = 0;
    for (int k = 0; k < 16; k++) {
        const int x = p[ring_dy[k] * stride + ring_dx[k]];
        if (x < v - threshold) dark |= 1u << k;
        if (x > v + threshold) bright |= 1u << k;
    }
    for (int pass = 0; pass < 2; pass++) {
        unsigned m = pass ? bright : dark;
        m |= m << 16;
        int run = 0;
        for (int k = 0; k < 25; k++) {
            if (m & (1u << k)) { if (++run > 8) return 1; }
            else run = 0;
        }
    }
    return 0;
}

/* returns number of keypoints written (or ORC_ECAPACITY) */
static int fast9_roi(const uint8_t *img, int stride, int w, int h, int threshold, int nms,
                     float *ox, float *oy, float *oresp, int cap)
{
    if (w < 7 || h < 7) return 0;
    uint8_t *score = (uint8_t *)calloc((size_t)w * h, 1);
    if (!score) return ORC_ENOMEM;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            const uint8_t *p = img + (size_t)y * stride + x;
            if (fast_is_corner(p, stride, threshold))
                score[(size_t)y * w + x] = (uint8_t)fast_corner_score(p, stride, threshold);
        }
    int n = 0;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            const int s = score[(size_t)y * w + x];
            if (!s) continue;
            if (nms) {
                const uint8_t *q = score + (size_t)y * w + x;
                if (!(s > q[-1] && s > q[1] && s > q[-w - 1] && s > q[-w] && s > q[-w + 1] &&
                      s > q[w - 1] && s > q[w] && s > q[w + 1]))
                    continue;
            }
            if (n >= cap) { free(score); return ORC_ECAPACITY; }
            ox[n] = (float)x; oy[n] = (float)y; oresp[n] = (float)s;
            n++;
        }
    free(score);
    return n;
}

ORC_API int orc_fast9(const uint8_t *img, int w, int h, int stride, int threshold, int nms,
                      float *ox, float *oy, float *oresp, int cap)
{
    return fast9_roi(img, stride, w, h, threshold, nms, ox, oy, oresp, cap);
}

/* ------------------------------------------------------------------------------------------ */
/* a3: per-cell detection over the lattice -- ORBExtractor.cpp:405-470.                        */
/* Output: candidates in lattice coordinates (absolute - 16), reference order (cell row-major, */
/* FAST row-major inside the cell).                                                            */
/* ------------------------------------------------------------------------------------------ */
ORC_API int orc_detect_level(const uint8_t *img, int cols, int rows, int stride, int cell_px,
                             float *ox, float *oy, float *oresp, int cap)
{
    const int minBX = EDGE_THRESHOLD - 3, minBY = minBX;
    const int maxBX = cols - EDGE_THRESHOLD + 3, maxBY = rows - EDGE_THRESHOLD + 3;
    const float W = (float)cell_px;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols < 1 || nRows < 1) return ORC_EUNSUPPORTED;   /* the reference divides by zero here */
    const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
    int n = 0;
    for (int i = 0; i < nRows; i++) {
        const int iniY = minBY + i * hCell;
        int maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;                    /* :435 */
        if (maxY > maxBY) maxY = maxBY;
        for (int j = 0; j < nCols; j++) {
            const int iniX = minBX + j * wCell;
            int maxX = iniX + wCell + 6;
            if (iniX >= maxBX - 6) continue;                /* :444 */
            if (maxX > maxBX) maxX = maxBX;
            /* :450-457: the "retry with minThFAST" is an identical no-op (threshold stuck at 20) */
            int m = fast9_roi(img + (size_t)iniY * stride + iniX, stride, maxX - iniX, maxY - iniY,
                              FAST_THRESHOLD, 1, ox + n, oy + n, oresp + n, cap - n);
            if (m < 0) return m;
            for (int k = 0; k < m; k++) { ox[n + k] += (float)(j * wCell); oy[n + k] += (float)(i * hCell); }
            n += m;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* a5: DistributeOctTree + DivideNode -- ORBExtractor.cpp:179-403, 121-177.                    */
/* Literal restatement with an index-linked list.  The reference sorts (count, node address);  */
/* CANONICAL TIE POLICY: address order == creation order, i.e. a monotonic     */
/* never-reusing allocator; `seq` below is the node's creation sequence number.                */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int ulx, uly, urx, bry;      /* UL.x, UL.y, UR.x, BR.y */
    int start, count;            /* slice of the per-pass index arena */
    int prev, next;              /* list links */
    int nomore;
} qnode;

typedef struct {
    qnode *nodes; int nnodes, capnodes;
    int head, tail, size;
    int *arena; size_t arena_used, arena_cap;
} qtree;

static int qt_new_node(qtree *t)
{
    if (t->nnodes == t->capnodes) {
        int nc = t->capnodes * 2;
        qnode *nn = (qnode *)realloc(t->nodes, sizeof(qnode) * nc);
        if (!nn) return -1;
        t->nodes = nn; t->capnodes = nc;
    }
    qnode *q = &t->nodes[t->nnodes];
    memset(q, 0, sizeof(*q));
    q->prev = q->next = -1;
    return t->nnodes++;
}
static long qt_alloc(qtree *t, int n)   /* returns arena offset or -1; arena may move */
{
    if (t->arena_used + (size_t)n > t->arena_cap) {
        size_t nc = t->arena_cap * 2 + (size_t)n;
        int *na = (int *)realloc(t->arena, sizeof(int) * nc);
        if (!na) return -1;
        t->arena = na; t->arena_cap = nc;
    }
    const long off = (long)t->arena_used;
    t->arena_used += (size_t)n;
    return off;
}
static void qt_push_front(qtree *t, int id)
{
    qnode *q = &t->nodes[id];
    q->prev = -1; q->next = t->head;
    if (t->head >= 0) t->nodes[t->head].prev = id; else t->tail = id;
    t->head = id; t->size++;
}
static void qt_push_back(qtree *t, int id)
{
    qnode *q = &t->nodes[id];
    q->next = -1; q->prev = t->tail;
    if (t->tail >= 0) t->nodes[t->tail].next = id; else t->head = id;
    t->tail = id; t->size++;
}
static int qt_erase(qtree *t, int id)   /* returns next */
{
    qnode *q = &t->nodes[id];
    const int nx = q->next;
    if (q->prev >= 0) t->nodes[q->prev].next = q->next; else t->head = q->next;
    if (q->next >= 0) t->nodes[q->next].prev = q->prev; else t->tail = q->prev;
    t->size--;
    return nx;
}

/* DivideNode (:121-177): creates up to 4 children (ids in ch[0..3], -1 if empty), does not link them */
static int qt_divide(qtree *t, int id, const float *px, const float *py, int ch[4])
{
    const qnode P = t->nodes[id];
    const int halfX = (int)ceilf((float)(P.urx - P.ulx) / 2);
    const int halfY = (int)ceilf((float)(P.bry - P.uly) / 2);
    const int midx = P.ulx + halfX, midy = P.uly + halfY;
    const long base = qt_alloc(t, 4 * P.count);
    if (base < 0) return -1;
    int cnt[4] = {0, 0, 0, 0};
    for (int i = 0; i < P.count; i++) {
        const int k = t->arena[P.start + i];
        int q;
        if (px[k] < (float)midx) q = (py[k] < (float)midy) ? 0 : 2;
        else q = (py[k] < (float)midy) ? 1 : 3;
        t->arena[base + (long)q * P.count + cnt[q]++] = k;
    }
    for (int q = 0; q < 4; q++) {
        ch[q] = -1;
        if (!cnt[q]) continue;
        const int c = qt_new_node(t);
        if (c < 0) return -1;
        qnode *C = &t->nodes[c];
        C->ulx = (q & 1) ? midx : P.ulx;
        C->urx = (q & 1) ? P.urx : midx;
        C->uly = (q & 2) ? midy : P.uly;
        C->bry = (q & 2) ? P.bry : midy;
        C->start = (int)(base + (long)q * P.count);
        C->count = cnt[q];
        C->nomore = (cnt[q] == 1);
        ch[q] = c;
    }
    return 0;
}

typedef struct { int size; int seq; } size_seq;
static int cmp_size_seq(const void *a, const void *b)
{
    const size_seq *x = (const size_seq *)a, *y = (const size_seq *)b;
    if (x->size != y->size) return x->size < y->size ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq ? 1 : 0);
}

/* Returns number of output indices (into the candidate arrays) written to out_idx, in list order. */
ORC_API int orc_distribute_octtree(const float *px, const float *py, const float *presp, int n,
                                   int minX, int maxX, int minY, int maxY, int N,
                                   int32_t *out_idx, int cap)
{
    if (n == 0) return 0;
    const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));   /* :183 */
    if (nIni < 1) return ORC_EUNSUPPORTED;                                 /* reference: UB (empty vpIniNodes) */
    const float hX = (float)(maxX - minX) / nIni;                          /* :185 */

    qtree t;
    t.capnodes = 64 + 8 * (N > 0 ? N : 1) + nIni; t.nnodes = 0;
    t.nodes = (qnode *)malloc(sizeof(qnode) * t.capnodes);
    t.head = t.tail = -1; t.size = 0;
    t.arena_cap = (size_t)n * 16 + 1024;   /* grows on demand */
    t.arena = (int *)malloc(sizeof(int) * t.arena_cap); t.arena_used = 0;
    size_seq *vcur = (size_seq *)malloc(sizeof(size_seq) * ((size_t)n + 16));
    size_seq *vprev = (size_seq *)malloc(sizeof(size_seq) * ((size_t)n + 16));
    if (!t.nodes || !t.arena || !vcur || !vprev) { free(t.nodes); free(t.arena); free(vcur); free(vprev); return ORC_ENOMEM; }
    int rc = 0;

    /* roots (:192-210) */
    int *rootcnt = (int *)calloc(nIni, sizeof(int));
    int *rootof = (int *)malloc(sizeof(int) * n);
    for (int i = 0; i < n; i++) {
        int r = (int)(px[i] / hX);
        if (r < 0 || r >= nIni) { rc = ORC_EUNSUPPORTED; goto done_roots; }
        rootof[i] = r; rootcnt[r]++;
    }
    for (int r = 0; r < nIni; r++) {
        const int id = qt_new_node(&t);
        qnode *q = &t.nodes[id];
        q->ulx = (int)(hX * (float)r); q->urx = (int)(hX * (float)(r + 1));
        q->uly = 0; q->bry = maxY - minY;
        q->start = (int)t.arena_used; q->count = 0;
        qt_alloc(&t, rootcnt[r]);
        qt_push_back(&t, id);
    }
    for (int i = 0; i < n; i++) {
        qnode *q = &t.nodes[rootof[i]];
        t.arena[q->start + q->count++] = i;
    }
done_roots:
    free(rootcnt); free(rootof);
    if (rc) goto done;

    /* :212-225 */
    for (int it = t.head; it >= 0;) {
        qnode *q = &t.nodes[it];
        if (q->count == 1) { q->nomore = 1; it = q->next; }
        else if (q->count == 0) it = qt_erase(&t, it);
        else it = q->next;
    }

    int finish = 0, ncur = 0;
    while (!finish) {
        int prevSize = t.size;
        int nToExpand = 0;
        ncur = 0;
        for (int it = t.head; it >= 0;) {
            if (t.nodes[it].nomore) { it = t.nodes[it].next; continue; }
            int ch[4];
            if (qt_divide(&t, it, px, py, ch)) { rc = ORC_ENOMEM; goto done; }
            for (int q = 0; q < 4; q++) {
                if (ch[q] < 0) continue;
                qt_push_front(&t, ch[q]);
                if (t.nodes[ch[q]].count > 1) {
                    nToExpand++;
                    vcur[ncur].size = t.nodes[ch[q]].count; vcur[ncur].seq = ch[q]; ncur++;
                }
            }
            it = qt_erase(&t, it);
        }
        if (t.size >= N || t.size == prevSize) finish = 1;                 /* :309 */
        else if (t.size + nToExpand * 3 > N) {                             /* :313 */
            while (!finish) {
                prevSize = t.size;
                size_seq *tmp = vprev; vprev = vcur; vcur = tmp;
                const int nprev = ncur; ncur = 0;
                qsort(vprev, nprev, sizeof(size_seq), cmp_size_seq);       /* :324 (count, address==seq) */
                for (int j = nprev - 1; j >= 0; j--) {
                    int ch[4];
                    const int id = vprev[j].seq;
                    if (qt_divide(&t, id, px, py, ch)) { rc = ORC_ENOMEM; goto done; }
                    for (int q = 0; q < 4; q++) {
                        if (ch[q] < 0) continue;
                        qt_push_front(&t, ch[q]);
                        if (t.nodes[ch[q]].count > 1) {
                            vcur[ncur].size = t.nodes[ch[q]].count; vcur[ncur].seq = ch[q]; ncur++;
                        }
                    }
                    qt_erase(&t, id);
                    if (t.size >= N) break;                                /* :370 */
                }
                if (t.size >= N || t.size == prevSize) finish = 1;         /* :374 */
            }
        }
    }

    /* :381-400 best response per node, first wins ties, list order */
    {
        int m = 0;
        for (int it = t.head; it >= 0; it = t.nodes[it].next) {
            const qnode *q = &t.nodes[it];
            const int *idx = t.arena + q->start;
            int best = idx[0];
            float br = presp[best];
            for (int k = 1; k < q->count; k++)
                if (presp[idx[k]] > br) { best = idx[k]; br = presp[best]; }
            if (m >= cap) { rc = ORC_ECAPACITY; goto done; }
            out_idx[m++] = best;
        }
        rc = m;
    }
done:
    free(t.nodes); free(t.arena); free(vcur); free(vprev);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* cv::fastAtan2 (scalar path of core/mathfuncs_core: atan_f32) -- called at ORBFinder.cpp:42  */
/* ------------------------------------------------------------------------------------------ */
ORC_API float orc_fast_atan2(float y, float x)
{
    const float k = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k;
    const float p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* umax -- ORBFinder.cpp:131-149 */
static void compute_umax(int *umax)
{
    int v, v0;
    const int vmax = (int)floor(HALF_PATCH_SIZE * sqrt(2.f) / 2 + 1);
    const int vmin = (int)ceil(HALF_PATCH_SIZE * sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}
ORC_API void orc_umax(int32_t *out) { int u[16]; compute_umax(u); for (int i = 0; i < 16; i++) out[i] = u[i]; }

/* a7: intensityCentroidAngle -- ORBFinder.cpp:16-43 (on the BLURRED level, ORBExtractor.cpp:536-541) */
ORC_API float orc_ic_angle(const uint8_t *img, int stride, float ptx, float pty)
{
    int umax[16];
    compute_umax(umax);
    int m_01 = 0, m_10 = 0;
    const uint8_t *center = img + (size_t)cv_round_f(pty) * stride + cv_round_f(ptx);
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int val_plus = center[u + v * stride], val_minus = center[u - v * stride];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return orc_fast_atan2((float)m_01, (float)m_10);
}

/* a8: computeOrbDescriptor -- ORBFinder.cpp:89-129.  cos/sin: the float argument promotes to
 * the double libm overload (global ::cos, no `using namespace std` in ORBFinder.cpp), result
 * narrowed to float. */
ORC_API void orc_brief(const uint8_t *img, int stride, float ptx, float pty, float angle_deg, uint8_t *desc)
{
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float angle = angle_deg * factorPI;
    const float a = (float)cos((double)angle), b = (float)sin((double)angle);
    const uint8_t *center = img + (size_t)cv_round_f(pty) * stride + cv_round_f(ptx);
    const int8_t *pat = brief_pattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            const float x0 = (float)pat[4 * k], y0 = (float)pat[4 * k + 1];
            const float x1 = (float)pat[4 * k + 2], y1 = (float)pat[4 * k + 3];
            const int t0 = center[cv_round_f(x0 * b + y0 * a) * stride + cv_round_f(x0 * a - y0 * b)];
            const int t1 = center[cv_round_f(x1 * b + y1 * a) * stride + cv_round_f(x1 * a - y1 * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* ORBExtractor::operator() -- ORBExtractor.cpp:496-562 (+ ComputePyramid :564-589,            */
/* ComputeKeyPointsOctTree :405-494).  Optional debug outputs for stage-by-stage parity.       */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    /* all optional (NULL = skip) */
    uint8_t *pyramid;           /* concatenated levels, tightly packed (w*h each)          */
    uint8_t *blurred;           /* concatenated blurred levels (only levels with keypoints) */
    float *cand_x, *cand_y, *cand_resp; int32_t *cand_count; int32_t cand_cap; /* per level capacity, lattice coords */
    int32_t *level_count;       /* keypoints per level after distribution */
} orc_debug;

ORC_API int orc_extract(const orc_params *p, const uint8_t *image, int W, int H, int stride,
                        orc_keypoint *kps, uint8_t *desc, int cap, int32_t *n_out, orc_debug *dbg)
{
    float scale[ORC_MAX_LEVELS], inv[ORC_MAX_LEVELS], s2[ORC_MAX_LEVELS], is2[ORC_MAX_LEVELS];
    int32_t quota[ORC_MAX_LEVELS];
    *n_out = 0;
    if (!image || W <= 0 || H <= 0) return ORC_OK;                        /* :499-500 empty -> silent return */
    int rc = orc_scale_tables(p, scale, inv, s2, is2, quota);
    if (rc) return rc;
    const int nl = p->nlevels;
    int32_t lw[ORC_MAX_LEVELS], lh[ORC_MAX_LEVELS];
    size_t off[ORC_MAX_LEVELS + 1]; off[0] = 0;
    for (int l = 0; l < nl; l++) {
        orc_level_size(W, H, inv[l], &lw[l], &lh[l]);
        if (lw[l] < 1 || lh[l] < 1) return ORC_EUNSUPPORTED;
        off[l + 1] = off[l] + (size_t)lw[l] * lh[l];
    }
    uint8_t *pyr = (uint8_t *)malloc(off[nl]);
    uint8_t *blur = (uint8_t *)malloc((size_t)lw[0] * lh[0]);
    const int ccap = lw[0] * lh[0] / 2 + 1024;
    float *cx = (float *)malloc(sizeof(float) * 3 * (size_t)ccap);
    int32_t *sel = (int32_t *)malloc(sizeof(int32_t) * (size_t)ccap);
    if (!pyr || !blur || !cx || !sel) { rc = ORC_ENOMEM; goto done; }
    float *cy = cx + ccap, *cr = cy + ccap;

    /* ComputePyramid :564-589 (the 19-px border is written but never read downstream) */
    for (int y = 0; y < H; y++) memcpy(pyr + (size_t)y * W, image + (size_t)y * stride, W);
    for (int l = 1; l < nl; l++) {
        rc = orc_resize_linear_u8(pyr + off[l - 1], lw[l - 1], lh[l - 1], lw[l - 1], pyr + off[l], lw[l], lh[l], lw[l]);
        if (rc) goto done;
    }
    if (dbg && dbg->pyramid) memcpy(dbg->pyramid, pyr, off[nl]);

    int total = 0;
    for (int l = 0; l < nl; l++) {
        const uint8_t *img = pyr + off[l];
        /* ComputeKeyPointsOctTree :405-494 */
        int nc = orc_detect_level(img, lw[l], lh[l], lw[l], p->cell_px, cx, cy, cr, ccap);
        if (nc < 0) { rc = nc; goto done; }
        if (dbg && dbg->cand_count) {
            dbg->cand_count[l] = nc;
            if (dbg->cand_x) {
                if (nc > dbg->cand_cap) { rc = ORC_ECAPACITY; goto done; }
                memcpy(dbg->cand_x + (size_t)l * dbg->cand_cap, cx, sizeof(float) * nc);
                memcpy(dbg->cand_y + (size_t)l * dbg->cand_cap, cy, sizeof(float) * nc);
                memcpy(dbg->cand_resp + (size_t)l * dbg->cand_cap, cr, sizeof(float) * nc);
            }
        }
        const int minB = EDGE_THRESHOLD - 3;
        int nk = orc_distribute_octtree(cx, cy, cr, nc, minB, lw[l] - EDGE_THRESHOLD + 3, minB,
                                        lh[l] - EDGE_THRESHOLD + 3, quota[l], sel, ccap);
        if (nk < 0) { rc = nk; goto done; }
        if (dbg && dbg->level_count) dbg->level_count[l] = nk;
        if (nk == 0) continue;                                             /* :532-533 */
        if (total + nk > cap) { rc = ORC_ECAPACITY; goto done; }
        const int scaledPatchSize = (int)(PATCH_SIZE * scale[l]);          /* :478 */
        /* :536-537 blur a borderless clone of the level */
        rc = orc_gaussian7_u8(img, lw[l], lh[l], lw[l], blur, lw[l]);
        if (rc) goto done;
        if (dbg && dbg->blurred) memcpy(dbg->blurred + off[l], blur, (size_t)lw[l] * lh[l]);
        for (int k = 0; k < nk; k++) {
            orc_keypoint *kp = &kps[total + k];
            kp->x = cx[sel[k]] + (float)minB;                              /* :484-485 */
            kp->y = cy[sel[k]] + (float)minB;
            kp->response = cr[sel[k]];
            kp->octave = l; kp->class_id = -1;
            kp->size = (float)scaledPatchSize;
            kp->angle = orc_ic_angle(blur, lw[l], kp->x, kp->y);           /* ORBFinder.cpp:73 */
            orc_brief(blur, lw[l], kp->x, kp->y, kp->angle, desc + (size_t)(total + k) * 32);
            if (l != 0) { kp->x *= scale[l]; kp->y *= scale[l]; }          /* :546-552 */
        }
        total += nk;
    }
    *n_out = total;
    rc = ORC_OK;
done:
    free(pyr); free(blur); free(cx); free(sel);
    return rc;
}

/* CPU-baseline helper: extract a batch of same-sized images, pthread workers pulling images
 * from a shared counter (the reference itself runs one extractor per thread,
 * ImageProcessing.cpp:82-84). */
typedef struct {
    const orc_params *p; const uint8_t *images; int n_images, W, H;
    orc_keypoint *kps; uint8_t *desc; int cap; int32_t *counts;
    int next; int err; pthread_mutex_t mu;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        const int i = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (i >= j->n_images) break;
        int rc = orc_extract(j->p, j->images + (size_t)i * j->W * j->H, j->W, j->H, j->W,
                             j->kps + (size_t)i * j->cap, j->desc + (size_t)i * j->cap * 32, j->cap,
                             &j->counts[i], NULL);
        if (rc) { pthread_mutex_lock(&j->mu); j->err = rc; pthread_mutex_unlock(&j->mu); }
    }
    return NULL;
}

ORC_API int orc_extract_batch(const orc_params *p, const uint8_t *images, int n_images, int W, int H,
                              orc_keypoint *kps, uint8_t *desc, int cap_per_image, int32_t *counts, int nthreads)
{
    batch_job j = {p, images, n_images, W, H, kps, desc, cap_per_image, counts, 0, 0, PTHREAD_MUTEX_INITIALIZER};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nthreads - 1; t++) if (pthread_create(&th[started], NULL, batch_worker, &j) == 0) started++;
    batch_worker(&j);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    return j.err;
}

/* ------------------------------------------------------------------------------------------ */
/* a10: ORBDistance::distance -- DescriptorDistance.cpp:9-25 (SWAR popcount over 8 x int32)    */
/* ------------------------------------------------------------------------------------------ */
ORC_API int orc_hamming(const uint8_t *a, const uint8_t *b)
{
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t x, y;
        memcpy(&x, a + 4 * i, 4); memcpy(&y, b + 4 * i, 4);
        uint32_t v = x ^ y;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

/* ------------------------------------------------------------------------------------------ */
/* a14/a15/a16 best / second-best scan + acceptance -- MatchCriteria.cpp:248-280 (scan),       */
/* :214-246 (mode 0, landmark/window rule), :601-635 (mode 1, BoW rule), :486-523 (mode 2,     */
/* mono-init rule).  Candidates as CSR lists of target indices; one list per query.            */
/* Outputs: best_idx (-1 if no candidate), best/second distance (FLT_MAX encoded as 65535),    */
/* accepted flag.                                                                              */
/* ------------------------------------------------------------------------------------------ */
static int accept_rule(int mode, float best, float second, float thr, float ratio)
{
    switch (mode) {
    case 0: return (best <= thr) && !(best > ratio * second);
    case 1: return (best < thr) && (best < ratio * second);
    case 2: return (best <= thr) && (best < second * ratio);
    default: return 0;
    }
}

ORC_API int orc_match_csr(const uint8_t *qdesc, int nq, const uint8_t *tdesc, int nt,
                          const int32_t *cand_off, const int32_t *cand_idx,   /* NULL,NULL = all targets in order */
                          int mode, float thr, float ratio,
                          int32_t *best_idx, uint16_t *best, uint16_t *second, uint8_t *accepted)
{
    for (int q = 0; q < nq; q++) {
        float bd = FLT_MAX, bd2 = FLT_MAX; int bi = -1;
        const int lo = cand_off ? cand_off[q] : 0, hi = cand_off ? cand_off[q + 1] : nt;
        for (int c = lo; c < hi; c++) {
            const int t = cand_idx ? cand_idx[c] : c;
            if (t < 0 || t >= nt) return ORC_EINVAL;
            const float d = (float)orc_hamming(qdesc + (size_t)q * 32, tdesc + (size_t)t * 32);
            if (d < bd) { bd2 = bd; bd = d; bi = t; }
            else if (d < bd2) bd2 = d;
        }
        best_idx[q] = bi;
        best[q] = bd == FLT_MAX ? 65535 : (uint16_t)bd;
        second[q] = bd2 == FLT_MAX ? 65535 : (uint16_t)bd2;
        accepted[q] = (uint8_t)(bi >= 0 && accept_rule(mode, bd, bd2, thr, ratio));
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* a19: EpipolarConsistencyBoWCriterion::CheckDistEpipolarLine -- src/features/MatchCriteria.cpp:659-676 */
/* with sigma2 = FeatureExtractorSettings::determineSigma2(kp2.size) (src/core/FeatureExtractorSettings.cpp:5-8). */
/* Plain float expressions evaluated left to right (this file is built with -ffp-contract=off); the   */
/* final comparison promotes to double because 3.84 is a double literal.  ok[c] for candidate pairs    */
/* (i1[c], i2[c]).                                                                                    */
/* ------------------------------------------------------------------------------------------ */
ORC_API int orc_epipolar_check(const orc_keypoint *kps1, const orc_keypoint *kps2, const int32_t *i1, const int32_t *i2, int n, const float *F12,
                               float sigma_ref, float size_ref, uint8_t *ok)
{
    for (int c = 0; c < n; c++) {
        const orc_keypoint *k1 = &kps1[i1[c]], *k2 = &kps2[i2[c]];
        const float sf = k2->size / size_ref;
        const float sigma2 = sigma_ref * (sf * sf);
        const float a = k1->x * F12[0] + k1->y * F12[3] + F12[6];
        const float b = k1->x * F12[1] + k1->y * F12[4] + F12[7];
        const float cc = k1->x * F12[2] + k1->y * F12[5] + F12[8];
        const float num = a * k2->x + b * k2->y + cc;
        const float den = a * a + b * b;
        if (den == 0) { ok[c] = 0; continue; }
        const float dsqr = num * num / den;
        ok[c] = (uint8_t)(dsqr < 3.84 * sigma2);
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* A.0: ImageProcessing::PreProcessImg -- src/main/ImageProcessing.cpp:118-138                     */
/* cv::resize(img, img, Size(), fscale, fscale) then cvtColor(RGB|BGR[A] -> GRAY).  Restated for    */
/* the two scales the reference's configurations use (pinned against cv2 4.13 by        */
/* tests/test_oracle_vs_cv2.py): 1.0 = copy; 0.5 = OpenCV's INTER_AREA fast path, the per-channel   */
/* 2x2 box mean (a+b+c+d+2)>>2 over dst = cvRound(src*0.5) pixels (needs 2*dst <= src).  Gray =      */
/* (R*9798 + G*19235 + B*3735 + 16384) >> 15 (OpenCV's 15-bit fixed-point BT.601 weights); a 4th     */
/* channel is ignored; 1-channel input passes through the colour step.                              */
/* ------------------------------------------------------------------------------------------ */
ORC_API int orc_preprocess_size(int w, int h, int half_scale, int *ow, int *oh)
{
    *ow = half_scale ? cv_round_d((double)w * 0.5) : w;
    *oh = half_scale ? cv_round_d((double)h * 0.5) : h;
    if (*ow < 1 || *oh < 1) return ORC_EINVAL;
    if (half_scale && (2 * *ow > w || 2 * *oh > h)) return ORC_EINVAL;
    return ORC_OK;
}

ORC_API int orc_preprocess(const uint8_t *src, int w, int h, int stride, int channels, int rgb_order, int half_scale, uint8_t *dst, int dst_stride)
{
    int ow, oh;
    if (channels != 1 && channels != 3 && channels != 4) return ORC_EINVAL;
    if (orc_preprocess_size(w, h, half_scale, &ow, &oh) != ORC_OK) return ORC_EINVAL;
    for (int y = 0; y < oh; y++)
        for (int x = 0; x < ow; x++) {
            int c[3] = {0, 0, 0};
            const int nc = channels == 1 ? 1 : 3;
            for (int k = 0; k < nc; k++) {
                if (half_scale) {
                    const uint8_t *p = src + (size_t)(2 * y) * stride + (size_t)(2 * x) * channels + k;
                    c[k] = (p[0] + p[channels] + p[stride] + p[stride + channels] + 2) >> 2;
                } else c[k] = src[(size_t)y * stride + (size_t)x * channels + k];
            }
            int g;
            if (channels == 1) g = c[0];
            else {
                const int r = rgb_order ? c[0] : c[2], b = rgb_order ? c[2] : c[0];
                g = (r * 9798 + c[1] * 19235 + b * 3735 + 16384) >> 15;
            }
            dst[(size_t)y * dst_stride + x] = (uint8_t)g;
        }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* f4: MapPointDBEntry::_computeDistinctiveDescriptor_ -- src/core/MapPointDB.cpp:127-171          */
/* All-pairs Hamming distances of one landmark's observation descriptors (row i, self distance 0), */
/* sort each row, median = sorted[(int)(0.5*(N-1))], keep the FIRST row with the smallest median   */
/* (strict <, :163).  Batched over landmarks through CSR offsets into one descriptor array.        */
/* ------------------------------------------------------------------------------------------ */
static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }
ORC_API int orc_distinctive_descriptor(const uint8_t *desc, const int32_t *lm_off, int n_landmarks, int32_t *best_idx, int32_t *best_median)
{
    for (int l = 0; l < n_landmarks; l++) {
        const int lo = lm_off[l], N = lm_off[l + 1] - lo;
        best_idx[l] = -1; best_median[l] = -1;
        if (N <= 0) continue;                               /* :141-143: nothing to do */
        int *row = (int *)malloc(sizeof(int) * (size_t)N);
        if (!row) return ORC_ENOMEM;
        float BestMedian = FLT_MAX; int BestIdx = 0;
        for (int i = 0; i < N; i++) {
            for (int j = 0; j < N; j++) row[j] = i == j ? 0 : orc_hamming(desc + (size_t)(lo + i) * 32, desc + (size_t)(lo + j) * 32);
            qsort(row, N, sizeof(int), cmp_int);
            const int median = row[(int)(0.5 * (N - 1))];
            if ((float)median < BestMedian) { BestMedian = (float)median; BestIdx = i; }
        }
        free(row);
        best_idx[l] = BestIdx; best_median[l] = (int32_t)BestMedian;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* a11: Stereomatcher::computeStereoMatches -- Stereomatcher.cpp:36-156                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    float mbf, fx;        /* Camera::mbf, Camera::fx() */
    int32_t n_rows;       /* (int)camera.mnMaxY */
    float th_high, th_low;/* FeatureMatcherSettings defaults 100 / 50 (ImageProcessing.cpp:101) */
    float size_ref;       /* FeatureExtractorSettings::size_ref = 31 */
} orc_stereo_params;

typedef struct { float d; int i; } dist_idx;
static int cmp_dist_idx(const void *a, const void *b)
{
    const dist_idx *x = (const dist_idx *)a, *y = (const dist_idx *)b;
    if (x->d != y->d) return x->d < y->d ? -1 : 1;
    return x->i < y->i ? -1 : (x->i > y->i ? 1 : 0);
}

ORC_API int orc_stereo_match(const orc_stereo_params *sp,
                             const orc_keypoint *kl, const uint8_t *dl, int nl,
                             const orc_keypoint *kr, const uint8_t *dr, int nr,
                             float *uR, float *depth, int32_t *best_r /* optional */, int32_t *best_d /* optional */)
{
    for (int i = 0; i < nl; i++) { uR[i] = -1.f; depth[i] = -1.f; if (best_r) best_r[i] = -1; if (best_d) best_d[i] = -1; }
    const float dist_threshold = (sp->th_high + sp->th_low) / 2;           /* :41 */
    const int nRows = sp->n_rows;
    const float mb = sp->mbf / sp->fx;                                      /* :23 */
    if (nRows <= 0) return ORC_EINVAL;
    /* row table :46-65 (CSR built in two passes; insertion order = iR ascending) */
    int *cnt = (int *)calloc((size_t)nRows + 1, sizeof(int));
    if (!cnt) return ORC_ENOMEM;
    for (int iR = 0; iR < nr; iR++) {
        const float r = 2.0f * kr[iR].size / sp->size_ref;
        const int maxr = (int)ceilf(kr[iR].y + r), minr = (int)floorf(kr[iR].y - r);
        for (int y = minr; y <= maxr; y++) {
            if (y < 0 || y >= nRows) { free(cnt); return ORC_EINVAL; }      /* reference: out-of-bounds write */
            cnt[y + 1]++;
        }
    }
    for (int y = 0; y < nRows; y++) cnt[y + 1] += cnt[y];
    int *tab = (int *)malloc(sizeof(int) * (size_t)(cnt[nRows] + 1));
    int *fill = (int *)calloc((size_t)nRows, sizeof(int));
    dist_idx *v = (dist_idx *)malloc(sizeof(dist_idx) * (size_t)(nl + 1));
    if (!tab || !fill || !v) { free(cnt); free(tab); free(fill); free(v); return ORC_ENOMEM; }
    for (int iR = 0; iR < nr; iR++) {
        const float r = 2.0f * kr[iR].size / sp->size_ref;
        const int maxr = (int)ceilf(kr[iR].y + r), minr = (int)floorf(kr[iR].y - r);
        for (int y = minr; y <= maxr; y++) tab[cnt[y] + fill[y]++] = iR;
    }
    const float minZ = mb, minD = 0, maxD = sp->mbf / minZ;                 /* :68-70 */
    int nv = 0, rc = ORC_OK;
    for (int iL = 0; iL < nl; iL++) {
        const float vL = kl[iL].y, uL = kl[iL].x;
        const int levelL = kl[iL].octave;
        const long row = (long)vL;                                           /* vRowIndices[vL]: float -> size_t */
        if (row < 0 || row >= nRows) { rc = ORC_EINVAL; break; }
        const int lo = cnt[row], hi = cnt[row + 1];
        if (lo == hi) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        float bestDist = sp->th_high; int bestIdxR = 0;
        for (int c = lo; c < hi; c++) {
            const int iR = tab[c];
            if (kr[iR].octave < levelL - 1 || kr[iR].octave > levelL + 1) continue;
            const float u = kr[iR].x;
            if (u >= minU && u <= maxU) {
                const float dist = (float)orc_hamming(dl + (size_t)iL * 32, dr + (size_t)iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < dist_threshold) {                                     /* :118 */
            float uR0 = kr[bestIdxR].x;
            float disparity = uL - uR0;
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01f; uR0 = (float)((double)uL - 0.01); }   /* :125-129 (0.01 double literals) */
                depth[iL] = sp->mbf / disparity;
                uR[iL] = uR0;
                if (best_r) best_r[iL] = bestIdxR;
                if (best_d) best_d[iL] = (int)bestDist;
                v[nv].d = bestDist; v[nv].i = iL; nv++;
            }
        }
    }
    if (rc == ORC_OK && nv > 0) {                                            /* empty: reference reads v[0] (UB) -> all -1 */
        qsort(v, nv, sizeof(dist_idx), cmp_dist_idx);                        /* :137 */
        const float median = v[nv / 2].d;
        const float thDist = 1.5f * 1.4f * median;
        for (int i = nv - 1; i >= 0; i--) {
            if (v[i].d < thDist) break;
            uR[v[i].i] = -1; depth[v[i].i] = -1;
            if (best_r) best_r[v[i].i] = -1;
            if (best_d) best_d[v[i].i] = -1;
        }
    }
    free(cnt); free(tab); free(fill); free(v);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* a12/a13: Frame grid -- Frame.cc:137-153 (assign), :459-469 (PosInGrid, round()),            */
/* :416-457 (GetFeaturesInAreaNEW).  Grid 64 x 48 (Frame.h:69-70), cell-major (ix*48+iy) CSR.  */
/* ------------------------------------------------------------------------------------------ */
#define GRID_COLS 64
#define GRID_ROWS 48
typedef struct { float min_x, max_x, min_y, max_y; } orc_bounds;

ORC_API int orc_grid_build(const orc_keypoint *kps, int n, const orc_bounds *b,
                           int32_t *cell_off /* 3073 */, int32_t *cell_idx /* n */)
{
    const float invW = (float)GRID_COLS / (b->max_x - b->min_x);            /* Frame.cc:65 */
    const float invH = (float)GRID_ROWS / (b->max_y - b->min_y);
    memset(cell_off, 0, sizeof(int32_t) * (GRID_COLS * GRID_ROWS + 1));
    int *cell = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    if (!cell) return ORC_ENOMEM;
    for (int i = 0; i < n; i++) {
        const int px = (int)roundf((kps[i].x - b->min_x) * invW);
        const int py = (int)roundf((kps[i].y - b->min_y) * invH);
        if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) { cell[i] = -1; continue; }
        cell[i] = px * GRID_ROWS + py;
        cell_off[cell[i] + 1]++;
    }
    for (int c = 0; c < GRID_COLS * GRID_ROWS; c++) cell_off[c + 1] += cell_off[c];
    int *fill = (int *)calloc(GRID_COLS * GRID_ROWS, sizeof(int));
    if (!fill) { free(cell); return ORC_ENOMEM; }
    for (int i = 0; i < n; i++) if (cell[i] >= 0) cell_idx[cell_off[cell[i]] + fill[cell[i]]++] = i;
    free(cell); free(fill);
    return ORC_OK;
}

/* window query used by the matcher below; returns count, writes candidate indices in the
 * reference's enumeration order (ix outer, iy inner, insertion order) */
static int grid_query(const orc_keypoint *kps, const orc_bounds *b, const int32_t *cell_off,
                      const int32_t *cell_idx, float x, float y, float r, int32_t *out, int cap)
{
    const float invW = (float)GRID_COLS / (b->max_x - b->min_x);
    const float invH = (float)GRID_ROWS / (b->max_y - b->min_y);
    int x0 = (int)floorf((x - b->min_x - r) * invW); if (x0 < 0) x0 = 0;
    if (x0 >= GRID_COLS) return 0;
    int x1 = (int)ceilf((x - b->min_x + r) * invW); if (x1 > GRID_COLS - 1) x1 = GRID_COLS - 1;
    if (x1 < 0) return 0;
    int y0 = (int)floorf((y - b->min_y - r) * invH); if (y0 < 0) y0 = 0;
    if (y0 >= GRID_ROWS) return 0;
    int y1 = (int)ceilf((y - b->min_y + r) * invH); if (y1 > GRID_ROWS - 1) y1 = GRID_ROWS - 1;
    if (y1 < 0) return 0;
    int n = 0;
    for (int ix = x0; ix <= x1; ix++)
        for (int iy = y0; iy <= y1; iy++) {
            const int c = ix * GRID_ROWS + iy;
            for (int j = cell_off[c]; j < cell_off[c + 1]; j++) {
                const int k = cell_idx[j];
                const float dx = kps[k].x - x, dy = kps[k].y - y;
                if (fabsf(dx) < r && fabsf(dy) < r) { if (n >= cap) return ORC_ECAPACITY; out[n++] = k; }
            }
        }
    return n;
}

ORC_API int orc_grid_query(const orc_keypoint *kps, const orc_bounds *b, const int32_t *cell_off,
                           const int32_t *cell_idx, float x, float y, float r, int32_t *out, int cap)
{
    return grid_query(kps, b, cell_off, cell_idx, x, y, r, out, cap);
}

/* a20 inner loop: window query (a13) -> PreviouslyMatched (a17, MatchCriteria.cpp:124-144) ->
 * FeatureSize (a17, :350-360) -> StereoConsistency (a17, :149-177; skipped when ur_radius < 0,
 * i.e. mono camera) -> BestScore (a14).  One query per landmark. */
typedef struct {
    float u, v, r;            /* projected position + search radius (FeatureMatcher.cc:88-92) */
    float size_lo, size_hi;   /* frac_smaller*projected_size, frac_larger*projected_size      */
    float ur, ur_radius;      /* projected right coordinate + radius; ur_radius < 0: skip     */
} orc_window_query;

ORC_API int orc_match_window(const orc_keypoint *kps, const uint8_t *tdesc, const float *t_uR,
                             const uint8_t *t_matched, int nt, const orc_bounds *b,
                             const int32_t *cell_off, const int32_t *cell_idx,
                             const orc_window_query *q, const uint8_t *qdesc, int nq,
                             float thr, float ratio,
                             int32_t *best_idx, uint16_t *best, uint16_t *second, uint8_t *accepted)
{
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nt + 1));
    if (!cand) return ORC_ENOMEM;
    for (int i = 0; i < nq; i++) {
        int n = grid_query(kps, b, cell_off, cell_idx, q[i].u, q[i].v, q[i].r, cand, nt);
        float bd = FLT_MAX, bd2 = FLT_MAX; int bi = -1;
        for (int c = 0; c < n; c++) {
            const int k = cand[c];
            if (t_matched && t_matched[k]) continue;
            if (!(kps[k].size > q[i].size_lo && kps[k].size < q[i].size_hi)) continue;
            if (q[i].ur_radius >= 0) {
                const float er = fabsf(q[i].ur - t_uR[k]);
                if (!(er < q[i].ur_radius && t_uR[k] > 0)) continue;
            }
            const float d = (float)orc_hamming(qdesc + (size_t)i * 32, tdesc + (size_t)k * 32);
            if (d < bd) { bd2 = bd; bd = d; bi = k; }
            else if (d < bd2) bd2 = d;
        }
        best_idx[i] = bi;
        best[i] = bd == FLT_MAX ? 65535 : (uint16_t)bd;
        second[i] = bd2 == FLT_MAX ? 65535 : (uint16_t)bd2;
        accepted[i] = (uint8_t)(bi >= 0 && accept_rule(0, bd, bd2, thr, ratio));
    }
    free(cand);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* f2: the per-landmark front half of FeatureMatcher::SearchByProjection(Frame&, landmarks, th)   */
/* (FeatureMatcher.cc:123-143 -> _SearchByProjection_ :57-121): ProjectionCriterion and           */
/* DistanceCriterion (MatchCriteria.cpp:13-28, 46-77), Frame::ProjectLandMark (Frame.cc:176-180), */
/* Camera::Project (Camera.cpp:116-153), Frame::landMarkSizePixels (Frame.cc:296-317) and the      */
/* window / size / stereo bounds the view criteria use (FeatureMatcher.cc:88-92,                   */
/* MatchCriteria.cpp:149-177, 350-360).  fp32 policy of the cv::Mat expressions (pinned against    */
/* cv2.gemm / cv2.norm by tests/test_oracle_vs_cv2.py): a 3x3 * 3x1 product is OpenCV's small-      */
/* matrix path, t = (a0*b0 + a1*b1) + a2*b2 in float, result = (float)((double)t + (double)c);      */
/* cv::norm accumulates squares in double.                                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    float Rcw[9], tcw[3];      /* Frame::mRcw, mtcw (row-major) */
    float Ow[3];               /* Frame::GetCameraCenter() */
    float K[9];                /* Camera::K */
    float mbf;                 /* Camera::mbf */
    int32_t stereo;            /* Camera::sensor == 1 */
    orc_bounds bounds;         /* Camera::mnMinX .. mnMaxY */
} orc_projection;
typedef struct {
    float Pw[3];               /* MapPoint::GetWorldPos() */
    float size;                /* MapPoint::getSize() (world units) */
    float min_dist, max_dist;  /* GetMinDistanceInvariance / GetMaxDistanceInvariance */
    int32_t assoc_idx;         /* Frame::hasAssociation(lm): keypoint already associated in this frame, or -1 */
} orc_landmark;

static void gemm31(const float *A, const float *b, const float *c, float *out)
{
    for (int i = 0; i < 3; i++) {
        const float p0 = A[3 * i] * b[0], p1 = A[3 * i + 1] * b[1], p2 = A[3 * i + 2] * b[2];
        const float t = (p0 + p1) + p2;
        out[i] = c ? (float)((double)t + (double)c[i]) : t;
    }
}
/* Frame::ProjectLandMark + Camera::Project: uv = (u, v, ur); returns the validity flag */
static int project_point(const orc_projection *pr, const float *Pw, float *uv)
{
    float Pc[3], Pch[3];
    gemm31(pr->Rcw, Pw, pr->tcw, Pc);
    const float PcZ = Pc[2];
    const float invz = 1.0f / PcZ;
    Pch[0] = Pc[0] / PcZ; Pch[1] = Pc[1] / PcZ; Pch[2] = Pc[2] / PcZ;
    gemm31(pr->K, Pch, NULL, uv);
    const float u = uv[0], v = uv[1];
    uv[2] = pr->stereo ? u - pr->mbf * invz : -1.0f;
    return PcZ > 0.0f && u >= pr->bounds.min_x && u <= pr->bounds.max_x && v >= pr->bounds.min_y && v <= pr->bounds.max_y;
}

ORC_API int orc_project_landmarks(const orc_projection *pr, const orc_landmark *lms, int n, const orc_keypoint *t_kps, int nt,
                                  float th, float size_ref, float frac_smaller, float frac_larger,
                                  orc_window_query *queries, uint8_t *passed)
{
    for (int i = 0; i < n; i++) {
        const orc_landmark *lm = &lms[i];
        float uv[3];
        const int valid = project_point(pr, lm->Pw, uv);                      /* ProjectionCriterion */
        const float PO[3] = {lm->Pw[0] - pr->Ow[0], lm->Pw[1] - pr->Ow[1], lm->Pw[2] - pr->Ow[2]};
        const float dist = (float)sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);   /* cv::norm */
        const int dist_ok = !(dist < lm->min_dist || dist > lm->max_dist);     /* DistanceCriterionCore */
        float size_px;
        if (lm->assoc_idx >= 0) {
            if (lm->assoc_idx >= nt) return ORC_EINVAL;
            size_px = t_kps[lm->assoc_idx].size;                               /* Frame.cc:298-300 */
        } else {
            float L[3] = {lm->Pw[0] - lm->size / 2, lm->Pw[1], lm->Pw[2]}, R[3] = {lm->Pw[0] + lm->size / 2, lm->Pw[1], lm->Pw[2]};
            float ul[3], ur[3];
            project_point(pr, L, ul); project_point(pr, R, ur);
            size_px = ur[0] - ul[0];
        }
        const float radius = th * size_px / size_ref;
        queries[i].u = uv[0]; queries[i].v = uv[1]; queries[i].r = radius;
        queries[i].size_lo = frac_smaller * size_px; queries[i].size_hi = frac_larger * size_px;
        queries[i].ur = uv[2]; queries[i].ur_radius = pr->stereo ? radius : -1.0f;
        passed[i] = (uint8_t)(valid && dist_ok);
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* f2 (motion model): RotationConsistencyCriterion::apply over the matches of _SearchByProjection_ */
/* (MatchCriteria.cpp:363-401 -> RotationConsistency :684-725).  The reference first re-keys the    */
/* matches by current keypoint index (`current_matches_idx[idx_curr] = idx_prev`, :382) while         */
/* walking a std::map ordered by MapPoint*: when several landmarks matched the same keypoint only    */
/* the LAST one in that order survives, whatever the histogram says.  Canonical policy here, as for  */
/* the quadtree: pointer order == the order the caller lists the landmarks in.                       */
/* accepted[] is updated in place; prev_angle[i] = angle of the keypoint landmark i is associated     */
/* with in the previous frame, curr_angle[k] = angle of current keypoint k.                           */
/* ------------------------------------------------------------------------------------------ */
ORC_API int orc_projection_rotation(const int32_t *best_idx, uint8_t *accepted, int n, const float *prev_angle,
                                    const orc_keypoint *t_kps, int nt)
{
    int *owner = (int *)malloc(sizeof(int) * (size_t)(nt > 0 ? nt : 1));
    if (!owner) return ORC_ENOMEM;
    for (int k = 0; k < nt; k++) owner[k] = -1;
    for (int i = 0; i < n; i++)
        if (accepted[i]) { if (best_idx[i] < 0 || best_idx[i] >= nt) { free(owner); return ORC_EINVAL; } owner[best_idx[i]] = i; }
    int hist[30] = {0};
    const float factor = 1.0f / 30;
    for (int i = 0; i < n; i++) {
        if (!accepted[i] || owner[best_idx[i]] != i) continue;
        float rot = prev_angle[i] - t_kps[best_idx[i]].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)roundf(rot * factor);
        if (bin == 30) bin = 0;
        if (bin < 0 || bin >= 30) { free(owner); return ORC_EINVAL; }
        hist[bin]++;
    }
    int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
    for (int i = 0; i < 30; i++) {
        const int sz = hist[i];
        if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; i3 = i2; i2 = i1; i1 = i; }
        else if (sz > max2) { max3 = max2; max2 = sz; i3 = i2; i2 = i; }
        else if (sz > max3) { max3 = sz; i3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { i2 = -1; i3 = -1; }
    else if (max3 < 0.1f * (float)max1) { i3 = -1; }
    for (int i = 0; i < n; i++) {
        if (!accepted[i]) continue;
        if (owner[best_idx[i]] != i) { accepted[i] = 0; continue; }
        float rot = prev_angle[i] - t_kps[best_idx[i]].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)roundf(rot * factor);
        if (bin == 30) bin = 0;
        accepted[i] = (uint8_t)(bin == i1 || bin == i2 || bin == i3);
    }
    free(owner);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* f3: bag-of-words quantisation.  hySLAM calls DBoW2 (ORBVocabulary::transform,                  */
/* src/features/low_level/ORBVocabulary.cpp:31-42; Frame::ComputeBoW Frame.cc:472-479, levelsup 4);  */
/* DBoW2 itself is an un-vendored, unversioned dependency (/usr/local/lib/libDBoW2.so,            */
/* src/CMakeLists.txt:32) and its vocabulary file is absent, so this restates the PUBLISHED        */
/* algorithm of DBoW2::TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)     */
/* (Galvez-Lopez & Tardos, DBoW2 TemplatedVocabulary.h): descend from the root, at every level     */
/* take the child with the smallest descriptor distance (first child wins ties: strict <),         */
/* remember the node reached at level L - levelsup, stop at a leaf.  PARITY UNPINNED by the        */
/* reference for this function (no DBoW2 source, no vocabulary, no fixtures).                      */
/* Tree: node 0 = root, children of node i = child_idx[child_off[i] .. child_off[i+1]).            */
/* ------------------------------------------------------------------------------------------ */
ORC_API int orc_bow_transform(int n_nodes, int L, const int32_t *child_off, const int32_t *child_idx, const uint8_t *node_desc,
                              const int32_t *word_of, const float *weight_of, const uint8_t *desc, int n, int levelsup,
                              int32_t *word_id, int32_t *node_id, float *weight)
{
    const int nid_level = L - levelsup;
    for (int i = 0; i < n; i++) {
        int node = 0, level = 0, nid = -1;
        if (nid_level <= 0) nid = 0;
        while (child_off[node + 1] > child_off[node]) {
            ++level;
            const int lo = child_off[node], hi = child_off[node + 1];
            int best = child_idx[lo];
            if (best < 0 || best >= n_nodes) return ORC_EINVAL;
            int best_d = orc_hamming(desc + (size_t)i * 32, node_desc + (size_t)best * 32);
            for (int c = lo + 1; c < hi; c++) {
                const int id = child_idx[c];
                if (id < 0 || id >= n_nodes) return ORC_EINVAL;
                const int d = orc_hamming(desc + (size_t)i * 32, node_desc + (size_t)id * 32);
                if (d < best_d) { best_d = d; best = id; }
            }
            node = best;
            if (level == nid_level) nid = node;
            if (level > 64) return ORC_EINVAL;      /* a cycle in the child lists */
        }
        word_id[i] = word_of[node];
        weight[i] = weight_of[node];
        node_id[i] = nid >= 0 ? nid : node;         /* leaf above level L - levelsup: DBoW2 leaves nid unset; we report the leaf */
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* a18: RotationConsistency + ComputeThreeMaxima -- MatchCriteria.cpp:684-767.                 */
/* pairs must be given in ascending idx_curr order (std::map iteration order).                 */
/* ------------------------------------------------------------------------------------------ */
ORC_API int orc_rotation_consistency(const float *angle_prev, const float *angle_curr, int n, uint8_t *keep)
{
    enum { L = 30 };
    int hist[L]; memset(hist, 0, sizeof(hist));
    int *bin = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    if (!bin) return ORC_ENOMEM;
    const float factor = 1.0f / L;
    for (int i = 0; i < n; i++) {
        float rot = angle_prev[i] - angle_curr[i];
        if (rot < 0.0) rot += 360.0f;
        int bi = (int)roundf(rot * factor);
        if (bi == L) bi = 0;
        if (bi < 0 || bi >= L) { free(bin); return ORC_EINVAL; }
        bin[i] = bi; hist[bi]++;
    }
    int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
    for (int i = 0; i < n; i++) keep[i] = (uint8_t)(bin[i] == ind1 || bin[i] == ind2 || bin[i] == ind3);
    free(bin);
    return ORC_OK;
}

/* ========================================================================================== */
/* Round 2: the remaining matcher entry points of FeatureMatcher.h:105-176.                    */
/* Every function below is pinned against the reference's own code (oracle/_ref) by            */
/* tests/test_oracle_match_vs_ref.py.                                                          */
/* ========================================================================================== */

/* ViewingAngleCriterionCore (MatchCriteria.cpp:94-110), used by FeatureMatcher::Fuse (:469):   */
/* PO = Pw - Ow; d = (float)cv::norm(PO); PO = PO / d  (cv::Mat / scalar = convertTo with        */
/* alpha = 1/d: multiplication by (float)(1.0 / d) in float); pass iff PO.dot(Pn) > cos(angle),  */
/* dot accumulated in double, cos(float) = cosf (the <math.h> float overload).                   */
ORC_API int orc_viewing_angle(const float *Ow, const float *Pw, const float *normal, int n, float max_angle, uint8_t *pass)
{
    const float c = cosf(max_angle);
    for (int i = 0; i < n; i++) {
        const float PO[3] = {Pw[3 * i] - Ow[0], Pw[3 * i + 1] - Ow[1], Pw[3 * i + 2] - Ow[2]};
        const float dist = (float)sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);
        const float inv = (float)(1.0 / (double)dist);
        double dot = 0.0;
        for (int k = 0; k < 3; k++) dot += (double)(PO[k] * inv) * (double)normal[3 * i + k];
        pass[i] = (uint8_t)(dot > c);
    }
    return ORC_OK;
}

/* Window scan with the full set of view criteria of FeatureMatcher.cc: as orc_match_window plus         */
/*  * ProjectionViewCriterion (MatchCriteria.cpp:282-333 over KeyFrame::ReprojectionError,                  */
/*    KeyFrame.cc:548-575) when reproj_thr >= 0: err = (u-x)^2 + (v-y)^2 + (uR_view >= 0 ? (ur-uR_view)^2 : 0), */
/*    keep iff err / determineSigma2(kp.size) < (uR_view > 0 ? 1.3f : 1.0f) * reproj_thr;  views.uR(idx) = -1   */
/*    for monocular views (t_uR == NULL);                                                                      */
/*  * the acceptance rule as a parameter (0 landmark / 1 BoW / 2 mono-init).                                   */
ORC_API int orc_match_window_ex(const orc_keypoint *kps, const uint8_t *tdesc, const float *t_uR, const uint8_t *t_matched, int nt,
                                const orc_bounds *b, const int32_t *cell_off, const int32_t *cell_idx, const orc_window_query *q,
                                const uint8_t *qdesc, const uint8_t *q_active, int nq, float thr, float ratio, int rule,
                                float reproj_thr, float sigma_ref, float size_ref,
                                int32_t *best_idx, uint16_t *best, uint16_t *second, uint8_t *accepted)
{
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nt + 1));
    if (!cand) return ORC_ENOMEM;
    for (int i = 0; i < nq; i++) {
        best_idx[i] = -1; best[i] = 65535; second[i] = 65535; accepted[i] = 0;
        if (q_active && !q_active[i]) continue;
        int n = grid_query(kps, b, cell_off, cell_idx, q[i].u, q[i].v, q[i].r, cand, nt);
        float bd = FLT_MAX, bd2 = FLT_MAX; int bi = -1;
        for (int c = 0; c < n; c++) {
            const int k = cand[c];
            if (t_matched && t_matched[k]) continue;
            if (!(kps[k].size > q[i].size_lo && kps[k].size < q[i].size_hi)) continue;
            if (q[i].ur_radius >= 0) {
                const float er = fabsf(q[i].ur - t_uR[k]);
                if (!(er < q[i].ur_radius && t_uR[k] > 0)) continue;
            }
            if (reproj_thr >= 0) {
                const float ur_view = t_uR ? t_uR[k] : -1.0f;
                const float errX = q[i].u - kps[k].x, errY = q[i].v - kps[k].y;
                float errXr = 0.0f;
                if (ur_view >= 0.0f) errXr = q[i].ur - ur_view;
                const float sserr = errX * errX + errY * errY + errXr * errXr;
                const float sf = kps[k].size / size_ref;
                const float sigma2 = sigma_ref * (sf * sf);                   /* FeatureExtractorSettings.cpp:5-8 */
                const float stereo_factor = ur_view > 0 ? 1.30f : 1.00f;
                if (!((sserr / sigma2) < stereo_factor * reproj_thr)) continue;
            }
            const float d = (float)orc_hamming(qdesc + (size_t)i * 32, tdesc + (size_t)k * 32);
            if (d < bd) { bd2 = bd; bd = d; bi = k; }
            else if (d < bd2) bd2 = d;
        }
        best_idx[i] = bi;
        best[i] = bd == FLT_MAX ? 65535 : (uint16_t)bd;
        second[i] = bd2 == FLT_MAX ? 65535 : (uint16_t)bd2;
        accepted[i] = (uint8_t)(bi >= 0 && accept_rule(rule, bd, bd2, thr, ratio));
    }
    free(cand);
    return ORC_OK;
}

/* One direction of FeatureMatcher::SearchBySim3 (FeatureMatcher.cc:783-845 / 848-910): landmark of keyframe A ->      */
/* camera A (R_a * Pw + t_a) -> camera B through the similarity (sR_ba * p + t_ba) -> Camera::Project of B; the 3-D    */
/* distance in B must lie in the landmark's scale-invariance range; search radius = th * B.landMarkSizePixels(lm) /     */
/* size_ref, where landMarkSizePixels projects with B's OWN pose (pr_b), not the similarity.  No view criteria:        */
/* the window query carries open size bounds and no stereo radius.                                                     */
ORC_API int orc_project_sim3(const float *R_a, const float *t_a, const float *sR_ba, const float *t_ba, const orc_projection *pr_b,
                             const orc_landmark *lms, int n, const orc_keypoint *kps_b, int nb, float th, float size_ref,
                             orc_window_query *queries, uint8_t *passed)
{
    for (int i = 0; i < n; i++) {
        const orc_landmark *lm = &lms[i];
        float pa[3], pb[3], Pch[3], uv[3];
        gemm31(R_a, lm->Pw, t_a, pa);
        gemm31(sR_ba, pa, t_ba, pb);
        const float z = pb[2], invz = 1.0f / z;
        Pch[0] = pb[0] / z; Pch[1] = pb[1] / z; Pch[2] = pb[2] / z;
        gemm31(pr_b->K, Pch, NULL, uv);
        uv[2] = pr_b->stereo ? uv[0] - pr_b->mbf * invz : -1.0f;
        const int valid = z > 0.0f && uv[0] >= pr_b->bounds.min_x && uv[0] <= pr_b->bounds.max_x && uv[1] >= pr_b->bounds.min_y && uv[1] <= pr_b->bounds.max_y;
        const float dist = (float)sqrt((double)pb[0] * pb[0] + (double)pb[1] * pb[1] + (double)pb[2] * pb[2]);
        const int dist_ok = !(dist < lm->min_dist || dist > lm->max_dist);
        float size_px;
        if (lm->assoc_idx >= 0) {
            if (lm->assoc_idx >= nb) return ORC_EINVAL;
            size_px = kps_b[lm->assoc_idx].size;
        } else {
            float L[3] = {lm->Pw[0] - lm->size / 2, lm->Pw[1], lm->Pw[2]}, R[3] = {lm->Pw[0] + lm->size / 2, lm->Pw[1], lm->Pw[2]};
            float ul[3], ur[3];
            project_point(pr_b, L, ul); project_point(pr_b, R, ur);
            size_px = ur[0] - ul[0];
        }
        queries[i].u = uv[0]; queries[i].v = uv[1]; queries[i].r = th * size_px / size_ref;
        queries[i].size_lo = -FLT_MAX; queries[i].size_hi = FLT_MAX;
        queries[i].ur = uv[2]; queries[i].ur_radius = -1.0f;
        passed[i] = (uint8_t)(valid && dist_ok);
    }
    return ORC_OK;
}

/* FeatureMatcher::SearchForInitialization (FeatureMatcher.cc:404-462) with MonoInitScoreExceedsPrevious and MonoInitBestScore       */
/* (MatchCriteria.cpp:486-549).  Sequential by construction: feature i1 of frame 1 may only take a frame-2 feature away from an       */
/* earlier i1' if its distance is smaller than the stored one -- compared after TRUNCATION to int (`int dist = d1.distance(d2)`, :539) */
/* against the float distance of the previous match; accepted matches overwrite (matches[idx2] = i1).  Then RotationConsistency        */
/* (rot = angle(frame 1) - angle(frame 2), keyed by idx2), and the inverse map.  prev_xy (n1 x 2) is updated like vbPrevMatched.       */
ORC_API int orc_search_for_initialization(const orc_keypoint *k1, const uint8_t *d1, int n1, const orc_keypoint *k2, const uint8_t *d2, int n2,
                                          const orc_bounds *b2, const int32_t *cell_off2, const int32_t *cell_idx2, float *prev_xy, int window,
                                          float thr, float ratio, int32_t *matches12)
{
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n2 + 1));
    int32_t *owner = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n2 + 1));      /* matches: idx2 -> i1 */
    float *dprev = (float *)malloc(sizeof(float) * (size_t)(n2 + 1));            /* MonoCriteriaData::distances */
    uint8_t *keep = (uint8_t *)malloc((size_t)n2 + 1);
    if (!cand || !owner || !dprev || !keep) { free(cand); free(owner); free(dprev); free(keep); return ORC_ENOMEM; }
    for (int j = 0; j < n2; j++) { owner[j] = -1; dprev[j] = -1.0f; }
    for (int i1 = 0; i1 < n1; i1++) {
        const int n = grid_query(k2, b2, cell_off2, cell_idx2, prev_xy[2 * i1], prev_xy[2 * i1 + 1], (float)window, cand, n2);
        if (n <= 0) continue;
        float bd = FLT_MAX, bd2 = FLT_MAX; int bi = -1;
        for (int c = 0; c < n; c++) {
            const int i2 = cand[c];
            const int dist = orc_hamming(d1 + (size_t)i1 * 32, d2 + (size_t)i2 * 32);
            if (!(dprev[i2] < 0) && !((float)dist < dprev[i2])) continue;          /* MonoInitScoreExceedsPrevious */
            const float d = (float)dist;
            if (d < bd) { bd2 = bd; bd = d; bi = i2; }
            else if (d < bd2) bd2 = d;
        }
        if (bi >= 0 && accept_rule(2, bd, bd2, thr, ratio)) { owner[bi] = i1; dprev[bi] = bd; }
    }
    /* RotationConsistency(matches, views2, views1): idx_curr = idx2, idx_prev = i1 (MatchCriteria.cpp:684-725) */
    int hist[30]; memset(hist, 0, sizeof(hist));
    const float factor = 1.0f / 30;
    for (int j = 0; j < n2; j++) {
        if (owner[j] < 0) continue;
        float rot = k1[owner[j]].angle - k2[j].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)roundf(rot * factor);
        if (bin == 30) bin = 0;
        hist[bin]++;
    }
    int max1 = 0, max2 = 0, max3 = 0, i1m = -1, i2m = -1, i3m = -1;
    for (int i = 0; i < 30; i++) {
        const int sz = hist[i];
        if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; i3m = i2m; i2m = i1m; i1m = i; }
        else if (sz > max2) { max3 = max2; max2 = sz; i3m = i2m; i2m = i; }
        else if (sz > max3) { max3 = sz; i3m = i; }
    }
    if (max2 < 0.1f * (float)max1) { i2m = -1; i3m = -1; }
    else if (max3 < 0.1f * (float)max1) { i3m = -1; }
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    int nm = 0;
    for (int j = 0; j < n2; j++) {
        if (owner[j] < 0) continue;
        float rot = k1[owner[j]].angle - k2[j].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)roundf(rot * factor);
        if (bin == 30) bin = 0;
        keep[j] = (uint8_t)(bin == i1m || bin == i2m || bin == i3m);
        if (keep[j]) nm++;
    }
    /* matches_inverse[i1] = idx2: ascending idx2, later entries overwrite (:447-450) -- an i1 can own several idx2 */
    for (int j = 0; j < n2; j++)
        if (owner[j] >= 0 && keep[j]) matches12[owner[j]] = j;
    for (int i = 0; i < n1; i++)
        if (matches12[i] >= 0) { prev_xy[2 * i] = k2[matches12[i]].x; prev_xy[2 * i + 1] = k2[matches12[i]].y; }
    free(cand); free(owner); free(dprev); free(keep);
    return nm;
}
