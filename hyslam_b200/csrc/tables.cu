// tables.cu -- host-side, data-independent tables of the extractor: scale pyramid constants, per-level
// feature quotas, the FAST cell lattice, cv::resize coefficient tables and the quadtree path LUTs.
// Everything here depends only on (params, width, height) and is computed once per image shape.
#include <math.h>
#include <string.h>
#include <stdarg.h>

#include "common.cuh"

namespace hyorb {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *last_error() { return g_err; }

static inline int round_half_even(float v) { return (int)lrintf(v); }   // cvRound (SSE cvtss2si)

// ORBExtractor::ORBExtractor, src/features/ORBExtractor.cpp:76-119.  The reference keeps scaleFactor as a
// double initialised from the float setting and mixes float/double arithmetic; the casts below mirror it.
int scale_tables(const hyorb_extractor_params &p, float *scale, float *inv, float *sigma2, float *inv_sigma2, int *quota)
{
    const int L = p.nlevels;
    if (L < 1 || L > HYORB_MAX_LEVELS) { set_error("nlevels=%d outside 1..%d", L, HYORB_MAX_LEVELS); return HYORB_EINVAL; }
    if (!(p.scale_factor > 1.0f) && L > 1) { set_error("scale_factor must be > 1"); return HYORB_EINVAL; }
    const double sf = (double)p.scale_factor;
    float s[HYORB_MAX_LEVELS], s2[HYORB_MAX_LEVELS];
    s[0] = 1.0f; s2[0] = 1.0f;
    for (int i = 1; i < L; i++) {
        s[i] = (float)((double)s[i - 1] * sf);      // :92
        s2[i] = s[i] * s[i];                         // :93
    }
    for (int i = 0; i < L; i++) {
        if (scale) scale[i] = s[i];
        if (sigma2) sigma2[i] = s2[i];
        if (inv) inv[i] = 1.0f / s[i];               // :100
        if (inv_sigma2) inv_sigma2[i] = 1.0f / s2[i];// :101
    }
    if (quota) {
        const float factor = (float)(1.0 / sf);      // :107
        float want = (float)p.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)L));   // :108
        int sum = 0;
        for (int l = 0; l < L - 1; l++) {
            quota[l] = round_half_even(want);        // :113
            sum += quota[l];
            want *= factor;                          // :115
        }
        quota[L - 1] = p.nfeatures - sum > 0 ? p.nfeatures - sum : 0;   // :117
    }
    return HYORB_OK;
}

// cv::resize INTER_LINEAR coefficient tables (OpenCV imgproc/resize.cpp, called at ORBExtractor.cpp:577).
// Horizontal taps that fall outside the source reset the fraction; vertical taps keep the fraction and the
// kernel clips the two source rows instead -- that asymmetry is OpenCV's.
static void linear_table(int dst_n, int src_n, bool vertical, ResizeTab *t)
{
    const double inv_scale = (double)dst_n / src_n;
    const double sc = 1.0 / inv_scale;
    for (int d = 0; d < dst_n; d++) {
        float f = (float)((d + 0.5) * sc - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (!vertical) {
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= src_n - 1) { s = src_n - 1; f = 0.f; }
        }
        t[d].ofs = s;
        t[d].c0 = (short)round_half_even((1.f - f) * 2048.f);
        t[d].c1 = (short)round_half_even(f * 2048.f);
    }
}

// path bits of one lattice coordinate through the quadtree: ExtractorNode::DivideNode halves a node with
// ceil((hi-lo)/2) and sends a point left/up iff coordinate < lo+half (ORBExtractor.cpp:123-124, 151-166).
static uint32_t axis_path(int v, int lo, int hi)
{
    uint32_t bits = 0;
    for (int d = 0; d < QT_DMAX; d++) {
        const int half = (hi - lo + 1) >> 1;     // ceil of a non-negative integer halved
        const int mid = lo + half;
        bits <<= 1;
        if (v < mid) hi = mid;
        else { bits |= 1u; lo = mid; }
    }
    return bits;
}
static uint32_t spread_bits(uint32_t v)   // bit i -> bit 2i
{
    uint32_t r = 0;
    for (int i = 0; i < QT_DMAX; i++) r |= ((v >> i) & 1u) << (2 * i);
    return r;
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

int build_plan(const hyorb_extractor_params &p, int width, int height, HostPlan *out)
{
    float scale[HYORB_MAX_LEVELS], inv[HYORB_MAX_LEVELS];
    int quota[HYORB_MAX_LEVELS];
    HY_TRY(scale_tables(p, scale, inv, nullptr, nullptr, quota));
    if (p.cell_px < 1) { set_error("cell_px must be >= 1"); return HYORB_EINVAL; }
    if (width > MAX_DIM || height > MAX_DIM) { set_error("image %dx%d larger than %d px", width, height, MAX_DIM); return HYORB_EUNSUPPORTED; }
    PlanDev &P = out->dev;
    memset(&P, 0, sizeof(P));
    P.nlevels = p.nlevels; P.width = width; P.height = height;
    out->resize.clear(); out->lut.clear();
    unsigned long long off = 0; unsigned candOff = 0, selOff = 0; int tileBase = 0;
    for (int l = 0; l < p.nlevels; l++) {
        LevelDev &L = P.lv[l];
        L.w = round_half_even((float)width * inv[l]);        // ORBExtractor.cpp:569
        L.h = round_half_even((float)height * inv[l]);
        // the reference divides by zero / indexes out of range on levels this small; refuse instead
        if (L.w < 2 * LATTICE_MIN + p.cell_px || L.h < 2 * LATTICE_MIN + p.cell_px) {
            set_error("level %d of a %dx%d image is %dx%d: smaller than one FAST cell plus borders", l, width, height, L.w, L.h);
            return HYORB_EUNSUPPORTED;
        }
        L.pitch = (L.w + 15) & ~15;
        L.off = off;
        off += (unsigned long long)L.pitch * L.h;
        off = (off + 255) & ~255ull;
        L.maxBX = L.w - LATTICE_MIN; L.maxBY = L.h - LATTICE_MIN;
        const float fw = (float)(L.maxBX - LATTICE_MIN), fh = (float)(L.maxBY - LATTICE_MIN);
        const float W = (float)p.cell_px;
        L.nCols = (int)(fw / W); L.nRows = (int)(fh / W);               // :425-426
        L.wCell = (int)ceilf(fw / (float)L.nCols); L.hCell = (int)ceilf(fh / (float)L.nRows);   // :427-428
        L.mulW = (unsigned)(((1u << 20) + L.wCell - 1) / L.wCell); L.mulH = (unsigned)(((1u << 20) + L.hCell - 1) / L.hCell);
        if (L.wCell + 6 > 127 || L.hCell + 6 > 127) { set_error("cell of %dx%d px not supported", L.wCell, L.hCell); return HYORB_EUNSUPPORTED; }
        const int detW = (L.maxBX - 3) - DET_MIN, detH = (L.maxBY - 3) - DET_MIN;
        L.tilesX = detW > 0 ? (detW + FT_OW - 1) / FT_OW : 0;
        L.tilesY = detH > 0 ? (detH + FT_OH - 1) / FT_OH : 0;
        if (L.tilesX == 0 || L.tilesY == 0) L.tilesX = L.tilesY = 0;
        L.tileBase = tileBase; tileBase += L.tilesX * L.tilesY;
        L.quota = quota[l];
        // DistributeOctTree roots (:183-185); minX..maxX = lattice bounds
        const int lw = L.maxBX - LATTICE_MIN, lh = L.maxBY - LATTICE_MIN;
        L.nIni = (int)roundf((float)lw / (float)lh);
        if (L.nIni < 1 || L.nIni > QT_MAX_ROOTS) { set_error("level %d aspect %dx%d gives %d quadtree roots (supported 1..%d)", l, lw, lh, L.nIni, QT_MAX_ROOTS); return HYORB_EUNSUPPORTED; }
        L.hX = (float)lw / (float)L.nIni;
        // cell-local 3x3 NMS keeps at most ceil(wCell/2) * ceil(hCell/2) pixels of a cell: a quarter of the level plus one row / column per cell
        L.candCap = (L.w / 2 + L.nCols + 2) * (L.h / 2 + L.nRows + 2);
        L.candOff = candOff; candOff += (unsigned)((L.candCap + 63) & ~63);
        int need = 4 * L.quota; if (need < 4 * L.nIni) need = 4 * L.nIni; if (need < 64) need = 64;
        L.qtMaxN = next_pow2(need);
        if (L.qtMaxN > 8192) { set_error("level %d quota %d exceeds the quadtree kernel's limit of 2048 features per level", l, L.quota); return HYORB_EUNSUPPORTED; }
        L.selCap = L.qtMaxN;
        L.selOff = selOff; selOff += (unsigned)L.selCap;
        P.selTotalCap += L.selCap;
        L.scale = scale[l];
        L.kpSize = (float)(int)((float)PATCH_SIZE * scale[l]);          // :478 (int*float -> float -> int)
        // quadtree LUTs: candidates have lattice x in [3, lw-3), y likewise; index directly by lattice coordinate
        L.lutX = (int)out->lut.size();
        for (int x = 0; x <= lw; x++) {
            int r = (int)((float)x / L.hX);                             // :213 vpIniNodes[kp.pt.x/hX]
            if (r >= L.nIni) r = L.nIni - 1;                            // x == lw is never a candidate; keep the table in range
            const int ulx = (int)(L.hX * (float)r), urx = (int)(L.hX * (float)(r + 1));   // :197-198
            out->lut.push_back(((uint32_t)r << (2 * QT_DMAX)) | spread_bits(axis_path(x, ulx, urx)));
        }
        L.lutY = (int)out->lut.size();
        for (int y = 0; y <= lh; y++) out->lut.push_back(spread_bits(axis_path(y, 0, lh)) << 1);
        // resize tables from level l-1
        if (l > 0) {
            const LevelDev &S = P.lv[l - 1];
            L.area2x = (S.w == 2 * L.w && S.h == 2 * L.h);
            L.rsX = (int)out->resize.size(); out->resize.resize(out->resize.size() + L.w);
            linear_table(L.w, S.w, false, out->resize.data() + L.rsX);
            if (!L.area2x)      // k_resize picks a thread's 8 horizontal taps out of 8 consecutive source bytes
                for (int x = 0; x < L.w; x += 4) {
                    const ResizeTab *t = out->resize.data() + L.rsX;
                    if (t[x + 3 < L.w ? x + 3 : L.w - 1].ofs - t[x].ofs > 6) { set_error("scale_factor %.3f too large for the resize kernel (max 2)", p.scale_factor); return HYORB_EUNSUPPORTED; }
                }
            L.rsY = (int)out->resize.size(); out->resize.resize(out->resize.size() + L.h);
            linear_table(L.h, S.h, true, out->resize.data() + L.rsY);
        }
    }
    blur_tiles(&P);
    level_tiles(out);
    P.tilesPerImage = tileBase;
    P.pyrStride = off;
    P.candStride = candOff;
    P.selStride = selOff;
    return HYORB_OK;
}

}  // namespace hyorb
