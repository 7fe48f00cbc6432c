// stereo.cu -- K7: stereo association.  Replaces Stereomatcher::computeStereoMatches
// (src/features/Stereomatcher.cpp:36-156), constructed by ImageProcessing::ProcessStereoImage
// (src/main/ImageProcessing.cpp:100-103) with default thresholds 100 / 50 and size_ref 31:
//   row table of right keypoints (band r = 2*size/size_ref, :52-65), octave +-1 and uR in [uL-maxD, uL] filters
//   (:98-104), best Hamming distance below TH_HIGH (strict <, first right index wins ties, :110-114), accept below
//   (TH_HIGH+TH_LOW)/2 (:118), disparity = uL-uR0 with NO sub-pixel refinement (:120-133),
//   then drop matches with distance >= 1.5*1.4*median (:137-155).
// Three launches over a batch of pairs:
//   k_stereo_table  one CTA per pair: row table (CSR) of the right keypoints, shared-memory counters + block scan.
//                   Order inside a row is free: the winner below is the minimum of the packed key distance<<16 | iR,
//                   which is what "first wins" means for the reference's ascending list.
//   k_stereo_search one WARP per left keypoint: lanes stride over the row's candidates (uint4 XOR + POPC), warp
//                   min-reduction of the packed key; accepted distances go into a per-pair 257-bin histogram.
//   k_stereo_cut    one CTA per pair: median from the histogram (no sort), distance cut.
#include "common.cuh"

namespace hyorb {

constexpr int ST_THREADS = 256;
constexpr int ST_MAX_ROWS = MAX_DIM + 1;
constexpr int ST_ROWS_PER_KP = 20;     // default row-table budget: AVERAGE rows per right keypoint (a keypoint covers ceil(2r)+2 rows, r = 2*size/31);
                                       // callers that know the largest keypoint size pass a larger budget (launch_stereo's rows_budget)
constexpr int ST_HIST = 264;

struct StereoArgs {
    hyorb_stereo_params sp;
    float maxD;
    const hyorb_keypoint *kps; const uint4 *desc; const int32_t *counts; int capacity;
    int32_t *rowtab; int tabCap; int32_t *rowoff; int32_t *hist;
    float *uR; float *depth; int32_t *best_r; int32_t *best_d;
    int *status;
};

__device__ __forceinline__ bool right_band(const hyorb_keypoint &k, float size_ref, int nRows, int &minr, int &maxr)
{
    const float r = __fdiv_rn(__fmul_rn(2.0f, k.size), size_ref);                 // :57
    maxr = (int)ceilf(__fadd_rn(k.y, r)); minr = (int)floorf(__fsub_rn(k.y, r));  // :58-59
    return !(minr < 0 || maxr >= nRows);
}

__global__ void __launch_bounds__(ST_THREADS)
k_stereo_table(StereoArgs A)
{
    grid_dependency_wait();      // launch_dependent (common.cuh)
    __shared__ int s_off[ST_MAX_ROWS + 1];
    __shared__ int s_fill[ST_MAX_ROWS];
    __shared__ int s_part[ST_THREADS];
    const int tid = threadIdx.x, p = blockIdx.x;
    const int nRows = A.sp.n_rows;
    const hyorb_keypoint *kr = A.kps + (size_t)(2 * p + 1) * A.capacity;
    const int nr = min(A.counts[2 * p + 1], A.capacity);
    int32_t *tab = A.rowtab + (size_t)p * A.tabCap;
    int32_t *goff = A.rowoff + (size_t)p * (ST_MAX_ROWS + 1);
    for (int i = tid; i <= nRows; i += ST_THREADS) s_off[i] = 0;
    for (int i = tid; i < nRows; i += ST_THREADS) s_fill[i] = 0;
    for (int i = tid; i < ST_HIST; i += ST_THREADS) A.hist[(size_t)p * ST_HIST + i] = 0;
    __syncthreads();
    for (int iR = tid; iR < nr; iR += ST_THREADS) {
        int minr, maxr;
        if (!right_band(kr[iR], A.sp.size_ref, nRows, minr, maxr)) { atomicOr(A.status, ST_ROW_RANGE); continue; }   // reference: out-of-bounds write
        for (int yi = minr; yi <= maxr; yi++) atomicAdd(&s_off[yi + 1], 1);
    }
    __syncthreads();
    {   // inclusive scan of s_off[1..nRows] in place: s_off[i] = first entry of row i, s_off[i+1] = one past its last
        const int per = (nRows + ST_THREADS - 1) / ST_THREADS;
        const int lo = min(tid * per, nRows), hi = min(lo + per, nRows);
        int sum = 0;
        for (int i = lo; i < hi; i++) sum += s_off[i + 1];
        s_part[tid] = sum;
        __syncthreads();
        for (int o = 1; o < ST_THREADS; o <<= 1) {
            const int t = tid >= o ? s_part[tid - o] : 0;
            __syncthreads();
            s_part[tid] += t;
            __syncthreads();
        }
        int base = s_part[tid] - sum;
        for (int i = lo; i < hi; i++) { const int c = s_off[i + 1]; s_off[i + 1] = base + c; base += c; }
        __syncthreads();
    }
    const bool overflow = s_off[nRows] > A.tabCap;
    if (overflow && tid == 0) atomicOr(A.status, ST_ROWTAB_OVERFLOW);
    for (int i = tid; i <= nRows; i += ST_THREADS) goff[i] = overflow ? 0 : s_off[i];
    if (overflow) return;
    for (int iR = tid; iR < nr; iR += ST_THREADS) {
        int minr, maxr;
        if (!right_band(kr[iR], A.sp.size_ref, nRows, minr, maxr)) continue;
        // entry = right index | octave << 16: the search filters on the octave without touching the keypoint (capacity <= 65535)
        // (octaves are 0 .. nlevels-1; stored + 4 and clamped to [0, 32767] so that any value a caller could pass in [-4, 32763] stays exact)
        const int entry = iR | (min(max(kr[iR].octave + 4, 0), 0x7FFF) << 16);
        for (int yi = minr; yi <= maxr; yi++) tab[s_off[yi] + atomicAdd(&s_fill[yi], 1)] = entry;
    }
}

__global__ void __launch_bounds__(ST_THREADS)
k_stereo_search(StereoArgs A)
{
    grid_dependency_wait();      // launch_dependent (common.cuh)
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.y;
    const int iL = blockIdx.x * (ST_THREADS / 32) + (threadIdx.x >> 5);
    const int nl = min(A.counts[2 * p], A.capacity);
    if (iL >= nl) return;
    const int nRows = A.sp.n_rows;
    const hyorb_keypoint *kl = A.kps + (size_t)(2 * p) * A.capacity, *kr = A.kps + (size_t)(2 * p + 1) * A.capacity;
    const uint4 *dl = A.desc + (size_t)(2 * p) * A.capacity * 2, *dr = A.desc + (size_t)(2 * p + 1) * A.capacity * 2;
    const int32_t *tab = A.rowtab + (size_t)p * A.tabCap;
    const int32_t *goff = A.rowoff + (size_t)p * (ST_MAX_ROWS + 1);
    const size_t o = (size_t)p * A.capacity + iL;
    float out_uR = -1.f, out_depth = -1.f; int out_r = -1, out_d = -1;

    const hyorb_keypoint kp = kl[iL];
    const float vL = kp.y, uL = kp.x;
    const long long row = (long long)vL;                       // vRowIndices[vL]: float -> size_t (:85)
    if (row < 0 || row >= nRows) {
        if (lane == 0) atomicOr(A.status, ST_ROW_RANGE);
    } else {
        const int lo = goff[row], hi = goff[row + 1];
        const float minU = __fsub_rn(uL, A.maxD), maxU = uL;   // minD = 0 (:69, :90-91)
        uint32_t bestKey = 0xFFFFFFFFu;
        if (lo != hi && !(maxU < 0)) {
            const uint4 a0 = dl[2 * iL], a1 = dl[2 * iL + 1];
            for (int c = lo + lane; c < hi; c += 32) {
                const int entry = tab[c];
                const int iR = entry & 0xFFFF, oR = (entry >> 16) - 4;
                if (oR < kp.octave - 1 || oR > kp.octave + 1) continue;        // :100-101
                const float u = kr[iR].x;
                if (u >= minU && u <= maxU) {                                  // :105
                    const uint4 b0 = dr[2 * iR], b1 = dr[2 * iR + 1];
                    const int d = hamming256(a0, a1, b0, b1);
                    if ((float)d < A.sp.th_high) bestKey = min(bestKey, ((uint32_t)d << 16) | (uint32_t)iR);   // :110-114
                }
            }
        }
        bestKey = __reduce_min_sync(0xffffffffu, bestKey);
        if (bestKey != 0xFFFFFFFFu) {
            const int d = (int)(bestKey >> 16), iR = (int)(bestKey & 0xFFFFu);
            const float dist_threshold = __fdiv_rn(__fadd_rn(A.sp.th_high, A.sp.th_low), 2.0f);   // :41
            if ((float)d < dist_threshold) {                                   // :118
                float uR0 = kr[iR].x;
                float disparity = __fsub_rn(uL, uR0);
                if (disparity >= 0.f && disparity < A.maxD) {
                    if (disparity <= 0) { disparity = 0.01f; uR0 = (float)((double)uL - 0.01); }   // :125-129
                    out_depth = __fdiv_rn(A.sp.mbf, disparity);
                    out_uR = uR0; out_r = iR; out_d = d;
                    if (lane == 0) atomicAdd(&A.hist[(size_t)p * ST_HIST + d], 1);
                }
            }
        }
    }
    if (lane == 0) {
        A.uR[o] = out_uR; A.depth[o] = out_depth; A.best_d[o] = out_d;
        if (A.best_r) A.best_r[o] = out_r;
    }
}

__global__ void __launch_bounds__(ST_THREADS)
k_stereo_cut(StereoArgs A)
{
    grid_dependency_wait();      // launch_dependent (common.cuh)
    __shared__ float s_th;
    const int tid = threadIdx.x, p = blockIdx.x;
    const int nl = min(A.counts[2 * p], A.capacity);
    if (tid == 0) {      // median = element nv/2 of the sorted distances (:137-139)
        const int32_t *hist = A.hist + (size_t)p * ST_HIST;
        int nv = 0;
        for (int i = 0; i <= 256; i++) nv += hist[i];
        float th = 3.0e38f;
        if (nv > 0) {    // (empty: the reference reads v[0] of an empty vector -- undefined; nothing to cut here)
            int k = nv / 2, acc = 0, med = 0;
            for (int i = 0; i <= 256; i++) { acc += hist[i]; if (acc > k) { med = i; break; } }
            th = __fmul_rn(1.5f * 1.4f, (float)med);           // :140
        }
        s_th = th;
    }
    __syncthreads();
    const float thDist = s_th;
    const size_t o = (size_t)p * A.capacity;
    for (int iL = tid; iL < nl; iL += ST_THREADS) {
        const int d = A.best_d[o + iL];
        if (d >= 0 && (float)d >= thDist) {                    // :142-153
            A.uR[o + iL] = -1.f; A.depth[o + iL] = -1.f; A.best_d[o + iL] = -1;
            if (A.best_r) A.best_r[o + iL] = -1;
        }
    }
}

int stereo_rows_budget(float max_kp_size, float size_ref)
{
    int rows = ST_ROWS_PER_KP;
    if (max_kp_size > 0 && size_ref > 0) {
        const float r = 2.0f * max_kp_size / size_ref;
        const int need = (int)ceilf(2.0f * r) + 3;
        if (need > rows) rows = need;
    }
    return rows > ST_MAX_ROWS ? ST_MAX_ROWS : rows;
}
size_t stereo_scratch_ints_per_pair(int capacity, int rows_budget) { return (size_t)capacity * (rows_budget > 0 ? rows_budget : ST_ROWS_PER_KP) + (ST_MAX_ROWS + 1) + ST_HIST; }

// scratch: stereo_scratch_ints_per_pair(capacity, rows_budget) * n_pairs int32
int launch_stereo(const hyorb_stereo_params &sp, int n_pairs, const hyorb_keypoint *kps, const uint8_t *desc, const int32_t *counts, int capacity,
                  int32_t *scratch, float *uR, float *depth, int32_t *best_r, int32_t *best_d, int *status, cudaStream_t st, long *launches, int rows_budget)
{
    if (n_pairs <= 0) return HYORB_OK;
    if (sp.n_rows < 1 || sp.n_rows > ST_MAX_ROWS) { set_error("n_rows=%d outside 1..%d", sp.n_rows, ST_MAX_ROWS); return HYORB_EINVAL; }
    if (capacity > 65535) { set_error("stereo capacity %d > 65535", capacity); return HYORB_EUNSUPPORTED; }
    if (!best_d) { set_error("internal: best_d scratch required"); return HYORB_EINVAL; }
    StereoArgs A;
    A.sp = sp;
    const float mb = sp.mbf / sp.fx;            // Camera: mb = mbf / fx
    A.maxD = sp.mbf / mb;                       // Stereomatcher.cpp:68-70 (minZ = mb)
    A.kps = kps; A.desc = (const uint4 *)desc; A.counts = counts; A.capacity = capacity;
    A.tabCap = capacity * (rows_budget > 0 ? rows_budget : ST_ROWS_PER_KP);
    A.rowtab = scratch;
    A.rowoff = scratch + (size_t)A.tabCap * n_pairs;
    A.hist = A.rowoff + (size_t)(ST_MAX_ROWS + 1) * n_pairs;
    A.uR = uR; A.depth = depth; A.best_r = best_r; A.best_d = best_d; A.status = status;
    HY_CUDA(launch_dependent(k_stereo_table, dim3(n_pairs), dim3(ST_THREADS), 0, st, A));
    dim3 grd((capacity + ST_THREADS / 32 - 1) / (ST_THREADS / 32), n_pairs);
    HY_CUDA(launch_dependent(k_stereo_search, grd, dim3(ST_THREADS), 0, st, A));
    HY_CUDA(launch_dependent(k_stereo_cut, dim3(n_pairs), dim3(ST_THREADS), 0, st, A));
    *launches += 3;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

}  // namespace hyorb
