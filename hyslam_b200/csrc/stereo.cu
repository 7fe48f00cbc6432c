// stereo.cu -- K7: stereo association.  Replaces Stereomatcher::computeStereoMatches
// (src/features/Stereomatcher.cpp:36-156), constructed by ImageProcessing::ProcessStereoImage
// (src/main/ImageProcessing.cpp:100-103) with default thresholds 100 / 50 and size_ref 31:
//   row table of right keypoints (band r = 2*size/size_ref, :52-65), octave +-1 and uR in [uL-maxD, uL] filters
//   (:98-104), best Hamming distance below TH_HIGH (strict <, first right index wins ties, :110-114), accept below
//   (TH_HIGH+TH_LOW)/2 (:118), disparity = uL-uR0 with NO sub-pixel refinement (:120-133),
//   then drop matches with distance >= 1.5*1.4*median (:137-155).
// One CTA per stereo pair.  The row table is built with shared-memory counters (order inside a row is free: the
// winner is the minimum of the packed key distance<<16 | iR, which is what "first wins" means for an ascending
// list); the median comes from a 257-bin histogram instead of a sort.
#include "common.cuh"

namespace hyorb {

constexpr int ST_THREADS = 256;
constexpr int ST_MAX_ROWS = MAX_DIM + 1;
constexpr int ST_ROWS_PER_KP = 20;     // ceil(2*r)+2 with r = 2*size/31 and size <= 31*scale_max; checked at run time

__global__ void __launch_bounds__(ST_THREADS)
k_stereo(hyorb_stereo_params sp, float maxD, const hyorb_keypoint *__restrict__ kps, const uint4 *__restrict__ desc,
         const int32_t *__restrict__ counts, int capacity, int32_t *__restrict__ rowtab, int tabCap,
         float *__restrict__ uR_out, float *__restrict__ depth_out, int32_t *__restrict__ best_r, int32_t *__restrict__ best_d,
         int *__restrict__ status)
{
    __shared__ int s_off[ST_MAX_ROWS + 1];
    __shared__ int s_fill[ST_MAX_ROWS];
    __shared__ int s_part[ST_THREADS];
    __shared__ int s_hist[260];
    __shared__ float s_th;
    const int tid = threadIdx.x, p = blockIdx.x;
    const int nRows = sp.n_rows;
    const hyorb_keypoint *kl = kps + (size_t)(2 * p) * capacity, *kr = kps + (size_t)(2 * p + 1) * capacity;
    const uint4 *dl = desc + (size_t)(2 * p) * capacity * 2, *dr = desc + (size_t)(2 * p + 1) * capacity * 2;
    const int nl = min(counts[2 * p], capacity), nr = min(counts[2 * p + 1], capacity);
    float *uR = uR_out + (size_t)p * capacity, *depth = depth_out + (size_t)p * capacity;
    int32_t *br = best_r ? best_r + (size_t)p * capacity : nullptr, *bd = best_d ? best_d + (size_t)p * capacity : nullptr;
    int32_t *tab = rowtab + (size_t)p * tabCap;

    for (int i = tid; i <= nRows; i += ST_THREADS) s_off[i] = 0;
    for (int i = tid; i < nRows; i += ST_THREADS) s_fill[i] = 0;
    for (int i = tid; i < 260; i += ST_THREADS) s_hist[i] = 0;
    for (int i = tid; i < nl; i += ST_THREADS) { uR[i] = -1.f; depth[i] = -1.f; if (br) br[i] = -1; if (bd) bd[i] = -1; }
    __syncthreads();
    // ---- row table, pass 1: counts (:52-65)
    for (int iR = tid; iR < nr; iR += ST_THREADS) {
        const float y = kr[iR].y;
        const float r = __fdiv_rn(__fmul_rn(2.0f, kr[iR].size), sp.size_ref);
        const int maxr = (int)ceilf(__fadd_rn(y, r)), minr = (int)floorf(__fsub_rn(y, r));
        if (minr < 0 || maxr >= nRows || maxr - minr + 1 > ST_ROWS_PER_KP) { atomicOr(status, ST_ROW_RANGE); continue; }   // reference: out-of-bounds write
        for (int yi = minr; yi <= maxr; yi++) atomicAdd(&s_off[yi + 1], 1);
    }
    __syncthreads();
    // exclusive scan of s_off[1..nRows] in place (chunk per thread)
    {
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
        for (int i = lo; i < hi; i++) { const int c = s_off[i + 1]; s_off[i + 1] = base + c; base += c; }   // inclusive -> s_off[i+1] = end of row i
        __syncthreads();
    }
    if (s_off[nRows] > tabCap) { if (tid == 0) atomicOr(status, ST_OUT_OVERFLOW); return; }
    // pass 2: fill
    for (int iR = tid; iR < nr; iR += ST_THREADS) {
        const float y = kr[iR].y;
        const float r = __fdiv_rn(__fmul_rn(2.0f, kr[iR].size), sp.size_ref);
        const int maxr = (int)ceilf(__fadd_rn(y, r)), minr = (int)floorf(__fsub_rn(y, r));
        if (minr < 0 || maxr >= nRows || maxr - minr + 1 > ST_ROWS_PER_KP) continue;
        for (int yi = minr; yi <= maxr; yi++) tab[s_off[yi] + atomicAdd(&s_fill[yi], 1)] = iR;
    }
    __syncthreads();
    // ---- per left keypoint search (:76-134)
    const float dist_threshold = __fdiv_rn(__fadd_rn(sp.th_high, sp.th_low), 2.0f);    // :41
    for (int iL = tid; iL < nl; iL += ST_THREADS) {
        const hyorb_keypoint kp = kl[iL];
        const float vL = kp.y, uL = kp.x;
        const long long row = (long long)vL;              // vRowIndices[vL]: float -> size_t
        if (row < 0 || row >= nRows) { atomicOr(status, ST_ROW_RANGE); continue; }
        const int lo = s_off[row], hi = s_off[row + 1];
        if (lo == hi) continue;
        const float minU = __fsub_rn(uL, maxD), maxU = uL;   // minD = 0
        if (maxU < 0) continue;
        const uint4 a0 = dl[2 * iL], a1 = dl[2 * iL + 1];
        uint32_t bestKey = 0xFFFFFFFFu;
        for (int c = lo; c < hi; c++) {
            const int iR = tab[c];
            const int oR = kr[iR].octave;
            if (oR < kp.octave - 1 || oR > kp.octave + 1) continue;
            const float u = kr[iR].x;
            if (u >= minU && u <= maxU) {
                const uint4 b0 = dr[2 * iR], b1 = dr[2 * iR + 1];
                const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                              __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
                if ((float)d < sp.th_high) bestKey = min(bestKey, ((uint32_t)d << 16) | (uint32_t)iR);
            }
        }
        if (bestKey == 0xFFFFFFFFu) continue;
        const int d = (int)(bestKey >> 16), iR = (int)(bestKey & 0xFFFFu);
        if ((float)d < dist_threshold) {                                    // :118
            float uR0 = kr[iR].x;
            float disparity = __fsub_rn(uL, uR0);
            if (disparity >= 0.f && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01f; uR0 = (float)((double)uL - 0.01); }   // :125-129
                depth[iL] = __fdiv_rn(sp.mbf, disparity);
                uR[iL] = uR0;
                if (br) br[iL] = iR;
                if (bd) bd[iL] = d;
                atomicAdd(&s_hist[d], 1);
            }
        }
    }
    __syncthreads();
    // ---- median cut (:137-155): median = element nv/2 of the sorted distances
    if (tid == 0) {
        int nv = 0;
        for (int i = 0; i <= 256; i++) nv += s_hist[i];
        float th = 3.0e38f;
        if (nv > 0) {
            int k = nv / 2, acc = 0, med = 0;
            for (int i = 0; i <= 256; i++) { acc += s_hist[i]; if (acc > k) { med = i; break; } }
            th = __fmul_rn(1.5f * 1.4f, (float)med);
        }
        s_th = th;
    }
    __syncthreads();
    const float thDist = s_th;
    for (int iL = tid; iL < nl; iL += ST_THREADS) {
        // matched entries carry their distance in best_d when requested; recompute from uR otherwise is impossible,
        // so the distance is kept in depth's sibling array below
        if (uR[iL] == -1.f && depth[iL] == -1.f) continue;
        const int d = bd ? bd[iL] : -1;
        if (d >= 0 && (float)d >= thDist) { uR[iL] = -1.f; depth[iL] = -1.f; if (br) br[iL] = -1; bd[iL] = -1; }
    }
}

int launch_stereo(const hyorb_stereo_params &sp, int n_pairs, const hyorb_keypoint *kps, const uint8_t *desc, const int32_t *counts, int capacity,
                  int32_t *rowtab, int tabCap, float *uR, float *depth, int32_t *best_r, int32_t *best_d, int *status, cudaStream_t st, long *launches)
{
    if (n_pairs <= 0) return HYORB_OK;
    if (sp.n_rows < 1 || sp.n_rows > ST_MAX_ROWS) { set_error("n_rows=%d outside 1..%d", sp.n_rows, ST_MAX_ROWS); return HYORB_EINVAL; }
    if (capacity > 65535) { set_error("stereo capacity %d > 65535", capacity); return HYORB_EUNSUPPORTED; }
    if (!best_d) { set_error("internal: best_d scratch required"); return HYORB_EINVAL; }
    const float mb = sp.mbf / sp.fx;            // Camera: mb = mbf / fx
    const float maxD = sp.mbf / mb;             // Stereomatcher.cpp:68-70 (minZ = mb)
    k_stereo<<<n_pairs, ST_THREADS, 0, st>>>(sp, maxD, kps, (const uint4 *)desc, counts, capacity, rowtab, tabCap, uR, depth, best_r, best_d, status);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

}  // namespace hyorb
