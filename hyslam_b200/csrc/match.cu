// match.cu -- K6: binary-descriptor matching.  Replaces the inner loops of HYSLAM::FeatureMatcher / MatchCriteria:
//   * ORBDistance::distance (src/features/low_level/DescriptorDistance.cpp:9-25): 256-bit Hamming distance, here
//     uint4 loads, XOR and __popc;
//   * BestScoreCriterionCore (src/features/MatchCriteria.cpp:248-280): best / second-best scan with strict `<`
//     (first candidate wins ties, the second best may equal the best);
//   * the acceptance rules of BestScoreCriterion (:214-246), BestMatchBoWCriterion (:601-635), MonoInitBestScore (:486-523);
//   * Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInAreaNEW (src/core/Frame.cc:137-153, 459-469, 416-457);
//   * PreviouslyMatchedCriterion (:124-144), FeatureSizeCriterion (:350-360), StereoConsistencyCriterion (:149-177);
//   * RotationConsistency + ComputeThreeMaxima (:684-767).
// INT-pipe bound (XOR/POPC), HBM traffic negligible; no tensor cores (bit work, not a contraction).
//
// Order semantics without a sequential scan: (best, first index) = min over candidates of the packed key
// (distance << 22 | position); second = 2nd order statistic of the distances with multiplicity.  Both merge
// associatively, so lanes / CTAs scan slices and combine with warp shuffles.
#include <float.h>
#include <cub/device/device_radix_sort.cuh>
#include "common.cuh"

namespace hyorb {

constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;
constexpr int POS_BITS = 22;
constexpr uint32_t POS_MASK = (1u << POS_BITS) - 1;
constexpr int DIST_NONE = 1023;    // > 256: "no candidate" (the reference's FLT_MAX)

// running (best key, second distance) update with one more candidate
__device__ __forceinline__ void scan_update(uint32_t &bkey, int &second, int d, uint32_t pos)
{
    const uint32_t key = ((uint32_t)d << POS_BITS) | pos;
    if (key < bkey) { second = min(second, (int)(bkey >> POS_BITS)); bkey = key; }
    else second = min(second, d);
}
// merge two partial scans
__device__ __forceinline__ void scan_merge(uint32_t &bkey, int &second, uint32_t okey, int osecond)
{
    const uint32_t lo = min(bkey, okey), hi = max(bkey, okey);
    second = min(min(second, osecond), (int)(hi >> POS_BITS));
    bkey = lo;
}
__device__ __forceinline__ void warp_merge(uint32_t &bkey, int &second)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t ok = __shfl_xor_sync(0xffffffffu, bkey, o);
        const int os = __shfl_xor_sync(0xffffffffu, second, o);
        scan_merge(bkey, second, ok, os);
    }
}

__device__ __forceinline__ void write_result(int q, uint32_t bkey, int second, int best_target, int rule, float thr, float ratio,
                                             int32_t *best_idx, uint16_t *best, uint16_t *sec, uint8_t *accepted)
{
    const int bd = (int)(bkey >> POS_BITS);
    const bool has = bd < DIST_NONE;
    const float fb = has ? (float)bd : FLT_MAX, fs = second < DIST_NONE ? (float)second : FLT_MAX;
    best_idx[q] = has ? best_target : -1;
    best[q] = has ? (uint16_t)bd : (uint16_t)65535;
    sec[q] = second < DIST_NONE ? (uint16_t)second : (uint16_t)65535;
    accepted[q] = (uint8_t)(has && accept_rule(rule, fb, fs, thr, ratio));
}

// ---------------- brute force: every target in index order (C4: 8000 x 8000)
constexpr int BF_THREADS = 128, BF_TILE = 128;

// Two queries per thread: a target's two 128-bit shared-memory loads and its position bookkeeping are shared by two distances.
#ifndef HYORB_BF_QPT
#define HYORB_BF_QPT 2
#endif
constexpr int BF_QPT = HYORB_BF_QPT;
int bf_queries_per_cta() { return BF_THREADS * BF_QPT; }
__global__ void __launch_bounds__(BF_THREADS)
k_bf_partial(const uint4 *__restrict__ q, int nq, const uint4 *__restrict__ t, int nt, int chunk,
             uint32_t *__restrict__ pkey, uint16_t *__restrict__ psecond)
{
    __shared__ uint4 s_t[BF_TILE * 2];
    const int q0 = (blockIdx.x * BF_THREADS + threadIdx.x) * BF_QPT;
    const int t0 = blockIdx.y * chunk, t1 = min(t0 + chunk, nt);
    uint4 a0[BF_QPT], a1[BF_QPT];
    // running smallest and second-smallest KEY (distance << 22 | target): three min / max per distance, no branch; the second-best
    // distance of the rules (counted with multiplicity, like the reference's best / second-best update) is the second key's distance
    uint32_t bkey[BF_QPT], skey[BF_QPT];
#pragma unroll
    for (int k = 0; k < BF_QPT; k++) {
        a0[k] = a1[k] = make_uint4(0, 0, 0, 0);
        if (q0 + k < nq) { a0[k] = q[2 * (size_t)(q0 + k)]; a1[k] = q[2 * (size_t)(q0 + k) + 1]; }
        bkey[k] = KEY_NONE; skey[k] = KEY_NONE;
    }
    for (int base = t0; base < t1; base += BF_TILE) {
        const int cnt = min(BF_TILE, t1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * cnt; i += BF_THREADS) s_t[i] = t[2 * (size_t)base + i];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cnt; j++) {
            const uint4 b0 = s_t[2 * j], b1 = s_t[2 * j + 1];
#pragma unroll
            for (int k = 0; k < BF_QPT; k++) {
                const uint32_t key = ((uint32_t)hamming256(a0[k], a1[k], b0, b1) << POS_BITS) + (uint32_t)(base + j);
                skey[k] = min(skey[k], max(bkey[k], key));
                bkey[k] = min(bkey[k], key);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < BF_QPT; k++)
        if (q0 + k < nq) {
            pkey[(size_t)blockIdx.y * nq + q0 + k] = bkey[k];
            psecond[(size_t)blockIdx.y * nq + q0 + k] = (uint16_t)min(skey[k] >> POS_BITS, (uint32_t)DIST_NONE);
        }
}

__global__ void k_bf_finalize(const uint32_t *__restrict__ pkey, const uint16_t *__restrict__ psecond, int nsplit, int nq,
                              int rule, float thr, float ratio, int32_t *best_idx, uint16_t *best, uint16_t *sec, uint8_t *accepted)
{
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    uint32_t bkey = KEY_NONE; int second = DIST_NONE;
    for (int s = 0; s < nsplit; s++) scan_merge(bkey, second, pkey[(size_t)s * nq + qi], (int)psecond[(size_t)s * nq + qi]);
    write_result(qi, bkey, second, (int)(bkey & POS_MASK), rule, thr, ratio, best_idx, best, sec, accepted);
}

// ---------------- EpipolarConsistencyBoWCriterion::CheckDistEpipolarLine (MatchCriteria.cpp:659-676), plain fp32, left to right, unfused
struct EpiLine { float a, b, c, den; };
__device__ __forceinline__ EpiLine epi_line(const EpipolarDev &epi, int qi)       // l = x1' F12 = [a b c]
{
    const float x1 = epi.kps1[qi].x, y1 = epi.kps1[qi].y;
    EpiLine l;
    l.a = x1 * epi.F[0] + y1 * epi.F[3] + epi.F[6];
    l.b = x1 * epi.F[1] + y1 * epi.F[4] + epi.F[7];
    l.c = x1 * epi.F[2] + y1 * epi.F[5] + epi.F[8];
    l.den = l.a * l.a + l.b * l.b;
    return l;
}
__device__ __forceinline__ bool epi_pass(const EpipolarDev &epi, const EpiLine &l, int ti)
{
    const hyorb_keypoint k2 = epi.kps2[ti];
    const float sf = k2.size / epi.size_ref;                    // FeatureExtractorSettings::determineSigma2 (FeatureExtractorSettings.cpp:5-8)
    const float sigma2 = epi.sigma_ref * (sf * sf);
    const float num = l.a * k2.x + l.b * k2.y + l.c;
    if (l.den == 0.f) return false;
    const float dsqr = num * num / l.den;
    return (double)dsqr < 3.84 * (double)sigma2;                // `dsqr < 3.84*sigma2`: the literal is a double
}

// ---------------- candidate lists (CSR): one warp per query, lanes stride over the list
__global__ void __launch_bounds__(256)
k_match_csr(const uint4 *__restrict__ q, int nq, const uint4 *__restrict__ t, int nt, const int32_t *__restrict__ off,
            const int32_t *__restrict__ idx, int rule, float thr, float ratio, int32_t *best_idx, uint16_t *best, uint16_t *sec,
            uint8_t *accepted, int *status, const EpipolarDev epi)
{
    const int lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= nq) return;
    const uint4 a0 = q[2 * qi], a1 = q[2 * qi + 1];
    const int lo = off[qi], hi = off[qi + 1];
    uint32_t bkey = KEY_NONE; int second = DIST_NONE;
    EpiLine el{0.f, 0.f, 0.f, 0.f};
    if (epi.kps1) el = epi_line(epi, qi);
    for (int c = lo + lane; c < hi; c += 32) {
        const int ti = idx[c];
        if (ti < 0 || ti >= nt) { atomicOr(status, ST_BAD_INDEX); continue; }
        if (epi.kps1 && !epi_pass(epi, el, ti)) continue;
        const int d = hamming256(a0, a1, t[2 * (size_t)ti], t[2 * (size_t)ti + 1]);
        scan_update(bkey, second, d, (uint32_t)(c - lo) & POS_MASK);
    }
    warp_merge(bkey, second);
    if (lane == 0) {
        const int bt = bkey != KEY_NONE ? idx[lo + (int)(bkey & POS_MASK)] : -1;
        write_result(qi, bkey, second, bt, rule, thr, ratio, best_idx, best, sec, accepted);
    }
}

// ---------------- Frame grid (Frame.cc:137-153, 459-469)
__global__ void k_grid_cells(const hyorb_keypoint *__restrict__ kps, int n, hyorb_bounds b, int32_t *__restrict__ cell_of,
                             int32_t *__restrict__ cell_cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float invW = __fdiv_rn((float)HYORB_GRID_COLS, __fsub_rn(b.max_x, b.min_x));   // Frame.cc:65-66
    const float invH = __fdiv_rn((float)HYORB_GRID_ROWS, __fsub_rn(b.max_y, b.min_y));
    const int px = (int)roundf(__fmul_rn(__fsub_rn(kps[i].x, b.min_x), invW));            // PosInGrid uses round()
    const int py = (int)roundf(__fmul_rn(__fsub_rn(kps[i].y, b.min_y), invH));
    int c = -1;
    if (px >= 0 && px < HYORB_GRID_COLS && py >= 0 && py < HYORB_GRID_ROWS) { c = px * HYORB_GRID_ROWS + py; atomicAdd(&cell_cnt[c], 1); }
    cell_of[i] = c;
}

// single CTA: exclusive scan of the 3072 cell counts, then a stable fill (ascending keypoint index inside a cell,
// i.e. the reference's push_back order) -- warps take turns so no atomics are needed.
__global__ void __launch_bounds__(1024)
k_grid_fill(const int32_t *__restrict__ cell_of, int n, const int32_t *__restrict__ cell_cnt, int32_t *__restrict__ cell_off,
            int32_t *__restrict__ cell_idx)
{
    constexpr int NC = HYORB_GRID_COLS * HYORB_GRID_ROWS;
    __shared__ int s_off[NC + 1];
    __shared__ int s_fill[NC];
    __shared__ int s_part[1024];
    const int tid = threadIdx.x;
    // 3 cells per thread
    int v[3], sum = 0;
    for (int k = 0; k < 3; k++) { const int c = tid * 3 + k; v[k] = c < NC ? cell_cnt[c] : 0; sum += v[k]; }
    s_part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int t = tid >= o ? s_part[tid - o] : 0;
        __syncthreads();
        s_part[tid] += t;
        __syncthreads();
    }
    int base = s_part[tid] - sum;
    for (int k = 0; k < 3; k++) { const int c = tid * 3 + k; if (c < NC) { s_off[c] = base; s_fill[c] = 0; cell_off[c] = base; } base += v[k]; }
    if (tid == 1023) { s_off[NC] = s_part[1023]; cell_off[NC] = s_part[1023]; }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    for (int chunk = 0; chunk < n; chunk += 1024) {
        const int i = chunk + tid;
        const int c = i < n ? cell_of[i] : -1;
        for (int w = 0; w < 32; w++) {
            if (warp == w) {
                const unsigned peers = __match_any_sync(0xffffffffu, c);
                const int rank = __popc(peers & ((1u << lane) - 1));
                const int first = c >= 0 ? s_fill[c] : 0;
                __syncwarp();                      // every lane: all reads of s_fill precede the leaders' updates
                if (c >= 0) {
                    if (rank == 0) s_fill[c] = first + __popc(peers);
                    cell_idx[s_off[c] + first + rank] = i;
                }
            }
            __syncthreads();
        }
    }
}

// ---------------- windowed matching (one warp per landmark query)
__global__ void __launch_bounds__(256)
k_match_window(const hyorb_keypoint *__restrict__ kps, const uint4 *__restrict__ tdesc, const float *__restrict__ t_uR,
               const uint8_t *__restrict__ t_matched, int nt, hyorb_bounds b, const int32_t *__restrict__ cell_off,
               const int32_t *__restrict__ cell_idx, const hyorb_window_query *__restrict__ qs, const uint4 *__restrict__ qdesc, int nq,
               float thr, float ratio, const uint8_t *__restrict__ q_active, int32_t *best_idx, uint16_t *best, uint16_t *sec, uint8_t *accepted,
               WindowCriteria wc)
{
    const int lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= nq) return;
    if (q_active && !q_active[qi]) {        // landmark rejected by the landmark criteria: no candidates, no match
        if (lane == 0) write_result(qi, KEY_NONE, DIST_NONE, -1, wc.rule, thr, ratio, best_idx, best, sec, accepted);
        return;
    }
    const hyorb_window_query Q = qs[qi];
    const uint4 a0 = qdesc[2 * qi], a1 = qdesc[2 * qi + 1];
    // Frame::GetFeaturesInAreaNEW (Frame.cc:416-457)
    const float invW = __fdiv_rn((float)HYORB_GRID_COLS, __fsub_rn(b.max_x, b.min_x));
    const float invH = __fdiv_rn((float)HYORB_GRID_ROWS, __fsub_rn(b.max_y, b.min_y));
    int x0 = (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(Q.u, b.min_x), Q.r), invW));
    int x1 = (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(Q.u, b.min_x), Q.r), invW));
    int y0 = (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(Q.v, b.min_y), Q.r), invH));
    int y1 = (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(Q.v, b.min_y), Q.r), invH));
    x0 = max(x0, 0); x1 = min(x1, HYORB_GRID_COLS - 1); y0 = max(y0, 0); y1 = min(y1, HYORB_GRID_ROWS - 1);
    uint32_t bkey = KEY_NONE; int second = DIST_NONE;
    int bestT = -1;
    uint32_t pos0 = 0;
    if (x0 < HYORB_GRID_COLS && x1 >= 0 && y0 < HYORB_GRID_ROWS && y1 >= 0) {
        for (int ix = x0; ix <= x1; ix++)
            for (int iy = y0; iy <= y1; iy++) {
                const int c = ix * HYORB_GRID_ROWS + iy;
                const int lo = cell_off[c], hi = cell_off[c + 1];
                for (int j = lo + lane; j < hi; j += 32) {
                    const int k = cell_idx[j];
                    const hyorb_keypoint kp = kps[k];
                    const float dx = __fsub_rn(kp.x, Q.u), dy = __fsub_rn(kp.y, Q.v);
                    if (!(fabsf(dx) < Q.r && fabsf(dy) < Q.r)) continue;
                    if (t_matched && t_matched[k]) continue;                          // PreviouslyMatchedCriterion
                    if (!(kp.size > Q.size_lo && kp.size < Q.size_hi)) continue;      // FeatureSizeCriterion
                    if (Q.ur_radius >= 0) {                                           // StereoConsistencyCriterion
                        const float ur = t_uR[k];
                        if (!(fabsf(__fsub_rn(Q.ur, ur)) < Q.ur_radius && ur > 0)) continue;
                    }
                    if (wc.reproj_thr >= 0.0f) {                                      // ProjectionViewCriterion (MatchCriteria.cpp:282-333, KeyFrame.cc:548-575)
                        const float ur_view = t_uR ? t_uR[k] : -1.0f;                 // FeatureViews::uR: -1 for monocular views
                        const float errX = __fsub_rn(Q.u, kp.x), errY = __fsub_rn(Q.v, kp.y);
                        const float errXr = ur_view >= 0.0f ? __fsub_rn(Q.ur, ur_view) : 0.0f;
                        const float sserr = __fadd_rn(__fadd_rn(__fmul_rn(errX, errX), __fmul_rn(errY, errY)), __fmul_rn(errXr, errXr));
                        const float sf = __fdiv_rn(kp.size, wc.size_ref);
                        const float sigma2 = __fmul_rn(wc.sigma_ref, __fmul_rn(sf, sf));   // FeatureExtractorSettings::determineSigma2
                        const float stereo_factor = ur_view > 0.0f ? 1.30f : 1.00f;
                        if (!(__fdiv_rn(sserr, sigma2) < __fmul_rn(stereo_factor, wc.reproj_thr))) continue;
                    }
                    const int d = hamming256(a0, a1, tdesc[2 * (size_t)k], tdesc[2 * (size_t)k + 1]);
                    const uint32_t pos = (pos0 + (uint32_t)(j - lo)) & POS_MASK;
                    const uint32_t key = ((uint32_t)d << POS_BITS) | pos;
                    if (key < bkey) bestT = k;
                    scan_update(bkey, second, d, pos);
                }
                pos0 += (uint32_t)(hi - lo);
            }
    }
    // merge lanes, carrying the target index of the winner
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t ok = __shfl_xor_sync(0xffffffffu, bkey, o);
        const int os = __shfl_xor_sync(0xffffffffu, second, o);
        const int ot = __shfl_xor_sync(0xffffffffu, bestT, o);
        if (ok < bkey) bestT = ot;
        scan_merge(bkey, second, ok, os);
    }
    if (lane == 0) write_result(qi, bkey, second, bestT, wc.rule, thr, ratio, best_idx, best, sec, accepted);
}

// ---------------- rotation histogram (single CTA)
__global__ void __launch_bounds__(256)
k_rotation(const float *__restrict__ a_prev, const float *__restrict__ a_curr, int n, uint8_t *__restrict__ keep, int *status)
{
    constexpr int HL = 30;
    __shared__ int hist[HL];
    __shared__ int ind[3];
    if (threadIdx.x < HL) hist[threadIdx.x] = 0;
    __syncthreads();
    const float factor = 1.0f / HL;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float rot = __fsub_rn(a_prev[i], a_curr[i]);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, factor));
        if (bin == HL) bin = 0;
        if (bin < 0 || bin >= HL) { atomicOr(status, ST_BAD_INDEX); continue; }   // the reference asserts
        atomicAdd(&hist[bin], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {   // ComputeThreeMaxima (:727-767)
        int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
        for (int i = 0; i < HL; i++) {
            const int s = hist[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; i3 = i2; i2 = i1; i1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; i3 = i2; i2 = i; }
            else if (s > max3) { max3 = s; i3 = i; }
        }
        if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { i2 = -1; i3 = -1; }
        else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { i3 = -1; }
        ind[0] = i1; ind[1] = i2; ind[2] = i3;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float rot = __fsub_rn(a_prev[i], a_curr[i]);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, factor));
        if (bin == HL) bin = 0;
        keep[i] = (uint8_t)(bin == ind[0] || bin == ind[1] || bin == ind[2]);
    }
}

// ---------------- launchers
int launch_match_bruteforce(const uint8_t *q, int nq, const uint8_t *t, int nt, int rule, float thr, float ratio,
                            int32_t *best_idx, uint16_t *best, uint16_t *second, uint8_t *accepted,
                            uint32_t *pkey, uint16_t *psecond, int nsplit, cudaStream_t st, long *launches)
{
    if (nq <= 0) return HYORB_OK;
    if (nt >= (1 << POS_BITS)) { set_error("more than %d targets", (1 << POS_BITS) - 1); return HYORB_EUNSUPPORTED; }
    const int chunk = nt > 0 ? (nt + nsplit - 1) / nsplit : 1;
    dim3 grd((nq + BF_THREADS * BF_QPT - 1) / (BF_THREADS * BF_QPT), nsplit);
    k_bf_partial<<<grd, BF_THREADS, 0, st>>>((const uint4 *)q, nq, (const uint4 *)t, nt, chunk, pkey, psecond);
    k_bf_finalize<<<(nq + 255) / 256, 256, 0, st>>>(pkey, psecond, nsplit, nq, rule, thr, ratio, best_idx, best, second, accepted);
    *launches += 2;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

int launch_match_csr(const uint8_t *q, int nq, const uint8_t *t, int nt, const int32_t *off, const int32_t *idx, int rule, float thr,
                     float ratio, int32_t *best_idx, uint16_t *best, uint16_t *second, uint8_t *accepted, int *status, const EpipolarDev &epi,
                     cudaStream_t st, long *launches)
{
    if (nq <= 0) return HYORB_OK;
    k_match_csr<<<(nq + 7) / 8, 256, 0, st>>>((const uint4 *)q, nq, (const uint4 *)t, nt, off, idx, rule, thr, ratio, best_idx, best, second, accepted, status, epi);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

int launch_grid_build(const hyorb_keypoint *kps, int n, hyorb_bounds b, int32_t *cell_off, int32_t *cell_idx, int32_t *cell_of, int32_t *cell_cnt,
                      cudaStream_t st, long *launches)
{
    HY_CUDA(cudaMemsetAsync(cell_cnt, 0, sizeof(int32_t) * HYORB_GRID_COLS * HYORB_GRID_ROWS, st));
    if (n > 0) { k_grid_cells<<<(n + 255) / 256, 256, 0, st>>>(kps, n, b, cell_of, cell_cnt); ++*launches; }
    k_grid_fill<<<1, 1024, 0, st>>>(cell_of, n, cell_cnt, cell_off, cell_idx);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

int launch_match_window(const hyorb_keypoint *kps, const uint8_t *tdesc, const float *t_uR, const uint8_t *t_matched, int nt, hyorb_bounds b,
                        const int32_t *cell_off, const int32_t *cell_idx, const hyorb_window_query *q, const uint8_t *qdesc, int nq,
                        float thr, float ratio, int32_t *best_idx, uint16_t *best, uint16_t *second, uint8_t *accepted, cudaStream_t st, long *launches,
                        const uint8_t *q_active, WindowCriteria wc)
{
    if (nq <= 0) return HYORB_OK;
    k_match_window<<<(nq + 7) / 8, 256, 0, st>>>(kps, (const uint4 *)tdesc, t_uR, t_matched, nt, b, cell_off, cell_idx, q, (const uint4 *)qdesc, nq,
                                                 thr, ratio, q_active, best_idx, best, second, accepted, wc);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

// ---------------- landmark projection: the per-landmark front half of FeatureMatcher::SearchByProjection(Frame&, landmarks, th)
// (FeatureMatcher.cc:123-143, 57-121).  One thread per landmark: ProjectionCriterion + DistanceCriterion (MatchCriteria.cpp:13-28,
// 46-77), Frame::ProjectLandMark / Camera::Project (Frame.cc:176-180, Camera.cpp:116-153), Frame::landMarkSizePixels (Frame.cc:296-317)
// and the window / size / stereo bounds of the view criteria.  fp32 policy of the reference's cv::Mat expressions (pinned on the CPU
// against cv2.gemm / cv2.norm): 3x3 * 3x1 = OpenCV's small-matrix path, t = (a0*b0 + a1*b1) + a2*b2 in float (no FMA), + C rounded
// once; cv::norm sums squares in double.
__device__ __forceinline__ void gemm31(const float *A, const float *b, const float *c, float *out)
{
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float t = __fadd_rn(__fadd_rn(__fmul_rn(A[3 * i], b[0]), __fmul_rn(A[3 * i + 1], b[1])), __fmul_rn(A[3 * i + 2], b[2]));
        out[i] = c ? __fadd_rn(t, c[i]) : t;      // (float)((double)t + (double)c) == the correctly rounded float sum
    }
}
__device__ __forceinline__ bool project_point(const hyorb_projection &pr, const float *Pw, float *uv)
{
    float Pc[3], Pch[3];
    gemm31(pr.Rcw, Pw, pr.tcw, Pc);
    const float PcZ = Pc[2];
    const float invz = __fdiv_rn(1.0f, PcZ);
    Pch[0] = __fdiv_rn(Pc[0], PcZ); Pch[1] = __fdiv_rn(Pc[1], PcZ); Pch[2] = __fdiv_rn(Pc[2], PcZ);
    gemm31(pr.K, Pch, nullptr, uv);
    const float u = uv[0], v = uv[1];
    uv[2] = pr.stereo ? __fsub_rn(u, __fmul_rn(pr.mbf, invz)) : -1.0f;
    return PcZ > 0.0f && u >= pr.bounds.min_x && u <= pr.bounds.max_x && v >= pr.bounds.min_y && v <= pr.bounds.max_y;
}

__global__ void __launch_bounds__(128)
k_project_landmarks(hyorb_projection pr, const hyorb_landmark *__restrict__ lms, int n, const hyorb_keypoint *__restrict__ t_kps, int nt,
                    float th, float size_ref, float frac_smaller, float frac_larger, unsigned flags, hyorb_window_query *__restrict__ queries,
                    uint8_t *__restrict__ passed, int *status, const float *__restrict__ normals, float cos_max_angle)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const hyorb_landmark lm = lms[i];
    float uv[3];
    const bool valid = project_point(pr, lm.Pw, uv);
    const float PO[3] = {__fsub_rn(lm.Pw[0], pr.Ow[0]), __fsub_rn(lm.Pw[1], pr.Ow[1]), __fsub_rn(lm.Pw[2], pr.Ow[2])};
    const double sq = __dadd_rn(__dadd_rn(__dmul_rn((double)PO[0], (double)PO[0]), __dmul_rn((double)PO[1], (double)PO[1])), __dmul_rn((double)PO[2], (double)PO[2]));
    const float dist = (float)__dsqrt_rn(sq);
    bool dist_ok = !(flags & HYORB_SBP_DISTANCE) || !(dist < lm.min_dist || dist > lm.max_dist);
    if (flags & HYORB_SBP_VIEWANGLE) {
        // ViewingAngleCriterionCore (MatchCriteria.cpp:94-110): PO / d is cv::Mat::convertTo with alpha = 1/d (a float multiply by
        // (float)(1.0 / d)); Mat::dot accumulates in double; the bound is cosf(max_angle), evaluated by the host's libm
        const float inv = (float)__ddiv_rn(1.0, (double)dist);
        double dot = 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) dot = __dadd_rn(dot, __dmul_rn((double)__fmul_rn(PO[k], inv), (double)normals[3 * i + k]));
        dist_ok = dist_ok && (dot > (double)cos_max_angle);
    }
    float size_px;
    if (lm.assoc_idx >= 0) {
        if (lm.assoc_idx >= nt) { atomicOr(status, ST_BAD_INDEX); size_px = 0.f; }
        else size_px = t_kps[lm.assoc_idx].size;
    } else {
        const float half = __fdiv_rn(lm.size, 2.0f);
        const float L[3] = {__fsub_rn(lm.Pw[0], half), lm.Pw[1], lm.Pw[2]}, R[3] = {__fadd_rn(lm.Pw[0], half), lm.Pw[1], lm.Pw[2]};
        float ul[3], ur[3];
        project_point(pr, L, ul); project_point(pr, R, ur);
        size_px = __fsub_rn(ur[0], ul[0]);
    }
    const float radius = __fdiv_rn(__fmul_rn(th, size_px), size_ref);
    hyorb_window_query q;
    q.u = uv[0]; q.v = uv[1]; q.r = radius;
    q.size_lo = __fmul_rn(frac_smaller, size_px); q.size_hi = __fmul_rn(frac_larger, size_px);
    q.ur = uv[2]; q.ur_radius = (pr.stereo && (flags & HYORB_SBP_STEREO)) ? radius : -1.0f;
    queries[i] = q;
    passed[i] = (uint8_t)(valid && dist_ok);
}

// RotationConsistencyCriterion over the matches of a projection search (MatchCriteria.cpp:363-401, 684-767), single CTA.  The
// reference re-keys the matches by current keypoint index before the histogram (:382): of several landmarks that matched the
// same keypoint only the last one in its map order survives -- here: the one listed last by the caller.
__global__ void __launch_bounds__(256)
k_projection_rotation(const int32_t *__restrict__ best_idx, uint8_t *__restrict__ accepted, int n, const float *__restrict__ prev_angle,
                      const hyorb_keypoint *__restrict__ t_kps, int nt, int32_t *__restrict__ owner, int *status)
{
    constexpr int HL = 30;
    __shared__ int hist[HL];
    __shared__ int ind[3];
    if (threadIdx.x < HL) hist[threadIdx.x] = 0;
    for (int k = threadIdx.x; k < nt; k += blockDim.x) owner[k] = -1;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (accepted[i]) atomicMax(&owner[best_idx[i]], i);
    __syncthreads();
    const float factor = 1.0f / HL;
    auto bin_of = [&](int i) {
        float rot = __fsub_rn(prev_angle[i], t_kps[best_idx[i]].angle);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, factor));
        return bin == HL ? 0 : bin;
    };
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (!accepted[i] || owner[best_idx[i]] != i) continue;
        const int bin = bin_of(i);
        if (bin < 0 || bin >= HL) { atomicOr(status, ST_BAD_INDEX); continue; }   // the reference asserts
        atomicAdd(&hist[bin], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {   // ComputeThreeMaxima (:727-767)
        int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
        for (int i = 0; i < HL; i++) {
            const int s = hist[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; i3 = i2; i2 = i1; i1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; i3 = i2; i2 = i; }
            else if (s > max3) { max3 = s; i3 = i; }
        }
        if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { i2 = -1; i3 = -1; }
        else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { i3 = -1; }
        ind[0] = i1; ind[1] = i2; ind[2] = i3;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (!accepted[i]) continue;
        if (owner[best_idx[i]] != i) { accepted[i] = 0; continue; }
        const int bin = bin_of(i);
        accepted[i] = (uint8_t)(bin == ind[0] || bin == ind[1] || bin == ind[2]);
    }
}

int launch_projection_rotation(const int32_t *best_idx, uint8_t *accepted, int n, const float *prev_angle, const hyorb_keypoint *t_kps, int nt,
                               int32_t *owner, int *status, cudaStream_t st, long *launches)
{
    if (n <= 0) return HYORB_OK;
    k_projection_rotation<<<1, 256, 0, st>>>(best_idx, accepted, n, prev_angle, t_kps, nt, owner, status);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

int launch_project_landmarks(const hyorb_projection &pr, const hyorb_landmark *lms, int n, const hyorb_keypoint *t_kps, int nt, float th, float size_ref,
                             float frac_smaller, float frac_larger, unsigned flags, hyorb_window_query *queries, uint8_t *passed, int *status, cudaStream_t st,
                             long *launches, const float *normals, float cos_max_angle)
{
    if (n <= 0) return HYORB_OK;
    k_project_landmarks<<<(n + 127) / 128, 128, 0, st>>>(pr, lms, n, t_kps, nt, th, size_ref, frac_smaller, frac_larger, flags, queries, passed, status,
                                                         normals, cos_max_angle);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

// ---------------- one direction of FeatureMatcher::SearchBySim3 (FeatureMatcher.cc:783-845 / 848-910): landmark of keyframe A -> camera A
// (R_a * Pw + t_a) -> camera B through the similarity (sR_ba * p + t_ba) -> Camera::Project of B; the 3-D distance in B must lie in the
// landmark's scale-invariance range; radius = th * B.landMarkSizePixels(lm) / size_ref with B's OWN pose.  No view criteria follow.
struct Sim3Dev { float R_a[9], t_a[3], sR_ba[9], t_ba[3]; };
__global__ void __launch_bounds__(128)
k_project_sim3(Sim3Dev T, hyorb_projection prb, const hyorb_landmark *__restrict__ lms, int n, const hyorb_keypoint *__restrict__ kps_b, int nb, float th,
               float size_ref, hyorb_window_query *__restrict__ queries, uint8_t *__restrict__ passed, int *status)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const hyorb_landmark lm = lms[i];
    float pa[3], pb[3], Pch[3], uv[3];
    gemm31(T.R_a, lm.Pw, T.t_a, pa);
    gemm31(T.sR_ba, pa, T.t_ba, pb);
    const float z = pb[2], invz = __fdiv_rn(1.0f, z);
    Pch[0] = __fdiv_rn(pb[0], z); Pch[1] = __fdiv_rn(pb[1], z); Pch[2] = __fdiv_rn(pb[2], z);
    gemm31(prb.K, Pch, nullptr, uv);
    uv[2] = prb.stereo ? __fsub_rn(uv[0], __fmul_rn(prb.mbf, invz)) : -1.0f;
    const bool valid = z > 0.0f && uv[0] >= prb.bounds.min_x && uv[0] <= prb.bounds.max_x && uv[1] >= prb.bounds.min_y && uv[1] <= prb.bounds.max_y;
    const double sq = __dadd_rn(__dadd_rn(__dmul_rn((double)pb[0], (double)pb[0]), __dmul_rn((double)pb[1], (double)pb[1])), __dmul_rn((double)pb[2], (double)pb[2]));
    const float dist = (float)__dsqrt_rn(sq);
    const bool dist_ok = !(dist < lm.min_dist || dist > lm.max_dist);
    float size_px;
    if (lm.assoc_idx >= 0) {
        if (lm.assoc_idx >= nb) { atomicOr(status, ST_BAD_INDEX); size_px = 0.f; }
        else size_px = kps_b[lm.assoc_idx].size;
    } else {
        const float half = __fdiv_rn(lm.size, 2.0f);
        const float L[3] = {__fsub_rn(lm.Pw[0], half), lm.Pw[1], lm.Pw[2]}, R[3] = {__fadd_rn(lm.Pw[0], half), lm.Pw[1], lm.Pw[2]};
        float ul[3], ur[3];
        project_point(prb, L, ul); project_point(prb, R, ur);
        size_px = __fsub_rn(ur[0], ul[0]);
    }
    hyorb_window_query q;
    q.u = uv[0]; q.v = uv[1]; q.r = __fdiv_rn(__fmul_rn(th, size_px), size_ref);
    q.size_lo = -FLT_MAX; q.size_hi = FLT_MAX; q.ur = uv[2]; q.ur_radius = -1.0f;
    queries[i] = q;
    passed[i] = (uint8_t)(valid && dist_ok);
}

int launch_project_sim3(const float *R_a, const float *t_a, const float *sR_ba, const float *t_ba, const hyorb_projection &prb, const hyorb_landmark *lms, int n,
                        const hyorb_keypoint *kps_b, int nb, float th, float size_ref, hyorb_window_query *queries, uint8_t *passed, int *status,
                        cudaStream_t st, long *launches)
{
    if (n <= 0) return HYORB_OK;
    Sim3Dev T;
    for (int i = 0; i < 9; i++) { T.R_a[i] = R_a[i]; T.sR_ba[i] = sR_ba[i]; }
    for (int i = 0; i < 3; i++) { T.t_a[i] = t_a[i]; T.t_ba[i] = t_ba[i]; }
    k_project_sim3<<<(n + 127) / 128, 128, 0, st>>>(T, prb, lms, n, kps_b, nb, th, size_ref, queries, passed, status);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

// ---------------- FeatureMatcher::SearchForInitialization (FeatureMatcher.cc:404-462) with MonoInitScoreExceedsPrevious + MonoInitBestScore
// (MatchCriteria.cpp:486-549).  The reference walks the frame-1 features in order; feature i1 may only consider a frame-2 feature i2 that an
// earlier i1' matched if its own distance is smaller than that match's.  The decision of i1 is therefore a function of the decisions of all
// i1' < i1 -- a triangular system with exactly one solution.  It is solved by fixed-point iteration: every pass recomputes ALL decisions in
// parallel from the previous pass's claims (per i2: the smallest distance claimed by an EARLIER feature); after pass k the first k features are
// final, and in practice the dependency chains are a handful long.  The host stops at the first pass that changes nothing.
//   claim2[i1] / claimd[i1]: frame-2 feature and distance i1 currently claims (-1: none); head[i2] / next[i1]: per-i2 lists of claimers.
__global__ void k_mono_lists(const int32_t *__restrict__ claim2, int n1, int32_t *__restrict__ head, int32_t *__restrict__ next)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const int j = claim2[i];
    next[i] = j >= 0 ? atomicExch(&head[j], i) : -1;
}

__global__ void __launch_bounds__(256)
k_mono_pass(const uint4 *__restrict__ d1, int n1, const hyorb_keypoint *__restrict__ k2, const uint4 *__restrict__ d2, hyorb_bounds b,
            const int32_t *__restrict__ cell_off, const int32_t *__restrict__ cell_idx, const float *__restrict__ prev_xy, float r, float thr, float ratio,
            const int32_t *__restrict__ claim2, const int32_t *__restrict__ claimd, const int32_t *__restrict__ head, const int32_t *__restrict__ next,
            int32_t *__restrict__ out2, int32_t *__restrict__ outd, int *__restrict__ changed)
{
    const int lane = threadIdx.x & 31;
    const int i1 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i1 >= n1) return;
    const float qu = prev_xy[2 * i1], qv = prev_xy[2 * i1 + 1];
    const uint4 a0 = d1[2 * i1], a1 = d1[2 * i1 + 1];
    const float invW = __fdiv_rn((float)HYORB_GRID_COLS, __fsub_rn(b.max_x, b.min_x));
    const float invH = __fdiv_rn((float)HYORB_GRID_ROWS, __fsub_rn(b.max_y, b.min_y));
    int x0 = (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(qu, b.min_x), r), invW));
    int x1 = (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(qu, b.min_x), r), invW));
    int y0 = (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(qv, b.min_y), r), invH));
    int y1 = (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(qv, b.min_y), r), invH));
    x0 = max(x0, 0); x1 = min(x1, HYORB_GRID_COLS - 1); y0 = max(y0, 0); y1 = min(y1, HYORB_GRID_ROWS - 1);
    uint32_t bkey = KEY_NONE; int second = DIST_NONE;
    int bestT = -1;
    uint32_t pos0 = 0;
    if (x0 < HYORB_GRID_COLS && x1 >= 0 && y0 < HYORB_GRID_ROWS && y1 >= 0) {
        for (int ix = x0; ix <= x1; ix++)
            for (int iy = y0; iy <= y1; iy++) {
                const int c = ix * HYORB_GRID_ROWS + iy;
                const int lo = cell_off[c], hi = cell_off[c + 1];
                for (int j = lo + lane; j < hi; j += 32) {
                    const int k = cell_idx[j];
                    const hyorb_keypoint kp = k2[k];
                    if (!(fabsf(__fsub_rn(kp.x, qu)) < r && fabsf(__fsub_rn(kp.y, qv)) < r)) continue;
                    const int d = hamming256(a0, a1, d2[2 * (size_t)k], d2[2 * (size_t)k + 1]);
                    // MonoInitScoreExceedsPrevious: the smallest distance an EARLIER feature claims on k (claims on one i2 decrease with i1)
                    int dprev = DIST_NONE;
                    for (int e = head[k]; e >= 0; e = next[e])
                        if (e < i1) dprev = min(dprev, claimd[e]);
                    if (dprev != DIST_NONE && !(d < dprev)) continue;
                    const uint32_t pos = (pos0 + (uint32_t)(j - lo)) & POS_MASK;
                    const uint32_t key = ((uint32_t)d << POS_BITS) | pos;
                    if (key < bkey) bestT = k;
                    scan_update(bkey, second, d, pos);
                }
                pos0 += (uint32_t)(hi - lo);
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t ok = __shfl_xor_sync(0xffffffffu, bkey, o);
        const int os = __shfl_xor_sync(0xffffffffu, second, o);
        const int ot = __shfl_xor_sync(0xffffffffu, bestT, o);
        if (ok < bkey) bestT = ot;
        scan_merge(bkey, second, ok, os);
    }
    if (lane == 0) {
        const int bd = (int)(bkey >> POS_BITS);
        const bool has = bd < DIST_NONE;
        const bool acc = has && accept_rule(HYORB_RULE_MONOINIT, (float)bd, second < DIST_NONE ? (float)second : FLT_MAX, thr, ratio);
        const int c2 = acc ? bestT : -1, cd = acc ? bd : -1;
        if (c2 != claim2[i1] || cd != claimd[i1]) *changed = 1;
        out2[i1] = c2; outd[i1] = cd;
    }
}

// owner of every claimed frame-2 feature = the last claimer (matches[idx2] = i1 overwrites), RotationConsistency keyed by idx2
// (rot = angle(frame 1) - angle(frame 2)), then matches12 / vbPrevMatched (FeatureMatcher.cc:441-458).  Single CTA.
__global__ void __launch_bounds__(256)
k_mono_finish(const int32_t *__restrict__ claim2, int n1, const hyorb_keypoint *__restrict__ k1, const hyorb_keypoint *__restrict__ k2, int n2,
              int32_t *__restrict__ owner, int32_t *__restrict__ matches12, float *__restrict__ prev_xy, int *__restrict__ n_matches, int *status)
{
    constexpr int HL = 30;
    __shared__ int hist[HL];
    __shared__ int ind[3];
    __shared__ int s_n;
    if (threadIdx.x < HL) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_n = 0;
    for (int j = threadIdx.x; j < n2; j += blockDim.x) owner[j] = -1;
    for (int i = threadIdx.x; i < n1; i += blockDim.x) matches12[i] = -1;
    __syncthreads();
    for (int i = threadIdx.x; i < n1; i += blockDim.x)
        if (claim2[i] >= 0) atomicMax(&owner[claim2[i]], i);
    __syncthreads();
    const float factor = 1.0f / HL;
    auto bin_of = [&](int j) {
        float rot = __fsub_rn(k1[owner[j]].angle, k2[j].angle);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, factor));
        return bin == HL ? 0 : bin;
    };
    for (int j = threadIdx.x; j < n2; j += blockDim.x) {
        if (owner[j] < 0) continue;
        const int bin = bin_of(j);
        if (bin < 0 || bin >= HL) { atomicOr(status, ST_BAD_INDEX); continue; }
        atomicAdd(&hist[bin], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
        for (int i = 0; i < HL; i++) {
            const int s = hist[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; i3 = i2; i2 = i1; i1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; i3 = i2; i2 = i; }
            else if (s > max3) { max3 = s; i3 = i; }
        }
        if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { i2 = -1; i3 = -1; }
        else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { i3 = -1; }
        ind[0] = i1; ind[1] = i2; ind[2] = i3;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n2; j += blockDim.x) {
        if (owner[j] < 0) continue;
        const int bin = bin_of(j);
        if (bin == ind[0] || bin == ind[1] || bin == ind[2]) {      // a feature of frame 1 owns at most one feature of frame 2
            const int i = owner[j];
            matches12[i] = j;
            prev_xy[2 * i] = k2[j].x; prev_xy[2 * i + 1] = k2[j].y;
            atomicAdd(&s_n, 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *n_matches = s_n;
}

int launch_mono_pass(const uint8_t *d1, int n1, const hyorb_keypoint *k2, const uint8_t *d2, int n2, hyorb_bounds b, const int32_t *cell_off, const int32_t *cell_idx,
                     const float *prev_xy, float r, float thr, float ratio, const int32_t *claim2, const int32_t *claimd, int32_t *head, int32_t *next,
                     int32_t *out2, int32_t *outd, int *changed, cudaStream_t st, long *launches)
{
    HY_CUDA(cudaMemsetAsync(head, 0xFF, sizeof(int32_t) * (size_t)std::max(n2, 1), st));
    HY_CUDA(cudaMemsetAsync(changed, 0, sizeof(int), st));
    k_mono_lists<<<(n1 + 255) / 256, 256, 0, st>>>(claim2, n1, head, next);
    k_mono_pass<<<(n1 + 7) / 8, 256, 0, st>>>((const uint4 *)d1, n1, k2, (const uint4 *)d2, b, cell_off, cell_idx, prev_xy, r, thr, ratio, claim2, claimd, head, next,
                                             out2, outd, changed);
    *launches += 2;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}
int launch_mono_finish(const int32_t *claim2, int n1, const hyorb_keypoint *k1, const hyorb_keypoint *k2, int n2, int32_t *owner, int32_t *matches12,
                       float *prev_xy, int *n_matches, int *status, cudaStream_t st, long *launches)
{
    k_mono_finish<<<1, 256, 0, st>>>(claim2, n1, k1, k2, n2, owner, matches12, prev_xy, n_matches, status);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

// ---------------- landmark representative descriptor (MapPointDBEntry::_computeDistinctiveDescriptor_, src/core/MapPointDB.cpp:127-171)
// One warp per landmark.  For every observation i the warp histograms the Hamming distances to all observations (self
// distance 0) into 257 shared-memory bins and reads the order statistic sorted[(int)(0.5*(N-1))] off a warp scan of the
// bins -- the reference sorts the row; the first row with the smallest median wins (strict <, :163).
constexpr int DD_WARPS = 4, DD_BINS = 288;      // 257 bins rounded up to 9 per lane

__global__ void __launch_bounds__(DD_WARPS * 32)
k_distinctive(const uint4 *__restrict__ desc, const int32_t *__restrict__ off, int n_lm, int32_t *__restrict__ best_idx, int32_t *__restrict__ best_median)
{
    __shared__ int s_hist[DD_WARPS][DD_BINS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int lm = blockIdx.x * DD_WARPS + w;
    if (lm >= n_lm) return;
    const int lo = off[lm], N = off[lm + 1] - lo;
    if (N <= 0) { if (lane == 0) { best_idx[lm] = -1; best_median[lm] = -1; } return; }
    const int k = (N - 1) >> 1;                  // (int)(0.5*(N-1))
    int *hist = s_hist[w];
    int bestMed = 0x7fffffff, bestI = 0;
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int t = 0; t < DD_BINS / 32; t++) hist[lane * (DD_BINS / 32) + t] = 0;
        __syncwarp();
        const uint4 a0 = desc[2 * (size_t)(lo + i)], a1 = desc[2 * (size_t)(lo + i) + 1];
        for (int j = lane; j < N; j += 32) {
            const int d = j == i ? 0 : hamming256(a0, a1, desc[2 * (size_t)(lo + j)], desc[2 * (size_t)(lo + j) + 1]);
            atomicAdd(&hist[d], 1);
        }
        __syncwarp();
        // lane owns bins [9 lane, 9 lane + 9): chunk sums -> inclusive warp scan -> the chunk where the running count reaches k+1
        int c = 0;
#pragma unroll
        for (int t = 0; t < DD_BINS / 32; t++) c += hist[lane * (DD_BINS / 32) + t];
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        const unsigned reach = __ballot_sync(0xffffffffu, inc >= k + 1);
        const int L = __ffs(reach) - 1;          // exists: the counts sum to N > k
        int med = 0;
        if (lane == L) {
            int run = inc - c;
#pragma unroll
            for (int t = 0; t < DD_BINS / 32; t++) {
                run += hist[lane * (DD_BINS / 32) + t];
                if (run >= k + 1) { med = lane * (DD_BINS / 32) + t; break; }
            }
        }
        med = __shfl_sync(0xffffffffu, med, L);
        if (med < bestMed) { bestMed = med; bestI = i; }
        __syncwarp();
    }
    if (lane == 0) { best_idx[lm] = bestI; best_median[lm] = bestMed; }
}

int launch_distinctive(const uint8_t *desc, const int32_t *off, int n_lm, int32_t *best_idx, int32_t *best_median, cudaStream_t st, long *launches)
{
    if (n_lm <= 0) return HYORB_OK;
    k_distinctive<<<(n_lm + DD_WARPS - 1) / DD_WARPS, DD_WARPS * 32, 0, st>>>((const uint4 *)desc, off, n_lm, best_idx, best_median);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

// ---------------- bag-of-words quantisation and BoW-gated matching
// DBoW2::TemplatedVocabulary::transform as called by ORBVocabulary::transform (src/features/low_level/ORBVocabulary.cpp:31-42):
// one warp per feature descends the vocabulary tree; at every level the lanes take one child each (k <= 32 per pass), Hamming
// distance to the child's descriptor, warp argmin with the child's list position as tie-break (DBoW2 keeps the first minimum).
__global__ void __launch_bounds__(256)
k_bow_descend(const int32_t *__restrict__ child_off, const int32_t *__restrict__ child_idx, const uint4 *__restrict__ node_desc,
              const int32_t *__restrict__ word_of, const float *__restrict__ weight_of, int nid_level, const uint4 *__restrict__ desc, int n,
              int32_t *__restrict__ word_id, int32_t *__restrict__ node_id, float *__restrict__ weight)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const uint4 a0 = desc[2 * (size_t)i], a1 = desc[2 * (size_t)i + 1];
    int node = 0, level = 0, nid = nid_level <= 0 ? 0 : -1;
    while (true) {
        const int lo = child_off[node], hi = child_off[node + 1];
        if (hi <= lo || level > 64) break;
        ++level;
        uint32_t key = 0xFFFFFFFFu;
        for (int c = lo + lane; c < hi; c += 32) {
            const int id = child_idx[c];
            const uint32_t d = (uint32_t)hamming256(a0, a1, node_desc[2 * (size_t)id], node_desc[2 * (size_t)id + 1]);
            key = min(key, (d << 20) | (uint32_t)(c - lo));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
        node = child_idx[lo + (int)(key & 0xFFFFFu)];
        if (level == nid_level) nid = node;
    }
    if (lane == 0) { word_id[i] = word_of[node]; weight[i] = weight_of[node]; node_id[i] = nid >= 0 ? nid : node; }
}

// candidate ranges of BoW-gated matching: features of keyframe 2 sorted by (node, index) -> for every feature of keyframe 1 the
// run of keyframe-2 features under the same node (FeatureMatcher::_SearchByBoW_'s merge walk, FeatureMatcher.cc:281-345)
__global__ void k_bow_ranges(const int32_t *__restrict__ node1, int n1, const int32_t *__restrict__ sorted_node2, int n2,
                             int32_t *__restrict__ cbegin, int32_t *__restrict__ cend)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const int key = node1[i];
    int lo = 0, hi = n2;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted_node2[mid] < key) lo = mid + 1; else hi = mid; }
    const int b = lo;
    hi = n2;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted_node2[mid] <= key) lo = mid + 1; else hi = mid; }
    cbegin[i] = b; cend[i] = lo;
}
__global__ void k_iota(int32_t *v, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) v[i] = i; }

// one warp per query over the candidate run idx[cbegin[q] .. cend[q]) (shared by all queries under the same node); mask1 / mask2:
// the BoWIndexCriterion filters (0 = not eligible), may be NULL
__global__ void __launch_bounds__(256)
k_match_ranges(const uint4 *__restrict__ q, const uint8_t *__restrict__ mask1, int nq, const uint4 *__restrict__ t, const uint8_t *__restrict__ mask2,
               const int32_t *__restrict__ cbegin, const int32_t *__restrict__ cend, const int32_t *__restrict__ idx, int rule, float thr,
               float ratio, int32_t *best_idx, uint16_t *best, uint16_t *sec, uint8_t *accepted, const EpipolarDev epi)
{
    const int lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= nq) return;
    uint32_t bkey = KEY_NONE; int second = DIST_NONE;
    const int lo = cbegin[qi], hi = (mask1 && !mask1[qi]) ? lo : cend[qi];
    if (hi > lo) {
        const uint4 a0 = q[2 * qi], a1 = q[2 * qi + 1];
        EpiLine el{0.f, 0.f, 0.f, 0.f};
        if (epi.kps1) el = epi_line(epi, qi);
        for (int c = lo + lane; c < hi; c += 32) {
            const int ti = idx[c];
            if (mask2 && !mask2[ti]) continue;
            if (epi.kps1 && !epi_pass(epi, el, ti)) continue;
            const int d = hamming256(a0, a1, t[2 * (size_t)ti], t[2 * (size_t)ti + 1]);
            scan_update(bkey, second, d, (uint32_t)(c - lo) & POS_MASK);
        }
    }
    warp_merge(bkey, second);
    if (lane == 0) {
        const int bt = bkey != KEY_NONE ? idx[lo + (int)(bkey & POS_MASK)] : -1;
        write_result(qi, bkey, second, bt, rule, thr, ratio, best_idx, best, sec, accepted);
    }
}

int launch_bow_descend(const int32_t *child_off, const int32_t *child_idx, const uint8_t *node_desc, const int32_t *word_of, const float *weight_of,
                       int nid_level, const uint8_t *desc, int n, int32_t *word_id, int32_t *node_id, float *weight, cudaStream_t st, long *launches)
{
    if (n <= 0) return HYORB_OK;
    k_bow_descend<<<(n + 7) / 8, 256, 0, st>>>(child_off, child_idx, (const uint4 *)node_desc, word_of, weight_of, nid_level, (const uint4 *)desc, n,
                                               word_id, node_id, weight);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

size_t bow_sort_temp_bytes(int n)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t *)nullptr, (int32_t *)nullptr, (const int32_t *)nullptr, (int32_t *)nullptr, n);
    return bytes;
}

// node2 -> (sorted_node2, sorted_idx2) by a stable radix sort (index order inside a node, as FeatureVector::addFeature appends), then ranges + scan
int launch_bow_match(const uint8_t *desc1, const uint8_t *mask1, const int32_t *node1, int n1, const uint8_t *desc2, const uint8_t *mask2,
                     const int32_t *node2, int n2, int32_t *iota, int32_t *sorted_node2, int32_t *sorted_idx2, void *temp, size_t temp_bytes,
                     int32_t *cbegin, int32_t *cend, int rule, float thr, float ratio, int32_t *best_idx, uint16_t *best, uint16_t *second,
                     uint8_t *accepted, const EpipolarDev &epi, cudaStream_t st, long *launches)
{
    if (n1 <= 0) return HYORB_OK;
    if (n2 > 0) {
        k_iota<<<(n2 + 255) / 256, 256, 0, st>>>(iota, n2);
        HY_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, node2, sorted_node2, iota, sorted_idx2, n2, 0, 32, st));
        *launches += 2;
    }
    k_bow_ranges<<<(n1 + 255) / 256, 256, 0, st>>>(node1, n1, sorted_node2, n2, cbegin, cend);
    k_match_ranges<<<(n1 + 7) / 8, 256, 0, st>>>((const uint4 *)desc1, mask1, n1, (const uint4 *)desc2, mask2, cbegin, cend, sorted_idx2, rule, thr, ratio,
                                                 best_idx, best, second, accepted, epi);
    *launches += 2;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

int launch_rotation(const float *a_prev, const float *a_curr, int n, uint8_t *keep, int *status, cudaStream_t st, long *launches)
{
    if (n <= 0) return HYORB_OK;
    k_rotation<<<1, 256, 0, st>>>(a_prev, a_curr, n, keep, status);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

}  // namespace hyorb
