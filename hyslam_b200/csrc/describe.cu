// describe.cu -- K5: orientation + rotated BRIEF + final keypoint packaging.
// Replaces ORBFinder::compute (src/features/low_level/ORBFinder.cpp:70-78): intensityCentroidAngle (:16-43) and
// computeOrbDescriptor (:89-129), both evaluated on the BLURRED level (ORBExtractor.cpp:536-541), and
// the tail of ORBExtractor::operator() (:546-561): pt *= scale[level], levels concatenated 0..L-1.
// One warp per keypoint: lanes 0..30 take one row of the intensity-centroid disc each (integer moments, exact in any
// order: masked DP4A against the column weights), then lane i evaluates the 8 tests of descriptor byte i.  Floating point follows the reference's
// x86-64 baseline build: every fp32 operation individually rounded (no FMA contraction), cos/sin in double then
// narrowed, cvRound = round-half-even.
#include <float.h>
#include <mutex>
#include "common.cuh"

namespace hyorb {

__device__ const int8_t d_brief_pattern[1024] = {
#include "../../include/hyorb_brief_pattern.inc"
};
// the same table as floats (x0, y0, x1, y1 per test), filled once per device by k_pattern_to_float so that a lane fetches
// its 8 tests with 8 x 128-bit loads and no int->float conversions.  Lane i owns descriptor byte i = tests 8i .. 8i+7; the
// table is stored test-slot major ([slot t][lane i]) so that one warp-wide load reads 512 contiguous bytes (4 L1
// wavefronts) instead of 32 different cache lines.
__device__ float4 d_brief_pattern_f[256];
__global__ void k_pattern_to_float()
{
    const int i = threadIdx.x;                 // test index 8*lane + slot
    const int lane = i >> 3, slot = i & 7;
    d_brief_pattern_f[slot * 32 + lane] = make_float4((float)d_brief_pattern[4 * i], (float)d_brief_pattern[4 * i + 1], (float)d_brief_pattern[4 * i + 2],
                                                      (float)d_brief_pattern[4 * i + 3]);
}
// cvRound for |v| < 2^22: adding 1.5 * 2^23 leaves round-half-even(v) in the low mantissa bits (one FADD + one IADD on the
// full-rate pipes instead of an F2I on the quarter-rate one)
__device__ __forceinline__ int cv_round_small(float v)
{
    return __float_as_int(__fadd_rn(v, 12582912.f)) - 0x4B400000;
}

constexpr int DS_WARPS = 8;
#ifndef HYORB_DS_MINB
#define HYORB_DS_MINB 8      // 32 registers, 64 warps per SM (measured 0.48 -> 0.44 ms per 256 images against 48 registers)
#endif

// cv::fastAtan2 (OpenCV core/mathfuncs_core, scalar atan_f32), called at ORBFinder.cpp:42
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float k = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k;
    const float p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, (float)DBL_EPSILON));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, (float)DBL_EPSILON));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

constexpr int PT_R = 18;                 // patch radius: |rotated pattern coordinate| <= 18 (pattern radius 13*sqrt2), disc radius 15
constexpr int PT_ROWS = 2 * PT_R + 1;    // 37
constexpr int PT_WORDS = 10;             // 37 columns + up to 3 bytes of alignment slack, as 32-bit words

// byte masks of the disc: row |v| keeps columns |u| <= umax[|v|] (ORBFinder::orientationSetup, ORBFinder.cpp:131-149); word k of a
// row covers u = -15 + 4k .. -12 + 4k (the 32nd byte, u = 16, is never inside)
__device__ __forceinline__ uint32_t disc_mask_word(int av, int k)
{
    const int umax = av <= 3 ? 15 : av <= 6 ? 14 : av <= 8 ? 13 : av == 9 ? 12 : av == 10 ? 11 : av == 11 ? 10 : av == 12 ? 9 : av == 13 ? 8 : av == 14 ? 6 : 3;
    uint32_t m = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const int u = -15 + 4 * k + b, au = u < 0 ? -u : u;
        if (au <= umax) m |= 0xFFu << (8 * b);
    }
    return m;
}
__device__ __forceinline__ int dp4a_u8_s8(uint32_t a_unsigned, uint32_t b_signed, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_unsigned), "r"(b_signed), "r"(c));
    return d;
}
__host__ __device__ constexpr uint32_t pack_s8(int a, int b, int c, int d) { return (uint32_t)(uint8_t)a | ((uint32_t)(uint8_t)b << 8) | ((uint32_t)(uint8_t)c << 16) | ((uint32_t)(uint8_t)d << 24); }

__global__ void __launch_bounds__(DS_WARPS * 32, HYORB_DS_MINB)
k_describe(const PlanDev *__restrict__ plan, const uint8_t *__restrict__ blur, const uint32_t *__restrict__ sel_all,
           const int *__restrict__ selCount, hyorb_keypoint *__restrict__ kps, uint8_t *__restrict__ desc, int capacity,
           int *__restrict__ counts, int *__restrict__ status)
{
    // the 37x37 window of the blurred level around the keypoint, one per warp, staged with coalesced 32-bit loads: the
    // 31 disc rows and the 512 rotated-pattern taps then hit shared memory instead of issuing byte gathers to L1
    __shared__ uint32_t s_patch[DS_WARPS][PT_ROWS * PT_WORDS];
    __shared__ uint32_t s_mask[16][8];
    __shared__ int s_end[HYORB_MAX_LEVELS + 1];       // s_end[l] = output slots of levels < l (exclusive ends), s_end[nl] = total
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 128) s_mask[threadIdx.x >> 3][threadIdx.x & 7] = disc_mask_word(threadIdx.x >> 3, threadIdx.x & 7);
    const int b = blockIdx.y;
    const int nl = plan->nlevels;
    // locate the output slots: levels are concatenated in order (ORBExtractor.cpp:523-555).  ONE warp of the CTA scans the level counts
    // (lane i holds level i's count) and publishes the level boundaries; the other warps used to repeat the scan behind its global loads.
    if (warp == DS_WARPS - 1) {
        int cnt = 0;
        if (lane < nl) { cnt = selCount[b * HYORB_MAX_LEVELS + lane]; cnt = min(cnt, plan->lv[lane].selCap); }
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < HYORB_MAX_LEVELS; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane < HYORB_MAX_LEVELS) s_end[lane + 1] = inc;      // lanes >= nl add 0: s_end[nl..] = total
        if (lane == 0) s_end[0] = 0;
    }
    __syncthreads();
    const int k = blockIdx.x * DS_WARPS + warp;
    const int total = s_end[nl];
    int l = -1, j = 0;
    if (k < total) {
        l = 0;
#pragma unroll
        for (int t = 1; t < HYORB_MAX_LEVELS; t++) l += (t < nl && s_end[t] <= k);       // number of levels that end at or before slot k
        j = k - s_end[l];
    }
    if (k == 0 && lane == 0) {
        // every slot needs a warp: the grid covers min(capacity, the extractor's keypoint bound) slots (launch_describe)
        if (total > capacity || total > (int)gridDim.x * DS_WARPS) { atomicOr(status, ST_OUT_OVERFLOW); counts[b] = min(total, capacity); }
        else counts[b] = total;
    }
    if (l < 0 || k >= capacity) return;
    const LevelDev &L = plan->lv[l];
    const uint32_t c = sel_all[(size_t)b * plan->selStride + L.selOff + j];
    const int px = cand_x(c) + LATTICE_MIN, py = cand_y(c) + LATTICE_MIN;     // :484-485
    const int pitch = L.pitch;                                                  // multiple of 16; level base 256-aligned
    const uint8_t *level = blur + (size_t)b * plan->pyrStride + L.off;
    // stage rows py-18 .. py+18, bytes gx0 .. gx0+39 where gx0 = (px-18) rounded down to 4 (px >= 19 keeps it >= 0; the
    // last word may run <= 2 bytes past the row, still inside the level's padded block, and is never read back)
    const int gx0 = (px - PT_R) & ~3, off = (px - PT_R) - gx0;
    uint32_t *patch = s_patch[warp];
    {
        // lanes 0..29 cover 3 rows x 10 words per step; all 13 loads are issued before the first store
        const int lr = lane / PT_WORDS, wd = lane - lr * PT_WORDS;
        // 32-bit word offsets from ONE 64-bit base (a level is far below 2^31 words): one IMAD.WIDE per load instead of 64-bit pointer chains
        const uint32_t *base = (const uint32_t *)(level + gx0);
        const int pw = pitch >> 2;                               // pitch is a multiple of 16
        const int o0 = (py - PT_R + lr) * pw + wd;
        const bool act = lane < 30;
        uint32_t v[13];
#pragma unroll
        for (int t = 0; t < 13; t++) {
            const int r = 3 * t + lr;
            v[t] = (act && r < PT_ROWS) ? __ldg(base + (o0 + 3 * t * pw)) : 0u;
        }
        uint32_t *dstw = patch + lr * PT_WORDS + wd;
#pragma unroll
        for (int t = 0; t < 13; t++) {
            const int r = 3 * t + lr;
            if (act && r < PT_ROWS) dstw[3 * t * PT_WORDS] = v[t];
        }
    }
    __syncwarp();
    const uint8_t *center = (const uint8_t *)patch + PT_R * (PT_WORDS * 4) + off + PT_R;
    constexpr int PP = PT_WORDS * 4;      // patch pitch in bytes

    // ---- intensity centroid (ORBFinder.cpp:16-43): lane = disc row v; m10 += sum u*I and m01 += v * sum I over |u| <= umax[|v|]
    // (ORBFinder::orientationSetup, :131-149) as masked DP4As against the column weights -- integer sums, exact in any order
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int v = lane - HALF_PATCH, av = v < 0 ? -v : v;
        const int o = off + PT_R - HALF_PATCH;                        // patch byte of column u = -15 in this row
        const uint32_t *rowp = patch + (PT_R + v) * PT_WORDS + (o >> 2);
        const unsigned sh = (unsigned)(o & 3) * 8;
        uint32_t x[9];
#pragma unroll
        for (int t = 0; t < 9; t++) x[t] = rowp[t];
        int su = 0, s1 = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const uint32_t w = __funnelshift_r(x[t], x[t + 1], sh) & s_mask[av][t];
            su = dp4a_u8_s8(w, pack_s8(-15 + 4 * t, -14 + 4 * t, -13 + 4 * t, -12 + 4 * t), su);
            s1 = (int)__dp4a(w, 0x01010101u, (unsigned)s1);
        }
        m10 = su;
        m01 = v * s1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- rotated BRIEF (ORBFinder.cpp:89-129)
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float rad = __fmul_rn(angle, factorPI);
    double sn, cs;
    sincos((double)rad, &sn, &cs);
    const float a = (float)cs, bb = (float)sn;
    int val = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const float4 pt = d_brief_pattern_f[t * 32 + lane];
        const float x0 = pt.x, y0 = pt.y, x1 = pt.z, y1 = pt.w;
        const int r0 = cv_round_small(__fadd_rn(__fmul_rn(x0, bb), __fmul_rn(y0, a)));
        const int c0 = cv_round_small(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, bb)));
        const int r1 = cv_round_small(__fadd_rn(__fmul_rn(x1, bb), __fmul_rn(y1, a)));
        const int c1 = cv_round_small(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, bb)));
        const int t0 = center[r0 * PP + c0], t1 = center[r1 * PP + c1];
        val |= (t0 < t1) << t;
    }
    const size_t o = (size_t)b * capacity + k;
    desc[o * HYORB_DESC_BYTES + lane] = (uint8_t)val;
    if (lane == 0) {
        hyorb_keypoint kp;
        kp.x = (float)px; kp.y = (float)py;
        if (l != 0) { kp.x = __fmul_rn(kp.x, L.scale); kp.y = __fmul_rn(kp.y, L.scale); }   // ORBExtractor.cpp:546-552
        kp.size = L.kpSize; kp.angle = angle; kp.response = (float)cand_resp(c);
        kp.octave = l; kp.class_id = -1;
        kps[o] = kp;
    }
}

int launch_describe(const PlanDev &hp, const PlanDev *dp, const uint8_t *blur, const uint32_t *sel, const int *selCount,
                    hyorb_keypoint *kps, uint8_t *desc, int capacity, int *counts, int *status, int B, cudaStream_t st, long *launches, int max_kps)
{
    // the float copy of the pattern is filled once per device; the mutex blocks concurrent first callers (the reference
    // runs the left and the right extractor on two threads) until the table is complete
    static std::mutex pattern_mu;
    static bool pattern_ready[64];      // set only after the table has been written successfully: a failed first call is retried by the next one
    int dev = 0;
    HY_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device ordinal %d not supported", dev); return HYORB_EUNSUPPORTED; }
    {
        std::lock_guard<std::mutex> lock(pattern_mu);
        if (!pattern_ready[dev]) {
            k_pattern_to_float<<<1, 256, 0, st>>>();
            cudaError_t perr = cudaGetLastError();
            if (perr == cudaSuccess) perr = cudaStreamSynchronize(st);
            ++*launches;
            if (perr != cudaSuccess) { set_error("pattern table: %s", cudaGetErrorString(perr)); return HYORB_ECUDA; }
            pattern_ready[dev] = true;
        }
    }
    // one warp per output slot: at most min(capacity, what the quadtree can produce for these quotas (max_kps; 0 = unknown), the leaf
    // capacity); the kernel reports ST_OUT_OVERFLOW if an image ever produced more than the grid covers
    int slots = hp.selTotalCap < capacity ? hp.selTotalCap : capacity;
    if (max_kps > 0 && max_kps < slots) slots = max_kps;
    if (slots < 1) slots = 1;
    dim3 grd((slots + DS_WARPS - 1) / DS_WARPS, B);
    k_describe<<<grd, DS_WARPS * 32, 0, st>>>(dp, blur, sel, selCount, kps, desc, capacity, counts, status);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

}  // namespace hyorb
