// api.cu -- the C ABI of libhyorb (include/hyorb.h): handles, workspaces, staging, status plumbing.
// Host side of the drop-in: what ORBExtractor / Stereomatcher / FeatureMatcher objects own in the reference
// (per-object pyramid + scratch, src/features/ORBExtractor.h:102) lives in a handle here; nothing is global.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <vector>

#include "common.cuh"
#include "tma.cuh"

// A batch call drives up to 2 x MAX_LANES streams and throughput users run several handles from several threads.  The
// driver multiplexes streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8); streams that share a queue
// serialise on each other (measured: two host threads x 8 lanes end to end 25-36k pairs/s with 8 queues, 39k with 32).
// The variable is read when the CUDA context is created and belongs to the host process: the library does NOT touch the
// environment; INTEGRATION.md recommends CUDA_DEVICE_MAX_CONNECTIONS=32 (bench.py and the Python package set it for themselves).

namespace hyorb {
struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    cudaStream_t home = nullptr; bool has_home = false;     // the owning handle's stream: fresh memory is zeroed on it
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return HYORB_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) { set_error("cudaMalloc(%zu bytes) -> %s", bytes, cudaGetErrorString(e)); return e == cudaErrorMemoryAllocation ? HYORB_ENOMEM : HYORB_ECUDA; }
        // fresh buffers are zeroed once: row padding of the image planes is read (and then multiplied by a zero coefficient or
        // masked) without ever being written, and this keeps those reads initialised
        // -- on the handle's own stream (every use of the buffer is enqueued on it, or on a lane stream that first waits for it), so no
        // other stream of the host application is stalled; buffers without a home stream fall back to a synchronous memset
        if (has_home) { if (cudaMemsetAsync(p, 0, bytes, home) != cudaSuccess) cudaGetLastError(); }
        else if (cudaMemset(p, 0, bytes) != cudaSuccess || cudaStreamSynchronize(cudaStreamLegacy) != cudaSuccess) { cudaGetLastError(); }
        cap = bytes;
        return HYORB_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return (T *)p; }
};

static int status_to_rc(int st)
{
    if (st == 0) return HYORB_OK;
    if (st & (ST_CAND_OVERFLOW | ST_SEL_OVERFLOW | ST_OUT_OVERFLOW)) {
        set_error("device buffer overflow (status 0x%x): %s%s%s", st, (st & ST_CAND_OVERFLOW) ? "FAST candidate list " : "",
                  (st & ST_SEL_OVERFLOW) ? "quadtree leaf list " : "", (st & ST_OUT_OVERFLOW) ? "output capacity" : "");
        return HYORB_ECAPACITY;
    }
    if (st & (ST_QT_LIMIT | ST_QT_MISMATCH)) { set_error("quadtree kernel limit / consistency check failed (status 0x%x)", st); return HYORB_EUNSUPPORTED; }
    if (st & ST_BAD_INDEX) { set_error("candidate index or rotation bin out of range"); return HYORB_EINVAL; }
    if (st & ST_ROWTAB_OVERFLOW) {
        set_error("stereo: the right keypoints cover more image rows than the row table holds (20 rows per keypoint on average for device batches; "
                  "the extractor and host entry points size it from the largest keypoint)");
        return HYORB_ECAPACITY;
    }
    if (st & ST_ROW_RANGE) { set_error("stereo: keypoint row band outside [0, n_rows) (the reference writes out of bounds here)"); return HYORB_EINVAL; }
    set_error("device status 0x%x", st);
    return HYORB_ECUDA;
}

}  // namespace hyorb

using namespace hyorb;

struct hyorb_extractor {
    hyorb_extractor_params params;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // A batch is cut into `lanes` independent sub-batches, each enqueued on its own stream, so that the latency-bound
    // kernels of one sub-batch (quadtree, stereo table) overlap the throughput-bound ones (FAST, blur) of another.
    // Lane 0 runs on `stream`.  Inside a lane the blur runs on a side stream next to FAST + quadtree (it only needs the pyramid).
    static constexpr int MAX_LANES = 16;
    int lanes = 3;           // device-pointer entry points (measured 1 / 2 / 3 / 4 / 6 lanes: 43.7k / 48.4k / 49.1k / 49.0k / 48.7k pairs per second)
    int host_lanes = 8;      // host-buffer entry points: more, smaller lanes so the PCIe copies pipeline against the kernels
    int side_blur = 2;        // 1: blur on a side stream next to FAST + quadtree; 2: next to the quadtree only (HYORB_SIDE_BLUR)
    cudaStream_t lane_stream[MAX_LANES] = {}, side[MAX_LANES] = {};
    cudaEvent_t ev_start = nullptr, ev_pyr[MAX_LANES] = {}, ev_blur[MAX_LANES] = {}, ev_done[MAX_LANES] = {};
    float scale[HYORB_MAX_LEVELS], inv[HYORB_MAX_LEVELS], sigma2[HYORB_MAX_LEVELS], inv_sigma2[HYORB_MAX_LEVELS];
    int quota[HYORB_MAX_LEVELS];
    HostPlan plan;
    bool have_plan = false;
    DevBuf d_plan, d_resize, d_lut;
    int Bcap = 0;
    DevBuf d_pyr, d_blur, d_cand, d_qcode, d_qnode, d_qleaf, d_sel, d_candCount, d_selCount, d_status;
    DevBuf d_in, d_raw, d_kps, d_desc, d_counts;   // staging of the _host entry points: d_raw = byte-for-byte mirror of a dense host batch,
                                                   // d_in = level 0 in TMA-compatible layout (repacked on the device when the source is not)
    // TMA: tensor maps of the pyramid levels (FAST tile boxes) live in device memory, level 0's travels as a kernel parameter
    DevBuf d_tmaps, d_tmaps_lv, d_lvtab;
    CUtensorMap h_tmaps[HYORB_MAX_LEVELS], h_tmaps_lv[HYORB_MAX_LEVELS];
    CUtensorMap tm0, tmL0;    // level 0 with the FAST box / with the fused-level box (level.cu)
    int dl_bound = 0;         // entries per image a large host batch downloads while it runs (= kp_bound; HYORB_KP_BOUND overrides it for tests)
    int kp_bound = 0;         // sum over levels of (quota + 3): the most keypoints DistributeOctTree returns per image (ORBExtractor.cpp:107-290 stops
                              // splitting at the first node count >= quota, one split adds at most 3); host downloads are sized by it
    int fused_levels = 1;     // 1: level.cu (pyramid + blur in one pass per level); 0: pyramid.cu + blur.cu (HYORB_FUSED_LEVELS)
    struct { const void *base; int pitch; unsigned long long stride; int B, w, h; } tm0_key = {nullptr, 0, 0, 0, 0, 0};
    int sm_count = 0;
    DevBuf d_rowtab, d_bestd, d_uR, d_depth;    // stereo stage of the ProcessStereoImage entry points
    long launches = 0;
    // per-stage CUDA-event timing (events on the launching stream; read back by hyorb_extractor_stage_times)
    bool profile = false;
    std::vector<cudaEvent_t> ev_free;
    std::vector<std::vector<cudaEvent_t>> ev_pending;   // HYORB_N_STAGES+1 events per profiled call
    double stage_ms[HYORB_N_STAGES] = {0, 0, 0, 0, 0, 0};
    long stage_calls = 0;
    // last call (for debug_read)
    int last_B = 0;
    Level0 last_l0{nullptr, 0, 0};
};

static int ex_ensure_plan(hyorb_extractor *h, int w, int hgt)
{
    if (h->have_plan && h->plan.dev.width == w && h->plan.dev.height == hgt) return HYORB_OK;
    HostPlan np;
    HY_TRY(build_plan(h->params, w, hgt, &np));
    HY_TRY(h->d_plan.ensure(sizeof(PlanDev)));
    HY_TRY(h->d_resize.ensure(sizeof(ResizeTab) * std::max<size_t>(np.resize.size(), 1)));
    HY_TRY(h->d_lut.ensure(sizeof(uint32_t) * std::max<size_t>(np.lut.size(), 1)));
    HY_TRY(h->d_lvtab.ensure(sizeof(int) * std::max<size_t>(np.lvtab.size(), 1)));
    // the tables must not be overwritten while earlier launches may still read them
    HY_CUDA(cudaStreamSynchronize(h->stream));
    h->plan = np;
    HY_CUDA(cudaMemcpyAsync(h->d_plan.p, &h->plan.dev, sizeof(PlanDev), cudaMemcpyHostToDevice, h->stream));
    if (!np.resize.empty()) HY_CUDA(cudaMemcpyAsync(h->d_resize.p, h->plan.resize.data(), sizeof(ResizeTab) * np.resize.size(), cudaMemcpyHostToDevice, h->stream));
    HY_CUDA(cudaMemcpyAsync(h->d_lut.p, h->plan.lut.data(), sizeof(uint32_t) * np.lut.size(), cudaMemcpyHostToDevice, h->stream));
    HY_CUDA(cudaMemcpyAsync(h->d_lvtab.p, h->plan.lvtab.data(), sizeof(int) * np.lvtab.size(), cudaMemcpyHostToDevice, h->stream));
    HY_CUDA(cudaStreamSynchronize(h->stream));
    h->have_plan = true;
    h->kp_bound = 0;
    for (int l = 0; l < np.dev.nlevels; l++) {      // the first pass splits every root unconditionally: at most 4 * round(width / height) nodes
        const int roots = np.dev.lv[l].h > 0 ? np.dev.lv[l].w / np.dev.lv[l].h + 2 : 2;
        h->kp_bound += std::max(h->quota[l] + 3, 4 * roots);
    }
    h->dl_bound = h->kp_bound;
    if (const char *v = getenv("HYORB_KP_BOUND")) { const int b = atoi(v); if (b > 0) h->dl_bound = b; }   // tests: force the overflow fetch of large host batches
    h->Bcap = 0;
    return HYORB_OK;
}

static int ex_ensure_workspace(hyorb_extractor *h, int B)
{
    if (B <= h->Bcap) return HYORB_OK;
    HY_CUDA(cudaStreamSynchronize(h->stream));
    const PlanDev &P = h->plan.dev;
    HY_TRY(h->d_pyr.ensure((size_t)P.pyrStride * B + 256));
    HY_TRY(h->d_blur.ensure((size_t)P.pyrStride * B + 256));
    HY_TRY(h->d_cand.ensure(sizeof(uint32_t) * (size_t)P.candStride * B));
    HY_TRY(h->d_qcode.ensure(sizeof(uint32_t) * (size_t)P.candStride * B));
    HY_TRY(h->d_qnode.ensure(sizeof(uint16_t) * (size_t)P.candStride * B));
    HY_TRY(h->d_qleaf.ensure(sizeof(uint2) * (size_t)P.selStride * B));
    HY_TRY(h->d_sel.ensure(sizeof(uint32_t) * (size_t)P.selStride * B));
    HY_TRY(h->d_candCount.ensure(sizeof(int) * HYORB_MAX_LEVELS * (size_t)B));
    HY_TRY(h->d_selCount.ensure(sizeof(int) * HYORB_MAX_LEVELS * (size_t)B));
    if (!h->d_status.p) {
        HY_TRY(h->d_status.ensure(sizeof(int)));
        HY_CUDA(cudaMemsetAsync(h->d_status.p, 0, sizeof(int), h->stream));
    }
    // tensor maps of pyramid levels 1.. over the whole workspace (image index = z coordinate); slot 0 is unused
    HY_TRY(h->d_tmaps.ensure(sizeof(CUtensorMap) * HYORB_MAX_LEVELS));
    memset(h->h_tmaps, 0, sizeof(h->h_tmaps));
    for (int l = 1; l < P.nlevels; l++)
        HY_TRY(tma_encode_u8_3d(&h->h_tmaps[l], h->d_pyr.as<uint8_t>() + P.lv[l].off, P.lv[l].w, P.lv[l].h, B, (size_t)P.lv[l].pitch,
                                (size_t)P.pyrStride, FT_BOXW, FT_PH));
    HY_CUDA(cudaMemcpyAsync(h->d_tmaps.p, h->h_tmaps, sizeof(h->h_tmaps), cudaMemcpyHostToDevice, h->stream));
    // the same levels with the box of the fused pyramid + blur kernel
    HY_TRY(h->d_tmaps_lv.ensure(sizeof(CUtensorMap) * HYORB_MAX_LEVELS));
    memset(h->h_tmaps_lv, 0, sizeof(h->h_tmaps_lv));
    for (int l = 1; l < P.nlevels; l++)
        HY_TRY(tma_encode_u8_3d(&h->h_tmaps_lv[l], h->d_pyr.as<uint8_t>() + P.lv[l].off, P.lv[l].w, P.lv[l].h, B, (size_t)P.lv[l].pitch,
                                (size_t)P.pyrStride, LV_BW, LV_BH));
    HY_CUDA(cudaMemcpyAsync(h->d_tmaps_lv.p, h->h_tmaps_lv, sizeof(h->h_tmaps_lv), cudaMemcpyHostToDevice, h->stream));
    HY_CUDA(cudaStreamSynchronize(h->stream));      // h_tmaps is pageable host memory
    h->Bcap = B;
    return HYORB_OK;
}

// level-0 tensor map of the current call (cached while the caller keeps passing the same buffer)
static int ex_ensure_tm0(hyorb_extractor *h, Level0 l0, int B, int w, int hgt)
{
    auto &k = h->tm0_key;
    if (k.base == l0.base && k.pitch == l0.pitch && k.stride == l0.stride && k.B == B && k.w == w && k.h == hgt) return HYORB_OK;
    HY_TRY(tma_encode_u8_3d(&h->tm0, l0.base, w, hgt, B, (size_t)l0.pitch, (size_t)l0.stride, FT_BOXW, FT_PH));
    HY_TRY(tma_encode_u8_3d(&h->tmL0, l0.base, w, hgt, B, (size_t)l0.pitch, (size_t)l0.stride, LV_BW, LV_BH));
    k.base = l0.base; k.pitch = l0.pitch; k.stride = l0.stride; k.B = B; k.w = w; k.h = hgt;
    return HYORB_OK;
}

static int ex_event(hyorb_extractor *h, cudaStream_t st, std::vector<cudaEvent_t> *set)
{
    if (!h->profile) return HYORB_OK;
    cudaEvent_t e;
    if (!h->ev_free.empty()) { e = h->ev_free.back(); h->ev_free.pop_back(); }
    else HY_CUDA(cudaEventCreate(&e));
    HY_CUDA(cudaEventRecord(e, st));
    set->push_back(e);
    return HYORB_OK;
}

// host buffers of a "_host" entry point: each lane uploads its own slice before its first kernel and downloads its own
// results after its last one, so PCIe traffic of one lane overlaps the kernels of the others
struct HostIO {
    const uint8_t *images; int stride; size_t image_stride;     // source images (host)
    hyorb_keypoint *kps; uint8_t *desc; int32_t *counts; float *uR; float *depth;   // destinations (host), uR/depth optional
    bool defer_results = false;     // small batches: the caller downloads exactly what was produced (ex_download_small) instead of capacity-sized blocks
};
constexpr int SMALL_BATCH = 8;      // images; below this a call is latency-bound and the extra round trip for the counts pays

// the whole extraction pipeline (+ optional stereo association over consecutive image pairs) for B images
static int ex_run(hyorb_extractor *h, Level0 l0, int B, int w, int hgt, hyorb_keypoint *d_kps, uint8_t *d_desc, int capacity, int32_t *d_counts,
                  const hyorb_stereo_params *sp = nullptr, float *d_uR = nullptr, float *d_depth = nullptr, const HostIO *io = nullptr)
{
    HY_CUDA(cudaSetDevice(h->device));
    HY_TRY(ex_ensure_plan(h, w, hgt));
    HY_TRY(ex_ensure_workspace(h, B));
    // the row table of the stereo stage is sized for this pyramid's largest keypoint (ORBExtractor.cpp:478: size = 31 * scale[level])
    const int st_rows = sp ? stereo_rows_budget(h->plan.dev.lv[h->plan.dev.nlevels - 1].kpSize, sp->size_ref) : 0;
    const size_t st_ints = sp ? stereo_scratch_ints_per_pair(capacity, st_rows) : 0;
    if (sp) {
        HY_TRY(h->d_rowtab.ensure(sizeof(int32_t) * st_ints * (B / 2)));
        HY_TRY(h->d_bestd.ensure(sizeof(int32_t) * (size_t)capacity * (B / 2)));
    }
    const PlanDev &P = h->plan.dev;
    const PlanDev *dp = h->d_plan.as<PlanDev>();
    // ---- level 0 must be TMA-addressable (16-byte aligned base, pitch, image stride): the staging copy of the host entry
    // points is; a caller's device batch that is not gets repacked once (each lane repacks its own images)
    Level0 src0 = l0;
    bool repack = false;
    if (!tma_compatible(l0.base, (size_t)l0.pitch, (size_t)l0.stride)) {
        const int pitch = P.lv[0].pitch;
        const size_t dstride = (size_t)pitch * hgt;
        HY_TRY(h->d_in.ensure(dstride * B + 512));
        l0 = Level0{h->d_in.as<uint8_t>(), pitch, (unsigned long long)dstride};
        repack = true;
    }
    HY_TRY(ex_ensure_tm0(h, l0, B, w, hgt));
    // ---- cut the batch into lanes (whole pairs when the stereo stage follows)
    const int unit = sp ? 2 : 1, units = B / unit;
    const int want = io ? h->host_lanes : h->lanes;
    int nl = want < 1 ? 1 : (want > hyorb_extractor::MAX_LANES ? hyorb_extractor::MAX_LANES : want);
    if (nl > units) nl = units;
    int first[hyorb_extractor::MAX_LANES + 1];
    for (int k = 0; k <= nl; k++) first[k] = unit * (int)((long long)units * k / nl);
    std::vector<cudaEvent_t> evs[hyorb_extractor::MAX_LANES];
    if (nl > 1 || h->side_blur) HY_CUDA(cudaEventRecord(h->ev_start, h->stream));
    // stage-major issue order so the hardware queues alternate between lanes
    for (int stage = 0; stage < 6; stage++) {
        for (int k = 0; k < nl; k++) {
            cudaStream_t st = h->lane_stream[k];
            const int i0 = first[k], Bk = first[k + 1] - first[k];
            Level0 l0k{l0.base + (size_t)i0 * l0.stride, l0.pitch, l0.stride};
            uint8_t *pyr = h->d_pyr.as<uint8_t>() + (size_t)i0 * P.pyrStride, *blur = h->d_blur.as<uint8_t>() + (size_t)i0 * P.pyrStride;
            uint32_t *cand = h->d_cand.as<uint32_t>() + (size_t)i0 * P.candStride;
            int *candCount = h->d_candCount.as<int>() + (size_t)i0 * HYORB_MAX_LEVELS, *selCount = h->d_selCount.as<int>() + (size_t)i0 * HYORB_MAX_LEVELS;
            uint32_t *sel = h->d_sel.as<uint32_t>() + (size_t)i0 * P.selStride;
            int *status = h->d_status.as<int>();
            // event slots per lane: 0 start, 1 pyramid done, 2 FAST done, 3 quadtree done, 4 describe start (blur joined), 5 describe done,
            // 6 stereo done, 7-8 blur (side stream)
            switch (stage) {
            case 0:
                if (k > 0) HY_CUDA(cudaStreamWaitEvent(st, h->ev_start, 0));
                if (io) {       // upload this lane's images into the staging copy: one flat copy when it mirrors a dense host batch
                    uint8_t *dst = const_cast<uint8_t *>(src0.base) + (size_t)i0 * src0.stride;
                    const uint8_t *src = io->images + (size_t)i0 * io->image_stride;
                    if (src0.pitch == io->stride && src0.stride == io->image_stride)
                        HY_CUDA(cudaMemcpyAsync(dst, src, (size_t)(Bk - 1) * io->image_stride + (size_t)io->stride * (hgt - 1) + w, cudaMemcpyHostToDevice, st));
                    else
                        for (int i = 0; i < Bk; i++)
                            HY_CUDA(cudaMemcpy2DAsync(dst + (size_t)i * src0.stride, src0.pitch, src + (size_t)i * io->image_stride, io->stride, w, hgt,
                                                      cudaMemcpyHostToDevice, st));
                }
                if (repack) {
                    HY_TRY(launch_repack(src0.base + (size_t)i0 * src0.stride, src0.pitch, (size_t)src0.stride, const_cast<uint8_t *>(l0k.base), l0.pitch,
                                         (size_t)l0.stride, w, hgt, Bk, st, &h->launches));
                }
                HY_CUDA(cudaMemsetAsync(candCount, 0, sizeof(int) * HYORB_MAX_LEVELS * (size_t)Bk, st));
                HY_TRY(ex_event(h, st, &evs[k]));
                if (h->fused_levels)      // pyramid + blur of every level in one pass per level (level.cu)
                    HY_TRY(launch_levels(P, dp, h->tmL0, h->d_tmaps_lv.as<CUtensorMap>(), i0, l0k, pyr, blur, h->d_resize.as<ResizeTab>(), h->d_lvtab.as<int>(), Bk,
                                         h->sm_count, st, &h->launches));
                else
                    HY_TRY(launch_pyramid(P, dp, l0k, pyr, h->d_resize.as<ResizeTab>(), Bk, st, &h->launches));
                HY_TRY(ex_event(h, st, &evs[k]));
                break;
            case 1:
                if (h->side_blur == 1 && !h->fused_levels) {
                    HY_CUDA(cudaEventRecord(h->ev_pyr[k], st));
                    HY_CUDA(cudaStreamWaitEvent(h->side[k], h->ev_pyr[k], 0));
                }
                HY_TRY(launch_fast(P, dp, h->tm0, h->d_tmaps.as<CUtensorMap>(), i0, cand, candCount, status, Bk, h->sm_count, st, &h->launches));
                if (h->side_blur == 2 && !h->fused_levels) {        // the latency-bound quadtree CTAs are dispatched first, the blur fills the rest of each SM
                    HY_CUDA(cudaEventRecord(h->ev_pyr[k], st));
                    HY_CUDA(cudaStreamWaitEvent(h->side[k], h->ev_pyr[k], 0));
                }
                HY_TRY(ex_event(h, st, &evs[k]));
                break;
            case 2:
                HY_TRY(launch_quadtree(P, dp, cand, candCount, h->d_lut.as<uint32_t>(), h->d_qcode.as<uint32_t>() + (size_t)i0 * P.candStride,
                                       h->d_qnode.as<uint16_t>() + (size_t)i0 * P.candStride, h->d_qleaf.as<uint2>() + (size_t)i0 * P.selStride, sel,
                                       selCount, status, Bk, st, &h->launches));
                HY_TRY(ex_event(h, st, &evs[k]));
                break;
            case 3: {
                const int side = h->fused_levels ? 0 : h->side_blur;
                cudaStream_t bs = side ? h->side[k] : st;
                std::vector<cudaEvent_t> evb;
                HY_TRY(ex_event(h, bs, &evb));
                if (!h->fused_levels) HY_TRY(launch_blur(P, dp, l0k, pyr, blur, Bk, bs, &h->launches));
                HY_TRY(ex_event(h, bs, &evb));
                if (side) {
                    HY_CUDA(cudaEventRecord(h->ev_blur[k], bs));
                    HY_CUDA(cudaStreamWaitEvent(st, h->ev_blur[k], 0));
                }
                HY_TRY(ex_event(h, st, &evs[k]));                   // slot 4
                if (h->profile) { evs[k].resize(9); evs[k][7] = evb[0]; evs[k][8] = evb[1]; }
                break;
            }
            case 4:
                HY_TRY(launch_describe(P, dp, blur, sel, selCount, d_kps + (size_t)i0 * capacity, d_desc + (size_t)i0 * capacity * HYORB_DESC_BYTES, capacity,
                                       d_counts + i0, status, Bk, st, &h->launches, h->kp_bound));
                if (h->profile) {
                    cudaEvent_t e;
                    if (!h->ev_free.empty()) { e = h->ev_free.back(); h->ev_free.pop_back(); } else HY_CUDA(cudaEventCreate(&e));
                    HY_CUDA(cudaEventRecord(e, st));
                    evs[k][5] = e;
                }
                break;
            case 5:
                if (sp) {
                    const int p0 = i0 / 2;
                    HY_TRY(launch_stereo(*sp, Bk / 2, d_kps + (size_t)i0 * capacity, d_desc + (size_t)i0 * capacity * HYORB_DESC_BYTES, d_counts + i0, capacity,
                                         h->d_rowtab.as<int32_t>() + st_ints * p0, d_uR + (size_t)p0 * capacity, d_depth + (size_t)p0 * capacity, nullptr,
                                         h->d_bestd.as<int32_t>() + (size_t)p0 * capacity, status, st, &h->launches, st_rows));
                }
                if (h->profile) {
                    cudaEvent_t e;
                    if (!h->ev_free.empty()) { e = h->ev_free.back(); h->ev_free.pop_back(); } else HY_CUDA(cudaEventCreate(&e));
                    HY_CUDA(cudaEventRecord(e, st));
                    evs[k][6] = e;
                }
                if (io && !io->defer_results) {
                    // download this lane's results: per image only the first `rows` entries of its capacity-sized block (2-D copies), where
                    // rows = what the quadtree can produce at most for these quotas; ex_fetch_overflow() covers an image that still exceeded it
                    const size_t rows = (size_t)std::min(capacity, h->dl_bound);
                    HY_CUDA(cudaMemcpyAsync(io->counts + i0, d_counts + i0, sizeof(int32_t) * Bk, cudaMemcpyDeviceToHost, st));
                    HY_CUDA(cudaMemcpy2DAsync(io->kps + (size_t)i0 * capacity, sizeof(hyorb_keypoint) * (size_t)capacity, d_kps + (size_t)i0 * capacity,
                                              sizeof(hyorb_keypoint) * (size_t)capacity, sizeof(hyorb_keypoint) * rows, Bk, cudaMemcpyDeviceToHost, st));
                    HY_CUDA(cudaMemcpy2DAsync(io->desc + (size_t)i0 * capacity * HYORB_DESC_BYTES, (size_t)HYORB_DESC_BYTES * capacity,
                                              d_desc + (size_t)i0 * capacity * HYORB_DESC_BYTES, (size_t)HYORB_DESC_BYTES * capacity, (size_t)HYORB_DESC_BYTES * rows, Bk,
                                              cudaMemcpyDeviceToHost, st));
                    if (sp && io->uR && io->depth) {
                        const int p0 = i0 / 2;
                        HY_CUDA(cudaMemcpy2DAsync(io->uR + (size_t)p0 * capacity, sizeof(float) * (size_t)capacity, d_uR + (size_t)p0 * capacity,
                                                  sizeof(float) * (size_t)capacity, sizeof(float) * rows, Bk / 2, cudaMemcpyDeviceToHost, st));
                        HY_CUDA(cudaMemcpy2DAsync(io->depth + (size_t)p0 * capacity, sizeof(float) * (size_t)capacity, d_depth + (size_t)p0 * capacity,
                                                  sizeof(float) * (size_t)capacity, sizeof(float) * rows, Bk / 2, cudaMemcpyDeviceToHost, st));
                    }
                }
                if (k > 0) {
                    HY_CUDA(cudaEventRecord(h->ev_done[k], st));
                    HY_CUDA(cudaStreamWaitEvent(h->stream, h->ev_done[k], 0));
                }
                break;
            }
        }
    }
    if (h->profile) { for (int k = 0; k < nl; k++) h->ev_pending.push_back(evs[k]); h->stage_calls++; }
    h->last_B = B; h->last_l0 = l0;
    return HYORB_OK;
}

static int ex_sync(hyorb_extractor *h);
// results of a small host batch: counts first, then only the keypoints / descriptors / stereo values each image produced
static int ex_download_small(hyorb_extractor *h, const HostIO &io, int B, int capacity, const hyorb_keypoint *d_kps, const uint8_t *d_desc,
                             const int32_t *d_counts, const float *d_uR, const float *d_depth)
{
    HY_CUDA(cudaMemcpyAsync(io.counts, d_counts, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, h->stream));
    HY_TRY(ex_sync(h));
    for (int i = 0; i < B; i++) {
        const int n = io.counts[i] < capacity ? io.counts[i] : capacity;
        if (n <= 0) continue;
        HY_CUDA(cudaMemcpyAsync(io.kps + (size_t)i * capacity, d_kps + (size_t)i * capacity, sizeof(hyorb_keypoint) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
        HY_CUDA(cudaMemcpyAsync(io.desc + (size_t)i * capacity * HYORB_DESC_BYTES, d_desc + (size_t)i * capacity * HYORB_DESC_BYTES, (size_t)HYORB_DESC_BYTES * n,
                                cudaMemcpyDeviceToHost, h->stream));
        if (d_uR && io.uR && io.depth && (i & 1) == 0) {
            const int p = i / 2;
            HY_CUDA(cudaMemcpyAsync(io.uR + (size_t)p * capacity, d_uR + (size_t)p * capacity, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
            HY_CUDA(cudaMemcpyAsync(io.depth + (size_t)p * capacity, d_depth + (size_t)p * capacity, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
        }
    }
    HY_CUDA(cudaStreamSynchronize(h->stream));
    return HYORB_OK;
}

// large host batches download min(capacity, kp_bound) entries per image while the batch is still running; an image that produced more
// (quotas the bound does not cover) gets the rest here, after the counts have arrived
static int ex_fetch_overflow(hyorb_extractor *h, const HostIO &io, int B, int capacity, const hyorb_keypoint *d_kps, const uint8_t *d_desc,
                             const float *d_uR, const float *d_depth)
{
    const int rows = std::min(capacity, h->dl_bound);
    bool any = false;
    for (int i = 0; i < B; i++) {
        const int n = std::min(io.counts[i], capacity);
        if (n <= rows) continue;
        any = true;
        const size_t o = (size_t)i * capacity + rows, m = (size_t)(n - rows);
        HY_CUDA(cudaMemcpyAsync(io.kps + o, d_kps + o, sizeof(hyorb_keypoint) * m, cudaMemcpyDeviceToHost, h->stream));
        HY_CUDA(cudaMemcpyAsync(io.desc + o * HYORB_DESC_BYTES, d_desc + o * HYORB_DESC_BYTES, (size_t)HYORB_DESC_BYTES * m, cudaMemcpyDeviceToHost, h->stream));
        if (d_uR && io.uR && io.depth && (i & 1) == 0) {
            const size_t q = (size_t)(i / 2) * capacity + rows;
            HY_CUDA(cudaMemcpyAsync(io.uR + q, d_uR + q, sizeof(float) * m, cudaMemcpyDeviceToHost, h->stream));
            HY_CUDA(cudaMemcpyAsync(io.depth + q, d_depth + q, sizeof(float) * m, cudaMemcpyDeviceToHost, h->stream));
        }
    }
    if (any) HY_CUDA(cudaStreamSynchronize(h->stream));
    return HYORB_OK;
}

static int ex_sync(hyorb_extractor *h)
{
    int st = 0;
    if (h->d_status.p) {
        HY_CUDA(cudaMemcpyAsync(&st, h->d_status.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        HY_CUDA(cudaStreamSynchronize(h->stream));
        if (st) HY_CUDA(cudaMemsetAsync(h->d_status.p, 0, sizeof(int), h->stream));
    } else {
        HY_CUDA(cudaStreamSynchronize(h->stream));
    }
    return status_to_rc(st);
}

extern "C" {

HYORB_API const char *hyorb_last_error(void) { return hyorb::last_error(); }
HYORB_API const char *hyorb_version(void) { return "hyorb-b200 0.1 (sm_100a)"; }
HYORB_API int hyorb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

HYORB_API int hyorb_extractor_create(const hyorb_extractor_params *params, int device, void *cuda_stream, hyorb_extractor **out)
{
    if (!params || !out) { set_error("null argument"); return HYORB_EINVAL; }
    *out = nullptr;
    if (params->flags != 0) { set_error("flags must be 0"); return HYORB_EINVAL; }
    if (params->nfeatures < 1) { set_error("nfeatures must be >= 1"); return HYORB_EINVAL; }
    hyorb_extractor *h = new (std::nothrow) hyorb_extractor();
    if (!h) return HYORB_ENOMEM;
    h->params = *params;
    int rc = scale_tables(*params, h->scale, h->inv, h->sigma2, h->inv_sigma2, h->quota);
    if (rc) { delete h; return rc; }
    if (hyorb_device_count() <= device || device < 0) { set_error("CUDA device %d not available (no CPU fallback exists)", device); delete h; return HYORB_ECUDA; }
    h->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) {
        if (cuda_stream) { h->stream = (cudaStream_t)cuda_stream; h->own_stream = false; }
        else { e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking); h->own_stream = true; }
    }
    h->lane_stream[0] = h->stream;
    {
        DevBuf *bufs[] = {&h->d_plan, &h->d_resize, &h->d_lut, &h->d_pyr, &h->d_blur, &h->d_cand, &h->d_qcode, &h->d_qnode, &h->d_qleaf, &h->d_sel,
                          &h->d_candCount, &h->d_selCount, &h->d_status, &h->d_in, &h->d_raw, &h->d_kps, &h->d_desc, &h->d_counts,
                          &h->d_rowtab, &h->d_bestd, &h->d_uR, &h->d_depth, &h->d_tmaps, &h->d_tmaps_lv, &h->d_lvtab};
        if (e == cudaSuccess) for (DevBuf *b : bufs) { b->home = h->stream; b->has_home = true; }
    }
    for (int k = 0; k < hyorb_extractor::MAX_LANES && e == cudaSuccess; k++) {
        if (k > 0) e = cudaStreamCreateWithFlags(&h->lane_stream[k], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->side[k], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_pyr[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_blur[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_done[k], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (const char *v = getenv("HYORB_LANES")) h->lanes = atoi(v);
    if (const char *v = getenv("HYORB_HOST_LANES")) h->host_lanes = atoi(v);
    if (const char *v = getenv("HYORB_SIDE_BLUR")) h->side_blur = atoi(v);
    if (const char *v = getenv("HYORB_FUSED_LEVELS")) h->fused_levels = atoi(v) != 0;
    if (e != cudaSuccess) {
        set_error("CUDA init: %s", cudaGetErrorString(e));
        cudaGetLastError();
        if (!h->own_stream) h->stream = nullptr;
        hyorb_extractor_destroy(h);          // releases the streams / events created so far
        return HYORB_ECUDA;
    }
    *out = h;
    return HYORB_OK;
}

HYORB_API int hyorb_extractor_destroy(hyorb_extractor *h)
{
    if (!h) return HYORB_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int k = 0; k < hyorb_extractor::MAX_LANES; k++) {
        if (k > 0 && h->lane_stream[k]) { cudaStreamSynchronize(h->lane_stream[k]); cudaStreamDestroy(h->lane_stream[k]); }
        if (h->side[k]) { cudaStreamSynchronize(h->side[k]); cudaStreamDestroy(h->side[k]); }
        if (h->ev_pyr[k]) cudaEventDestroy(h->ev_pyr[k]);
        if (h->ev_blur[k]) cudaEventDestroy(h->ev_blur[k]);
        if (h->ev_done[k]) cudaEventDestroy(h->ev_done[k]);
    }
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    DevBuf *bufs[] = {&h->d_plan, &h->d_resize, &h->d_lut, &h->d_pyr, &h->d_blur, &h->d_cand, &h->d_qcode, &h->d_qnode, &h->d_qleaf, &h->d_sel,
                      &h->d_candCount, &h->d_selCount, &h->d_status, &h->d_in, &h->d_raw, &h->d_kps, &h->d_desc, &h->d_counts,
                      &h->d_rowtab, &h->d_bestd, &h->d_uR, &h->d_depth, &h->d_tmaps, &h->d_tmaps_lv, &h->d_lvtab};
    for (DevBuf *b : bufs) b->release();
    for (auto &set : h->ev_pending) for (cudaEvent_t e : set) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_free) cudaEventDestroy(e);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return HYORB_OK;
}

HYORB_API int hyorb_extractor_get_levels(const hyorb_extractor *h) { return h ? h->params.nlevels : HYORB_EINVAL; }

HYORB_API int hyorb_extractor_get_scales(const hyorb_extractor *h, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2, int32_t *quota)
{
    if (!h) { set_error("null handle"); return HYORB_EINVAL; }
    for (int i = 0; i < h->params.nlevels; i++) {
        if (scale) scale[i] = h->scale[i];
        if (inv_scale) inv_scale[i] = h->inv[i];
        if (sigma2) sigma2[i] = h->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = h->inv_sigma2[i];
        if (quota) quota[i] = h->quota[i];
    }
    return HYORB_OK;
}

HYORB_API int hyorb_extract_batch_device(hyorb_extractor *h, const uint8_t *d_images, int n_images, int width, int height, int stride,
                                         size_t image_stride, hyorb_keypoint *d_kps, uint8_t *d_desc, int capacity, int32_t *d_counts)
{
    if (!h || !d_images || !d_kps || !d_desc || !d_counts) { set_error("null argument"); return HYORB_EINVAL; }
    if (n_images < 1 || width < 1 || height < 1 || stride < width || capacity < 1) { set_error("bad shape"); return HYORB_EINVAL; }
    if (n_images > 65535) { set_error("at most 65535 images per batch"); return HYORB_EUNSUPPORTED; }
    Level0 l0{d_images, stride, (unsigned long long)image_stride};
    return ex_run(h, l0, n_images, width, height, d_kps, d_desc, capacity, d_counts);
}

HYORB_API int hyorb_extractor_sync(hyorb_extractor *h)
{
    if (!h) { set_error("null handle"); return HYORB_EINVAL; }
    HY_CUDA(cudaSetDevice(h->device));
    return ex_sync(h);
}

// staging copy of a host batch.  Level 0 must end up TMA-addressable (16-byte aligned pitch and image stride).  A 2-D
// copy of odd-pitch rows runs at a fraction of the PCIe rate, so dense host batches are mirrored with one flat copy per
// lane and repacked by a kernel
static int ex_stage_geometry(hyorb_extractor *h, int n_images, int width, int height, int stride, size_t image_stride, Level0 *l0)
{
    const PlanDev &P = h->plan.dev;
    const bool dense = image_stride >= (size_t)stride * height && image_stride <= (size_t)stride * height + 4096;
    if (dense) {        // mirror: one flat copy per lane; ex_run repacks on the device unless the layout is already TMA-compatible
        DevBuf &buf = tma_compatible(nullptr, (size_t)stride, image_stride) ? h->d_in : h->d_raw;
        HY_TRY(buf.ensure(image_stride * n_images + 512));
        *l0 = Level0{buf.as<uint8_t>(), stride, (unsigned long long)image_stride};
    } else {            // scattered host images: the copy engine repacks rows to the plan's aligned pitch
        const int pitch = P.lv[0].pitch;
        const size_t dstride = (size_t)pitch * height;
        HY_TRY(h->d_in.ensure(dstride * n_images + 512));
        *l0 = Level0{h->d_in.as<uint8_t>(), pitch, (unsigned long long)dstride};
    }
    (void)width;
    return HYORB_OK;
}

HYORB_API int hyorb_extract_batch_host(hyorb_extractor *h, const uint8_t *images, int n_images, int width, int height, int stride,
                                       size_t image_stride, hyorb_keypoint *kps, uint8_t *desc, int capacity, int32_t *counts)
{
    if (!h || !images || !kps || !desc || !counts) { set_error("null argument"); return HYORB_EINVAL; }
    if (n_images < 1 || width < 1 || height < 1 || stride < width || capacity < 1) { set_error("bad shape"); return HYORB_EINVAL; }
    if (n_images > 65535) { set_error("at most 65535 images per batch"); return HYORB_EUNSUPPORTED; }
    HY_CUDA(cudaSetDevice(h->device));
    HY_TRY(ex_ensure_plan(h, width, height));
    Level0 l0;
    HY_TRY(ex_stage_geometry(h, n_images, width, height, stride, image_stride, &l0));
    HY_TRY(h->d_kps.ensure(sizeof(hyorb_keypoint) * (size_t)capacity * n_images));
    HY_TRY(h->d_desc.ensure((size_t)HYORB_DESC_BYTES * capacity * n_images));
    HY_TRY(h->d_counts.ensure(sizeof(int32_t) * n_images));
    HostIO io{images, stride, image_stride, kps, desc, counts, nullptr, nullptr};
    io.defer_results = n_images <= SMALL_BATCH;
    HY_TRY(ex_run(h, l0, n_images, width, height, h->d_kps.as<hyorb_keypoint>(), h->d_desc.as<uint8_t>(), capacity, h->d_counts.as<int32_t>(),
                  nullptr, nullptr, nullptr, &io));
    if (io.defer_results)
        return ex_download_small(h, io, n_images, capacity, h->d_kps.as<hyorb_keypoint>(), h->d_desc.as<uint8_t>(), h->d_counts.as<int32_t>(), nullptr, nullptr);
    HY_TRY(ex_sync(h));
    return ex_fetch_overflow(h, io, n_images, capacity, h->d_kps.as<hyorb_keypoint>(), h->d_desc.as<uint8_t>(), nullptr, nullptr);
}

HYORB_API int hyorb_extract_host(hyorb_extractor *h, const uint8_t *gray, int width, int height, int stride, hyorb_keypoint *kps,
                                 uint8_t *desc, int capacity, int *n)
{
    if (!h || !n) { set_error("null argument"); return HYORB_EINVAL; }
    *n = 0;
    if (!gray || width <= 0 || height <= 0) return HYORB_OK;     // ORBExtractor.cpp:499-500: empty image -> silent return
    if (!kps || !desc || capacity < 1 || stride < width) { set_error("bad argument"); return HYORB_EINVAL; }
    // a batch of one: one flat upload of the rows (repacked on the device when the pitch is not TMA-addressable) and a download of
    // exactly what was produced.  A narrow view of a much wider image is uploaded row by row instead (scattered layout).
    const size_t istride = (size_t)stride * height + ((size_t)stride > 2 * (size_t)width ? 8192 : 0);
    int32_t cnt = 0;
    HY_TRY(hyorb_extract_batch_host(h, gray, 1, width, height, stride, istride, kps, desc, capacity, &cnt));
    *n = cnt < capacity ? cnt : capacity;
    return HYORB_OK;
}

HYORB_API int hyorb_preprocess_size(int width, int height, int half_scale, int *out_width, int *out_height)
{
    if (!out_width || !out_height) { set_error("null argument"); return HYORB_EINVAL; }
    return preprocess_size(width, height, half_scale, out_width, out_height);
}

HYORB_API int hyorb_preprocess_device(hyorb_extractor *h, const uint8_t *d_src, int n_images, int width, int height, int stride, size_t image_stride,
                                      int channels, int rgb_order, int half_scale, uint8_t *d_gray, int gray_stride, size_t gray_image_stride)
{
    if (!h || !d_src || !d_gray) { set_error("null argument"); return HYORB_EINVAL; }
    HY_CUDA(cudaSetDevice(h->device));
    return launch_preprocess(d_src, stride, image_stride, width, height, channels, rgb_order != 0, half_scale != 0, d_gray, gray_stride, gray_image_stride,
                             n_images, h->stream, &h->launches);
}

HYORB_API int hyorb_extract_color_host(hyorb_extractor *h, const uint8_t *image, int width, int height, int stride, int channels, int rgb_order,
                                       int half_scale, uint8_t *gray_out, int gray_out_stride, hyorb_keypoint *kps, uint8_t *desc, int capacity, int *n)
{
    if (!h || !n) { set_error("null argument"); return HYORB_EINVAL; }
    *n = 0;
    if (!image || width <= 0 || height <= 0) return HYORB_OK;     // empty image -> silent return, like the gray entry point
    if (!kps || !desc || capacity < 1 || (channels != 1 && channels != 3 && channels != 4) || stride < width * channels) { set_error("bad argument"); return HYORB_EINVAL; }
    int gw, gh;
    HY_TRY(preprocess_size(width, height, half_scale, &gw, &gh));
    if (gray_out && gray_out_stride < gw) { set_error("bad argument"); return HYORB_EINVAL; }
    HY_CUDA(cudaSetDevice(h->device));
    HY_TRY(ex_ensure_plan(h, gw, gh));
    const PlanDev &P = h->plan.dev;
    const int pitch = P.lv[0].pitch;
    const size_t src_bytes = (size_t)stride * (height - 1) + (size_t)width * channels;
    HY_TRY(h->d_raw.ensure(src_bytes + 16));
    HY_TRY(h->d_in.ensure((size_t)pitch * gh + 512));
    HY_TRY(h->d_kps.ensure(sizeof(hyorb_keypoint) * (size_t)capacity));
    HY_TRY(h->d_desc.ensure((size_t)HYORB_DESC_BYTES * capacity));
    HY_TRY(h->d_counts.ensure(sizeof(int32_t)));
    HY_CUDA(cudaMemcpyAsync(h->d_raw.p, image, src_bytes, cudaMemcpyHostToDevice, h->stream));
    HY_TRY(launch_preprocess(h->d_raw.as<uint8_t>(), stride, 0, width, height, channels, rgb_order != 0, half_scale != 0, h->d_in.as<uint8_t>(), pitch,
                             (size_t)pitch * gh, 1, h->stream, &h->launches));
    if (gray_out)       // the reference keeps the gray frame (track_data.image, ImageProcessing.cpp:109)
        HY_CUDA(cudaMemcpy2DAsync(gray_out, gray_out_stride, h->d_in.p, pitch, gw, gh, cudaMemcpyDeviceToHost, h->stream));
    Level0 l0{h->d_in.as<uint8_t>(), pitch, (unsigned long long)pitch * gh};
    HY_TRY(ex_run(h, l0, 1, gw, gh, h->d_kps.as<hyorb_keypoint>(), h->d_desc.as<uint8_t>(), capacity, h->d_counts.as<int32_t>()));
    int32_t cnt = 0;
    HY_CUDA(cudaMemcpyAsync(&cnt, h->d_counts.p, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    HY_TRY(ex_sync(h));
    if (cnt > 0) {
        HY_CUDA(cudaMemcpyAsync(kps, h->d_kps.p, sizeof(hyorb_keypoint) * (size_t)cnt, cudaMemcpyDeviceToHost, h->stream));
        HY_CUDA(cudaMemcpyAsync(desc, h->d_desc.p, (size_t)HYORB_DESC_BYTES * cnt, cudaMemcpyDeviceToHost, h->stream));
        HY_CUDA(cudaStreamSynchronize(h->stream));
    }
    *n = cnt;
    return HYORB_OK;
}

HYORB_API int hyorb_process_stereo_batch_device(hyorb_extractor *h, const hyorb_stereo_params *sp, const uint8_t *d_images, int n_pairs,
                                                int width, int height, int stride, size_t image_stride, hyorb_keypoint *d_kps, uint8_t *d_desc,
                                                int capacity, int32_t *d_counts, float *d_uR, float *d_depth)
{
    if (!h || !sp || !d_images || !d_kps || !d_desc || !d_counts || !d_uR || !d_depth) { set_error("null argument"); return HYORB_EINVAL; }
    if (n_pairs < 1 || width < 1 || height < 1 || stride < width || capacity < 1) { set_error("bad shape"); return HYORB_EINVAL; }
    if (2 * n_pairs > 65535) { set_error("at most 32767 pairs per batch"); return HYORB_EUNSUPPORTED; }
    Level0 l0{d_images, stride, (unsigned long long)image_stride};
    return ex_run(h, l0, 2 * n_pairs, width, height, d_kps, d_desc, capacity, d_counts, sp, d_uR, d_depth);
}

HYORB_API int hyorb_process_stereo_batch_host(hyorb_extractor *h, const hyorb_stereo_params *sp, const uint8_t *images, int n_pairs, int width,
                                              int height, int stride, size_t image_stride, hyorb_keypoint *kps, uint8_t *desc, int capacity,
                                              int32_t *counts, float *uR, float *depth)
{
    if (!h || !sp || !images || !kps || !desc || !counts || !uR || !depth) { set_error("null argument"); return HYORB_EINVAL; }
    if (n_pairs < 1 || width < 1 || height < 1 || stride < width || capacity < 1) { set_error("bad shape"); return HYORB_EINVAL; }
    if (2 * n_pairs > 65535) { set_error("at most 32767 pairs per batch"); return HYORB_EUNSUPPORTED; }
    const int n_images = 2 * n_pairs;
    HY_CUDA(cudaSetDevice(h->device));
    HY_TRY(ex_ensure_plan(h, width, height));
    Level0 l0;
    HY_TRY(ex_stage_geometry(h, n_images, width, height, stride, image_stride, &l0));
    HY_TRY(h->d_kps.ensure(sizeof(hyorb_keypoint) * (size_t)capacity * n_images));
    HY_TRY(h->d_desc.ensure((size_t)HYORB_DESC_BYTES * capacity * n_images));
    HY_TRY(h->d_counts.ensure(sizeof(int32_t) * n_images));
    HY_TRY(h->d_uR.ensure(sizeof(float) * (size_t)capacity * n_pairs));
    HY_TRY(h->d_depth.ensure(sizeof(float) * (size_t)capacity * n_pairs));
    HostIO io{images, stride, image_stride, kps, desc, counts, uR, depth};
    io.defer_results = n_images <= SMALL_BATCH;
    HY_TRY(ex_run(h, l0, n_images, width, height, h->d_kps.as<hyorb_keypoint>(), h->d_desc.as<uint8_t>(), capacity, h->d_counts.as<int32_t>(), sp,
                  h->d_uR.as<float>(), h->d_depth.as<float>(), &io));
    if (io.defer_results)
        return ex_download_small(h, io, n_images, capacity, h->d_kps.as<hyorb_keypoint>(), h->d_desc.as<uint8_t>(), h->d_counts.as<int32_t>(),
                                 h->d_uR.as<float>(), h->d_depth.as<float>());
    HY_TRY(ex_sync(h));
    return ex_fetch_overflow(h, io, n_images, capacity, h->d_kps.as<hyorb_keypoint>(), h->d_desc.as<uint8_t>(), h->d_uR.as<float>(), h->d_depth.as<float>());
}

HYORB_API int hyorb_extractor_set_profiling(hyorb_extractor *h, int enable)
{
    if (!h) { set_error("null handle"); return HYORB_EINVAL; }
    h->profile = enable != 0;
    return HYORB_OK;
}

HYORB_API int hyorb_extractor_set_pipelining(hyorb_extractor *h, int device_lanes, int host_lanes, int side_blur)
{
    if (!h) { set_error("null handle"); return HYORB_EINVAL; }
    if (device_lanes > hyorb_extractor::MAX_LANES || host_lanes > hyorb_extractor::MAX_LANES || device_lanes == 0 || host_lanes == 0 || side_blur > 2) {
        set_error("pipelining: lanes must be 1..%d, side_blur 0..2", hyorb_extractor::MAX_LANES);
        return HYORB_EINVAL;
    }
    HY_CUDA(cudaSetDevice(h->device));
    HY_CUDA(cudaStreamSynchronize(h->stream));
    if (device_lanes > 0) h->lanes = device_lanes;
    if (host_lanes > 0) h->host_lanes = host_lanes;
    if (side_blur >= 0) h->side_blur = side_blur;
    return HYORB_OK;
}

HYORB_API int hyorb_extractor_keypoint_bound(hyorb_extractor *h, int width, int height)
{
    if (!h) { set_error("null handle"); return HYORB_EINVAL; }
    HY_CUDA(cudaSetDevice(h->device));
    HY_TRY(ex_ensure_plan(h, width, height));
    return h->dl_bound;
}

HYORB_API int hyorb_extractor_stage_times(hyorb_extractor *h, double *ms, long *calls, int reset)
{
    if (!h) { set_error("null handle"); return HYORB_EINVAL; }
    HY_CUDA(cudaSetDevice(h->device));
    HY_CUDA(cudaStreamSynchronize(h->stream));
    for (auto &set : h->ev_pending) {
        // slots: see ex_run.  pyramid 0-1, FAST 1-2, quadtree 2-3, describe 4-5, stereo 5-6, blur 7-8 (side stream, overlapped)
        static const int from[HYORB_N_STAGES] = {0, 1, 2, 7, 4, 5}, to[HYORB_N_STAGES] = {1, 2, 3, 8, 5, 6};
        if (set.size() >= 9)
            for (int i = 0; i < HYORB_N_STAGES; i++) {
                float t = 0.f;
                HY_CUDA(cudaEventElapsedTime(&t, set[from[i]], set[to[i]]));
                h->stage_ms[i] += t;
            }
        for (cudaEvent_t e : set) if (e) h->ev_free.push_back(e);
    }
    h->ev_pending.clear();
    for (int i = 0; i < HYORB_N_STAGES; i++) { if (ms) ms[i] = h->stage_ms[i]; if (reset) h->stage_ms[i] = 0; }
    if (calls) *calls = h->stage_calls;
    if (reset) h->stage_calls = 0;
    return HYORB_OK;
}

HYORB_API int hyorb_extractor_level_size(hyorb_extractor *h, int width, int height, int level, int *lw, int *lh)
{
    if (!h || level < 0 || level >= h->params.nlevels) { set_error("bad argument"); return HYORB_EINVAL; }
    const int w = (int)lrintf((float)width * h->inv[level]), hh = (int)lrintf((float)height * h->inv[level]);   // ORBExtractor.cpp:569
    if (lw) *lw = w;
    if (lh) *lh = hh;
    return HYORB_OK;
}

HYORB_API long hyorb_extractor_launch_count(const hyorb_extractor *h) { return h ? h->launches : 0; }

HYORB_API long hyorb_extractor_debug_read(hyorb_extractor *h, int image_index, int what, int level, void *dst, size_t dst_bytes)
{
    if (!h || !dst || !h->have_plan || image_index < 0 || image_index >= h->last_B || level < 0 || level >= h->plan.dev.nlevels) {
        set_error("debug_read: bad argument or no previous extract call");
        return HYORB_EINVAL;
    }
    if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess) { set_error("debug_read: CUDA error"); return HYORB_ECUDA; }
    const PlanDev &P = h->plan.dev;
    const LevelDev &L = P.lv[level];
    switch (what) {
    case HYORB_DBG_PYRAMID:
    case HYORB_DBG_BLURRED: {
        const size_t need = (size_t)L.w * L.h;
        if (dst_bytes < need) { set_error("debug_read: need %zu bytes", need); return HYORB_ECAPACITY; }
        const uint8_t *src; int pitch;
        if (what == HYORB_DBG_BLURRED) { src = h->d_blur.as<uint8_t>() + (size_t)P.pyrStride * image_index + L.off; pitch = L.pitch; }
        else if (level == 0) { src = h->last_l0.base + (size_t)h->last_l0.stride * image_index; pitch = h->last_l0.pitch; }
        else { src = h->d_pyr.as<uint8_t>() + (size_t)P.pyrStride * image_index + L.off; pitch = L.pitch; }
        if (cudaMemcpy2D(dst, L.w, src, pitch, L.w, L.h, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("debug_read: copy failed"); return HYORB_ECUDA; }
        return (long)need;
    }
    case HYORB_DBG_CANDIDATES: {
        int cnt = 0;
        if (cudaMemcpy(&cnt, h->d_candCount.as<int>() + image_index * HYORB_MAX_LEVELS + level, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return HYORB_ECUDA;
        if (cnt > L.candCap) cnt = L.candCap;
        const size_t need = sizeof(int32_t) * 3 * (size_t)cnt;
        if (dst_bytes < need) { set_error("debug_read: need %zu bytes", need); return HYORB_ECAPACITY; }
        std::vector<uint32_t> c(cnt);
        if (cnt && cudaMemcpy(c.data(), h->d_cand.as<uint32_t>() + (size_t)P.candStride * image_index + L.candOff, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost) != cudaSuccess)
            return HYORB_ECUDA;
        std::sort(c.begin(), c.end(), [&](uint32_t a, uint32_t b) {
            return cand_order_key(cand_x(a), cand_y(a), L.wCell, L.hCell, L.nCols) < cand_order_key(cand_x(b), cand_y(b), L.wCell, L.hCell, L.nCols);
        });
        int32_t *o = (int32_t *)dst;
        for (int i = 0; i < cnt; i++) { o[3 * i] = cand_x(c[i]); o[3 * i + 1] = cand_y(c[i]); o[3 * i + 2] = cand_resp(c[i]); }
        return (long)need;
    }
    case HYORB_DBG_LEVEL_COUNT: {
        if (dst_bytes < sizeof(int32_t)) return HYORB_ECAPACITY;
        if (cudaMemcpy(dst, h->d_selCount.as<int>() + image_index * HYORB_MAX_LEVELS + level, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return HYORB_ECUDA;
        return (long)sizeof(int32_t);
    }
    default: set_error("debug_read: unknown selector %d", what); return HYORB_EINVAL;
    }
}

}  // extern "C"

// =====================================================================================================
// matcher handle
// =====================================================================================================
struct hyorb_matcher {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    long launches = 0;
    DevBuf d_status, d_a, d_b, d_c, d_d, d_e, d_f, d_g, d_h, d_i, d_j, d_k, d_l;   // generic staging slots
    DevBuf d_pkey, d_psecond, d_rowtab, d_bestd, d_cellof, d_cellcnt, d_kp1, d_kp2;
    int stereo_rows = 0;      // row-table budget of the next stereo call (hyorb_stereo_match_host derives it from the keypoints it uploads)
};

static int m_prepare(hyorb_matcher *m)
{
    if (!m) { set_error("null matcher handle"); return HYORB_EINVAL; }
    HY_CUDA(cudaSetDevice(m->device));
    if (!m->d_status.p) {
        HY_TRY(m->d_status.ensure(sizeof(int)));
        HY_CUDA(cudaMemsetAsync(m->d_status.p, 0, sizeof(int), m->stream));
    }
    return HYORB_OK;
}
static int m_sync(hyorb_matcher *m)
{
    int st = 0;
    HY_CUDA(cudaMemcpyAsync(&st, m->d_status.p, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaStreamSynchronize(m->stream));
    if (st) HY_CUDA(cudaMemsetAsync(m->d_status.p, 0, sizeof(int), m->stream));
    return status_to_rc(st);
}
static int m_upload(hyorb_matcher *m, DevBuf &buf, const void *src, size_t bytes)
{
    HY_TRY(buf.ensure(std::max<size_t>(bytes, 16)));
    if (bytes) HY_CUDA(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, m->stream));
    return HYORB_OK;
}
static int bf_splits(int nq, int nt)
{
    // enough CTAs to fill 148 SMs a few times over, without slicing the target list below one smem tile
    const int qblocks = (nq + bf_queries_per_cta() - 1) / bf_queries_per_cta();      // k_bf_partial: 128 threads x BF_QPT queries
    int s = (148 * 8 + qblocks - 1) / qblocks;
    const int maxs = std::max(1, nt / 128);
    return std::max(1, std::min(s, std::min(maxs, 64)));
}

extern "C" {

HYORB_API int hyorb_matcher_create(int device, void *cuda_stream, hyorb_matcher **out)
{
    if (!out) { set_error("null argument"); return HYORB_EINVAL; }
    *out = nullptr;
    if (device < 0 || hyorb_device_count() <= device) { set_error("CUDA device %d not available (no CPU fallback exists)", device); return HYORB_ECUDA; }
    hyorb_matcher *m = new (std::nothrow) hyorb_matcher();
    if (!m) return HYORB_ENOMEM;
    m->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) {
        if (cuda_stream) m->stream = (cudaStream_t)cuda_stream;
        else { e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking); m->own_stream = true; }
    }
    if (e != cudaSuccess) { set_error("CUDA init: %s", cudaGetErrorString(e)); delete m; return HYORB_ECUDA; }
    {
        DevBuf *bufs[] = {&m->d_status, &m->d_a, &m->d_b, &m->d_c, &m->d_d, &m->d_e, &m->d_f, &m->d_g, &m->d_h, &m->d_i, &m->d_j, &m->d_k, &m->d_l,
                          &m->d_pkey, &m->d_psecond, &m->d_rowtab, &m->d_bestd, &m->d_cellof, &m->d_cellcnt, &m->d_kp1, &m->d_kp2};
        for (DevBuf *b : bufs) { b->home = m->stream; b->has_home = true; }
    }
    *out = m;
    return HYORB_OK;
}

HYORB_API int hyorb_matcher_destroy(hyorb_matcher *m)
{
    if (!m) return HYORB_OK;
    cudaSetDevice(m->device);
    cudaStreamSynchronize(m->stream);
    DevBuf *bufs[] = {&m->d_status, &m->d_a, &m->d_b, &m->d_c, &m->d_d, &m->d_e, &m->d_f, &m->d_g, &m->d_h, &m->d_i, &m->d_j, &m->d_k, &m->d_l,
                      &m->d_pkey, &m->d_psecond, &m->d_rowtab, &m->d_bestd, &m->d_cellof, &m->d_cellcnt, &m->d_kp1, &m->d_kp2};
    for (DevBuf *b : bufs) b->release();
    if (m->own_stream && m->stream) cudaStreamDestroy(m->stream);
    delete m;
    return HYORB_OK;
}

HYORB_API int hyorb_matcher_sync(hyorb_matcher *m) { HY_TRY(m_prepare(m)); return m_sync(m); }
HYORB_API long hyorb_matcher_launch_count(const hyorb_matcher *m) { return m ? m->launches : 0; }

HYORB_API int hyorb_match_bruteforce_device(hyorb_matcher *m, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt, int rule, float thr,
                                            float ratio, int32_t *d_best_idx, uint16_t *d_best, uint16_t *d_second, uint8_t *d_accepted)
{
    HY_TRY(m_prepare(m));
    if (nq < 0 || nt < 0 || rule < 0 || rule > 2) { set_error("bad argument"); return HYORB_EINVAL; }
    if (nq == 0) return HYORB_OK;
    const int ns = bf_splits(nq, nt);
    HY_TRY(m->d_pkey.ensure(sizeof(uint32_t) * (size_t)ns * nq));
    HY_TRY(m->d_psecond.ensure(sizeof(uint16_t) * (size_t)ns * nq));
    return launch_match_bruteforce(d_q, nq, d_t, nt, rule, thr, ratio, d_best_idx, d_best, d_second, d_accepted, m->d_pkey.as<uint32_t>(),
                                   m->d_psecond.as<uint16_t>(), ns, m->stream, &m->launches);
}

struct EpipolarHost { const hyorb_keypoint *kps1, *kps2; const float *F12; float sigma_ref, size_ref; };
static int m_epipolar_upload(hyorb_matcher *m, const EpipolarHost *eh, int n1, int n2, EpipolarDev *epi)
{
    *epi = EpipolarDev{};
    if (!eh) return HYORB_OK;
    HY_TRY(m_upload(m, m->d_kp1, eh->kps1, sizeof(hyorb_keypoint) * (size_t)n1));
    HY_TRY(m_upload(m, m->d_kp2, eh->kps2, sizeof(hyorb_keypoint) * (size_t)n2));
    epi->kps1 = m->d_kp1.as<hyorb_keypoint>(); epi->kps2 = m->d_kp2.as<hyorb_keypoint>();
    for (int i = 0; i < 9; i++) epi->F[i] = eh->F12[i];
    epi->sigma_ref = eh->sigma_ref; epi->size_ref = eh->size_ref;
    return HYORB_OK;
}

static int m_match_csr(hyorb_matcher *m, const uint8_t *q_desc, int nq, const uint8_t *t_desc, int nt, const int32_t *cand_off,
                       const int32_t *cand_idx, int rule, float thr, float ratio, int32_t *best_idx, uint16_t *best,
                       uint16_t *second, uint8_t *accepted, const EpipolarHost *eh)
{
    HY_TRY(m_prepare(m));
    if (nq < 0 || nt < 0 || rule < 0 || rule > 2 || (cand_off == nullptr) != (cand_idx == nullptr)) { set_error("bad argument"); return HYORB_EINVAL; }
    if (nq == 0) return HYORB_OK;
    if (!q_desc || (nt > 0 && !t_desc) || !best_idx || !best || !second || !accepted) { set_error("null argument"); return HYORB_EINVAL; }
    HY_TRY(m_upload(m, m->d_a, q_desc, (size_t)nq * 32));
    HY_TRY(m_upload(m, m->d_b, t_desc, (size_t)nt * 32));
    HY_TRY(m->d_c.ensure(sizeof(int32_t) * (size_t)nq));
    HY_TRY(m->d_d.ensure(sizeof(uint16_t) * (size_t)nq));
    HY_TRY(m->d_e.ensure(sizeof(uint16_t) * (size_t)nq));
    HY_TRY(m->d_f.ensure((size_t)nq));
    if (cand_off) {
        const int total = cand_off[nq];
        if (cand_off[0] != 0 || total < 0) { set_error("cand_off must start at 0 and be non-decreasing"); return HYORB_EINVAL; }
        for (int i = 0; i < nq; i++) {
            if (cand_off[i + 1] < cand_off[i]) { set_error("cand_off must be non-decreasing"); return HYORB_EINVAL; }
            if (cand_off[i + 1] - cand_off[i] >= (1 << 22)) { set_error("candidate list too long"); return HYORB_EUNSUPPORTED; }
        }
        HY_TRY(m_upload(m, m->d_g, cand_off, sizeof(int32_t) * ((size_t)nq + 1)));
        HY_TRY(m_upload(m, m->d_h, cand_idx, sizeof(int32_t) * (size_t)total));
        EpipolarDev epi;
        HY_TRY(m_epipolar_upload(m, eh, nq, nt, &epi));
        HY_TRY(launch_match_csr(m->d_a.as<uint8_t>(), nq, m->d_b.as<uint8_t>(), nt, m->d_g.as<int32_t>(), m->d_h.as<int32_t>(), rule, thr, ratio,
                                m->d_c.as<int32_t>(), m->d_d.as<uint16_t>(), m->d_e.as<uint16_t>(), m->d_f.as<uint8_t>(), m->d_status.as<int>(),
                                epi, m->stream, &m->launches));
    } else {
        if (eh) { set_error("the epipolar criterion needs candidate lists"); return HYORB_EINVAL; }
        HY_TRY(hyorb_match_bruteforce_device(m, m->d_a.as<uint8_t>(), nq, m->d_b.as<uint8_t>(), nt, rule, thr, ratio, m->d_c.as<int32_t>(),
                                             m->d_d.as<uint16_t>(), m->d_e.as<uint16_t>(), m->d_f.as<uint8_t>()));
    }
    HY_CUDA(cudaMemcpyAsync(best_idx, m->d_c.p, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(best, m->d_d.p, sizeof(uint16_t) * (size_t)nq, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(second, m->d_e.p, sizeof(uint16_t) * (size_t)nq, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(accepted, m->d_f.p, (size_t)nq, cudaMemcpyDeviceToHost, m->stream));
    return m_sync(m);
}

HYORB_API int hyorb_match_csr_host(hyorb_matcher *m, const uint8_t *q_desc, int nq, const uint8_t *t_desc, int nt, const int32_t *cand_off,
                                   const int32_t *cand_idx, int rule, float thr, float ratio, int32_t *best_idx, uint16_t *best,
                                   uint16_t *second, uint8_t *accepted)
{
    return m_match_csr(m, q_desc, nq, t_desc, nt, cand_off, cand_idx, rule, thr, ratio, best_idx, best, second, accepted, nullptr);
}

HYORB_API int hyorb_match_csr_epipolar_host(hyorb_matcher *m, const hyorb_keypoint *q_kps, const uint8_t *q_desc, int nq, const hyorb_keypoint *t_kps,
                                            const uint8_t *t_desc, int nt, const int32_t *cand_off, const int32_t *cand_idx, const float *F12,
                                            float sigma_ref, float size_ref, int rule, float thr, float ratio, int32_t *best_idx, uint16_t *best,
                                            uint16_t *second, uint8_t *accepted)
{
    if (!F12 || !cand_off || !cand_idx || (nq > 0 && !q_kps) || (nt > 0 && !t_kps)) { set_error("null argument"); return HYORB_EINVAL; }
    if (!(size_ref > 0.f)) { set_error("size_ref must be positive"); return HYORB_EINVAL; }
    const EpipolarHost eh{q_kps, t_kps, F12, sigma_ref, size_ref};
    return m_match_csr(m, q_desc, nq, t_desc, nt, cand_off, cand_idx, rule, thr, ratio, best_idx, best, second, accepted, &eh);
}

HYORB_API int hyorb_grid_build_host(hyorb_matcher *m, const hyorb_keypoint *kps, int n, const hyorb_bounds *b, int32_t *cell_off, int32_t *cell_idx)
{
    HY_TRY(m_prepare(m));
    if (n < 0 || !b || !cell_off || (n > 0 && (!kps || !cell_idx))) { set_error("bad argument"); return HYORB_EINVAL; }
    constexpr int NC = HYORB_GRID_COLS * HYORB_GRID_ROWS;
    HY_TRY(m_upload(m, m->d_a, kps, sizeof(hyorb_keypoint) * (size_t)n));
    HY_TRY(m->d_i.ensure(sizeof(int32_t) * (NC + 1)));
    HY_TRY(m->d_j.ensure(sizeof(int32_t) * std::max(n, 1)));
    HY_TRY(m->d_cellof.ensure(sizeof(int32_t) * std::max(n, 1)));
    HY_TRY(m->d_cellcnt.ensure(sizeof(int32_t) * NC));
    HY_TRY(launch_grid_build(m->d_a.as<hyorb_keypoint>(), n, *b, m->d_i.as<int32_t>(), m->d_j.as<int32_t>(), m->d_cellof.as<int32_t>(),
                             m->d_cellcnt.as<int32_t>(), m->stream, &m->launches));
    HY_CUDA(cudaMemcpyAsync(cell_off, m->d_i.p, sizeof(int32_t) * (NC + 1), cudaMemcpyDeviceToHost, m->stream));
    HY_TRY(m_sync(m));
    const int total = cell_off[NC];
    if (total > 0) {
        HY_CUDA(cudaMemcpyAsync(cell_idx, m->d_j.p, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, m->stream));
        HY_CUDA(cudaStreamSynchronize(m->stream));
    }
    return HYORB_OK;
}

HYORB_API int hyorb_match_window_host(hyorb_matcher *m, const hyorb_keypoint *t_kps, const uint8_t *t_desc, const float *t_uR,
                                      const uint8_t *t_matched, int nt, const hyorb_bounds *b, const hyorb_window_query *queries,
                                      const uint8_t *q_desc, int nq, float thr, float ratio, int32_t *best_idx, uint16_t *best,
                                      uint16_t *second, uint8_t *accepted)
{
    HY_TRY(m_prepare(m));
    if (nq < 0 || nt < 0 || !b) { set_error("bad argument"); return HYORB_EINVAL; }
    if (nq == 0) return HYORB_OK;
    if (!queries || !q_desc || !best_idx || !best || !second || !accepted || (nt > 0 && (!t_kps || !t_desc))) { set_error("null argument"); return HYORB_EINVAL; }
    for (int i = 0; i < nq; i++)
        if (queries[i].ur_radius >= 0 && !t_uR) { set_error("stereo consistency requested but t_uR is NULL"); return HYORB_EINVAL; }
    constexpr int NC = HYORB_GRID_COLS * HYORB_GRID_ROWS;
    HY_TRY(m_upload(m, m->d_a, t_kps, sizeof(hyorb_keypoint) * (size_t)nt));
    HY_TRY(m_upload(m, m->d_b, t_desc, (size_t)nt * 32));
    if (t_uR) HY_TRY(m_upload(m, m->d_k, t_uR, sizeof(float) * (size_t)nt));
    if (t_matched) HY_TRY(m_upload(m, m->d_l, t_matched, (size_t)nt));
    HY_TRY(m_upload(m, m->d_g, queries, sizeof(hyorb_window_query) * (size_t)nq));
    HY_TRY(m_upload(m, m->d_h, q_desc, (size_t)nq * 32));
    HY_TRY(m->d_i.ensure(sizeof(int32_t) * (NC + 1)));
    HY_TRY(m->d_j.ensure(sizeof(int32_t) * std::max(nt, 1)));
    HY_TRY(m->d_cellof.ensure(sizeof(int32_t) * std::max(nt, 1)));
    HY_TRY(m->d_cellcnt.ensure(sizeof(int32_t) * NC));
    HY_TRY(m->d_c.ensure(sizeof(int32_t) * (size_t)nq));
    HY_TRY(m->d_d.ensure(sizeof(uint16_t) * (size_t)nq));
    HY_TRY(m->d_e.ensure(sizeof(uint16_t) * (size_t)nq));
    HY_TRY(m->d_f.ensure((size_t)nq));
    HY_TRY(launch_grid_build(m->d_a.as<hyorb_keypoint>(), nt, *b, m->d_i.as<int32_t>(), m->d_j.as<int32_t>(), m->d_cellof.as<int32_t>(),
                             m->d_cellcnt.as<int32_t>(), m->stream, &m->launches));
    HY_TRY(launch_match_window(m->d_a.as<hyorb_keypoint>(), m->d_b.as<uint8_t>(), t_uR ? m->d_k.as<float>() : nullptr,
                               t_matched ? m->d_l.as<uint8_t>() : nullptr, nt, *b, m->d_i.as<int32_t>(), m->d_j.as<int32_t>(),
                               m->d_g.as<hyorb_window_query>(), m->d_h.as<uint8_t>(), nq, thr, ratio, m->d_c.as<int32_t>(), m->d_d.as<uint16_t>(),
                               m->d_e.as<uint16_t>(), m->d_f.as<uint8_t>(), m->stream, &m->launches));
    HY_CUDA(cudaMemcpyAsync(best_idx, m->d_c.p, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(best, m->d_d.p, sizeof(uint16_t) * (size_t)nq, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(second, m->d_e.p, sizeof(uint16_t) * (size_t)nq, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(accepted, m->d_f.p, (size_t)nq, cudaMemcpyDeviceToHost, m->stream));
    return m_sync(m);
}

static int m_project(hyorb_matcher *m, const hyorb_projection *pr, const hyorb_landmark *lms, int n, const hyorb_keypoint *t_kps, int nt, float th,
                     float size_ref, float frac_smaller, float frac_larger, unsigned flags = HYORB_SBP_DISTANCE | HYORB_SBP_STEREO)
{
    // landmarks -> d_e, target keypoints -> d_a; queries -> d_g, passed -> d_k (read back / consumed by the caller)
    HY_TRY(m_upload(m, m->d_e, lms, sizeof(hyorb_landmark) * (size_t)n));
    HY_TRY(m_upload(m, m->d_a, t_kps, sizeof(hyorb_keypoint) * (size_t)nt));
    HY_TRY(m->d_g.ensure(sizeof(hyorb_window_query) * (size_t)n));
    HY_TRY(m->d_k.ensure((size_t)n));
    return launch_project_landmarks(*pr, m->d_e.as<hyorb_landmark>(), n, m->d_a.as<hyorb_keypoint>(), nt, th, size_ref, frac_smaller, frac_larger, flags,
                                    m->d_g.as<hyorb_window_query>(), m->d_k.as<uint8_t>(), m->d_status.as<int>(), m->stream, &m->launches);
}

HYORB_API int hyorb_project_landmarks_host(hyorb_matcher *m, const hyorb_projection *pr, const hyorb_landmark *lms, int n, const hyorb_keypoint *t_kps,
                                           int nt, float th, float size_ref, float frac_smaller, float frac_larger, hyorb_window_query *queries,
                                           uint8_t *passed)
{
    HY_TRY(m_prepare(m));
    if (n < 0 || nt < 0 || !pr) { set_error("bad argument"); return HYORB_EINVAL; }
    if (n == 0) return HYORB_OK;
    if (!lms || !queries || !passed || (nt > 0 && !t_kps)) { set_error("null argument"); return HYORB_EINVAL; }
    HY_TRY(m_project(m, pr, lms, n, t_kps, nt, th, size_ref, frac_smaller, frac_larger));
    HY_CUDA(cudaMemcpyAsync(queries, m->d_g.p, sizeof(hyorb_window_query) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(passed, m->d_k.p, (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    return m_sync(m);
}

HYORB_API int hyorb_search_by_projection_host(hyorb_matcher *m, const hyorb_projection *pr, const hyorb_landmark *lms, const uint8_t *lm_desc, int n,
                                              const hyorb_keypoint *t_kps, const uint8_t *t_desc, const float *t_uR, const uint8_t *t_matched, int nt,
                                              float th, float size_ref, float thr, float ratio, int32_t *best_idx, uint16_t *best, uint16_t *second,
                                              uint8_t *accepted, uint8_t *passed)
{
    return hyorb_search_by_projection_ex_host(m, pr, lms, lm_desc, nullptr, n, t_kps, t_desc, t_uR, t_matched, nt, th, size_ref, thr, ratio,
                                              HYORB_SBP_DISTANCE | HYORB_SBP_STEREO, best_idx, best, second, accepted, passed);
}

HYORB_API int hyorb_search_by_projection_ex_host(hyorb_matcher *m, const hyorb_projection *pr, const hyorb_landmark *lms, const uint8_t *lm_desc,
                                                 const float *lm_prev_angle, int n, const hyorb_keypoint *t_kps, const uint8_t *t_desc,
                                                 const float *t_uR, const uint8_t *t_matched, int nt, float th, float size_ref, float thr, float ratio,
                                                 unsigned flags, int32_t *best_idx, uint16_t *best, uint16_t *second, uint8_t *accepted, uint8_t *passed)
{
    HY_TRY(m_prepare(m));
    if (n < 0 || nt < 0 || !pr || (flags & ~7u)) { set_error("bad argument"); return HYORB_EINVAL; }
    if ((flags & HYORB_SBP_ROTATION) && n > 0 && !lm_prev_angle) { set_error("HYORB_SBP_ROTATION needs lm_prev_angle"); return HYORB_EINVAL; }
    if (n == 0) return HYORB_OK;
    if (!lms || !lm_desc || !best_idx || !best || !second || !accepted || (nt > 0 && (!t_kps || !t_desc))) { set_error("null argument"); return HYORB_EINVAL; }
    const bool use_ur = pr->stereo && (flags & HYORB_SBP_STEREO);
    if (use_ur && !t_uR) { set_error("stereo consistency requested but t_uR is NULL"); return HYORB_EINVAL; }
    constexpr int NC = HYORB_GRID_COLS * HYORB_GRID_ROWS;
    HY_TRY(m_project(m, pr, lms, n, t_kps, nt, th, size_ref, 0.5f, 1.5f, flags));     // FeatureSizeCriterion(0.5, 1.5), FeatureMatcher.cc:132
    HY_TRY(m_upload(m, m->d_b, t_desc, (size_t)nt * 32));
    if (t_uR) HY_TRY(m_upload(m, m->d_l, t_uR, sizeof(float) * (size_t)nt));
    if (t_matched) HY_TRY(m_upload(m, m->d_bestd, t_matched, (size_t)nt));
    HY_TRY(m_upload(m, m->d_h, lm_desc, (size_t)n * 32));
    HY_TRY(m->d_i.ensure(sizeof(int32_t) * (NC + 1)));
    HY_TRY(m->d_j.ensure(sizeof(int32_t) * std::max(nt, 1)));
    HY_TRY(m->d_cellof.ensure(sizeof(int32_t) * std::max(nt, 1)));
    HY_TRY(m->d_cellcnt.ensure(sizeof(int32_t) * NC));
    HY_TRY(m->d_c.ensure(sizeof(int32_t) * (size_t)n));
    HY_TRY(m->d_d.ensure(sizeof(uint16_t) * (size_t)n));
    HY_TRY(m->d_pkey.ensure(sizeof(uint16_t) * (size_t)n));
    HY_TRY(m->d_f.ensure((size_t)n));
    HY_TRY(launch_grid_build(m->d_a.as<hyorb_keypoint>(), nt, pr->bounds, m->d_i.as<int32_t>(), m->d_j.as<int32_t>(), m->d_cellof.as<int32_t>(),
                             m->d_cellcnt.as<int32_t>(), m->stream, &m->launches));
    HY_TRY(launch_match_window(m->d_a.as<hyorb_keypoint>(), m->d_b.as<uint8_t>(), t_uR ? m->d_l.as<float>() : nullptr,
                               t_matched ? m->d_bestd.as<uint8_t>() : nullptr, nt, pr->bounds, m->d_i.as<int32_t>(), m->d_j.as<int32_t>(),
                               m->d_g.as<hyorb_window_query>(), m->d_h.as<uint8_t>(), n, thr, ratio, m->d_c.as<int32_t>(), m->d_d.as<uint16_t>(),
                               m->d_pkey.as<uint16_t>(), m->d_f.as<uint8_t>(), m->stream, &m->launches, m->d_k.as<uint8_t>()));
    if (flags & HYORB_SBP_ROTATION) {
        HY_TRY(m_upload(m, m->d_psecond, lm_prev_angle, sizeof(float) * (size_t)n));
        HY_TRY(m->d_rowtab.ensure(sizeof(int32_t) * std::max(nt, 1)));
        HY_TRY(launch_projection_rotation(m->d_c.as<int32_t>(), m->d_f.as<uint8_t>(), n, m->d_psecond.as<float>(), m->d_a.as<hyorb_keypoint>(), nt,
                                          m->d_rowtab.as<int32_t>(), m->d_status.as<int>(), m->stream, &m->launches));
    }
    HY_CUDA(cudaMemcpyAsync(best_idx, m->d_c.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(best, m->d_d.p, sizeof(uint16_t) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(second, m->d_pkey.p, sizeof(uint16_t) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(accepted, m->d_f.p, (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    if (passed) HY_CUDA(cudaMemcpyAsync(passed, m->d_k.p, (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    return m_sync(m);
}

// shared tail of the projection-style searches: grid over the target keypoints (already in d_a), window scan of the queries in d_g gated by
// d_k, results in d_c / d_d / d_pkey / d_f
static int m_window_scan(hyorb_matcher *m, const hyorb_bounds &bounds, const uint8_t *t_desc, const float *t_uR, const uint8_t *t_matched, int nt,
                         const uint8_t *q_desc, int n, float thr, float ratio, const WindowCriteria &wc)
{
    constexpr int NC = HYORB_GRID_COLS * HYORB_GRID_ROWS;
    HY_TRY(m_upload(m, m->d_b, t_desc, (size_t)nt * 32));
    if (t_uR) HY_TRY(m_upload(m, m->d_l, t_uR, sizeof(float) * (size_t)nt));
    if (t_matched) HY_TRY(m_upload(m, m->d_bestd, t_matched, (size_t)nt));
    HY_TRY(m_upload(m, m->d_h, q_desc, (size_t)n * 32));
    HY_TRY(m->d_i.ensure(sizeof(int32_t) * (NC + 1)));
    HY_TRY(m->d_j.ensure(sizeof(int32_t) * std::max(nt, 1)));
    HY_TRY(m->d_cellof.ensure(sizeof(int32_t) * std::max(nt, 1)));
    HY_TRY(m->d_cellcnt.ensure(sizeof(int32_t) * NC));
    HY_TRY(m->d_c.ensure(sizeof(int32_t) * (size_t)n));
    HY_TRY(m->d_d.ensure(sizeof(uint16_t) * (size_t)n));
    HY_TRY(m->d_pkey.ensure(sizeof(uint16_t) * (size_t)n));
    HY_TRY(m->d_f.ensure((size_t)n));
    HY_TRY(launch_grid_build(m->d_a.as<hyorb_keypoint>(), nt, bounds, m->d_i.as<int32_t>(), m->d_j.as<int32_t>(), m->d_cellof.as<int32_t>(),
                             m->d_cellcnt.as<int32_t>(), m->stream, &m->launches));
    return launch_match_window(m->d_a.as<hyorb_keypoint>(), m->d_b.as<uint8_t>(), t_uR ? m->d_l.as<float>() : nullptr,
                               t_matched ? m->d_bestd.as<uint8_t>() : nullptr, nt, bounds, m->d_i.as<int32_t>(), m->d_j.as<int32_t>(),
                               m->d_g.as<hyorb_window_query>(), m->d_h.as<uint8_t>(), n, thr, ratio, m->d_c.as<int32_t>(), m->d_d.as<uint16_t>(),
                               m->d_pkey.as<uint16_t>(), m->d_f.as<uint8_t>(), m->stream, &m->launches, m->d_k.as<uint8_t>(), wc);
}

HYORB_API int hyorb_fuse_host(hyorb_matcher *m, const hyorb_projection *pr, const hyorb_landmark *lms, const float *lm_normal, const uint8_t *lm_desc, int n,
                              const hyorb_keypoint *t_kps, const uint8_t *t_desc, const float *t_uR, int nt, float th, float size_ref, float sigma_ref,
                              float reproj_err, float cos_max_angle, float thr, float ratio, int32_t *best_idx, uint16_t *best, uint16_t *second,
                              uint8_t *accepted, uint8_t *passed)
{
    HY_TRY(m_prepare(m));
    if (n < 0 || nt < 0 || !pr) { set_error("bad argument"); return HYORB_EINVAL; }
    if (n == 0) return HYORB_OK;
    if (!lms || !lm_normal || !lm_desc || !best_idx || !best || !second || !accepted || (nt > 0 && (!t_kps || !t_desc))) { set_error("null argument"); return HYORB_EINVAL; }
    HY_TRY(m_upload(m, m->d_e, lms, sizeof(hyorb_landmark) * (size_t)n));
    HY_TRY(m_upload(m, m->d_a, t_kps, sizeof(hyorb_keypoint) * (size_t)nt));
    HY_TRY(m_upload(m, m->d_psecond, lm_normal, sizeof(float) * 3 * (size_t)n));
    HY_TRY(m->d_g.ensure(sizeof(hyorb_window_query) * (size_t)n));
    HY_TRY(m->d_k.ensure((size_t)n));
    // no StereoConsistencyCriterion in Fuse (FeatureMatcher.cc:471-474): HYORB_SBP_STEREO stays off, the queries carry ur for the reprojection error only
    HY_TRY(launch_project_landmarks(*pr, m->d_e.as<hyorb_landmark>(), n, m->d_a.as<hyorb_keypoint>(), nt, th, size_ref, 0.5f, 1.5f,
                                    HYORB_SBP_DISTANCE | HYORB_SBP_VIEWANGLE, m->d_g.as<hyorb_window_query>(), m->d_k.as<uint8_t>(), m->d_status.as<int>(),
                                    m->stream, &m->launches, m->d_psecond.as<float>(), cos_max_angle));
    WindowCriteria wc;
    wc.rule = HYORB_RULE_LANDMARK; wc.reproj_thr = reproj_err; wc.sigma_ref = sigma_ref; wc.size_ref = size_ref;
    HY_TRY(m_window_scan(m, pr->bounds, t_desc, t_uR, nullptr, nt, lm_desc, n, thr, ratio, wc));
    HY_CUDA(cudaMemcpyAsync(best_idx, m->d_c.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(best, m->d_d.p, sizeof(uint16_t) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(second, m->d_pkey.p, sizeof(uint16_t) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(accepted, m->d_f.p, (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    if (passed) HY_CUDA(cudaMemcpyAsync(passed, m->d_k.p, (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    return m_sync(m);
}

HYORB_API int hyorb_search_by_sim3_host(hyorb_matcher *m, const float *R_a, const float *t_a, const float *sR_ba, const float *t_ba, const hyorb_projection *pr_b,
                                        const hyorb_landmark *lms, const uint8_t *lm_desc, int n, const hyorb_keypoint *kps_b, const uint8_t *desc_b, int nb,
                                        float th, float size_ref, float thr, int32_t *best_idx, uint16_t *best, uint8_t *accepted, uint8_t *passed)
{
    HY_TRY(m_prepare(m));
    if (n < 0 || nb < 0 || !pr_b || !R_a || !t_a || !sR_ba || !t_ba) { set_error("bad argument"); return HYORB_EINVAL; }
    if (n == 0) return HYORB_OK;
    if (!lms || !lm_desc || !best_idx || !best || !accepted || (nb > 0 && (!kps_b || !desc_b))) { set_error("null argument"); return HYORB_EINVAL; }
    HY_TRY(m_upload(m, m->d_e, lms, sizeof(hyorb_landmark) * (size_t)n));
    HY_TRY(m_upload(m, m->d_a, kps_b, sizeof(hyorb_keypoint) * (size_t)nb));
    HY_TRY(m->d_g.ensure(sizeof(hyorb_window_query) * (size_t)n));
    HY_TRY(m->d_k.ensure((size_t)n));
    HY_TRY(launch_project_sim3(R_a, t_a, sR_ba, t_ba, *pr_b, m->d_e.as<hyorb_landmark>(), n, m->d_a.as<hyorb_keypoint>(), nb, th, size_ref,
                               m->d_g.as<hyorb_window_query>(), m->d_k.as<uint8_t>(), m->d_status.as<int>(), m->stream, &m->launches));
    // `bestDist <= TH_HIGH` is the whole acceptance test (:840): the landmark rule with an infinite ratio never rejects on the second best
    HY_TRY(m_window_scan(m, pr_b->bounds, desc_b, nullptr, nullptr, nb, lm_desc, n, thr, INFINITY, WindowCriteria()));
    HY_CUDA(cudaMemcpyAsync(best_idx, m->d_c.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(best, m->d_d.p, sizeof(uint16_t) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(accepted, m->d_f.p, (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    if (passed) HY_CUDA(cudaMemcpyAsync(passed, m->d_k.p, (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    return m_sync(m);
}

HYORB_API int hyorb_search_for_initialization_host(hyorb_matcher *m, const hyorb_keypoint *k1, const uint8_t *d1, int n1, const hyorb_keypoint *k2,
                                                   const uint8_t *d2, int n2, hyorb_bounds bounds2, float *prev_xy, int window, float thr, float ratio,
                                                   int32_t *matches12, int32_t *n_matches)
{
    HY_TRY(m_prepare(m));
    if (n1 < 0 || n2 < 0 || window < 0) { set_error("bad argument"); return HYORB_EINVAL; }
    if (n_matches) *n_matches = 0;
    if (n1 == 0) return HYORB_OK;
    if (!k1 || !d1 || !prev_xy || !matches12 || (n2 > 0 && (!k2 || !d2))) { set_error("null argument"); return HYORB_EINVAL; }
    constexpr int NC = HYORB_GRID_COLS * HYORB_GRID_ROWS;
    HY_TRY(m_upload(m, m->d_kp1, k1, sizeof(hyorb_keypoint) * (size_t)n1));
    HY_TRY(m_upload(m, m->d_a, k2, sizeof(hyorb_keypoint) * (size_t)n2));
    HY_TRY(m_upload(m, m->d_h, d1, (size_t)n1 * 32));
    HY_TRY(m_upload(m, m->d_b, d2, (size_t)n2 * 32));
    HY_TRY(m_upload(m, m->d_l, prev_xy, sizeof(float) * 2 * (size_t)n1));
    HY_TRY(m->d_i.ensure(sizeof(int32_t) * (NC + 1)));
    HY_TRY(m->d_j.ensure(sizeof(int32_t) * std::max(n2, 1)));
    HY_TRY(m->d_cellof.ensure(sizeof(int32_t) * std::max(n2, 1)));
    HY_TRY(m->d_cellcnt.ensure(sizeof(int32_t) * NC));
    // claims, double-buffered: [claim2 | claimd] x 2; per-i2 list heads, per-i1 links, owner, result, flags
    HY_TRY(m->d_c.ensure(sizeof(int32_t) * 4 * (size_t)n1));
    HY_TRY(m->d_d.ensure(sizeof(int32_t) * (size_t)std::max(n2, 1)));      // head
    HY_TRY(m->d_e.ensure(sizeof(int32_t) * (size_t)n1));                   // next
    HY_TRY(m->d_rowtab.ensure(sizeof(int32_t) * (size_t)std::max(n2, 1))); // owner
    HY_TRY(m->d_g.ensure(sizeof(int32_t) * (size_t)n1));                   // matches12
    HY_TRY(m->d_f.ensure(sizeof(int) * 2));                                // changed, n_matches
    HY_TRY(launch_grid_build(m->d_a.as<hyorb_keypoint>(), n2, bounds2, m->d_i.as<int32_t>(), m->d_j.as<int32_t>(), m->d_cellof.as<int32_t>(),
                             m->d_cellcnt.as<int32_t>(), m->stream, &m->launches));
    int32_t *buf = m->d_c.as<int32_t>();
    HY_CUDA(cudaMemsetAsync(buf, 0xFF, sizeof(int32_t) * 4 * (size_t)n1, m->stream));      // no claims yet
    int cur = 0, passes = 0;
    for (;;) {
        int32_t *c2 = buf + (size_t)cur * 2 * n1, *cd = c2 + n1, *o2 = buf + (size_t)(cur ^ 1) * 2 * n1, *od = o2 + n1;
        HY_TRY(launch_mono_pass(m->d_h.as<uint8_t>(), n1, m->d_a.as<hyorb_keypoint>(), m->d_b.as<uint8_t>(), n2, bounds2, m->d_i.as<int32_t>(), m->d_j.as<int32_t>(),
                                m->d_l.as<float>(), (float)window, thr, ratio, c2, cd, m->d_d.as<int32_t>(), m->d_e.as<int32_t>(), o2, od, m->d_f.as<int>(),
                                m->stream, &m->launches));
        int changed = 0;
        HY_CUDA(cudaMemcpyAsync(&changed, m->d_f.p, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
        HY_CUDA(cudaStreamSynchronize(m->stream));
        cur ^= 1;
        ++passes;
        if (!changed) break;
        if (passes > n1 + 1) { set_error("mono initialisation matcher did not converge"); return HYORB_EUNSUPPORTED; }   // cannot happen: pass k fixes feature k
    }
    HY_TRY(launch_mono_finish(buf + (size_t)cur * 2 * n1, n1, m->d_kp1.as<hyorb_keypoint>(), m->d_a.as<hyorb_keypoint>(), n2, m->d_rowtab.as<int32_t>(),
                              m->d_g.as<int32_t>(), m->d_l.as<float>(), m->d_f.as<int>() + 1, m->d_status.as<int>(), m->stream, &m->launches));
    HY_CUDA(cudaMemcpyAsync(matches12, m->d_g.p, sizeof(int32_t) * (size_t)n1, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(prev_xy, m->d_l.p, sizeof(float) * 2 * (size_t)n1, cudaMemcpyDeviceToHost, m->stream));
    int nm = 0;
    HY_CUDA(cudaMemcpyAsync(&nm, m->d_f.as<int>() + 1, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    HY_TRY(m_sync(m));
    if (n_matches) *n_matches = nm;
    return HYORB_OK;
}

HYORB_API int hyorb_rotation_consistency_host(hyorb_matcher *m, const float *angle_prev, const float *angle_curr, int n, uint8_t *keep)
{
    HY_TRY(m_prepare(m));
    if (n < 0) { set_error("bad argument"); return HYORB_EINVAL; }
    if (n == 0) return HYORB_OK;
    if (!angle_prev || !angle_curr || !keep) { set_error("null argument"); return HYORB_EINVAL; }
    HY_TRY(m_upload(m, m->d_a, angle_prev, sizeof(float) * (size_t)n));
    HY_TRY(m_upload(m, m->d_b, angle_curr, sizeof(float) * (size_t)n));
    HY_TRY(m->d_f.ensure((size_t)n));
    HY_TRY(launch_rotation(m->d_a.as<float>(), m->d_b.as<float>(), n, m->d_f.as<uint8_t>(), m->d_status.as<int>(), m->stream, &m->launches));
    HY_CUDA(cudaMemcpyAsync(keep, m->d_f.p, (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    return m_sync(m);
}

struct hyorb_vocabulary {
    int device = 0, n_nodes = 0, L = 0;
    DevBuf d_off, d_idx, d_desc, d_word, d_weight;
};

HYORB_API int hyorb_vocabulary_create(int device, int n_nodes, int L, const int32_t *child_off, const int32_t *child_idx, const uint8_t *node_desc,
                                      const int32_t *word_of, const float *weight_of, hyorb_vocabulary **out)
{
    if (!out) { set_error("null argument"); return HYORB_EINVAL; }
    *out = nullptr;
    if (n_nodes < 1 || L < 0 || !child_off || !node_desc || !word_of || !weight_of) { set_error("bad argument"); return HYORB_EINVAL; }
    if (device < 0 || hyorb_device_count() <= device) { set_error("CUDA device %d not available (no CPU fallback exists)", device); return HYORB_ECUDA; }
    if (child_off[0] != 0) { set_error("child_off must start at 0"); return HYORB_EINVAL; }
    for (int i = 0; i < n_nodes; i++)
        if (child_off[i + 1] < child_off[i]) { set_error("child_off must be non-decreasing"); return HYORB_EINVAL; }
    const int n_edges = child_off[n_nodes];
    if (n_edges > 0 && !child_idx) { set_error("null argument"); return HYORB_EINVAL; }
    for (int e = 0; e < n_edges; e++)
        if (child_idx[e] <= 0 || child_idx[e] >= n_nodes) { set_error("child_idx[%d] = %d outside 1..%d", e, child_idx[e], n_nodes - 1); return HYORB_EINVAL; }
    for (int i = 0; i < n_nodes; i++)
        if (child_off[i + 1] - child_off[i] > (1 << 20)) { set_error("node %d has too many children", i); return HYORB_EUNSUPPORTED; }
    hyorb_vocabulary *v = new (std::nothrow) hyorb_vocabulary();
    if (!v) return HYORB_ENOMEM;
    v->device = device; v->n_nodes = n_nodes; v->L = L;
    int rc = HYORB_OK;
    auto up = [&](DevBuf &b, const void *src, size_t bytes) {
        if (rc) return;
        rc = b.ensure(std::max<size_t>(bytes, 16));
        if (!rc && bytes && cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("vocabulary upload failed"); rc = HYORB_ECUDA; }
    };
    if (cudaSetDevice(device) != cudaSuccess) rc = HYORB_ECUDA;
    up(v->d_off, child_off, sizeof(int32_t) * ((size_t)n_nodes + 1));
    up(v->d_idx, child_idx, sizeof(int32_t) * (size_t)n_edges);
    up(v->d_desc, node_desc, (size_t)n_nodes * 32);
    up(v->d_word, word_of, sizeof(int32_t) * (size_t)n_nodes);
    up(v->d_weight, weight_of, sizeof(float) * (size_t)n_nodes);
    // a pageable cudaMemcpy may return once the bytes are staged; the matcher's streams are non-blocking (not ordered after the
    // legacy stream), so wait for the uploads here
    if (!rc && cudaStreamSynchronize(cudaStreamLegacy) != cudaSuccess) { set_error("vocabulary upload failed"); rc = HYORB_ECUDA; }
    if (rc) { hyorb_vocabulary_destroy(v); return rc; }
    *out = v;
    return HYORB_OK;
}

HYORB_API int hyorb_vocabulary_destroy(hyorb_vocabulary *v)
{
    if (!v) return HYORB_OK;
    cudaSetDevice(v->device);
    DevBuf *bufs[] = {&v->d_off, &v->d_idx, &v->d_desc, &v->d_word, &v->d_weight};
    for (DevBuf *b : bufs) b->release();
    delete v;
    return HYORB_OK;
}

static int m_bow_descend(hyorb_matcher *m, const hyorb_vocabulary *v, const uint8_t *d_desc, int n, int levelsup, DevBuf &word, DevBuf &node, DevBuf &weight)
{
    HY_TRY(word.ensure(sizeof(int32_t) * std::max(n, 1)));
    HY_TRY(node.ensure(sizeof(int32_t) * std::max(n, 1)));
    HY_TRY(weight.ensure(sizeof(float) * std::max(n, 1)));
    return launch_bow_descend(v->d_off.as<int32_t>(), v->d_idx.as<int32_t>(), v->d_desc.as<uint8_t>(), v->d_word.as<int32_t>(), v->d_weight.as<float>(),
                              v->L - levelsup, d_desc, n, word.as<int32_t>(), node.as<int32_t>(), weight.as<float>(), m->stream, &m->launches);
}

HYORB_API int hyorb_bow_transform_host(hyorb_matcher *m, const hyorb_vocabulary *v, const uint8_t *desc, int n, int levelsup, int32_t *word_id,
                                       int32_t *node_id, float *weight)
{
    HY_TRY(m_prepare(m));
    if (!v || n < 0) { set_error("bad argument"); return HYORB_EINVAL; }
    if (v->device != m->device) { set_error("vocabulary lives on device %d, matcher on %d", v->device, m->device); return HYORB_EINVAL; }
    if (n == 0) return HYORB_OK;
    if (!desc || !word_id || !node_id || !weight) { set_error("null argument"); return HYORB_EINVAL; }
    HY_TRY(m_upload(m, m->d_a, desc, (size_t)n * 32));
    HY_TRY(m_bow_descend(m, v, m->d_a.as<uint8_t>(), n, levelsup, m->d_c, m->d_g, m->d_d));
    HY_CUDA(cudaMemcpyAsync(word_id, m->d_c.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(node_id, m->d_g.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(weight, m->d_d.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    return m_sync(m);
}

static int m_search_by_bow(hyorb_matcher *m, const hyorb_vocabulary *v, const uint8_t *desc1, const uint8_t *mask1, int n1,
                           const uint8_t *desc2, const uint8_t *mask2, int n2, int levelsup, int rule, float thr, float ratio,
                           int32_t *node1, int32_t *node2, int32_t *best_idx, uint16_t *best, uint16_t *second, uint8_t *accepted,
                           const EpipolarHost *eh)
{
    HY_TRY(m_prepare(m));
    if (!v || n1 < 0 || n2 < 0 || rule < 0 || rule > 2) { set_error("bad argument"); return HYORB_EINVAL; }
    if (v->device != m->device) { set_error("vocabulary lives on device %d, matcher on %d", v->device, m->device); return HYORB_EINVAL; }
    if (n1 == 0) return HYORB_OK;
    if (!desc1 || (n2 > 0 && !desc2) || !best_idx || !best || !second || !accepted) { set_error("null argument"); return HYORB_EINVAL; }
    HY_TRY(m_upload(m, m->d_a, desc1, (size_t)n1 * 32));
    HY_TRY(m_upload(m, m->d_b, desc2, (size_t)n2 * 32));
    if (mask1) HY_TRY(m_upload(m, m->d_k, mask1, (size_t)n1));
    if (mask2) HY_TRY(m_upload(m, m->d_l, mask2, (size_t)n2));
    // quantise both sets (words / weights are by-products here)
    HY_TRY(m_bow_descend(m, v, m->d_a.as<uint8_t>(), n1, levelsup, m->d_c, m->d_g, m->d_d));
    HY_TRY(m_bow_descend(m, v, m->d_b.as<uint8_t>(), n2, levelsup, m->d_e, m->d_h, m->d_f));
    const size_t tb = bow_sort_temp_bytes(std::max(n2, 1));
    HY_TRY(m->d_i.ensure(sizeof(int32_t) * std::max(n2, 1)));          // iota
    HY_TRY(m->d_j.ensure(sizeof(int32_t) * std::max(n2, 1)));          // sorted node ids of set 2
    HY_TRY(m->d_cellof.ensure(sizeof(int32_t) * std::max(n2, 1)));     // set-2 feature indices in (node, index) order
    HY_TRY(m->d_rowtab.ensure(std::max<size_t>(tb, 16)));
    HY_TRY(m->d_cellcnt.ensure(sizeof(int32_t) * (size_t)n1));         // candidate range begin
    HY_TRY(m->d_bestd.ensure(sizeof(int32_t) * (size_t)n1));           // candidate range end
    HY_TRY(m->d_pkey.ensure(sizeof(int32_t) * (size_t)n1));            // best_idx
    HY_TRY(m->d_psecond.ensure(sizeof(uint16_t) * 2 * (size_t)n1 + (size_t)n1));   // best | second | accepted
    uint16_t *d_best = m->d_psecond.as<uint16_t>(), *d_sec = d_best + n1;
    uint8_t *d_acc = (uint8_t *)(d_sec + n1);
    EpipolarDev epi;
    HY_TRY(m_epipolar_upload(m, eh, n1, n2, &epi));
    HY_TRY(launch_bow_match(m->d_a.as<uint8_t>(), mask1 ? m->d_k.as<uint8_t>() : nullptr, m->d_g.as<int32_t>(), n1, m->d_b.as<uint8_t>(),
                            mask2 ? m->d_l.as<uint8_t>() : nullptr, m->d_h.as<int32_t>(), n2, m->d_i.as<int32_t>(), m->d_j.as<int32_t>(),
                            m->d_cellof.as<int32_t>(), m->d_rowtab.p, tb, m->d_cellcnt.as<int32_t>(), m->d_bestd.as<int32_t>(), rule, thr, ratio,
                            m->d_pkey.as<int32_t>(), d_best, d_sec, d_acc, epi, m->stream, &m->launches));
    HY_CUDA(cudaMemcpyAsync(best_idx, m->d_pkey.p, sizeof(int32_t) * (size_t)n1, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(best, d_best, sizeof(uint16_t) * (size_t)n1, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(second, d_sec, sizeof(uint16_t) * (size_t)n1, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(accepted, d_acc, (size_t)n1, cudaMemcpyDeviceToHost, m->stream));
    if (node1) HY_CUDA(cudaMemcpyAsync(node1, m->d_g.p, sizeof(int32_t) * (size_t)n1, cudaMemcpyDeviceToHost, m->stream));
    if (node2 && n2 > 0) HY_CUDA(cudaMemcpyAsync(node2, m->d_h.p, sizeof(int32_t) * (size_t)n2, cudaMemcpyDeviceToHost, m->stream));
    return m_sync(m);
}

HYORB_API int hyorb_search_by_bow_host(hyorb_matcher *m, const hyorb_vocabulary *v, const uint8_t *desc1, const uint8_t *mask1, int n1,
                                       const uint8_t *desc2, const uint8_t *mask2, int n2, int levelsup, int rule, float thr, float ratio,
                                       int32_t *node1, int32_t *node2, int32_t *best_idx, uint16_t *best, uint16_t *second, uint8_t *accepted)
{
    return m_search_by_bow(m, v, desc1, mask1, n1, desc2, mask2, n2, levelsup, rule, thr, ratio, node1, node2, best_idx, best, second, accepted, nullptr);
}

HYORB_API int hyorb_search_for_triangulation_host(hyorb_matcher *m, const hyorb_vocabulary *v, const hyorb_keypoint *kps1, const uint8_t *desc1,
                                                  const uint8_t *mask1, int n1, const hyorb_keypoint *kps2, const uint8_t *desc2,
                                                  const uint8_t *mask2, int n2, int levelsup, const float *F12, float sigma_ref, float size_ref,
                                                  float thr, float ratio, int32_t *node1, int32_t *node2, int32_t *best_idx, uint16_t *best,
                                                  uint16_t *second, uint8_t *accepted)
{
    if (!F12 || (n1 > 0 && !kps1) || (n2 > 0 && !kps2)) { set_error("null argument"); return HYORB_EINVAL; }
    if (!(size_ref > 0.f)) { set_error("size_ref must be positive"); return HYORB_EINVAL; }
    const EpipolarHost eh{kps1, kps2, F12, sigma_ref, size_ref};
    return m_search_by_bow(m, v, desc1, mask1, n1, desc2, mask2, n2, levelsup, HYORB_RULE_BOW, thr, ratio, node1, node2, best_idx, best, second,
                           accepted, &eh);
}

HYORB_API int hyorb_distinctive_descriptor_host(hyorb_matcher *m, const uint8_t *desc, const int32_t *lm_off, int n_landmarks, int32_t *best_idx,
                                                int32_t *best_median)
{
    HY_TRY(m_prepare(m));
    if (n_landmarks < 0) { set_error("bad argument"); return HYORB_EINVAL; }
    if (n_landmarks == 0) return HYORB_OK;
    if (!lm_off || !best_idx || !best_median) { set_error("null argument"); return HYORB_EINVAL; }
    if (lm_off[0] != 0) { set_error("lm_off must start at 0"); return HYORB_EINVAL; }
    for (int i = 0; i < n_landmarks; i++)
        if (lm_off[i + 1] < lm_off[i]) { set_error("lm_off must be non-decreasing"); return HYORB_EINVAL; }
    const int total = lm_off[n_landmarks];
    if (total > 0 && !desc) { set_error("null argument"); return HYORB_EINVAL; }
    HY_TRY(m_upload(m, m->d_a, desc, (size_t)total * 32));
    HY_TRY(m_upload(m, m->d_g, lm_off, sizeof(int32_t) * ((size_t)n_landmarks + 1)));
    HY_TRY(m->d_c.ensure(sizeof(int32_t) * (size_t)n_landmarks));
    HY_TRY(m->d_h.ensure(sizeof(int32_t) * (size_t)n_landmarks));
    HY_TRY(launch_distinctive(m->d_a.as<uint8_t>(), m->d_g.as<int32_t>(), n_landmarks, m->d_c.as<int32_t>(), m->d_h.as<int32_t>(), m->stream, &m->launches));
    HY_CUDA(cudaMemcpyAsync(best_idx, m->d_c.p, sizeof(int32_t) * (size_t)n_landmarks, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(best_median, m->d_h.p, sizeof(int32_t) * (size_t)n_landmarks, cudaMemcpyDeviceToHost, m->stream));
    return m_sync(m);
}

HYORB_API int hyorb_stereo_match_batch_device(hyorb_matcher *m, const hyorb_stereo_params *sp, int n_pairs, const hyorb_keypoint *d_kps,
                                              const uint8_t *d_desc, const int32_t *d_counts, int capacity, float *d_uR, float *d_depth,
                                              int32_t *d_best_r, int32_t *d_best_dist)
{
    HY_TRY(m_prepare(m));
    if (!sp || n_pairs < 0 || capacity < 1 || !d_kps || !d_desc || !d_counts || !d_uR || !d_depth) { set_error("bad argument"); return HYORB_EINVAL; }
    if (n_pairs == 0) return HYORB_OK;
    const int rows = m->stereo_rows > 0 ? m->stereo_rows : 0;      // hyorb_stereo_match_host sets it from the keypoints it uploads; else the default budget
    HY_TRY(m->d_rowtab.ensure(sizeof(int32_t) * stereo_scratch_ints_per_pair(capacity, rows) * n_pairs));
    if (!d_best_dist) {
        HY_TRY(m->d_bestd.ensure(sizeof(int32_t) * (size_t)capacity * n_pairs));
        d_best_dist = m->d_bestd.as<int32_t>();
    }
    return launch_stereo(*sp, n_pairs, d_kps, d_desc, d_counts, capacity, m->d_rowtab.as<int32_t>(), d_uR, d_depth, d_best_r, d_best_dist,
                         m->d_status.as<int>(), m->stream, &m->launches, rows);
}

HYORB_API int hyorb_stereo_match_host(hyorb_matcher *m, const hyorb_stereo_params *sp, const hyorb_keypoint *kps_l, const uint8_t *desc_l, int n_l,
                                      const hyorb_keypoint *kps_r, const uint8_t *desc_r, int n_r, float *uR, float *depth, int32_t *best_r,
                                      int32_t *best_dist)
{
    HY_TRY(m_prepare(m));
    if (!sp || n_l < 0 || n_r < 0) { set_error("bad argument"); return HYORB_EINVAL; }
    if (n_l == 0) return HYORB_OK;
    if (!kps_l || !desc_l || !uR || !depth || (n_r > 0 && (!kps_r || !desc_r))) { set_error("null argument"); return HYORB_EINVAL; }
    const int cap = std::max(std::max(n_l, n_r), 1);
    // lay the pair out the way hyorb_extract_batch_device does: image 0 = left, image 1 = right, stride = cap
    HY_TRY(m->d_a.ensure(sizeof(hyorb_keypoint) * (size_t)cap * 2));
    HY_TRY(m->d_b.ensure((size_t)32 * cap * 2));
    HY_TRY(m->d_c.ensure(sizeof(int32_t) * 2));
    HY_TRY(m->d_d.ensure(sizeof(float) * (size_t)cap));
    HY_TRY(m->d_e.ensure(sizeof(float) * (size_t)cap));
    HY_TRY(m->d_g.ensure(sizeof(int32_t) * (size_t)cap));
    HY_TRY(m->d_h.ensure(sizeof(int32_t) * (size_t)cap));
    const int32_t cnt[2] = {n_l, n_r};
    HY_CUDA(cudaMemcpyAsync(m->d_a.p, kps_l, sizeof(hyorb_keypoint) * (size_t)n_l, cudaMemcpyHostToDevice, m->stream));
    HY_CUDA(cudaMemcpyAsync(m->d_b.p, desc_l, (size_t)32 * n_l, cudaMemcpyHostToDevice, m->stream));
    if (n_r) {
        HY_CUDA(cudaMemcpyAsync(m->d_a.as<hyorb_keypoint>() + cap, kps_r, sizeof(hyorb_keypoint) * (size_t)n_r, cudaMemcpyHostToDevice, m->stream));
        HY_CUDA(cudaMemcpyAsync(m->d_b.as<uint8_t>() + (size_t)32 * cap, desc_r, (size_t)32 * n_r, cudaMemcpyHostToDevice, m->stream));
    }
    HY_CUDA(cudaMemcpyAsync(m->d_c.p, cnt, sizeof(cnt), cudaMemcpyHostToDevice, m->stream));
    HY_CUDA(cudaStreamSynchronize(m->stream));     // cnt is a stack array
    float max_size = 0.f;
    for (int i = 0; i < n_r; i++) max_size = std::max(max_size, kps_r[i].size);
    m->stereo_rows = stereo_rows_budget(max_size, sp->size_ref);
    const int rc = hyorb_stereo_match_batch_device(m, sp, 1, m->d_a.as<hyorb_keypoint>(), m->d_b.as<uint8_t>(), m->d_c.as<int32_t>(), cap,
                                                   m->d_d.as<float>(), m->d_e.as<float>(), m->d_g.as<int32_t>(), m->d_h.as<int32_t>());
    m->stereo_rows = 0;
    HY_TRY(rc);
    HY_CUDA(cudaMemcpyAsync(uR, m->d_d.p, sizeof(float) * (size_t)n_l, cudaMemcpyDeviceToHost, m->stream));
    HY_CUDA(cudaMemcpyAsync(depth, m->d_e.p, sizeof(float) * (size_t)n_l, cudaMemcpyDeviceToHost, m->stream));
    if (best_r) HY_CUDA(cudaMemcpyAsync(best_r, m->d_g.p, sizeof(int32_t) * (size_t)n_l, cudaMemcpyDeviceToHost, m->stream));
    if (best_dist) HY_CUDA(cudaMemcpyAsync(best_dist, m->d_h.p, sizeof(int32_t) * (size_t)n_l, cudaMemcpyDeviceToHost, m->stream));
    return m_sync(m);
}

}  // extern "C"
