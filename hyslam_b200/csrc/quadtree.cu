// quadtree.cu -- K3: keypoint distribution.  Replaces ORBExtractor::DistributeOctTree + ExtractorNode::DivideNode
// (src/features/ORBExtractor.cpp:179-403, 121-177) and the tail of ComputeKeyPointsOctTree (:472-490).
//
// The reference grows a std::list of nodes one DivideNode at a time.  This kernel computes the same leaves, the
// same kept keypoint per leaf and the same OUTPUT ORDER level-synchronously, one CTA per (image, pyramid level),
// from four facts about the reference (proved equal to a literal list implementation on the CPU by
// tests/test_quadtree_closed_form.py against oracle/quadtree_closed_form.c, which mirrors this file):
//   1. child bounds depend only on the parent's bounds, x and y split independently => a candidate's path is
//      root || interleave(rowbits, colbits), looked up from two data-independent tables (tables.cu);
//   2. every pass splits all multi-point leaves (phase 1) or a sorted prefix of them (phase 2), so they always
//      sit at one depth;
//   3. children are push_front'ed: the list is [deepest leaves, newest first] ... [root leaves], and a phase-1
//      pass visits the multi-point nodes newest first;
//   4. creation rank of a child = (visiting position of its parent, quadrant); the canonical tie policy
//      (node address order == creation order) makes the phase-2 sort key (count, creation rank).
// Per pass: warp-aggregated shared-memory histogram of candidates into 4 child slots per visited parent, block
// scan of the non-empty slots (= new creation ranks), ballot/scan compaction of singletons into the leaf list.
#include "common.cuh"

namespace hyorb {

constexpr int QT = QT_THREADS;
constexpr unsigned NODE_FINAL = 0xFFFFu;   // candidate already sits in an emitted leaf
constexpr unsigned NODE_STAY = 0x4000u;    // candidate stays in an unexpanded node of the previous depth (last pass only)
constexpr unsigned NODE_MASK = 0x1FFFu;

struct QtShared {
    int warp[QT / 32 + 1];
    int rootcnt[QT_MAX_ROOTS];
    int rootcrank[QT_MAX_ROOTS];
    int nleaf;
    int cut;
    int tmp;
};

// exclusive scan of one int per thread across the block; total returned to every thread
__device__ __forceinline__ int block_excl_scan(int v, int *s_warp, int &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();                       // protect s_warp from the previous use
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < QT / 32 ? s_warp[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < QT / 32) s_warp[lane] = winc - w;
        if (lane == QT / 32 - 1) s_warp[QT / 32] = winc;
    }
    __syncthreads();
    total = s_warp[QT / 32];
    return s_warp[wid] + inc - v;
}

// exclusive scan over i in [0,n) of value(i); emit(i, exclusive_prefix); returns the total.  Each thread owns a
// contiguous chunk.
template <typename V, typename E>
__device__ __forceinline__ int chunk_scan(int n, int *s_warp, V value, E emit)
{
    const int per = (n + QT - 1) / QT;
    const int lo = min((int)threadIdx.x * per, n), hi = min(lo + per, n);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += value(i);
    int total;
    int base = block_excl_scan(sum, s_warp, total);
    for (int i = lo; i < hi; i++) { emit(i, base); base += value(i); }
    return total;
}

// in-place bitonic sort, DESCENDING by key, n2 a power of two; all threads of the block must call
__device__ void bitonic_desc(uint32_t *key, uint32_t *val, int n2)
{
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += QT) {
                const int p = i ^ j;
                if (p > i) {
                    const uint32_t a = key[i], b = key[p];
                    const bool desc = ((i & k) == 0);
                    if ((a < b) == desc) {
                        key[i] = b; key[p] = a;
                        const uint32_t t = val[i]; val[i] = val[p]; val[p] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int pow2_at_least(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// warp-aggregated shared-memory counter increment: lanes that hit the same slot elect one leader.  Must be called by
// all 32 lanes; slot < 0 = this lane has nothing to add.
__device__ __forceinline__ void agg_inc(uint32_t *arr, int slot)
{
    const unsigned peers = __match_any_sync(0xffffffffu, slot);
    if (slot >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&arr[slot], (uint32_t)__popc(peers));
}

__global__ void __launch_bounds__(QT)
k_quadtree(const PlanDev *__restrict__ plan, const uint32_t *__restrict__ cand_all, const int *__restrict__ candCount,
           const uint32_t *__restrict__ lut, uint32_t *__restrict__ qcode_all, uint16_t *__restrict__ qnode_all,
           uint2 *__restrict__ qleaf_all, uint32_t *__restrict__ sel_all, int *__restrict__ selCount, int *__restrict__ status)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ QtShared S;

    const int l = blockIdx.x, b = blockIdx.y;
    const LevelDev &L = plan->lv[l];
    const int maxN = L.qtMaxN;
    const int tid = threadIdx.x;

    // dynamic shared memory carve-up (20 bytes per node slot)
    uint32_t *cnt = (uint32_t *)smem_raw;                    // [maxN] points per node of the current depth, by creation rank
    uint32_t *slotcnt = cnt + maxN;                          // [maxN] points per child slot (4 per visited parent)
    uint32_t *skey = slotcnt + maxN;                         // [maxN] sort keys          } alias: best[maxN] (u64)
    uint32_t *sval = skey + maxN;                            // [maxN] sort payload       }
    unsigned long long *best = (unsigned long long *)skey;
    uint16_t *pidx = (uint16_t *)(sval + maxN);              // [maxN] visiting position of a multi-point node
    uint16_t *slotnew = pidx + maxN;                         // [maxN] creation rank of the child in a slot

    int n = candCount[b * HYORB_MAX_LEVELS + l];
    if (n > L.candCap) n = L.candCap;
    int *outCount = selCount + b * HYORB_MAX_LEVELS + l;
    if (n == 0) { if (tid == 0) *outCount = 0; return; }

    const uint32_t *cand = cand_all + (size_t)b * plan->candStride + L.candOff;
    uint32_t *qcode = qcode_all + (size_t)b * plan->candStride + L.candOff;
    uint16_t *qnode = qnode_all + (size_t)b * plan->candStride + L.candOff;
    uint2 *qleaf = qleaf_all + (size_t)b * plan->selStride + L.selOff;
    uint32_t *sel = sel_all + (size_t)b * plan->selStride + L.selOff;
    const uint32_t *lutX = lut + L.lutX, *lutY = lut + L.lutY;
    const int N = L.quota;

    // ---------------- depth 0: roots (:192-225)
    if (tid < QT_MAX_ROOTS) S.rootcnt[tid] = 0;
    if (tid == 0) { S.nleaf = 0; }
    __syncthreads();
    for (int base = 0; base < n; base += QT) {
        const int i = base + tid;
        int r = -1;
        if (i < n) {
            const uint32_t c = cand[i];
            const uint32_t code = lutX[cand_x(c)] | lutY[cand_y(c)];
            qcode[i] = code;
            r = (int)(code >> (2 * QT_DMAX));
        }
        agg_inc((uint32_t *)S.rootcnt, r);   // neighbouring candidates mostly share the root
    }
    __syncthreads();
    if (tid == 0) {
        int F0 = 0;
        for (int r = L.nIni - 1; r >= 0; r--)
            if (S.rootcnt[r]) { S.rootcrank[r] = F0; cnt[F0] = (uint32_t)S.rootcnt[r]; F0++; }
        S.tmp = F0;
    }
    __syncthreads();
    int F = S.tmp;            // nodes of the current depth
    for (int i = tid; i < n; i += QT) {
        const int r = (int)(qcode[i] >> (2 * QT_DMAX));
        unsigned nd = (unsigned)S.rootcrank[r];
        if (S.rootcnt[r] == 1) {
            const int pos = atomicAdd(&S.nleaf, 1);
            if (pos < L.selCap) qleaf[pos] = make_uint2(((0u << 13) | nd) + 1u, (uint32_t)i);
            nd = NODE_FINAL;
        }
        qnode[i] = (uint16_t)nd;
    }
    __syncthreads();

    int size = F, depth = 0, M = 0, m = 0;
    bool phase2 = false, finish = false;
    while (!finish) {
        // ---- A. multi-point nodes of this depth and their visiting order
        M = chunk_scan(F, S.warp, [&](int c) { return cnt[c] > 1 ? 1 : 0; },
                       [&](int c, int pos) { if (cnt[c] > 1) { skey[pos] = ((phase2 ? cnt[c] : 0u) << 13) | (uint32_t)c; } });
        __syncthreads();
        if (phase2) {
            // sort(vPrevSizeAndPointerToNode) then walk from the back (:324-326): (count desc, creation rank desc)
            const int n2 = pow2_at_least(M);
            for (int i = M + tid; i < n2; i += QT) skey[i] = 0;
            __syncthreads();
            bitonic_desc(skey, sval, n2);
            for (int p = tid; p < M; p += QT) pidx[skey[p] & NODE_MASK] = (uint16_t)p;
        } else {
            // list order: newest first (:257)
            for (int p = tid; p < M; p += QT) pidx[skey[p] & NODE_MASK] = (uint16_t)(M - 1 - p);
        }
        for (int s = tid; s < 4 * M; s += QT) slotcnt[s] = 0;
        __syncthreads();
        // ---- C. histogram candidates into the 4 child slots of their node (DivideNode :151-166)
        const int sh = 2 * (QT_DMAX - 1 - depth);
        for (int base = 0; base < n; base += QT) {
            const int i = base + tid;
            int slot = -1;
            if (i < n) {
                const unsigned nd = qnode[i];
                if (nd != NODE_FINAL) slot = 4 * (int)pidx[nd] + (int)((qcode[i] >> sh) & 3u);
            }
            agg_inc(slotcnt, slot);
        }
        __syncthreads();
        // ---- D. how many parents are expanded: all (phase 1) or until the list holds N nodes (:370-371)
        if (tid == 0) S.cut = M;
        const int grow = chunk_scan(M, S.warp,
            [&](int p) { return (int)(slotcnt[4 * p] > 0) + (int)(slotcnt[4 * p + 1] > 0) + (int)(slotcnt[4 * p + 2] > 0) + (int)(slotcnt[4 * p + 3] > 0) - 1; },
            [&](int p, int before) {
                if (phase2) {
                    const int k1 = (int)(slotcnt[4 * p] > 0) + (int)(slotcnt[4 * p + 1] > 0) + (int)(slotcnt[4 * p + 2] > 0) + (int)(slotcnt[4 * p + 3] > 0) - 1;
                    if (size + before < N && size + before + k1 >= N) S.cut = p + 1;   // unique p: the prefix is monotone
                }
            });
        __syncthreads();
        m = S.cut;
        // ---- E. creation ranks of the children, in (visiting order, quadrant) order
        int nToExpand = 0;
        {
            int multi_local = 0;
            const int F2 = chunk_scan(4 * m, S.warp, [&](int s) { return slotcnt[s] > 0 ? 1 : 0; },
                                      [&](int s, int pos) {
                                          if (slotcnt[s] > 0) { slotnew[s] = (uint16_t)pos; cnt[pos] = slotcnt[s]; if (slotcnt[s] > 1) multi_local++; }
                                      });
            // cnt[] (old depth) is dead after step A, so it was safe to overwrite above
            int tot;
            block_excl_scan(multi_local, S.warp, tot);
            nToExpand = tot;
            // size after this pass
            int nsize;
            if (m == M) nsize = size + grow;
            else {
                // sum of (k-1) over the first m parents = F2 - m
                nsize = size + (F2 - m);
            }
            __syncthreads();
            // ---- F. move candidates to their child; singletons become leaves (bNoMore, :168-175)
            for (int i = tid; i < n; i += QT) {
                unsigned nd = qnode[i];
                if (nd == NODE_FINAL) continue;
                const int p = pidx[nd];
                if (p >= m) { qnode[i] = (uint16_t)(nd | NODE_STAY); continue; }
                const int slot = 4 * p + (int)((qcode[i] >> sh) & 3u);
                nd = slotnew[slot];
                if (slotcnt[slot] == 1) {
                    const int pos = atomicAdd(&S.nleaf, 1);
                    if (pos < L.selCap) qleaf[pos] = make_uint2((((uint32_t)(depth + 1) << 13) | nd) + 1u, (uint32_t)i);
                    nd = NODE_FINAL;
                }
                qnode[i] = (uint16_t)nd;
            }
            __syncthreads();
            const int prevSize = size;
            size = nsize; depth++; F = F2;
            if (size >= N || size == prevSize) finish = true;                 // :309, :374
            else if (!phase2 && size + 3 * nToExpand > N) phase2 = true;       // :313
            if (!finish && (depth >= QT_DMAX || F > maxN || 4 * F > 4 * maxN)) { if (tid == 0) atomicOr(status, ST_QT_LIMIT); finish = true; }
        }
    }

    // ---------------- leaves that still hold several candidates: keep the max response, first in the reference's
    // candidate order on ties (:381-400)
    const int nStay = M - m;                      // unexpanded nodes of depth-1 (phase-2 cut)
    const int nMulti = nStay + F;                 // ids: [0,nStay) unexpanded parents, [nStay, nStay+F) nodes of the last depth
    if (nMulti > maxN) { if (tid == 0) { atomicOr(status, ST_QT_LIMIT); *outCount = 0; } return; }
    for (int i = tid; i < nMulti; i += QT) best[i] = 0ull;
    __syncthreads();
    for (int i = tid; i < n; i += QT) {
        const unsigned nd = qnode[i];
        if (nd == NODE_FINAL) continue;
        const int id = (nd & NODE_STAY) ? (int)pidx[nd & NODE_MASK] - m : nStay + (int)nd;
        const uint32_t c = cand[i];
        const uint32_t ok = cand_order_key(cand_x(c), cand_y(c), L.wCell, L.hCell, L.nCols);
        const unsigned long long v = ((unsigned long long)cand_resp(c) << 32) | (unsigned long long)(0xFFFFFFFFu - ok);
        atomicMax(&best[id], v);
    }
    __syncthreads();
    for (int i = tid; i < n; i += QT) {
        const unsigned nd = qnode[i];
        if (nd == NODE_FINAL) continue;
        const bool stay = (nd & NODE_STAY) != 0;
        const int id = stay ? (int)pidx[nd & NODE_MASK] - m : nStay + (int)nd;
        const uint32_t c = cand[i];
        const uint32_t ok = cand_order_key(cand_x(c), cand_y(c), L.wCell, L.hCell, L.nCols);
        const unsigned long long v = ((unsigned long long)cand_resp(c) << 32) | (unsigned long long)(0xFFFFFFFFu - ok);
        if (best[id] == v) {
            const int pos = atomicAdd(&S.nleaf, 1);
            const uint32_t key = stay ? (((uint32_t)(depth - 1) << 13) | (nd & NODE_MASK)) : (((uint32_t)depth << 13) | nd);
            if (pos < L.selCap) qleaf[pos] = make_uint2(key + 1u, (uint32_t)i);
        }
    }
    __syncthreads();

    // ---------------- output in list order: deepest leaves first, newest first inside a depth
    const int nleaf = S.nleaf;
    if (nleaf != size && tid == 0) atomicOr(status, ST_QT_MISMATCH);   // internal consistency: one kept keypoint per leaf
    if (nleaf > L.selCap || nleaf > maxN) { if (tid == 0) { atomicOr(status, ST_SEL_OVERFLOW); *outCount = 0; } return; }
    const int n2 = pow2_at_least(nleaf);
    for (int i = tid; i < n2; i += QT) {
        if (i < nleaf) { const uint2 e = qleaf[i]; skey[i] = e.x; sval[i] = e.y; }
        else { skey[i] = 0; sval[i] = 0; }
    }
    __syncthreads();
    bitonic_desc(skey, sval, n2);
    for (int i = tid; i < nleaf; i += QT) sel[i] = cand[sval[i]];
    if (tid == 0) *outCount = nleaf;
}

int launch_quadtree(const PlanDev &hp, const PlanDev *dp, const uint32_t *cand, const int *candCount, const uint32_t *lut,
                    uint32_t *qcode, uint16_t *qnode, uint2 *qleaf, uint32_t *sel, int *selCount, int *status, int B,
                    cudaStream_t st, long *launches)
{
    int maxN = 0;
    for (int l = 0; l < hp.nlevels; l++) if (hp.lv[l].qtMaxN > maxN) maxN = hp.lv[l].qtMaxN;
    const size_t smem = (size_t)maxN * 20;
    static bool attr_set = false;   // idempotent: raising the limit again is harmless
    if (!attr_set || smem > 48 * 1024) {
        HY_CUDA(cudaFuncSetAttribute(k_quadtree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        attr_set = true;
    }
    dim3 grd(hp.nlevels, B);
    k_quadtree<<<grd, QT, smem, st>>>(dp, cand, candCount, lut, qcode, qnode, qleaf, sel, selCount, status);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

}  // namespace hyorb
