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

#ifndef HYORB_QT_MINB
#define HYORB_QT_MINB 3      // 40 registers, 3 CTAs per SM; 4 x 32 registers measured the same (0.486 vs 0.482 ms): the kernel is not occupancy-bound
#endif
constexpr int QT_LAT = 1024;      // threads per CTA when only a few images are in flight: per-frame latency over occupancy
constexpr unsigned NODE_FINAL = 0xFFFFu;   // candidate already sits in an emitted leaf
constexpr unsigned NODE_STAY = 0x4000u;    // candidate stays in an unexpanded node of the previous depth (last pass only)
constexpr unsigned NODE_MASK = 0x1FFFu;

struct QtShared {
    int warp[QT_LAT / 32 + 1];
    int rootcnt[QT_MAX_ROOTS];
    int rootcrank[QT_MAX_ROOTS];
    int rootmi[QT_MAX_ROOTS];
    int nleaf;
    int cut;
    int tmp;
};

// exclusive scan of one int per thread across the block; total returned to every thread
template <int QT>
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
template <int QT, typename V, typename E>
__device__ __forceinline__ int chunk_scan(int n, int *s_warp, V value, E emit)
{
    const int per = (n + QT - 1) / QT;
    const int lo = min((int)threadIdx.x * per, n), hi = min(lo + per, n);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += value(i);
    int total;
    int base = block_excl_scan<QT>(sum, s_warp, total);
    for (int i = lo; i < hi; i++) { emit(i, base); base += value(i); }
    return total;
}

// in-place bitonic sort, DESCENDING by key, n2 a power of two; all threads of the block must call
template <int QT>
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

// shared-memory counter increment; slot < 0 = nothing to add.  (A warp-aggregated version -- __match_any_sync, one atomic per
// distinct slot -- measured slower: 0.54 vs 0.47 ms per 256 C2 images; Blackwell's shared-memory atomics absorb the
// same-address runs of neighbouring candidates.)
__device__ __forceinline__ void agg_inc(uint32_t *arr, int slot)
{
    if (slot >= 0) atomicAdd(&arr[slot], 1u);
}

template <int QT>
__global__ void __launch_bounds__(QT, QT == 512 ? HYORB_QT_MINB : 1)
k_quadtree(const PlanDev *__restrict__ plan, const uint32_t *__restrict__ cand_all, const int *__restrict__ candCount,
           const uint32_t *__restrict__ lut, uint32_t *__restrict__ qcode_all, uint16_t *__restrict__ qnode_all,
           uint2 *__restrict__ qleaf_all, uint32_t *__restrict__ sel_all, int *__restrict__ selCount, int *__restrict__ status)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ QtShared S;

    grid_dependency_wait();      // launch_dependent (common.cuh): follows k_fast
    const int l = blockIdx.x, b = blockIdx.y;
    const LevelDev &L = plan->lv[l];
    const int maxN = L.qtMaxN;
    const int tid = threadIdx.x;

    // dynamic shared memory carve-up (QT_SMEM_PER_SLOT = 26 bytes per node slot).  "mi" = index of a multi-point node
    // among the multi-point nodes of its depth, in creation order; a node's four child slots are 4*mi + quadrant.
    uint32_t *Hbase = (uint32_t *)smem_raw;                  // [maxN] x2: points per child slot, current / next depth
    uint32_t *skey = (uint32_t *)smem_raw + 2 * maxN;        // [maxN] sort keys          } alias: best[maxN] (u64)
    uint32_t *sval = skey + maxN;                            // [maxN] sort payload       }
    unsigned long long *best = (unsigned long long *)skey;
    uint16_t *crankBase = (uint16_t *)(sval + maxN);         // [maxN] x2: creation rank of multi node mi, current / next depth
    uint16_t *P = crankBase + 2 * maxN;                      // [maxN] visiting position of multi node mi in this pass
    uint16_t *ncrank = P + maxN;                             // [maxN] creation rank of the child in a slot
    uint16_t *nmi = ncrank + maxN;                           // [maxN] mi of the child in a slot (if it holds > 1 point)

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
    const int nIni = L.nIni;

    // ---------------- depth 0: roots (:192-225).  One sweep computes every candidate's path code and histograms it
    // into its root and into the root's four children.
    if (tid < QT_MAX_ROOTS) S.rootcnt[tid] = 0;
    for (int i = tid; i < 4 * nIni; i += QT) Hbase[maxN + i] = 0;
    if (tid == 0) S.nleaf = 0;
    __syncthreads();
    for (int base = 0; base < n; base += QT) {
        const int i = base + tid;
        int r = -1, s1 = -1;
        if (i < n) {
            const uint32_t c = cand[i];
            const uint32_t code = lutX[cand_x(c)] | lutY[cand_y(c)];
            qcode[i] = code;
            r = (int)(code >> (2 * QT_DMAX));
            s1 = 4 * r + (int)((code >> (2 * (QT_DMAX - 1))) & 3u);
        }
        agg_inc((uint32_t *)S.rootcnt, r);   // neighbouring candidates mostly share the root
        agg_inc(Hbase + maxN, s1);
    }
    __syncthreads();
    if (tid == 0) {
        // creation rank of the roots counts from the right, so that "visit newest first" is ascending root index, which is what
        // the reference's first pass does; mi enumerates the multi-point roots in creation order
        int F0 = 0, M0 = 0;
        for (int r = nIni - 1; r >= 0; r--) {
            S.rootmi[r] = -1;
            if (S.rootcnt[r]) {
                S.rootcrank[r] = F0;
                if (S.rootcnt[r] > 1) { S.rootmi[r] = M0; crankBase[M0] = (uint16_t)F0; M0++; }
                F0++;
            }
        }
        S.tmp = F0; S.cut = M0;
    }
    __syncthreads();
    for (int i = tid; i < 4 * nIni; i += QT) {          // child histograms re-indexed by mi
        const int r = i >> 2;
        if (S.rootmi[r] >= 0) Hbase[4 * S.rootmi[r] + (i & 3)] = Hbase[maxN + i];
    }
    __syncthreads();

    int size = S.tmp, depth = 0, M = S.cut, m = 0, cur = 0, M2 = 0;
    bool phase2 = false, finish = false;
    __syncthreads();          // S.tmp / S.cut are reused below
    while (true) {
        uint32_t *H = Hbase + cur * maxN, *Hn = Hbase + (cur ^ 1) * maxN;
        // ---- A. visiting order of the multi-point nodes: skey[p] = mi of the p-th visited node
        if (phase2) {
            // sort(vPrevSizeAndPointerToNode) then walk from the back (:324-326): (count desc, creation rank desc)
            const int n2 = pow2_at_least(M);
            for (int i = tid; i < n2; i += QT)
                skey[i] = i < M ? (((H[4 * i] + H[4 * i + 1] + H[4 * i + 2] + H[4 * i + 3]) << 13) | (uint32_t)i) : 0u;
            __syncthreads();
            bitonic_desc<QT>(skey, sval, n2);
            for (int p = tid; p < M; p += QT) { const uint32_t mi = skey[p] & NODE_MASK; skey[p] = mi; P[mi] = (uint16_t)p; }
        } else {
            // list order: newest first (:257)
            for (int p = tid; p < M; p += QT) { skey[p] = (uint32_t)(M - 1 - p); P[M - 1 - p] = (uint16_t)p; }
        }
        __syncthreads();
        // ---- B. how many parents are expanded: all (phase 1) or until the list holds N nodes (:370-371)
        if (tid == 0) S.cut = M;
        auto kids = [&](int p) { const uint32_t *h4 = H + 4 * skey[p]; return (int)(h4[0] > 0) + (int)(h4[1] > 0) + (int)(h4[2] > 0) + (int)(h4[3] > 0); };
        chunk_scan<QT>(M, S.warp, [&](int p) { return kids(p) - 1; },
                   [&](int p, int before) {
                       if (phase2) { const int k1 = kids(p) - 1; if (size + before < N && size + before + k1 >= N) S.cut = p + 1; }   // unique: the prefix is monotone
                   });
        __syncthreads();
        m = S.cut;
        // ---- C. creation ranks of the children, in (visiting order, quadrant) order; low half counts non-empty slots,
        // high half counts slots with > 1 point (= the next depth's mi)
        uint16_t *crankNext = crankBase + (cur ^ 1) * maxN;
        const int tot = chunk_scan<QT>(4 * m, S.warp,
            [&](int t) { const uint32_t c = H[4 * skey[t >> 2] + (t & 3)]; return (int)(c > 0) + ((int)(c > 1) << 16); },
            [&](int t, int pos) {
                const int slot = 4 * (int)skey[t >> 2] + (t & 3);
                const uint32_t c = H[slot];
                if (c > 0) ncrank[slot] = (uint16_t)(pos & 0xFFFF);
                if (c > 1) { nmi[slot] = (uint16_t)(pos >> 16); crankNext[pos >> 16] = (uint16_t)(pos & 0xFFFF); }
            });
        const int F2 = tot & 0xFFFF;
        M2 = tot >> 16;
        // the list size after this pass is known before the candidates move: decide now whether another pass follows
        const int prevSize = size;
        size += F2 - m;                      // sum over expanded parents of (children - 1)
        if (size >= N || size == prevSize) finish = true;                 // :309, :374
        else if (!phase2 && size + 3 * M2 > N) phase2 = true;              // :313 (nToExpand = children with > 1 point)
        if (!finish && (depth + 1 >= QT_DMAX || 4 * M2 > maxN)) { if (tid == 0) atomicOr(status, ST_QT_LIMIT); finish = true; }
        // another pass => size < N => M2 < N => the next histogram (4*M2 slots) fits
        if (!finish) for (int i = tid; i < 4 * M2; i += QT) Hn[i] = 0;
        __syncthreads();
        // ---- D. one sweep: move every candidate to its child (DivideNode :151-166), make singletons leaves (bNoMore,
        // :168-175) and histogram the survivors into their child's four children for the next pass
        const int sh = 2 * (QT_DMAX - 1 - depth);
        for (int base = 0; base < n; base += QT) {
            const int i = base + tid;
            int nslot = -1;
            if (i < n) {
                const uint32_t code = qcode[i];
                unsigned nd;
                if (depth == 0) {
                    const int r = (int)(code >> (2 * QT_DMAX));
                    nd = S.rootmi[r] >= 0 ? (unsigned)S.rootmi[r] : NODE_FINAL;
                    if (S.rootcnt[r] == 1) {
                        const int pos = atomicAdd(&S.nleaf, 1);
                        if (pos < L.selCap) qleaf[pos] = make_uint2(((0u << 13) | (uint32_t)S.rootcrank[r]) + 1u, (uint32_t)i);
                    }
                    if (nd == NODE_FINAL) qnode[i] = (uint16_t)NODE_FINAL;
                } else nd = qnode[i];
                if (nd != NODE_FINAL) {
                    if ((int)P[nd] >= m) qnode[i] = (uint16_t)(nd | NODE_STAY);      // parent not expanded: only in the last pass
                    else {
                        const int slot = 4 * (int)nd + (int)((code >> sh) & 3u);
                        if (H[slot] == 1) {
                            const int pos = atomicAdd(&S.nleaf, 1);
                            if (pos < L.selCap) qleaf[pos] = make_uint2((((uint32_t)(depth + 1) << 13) | ncrank[slot]) + 1u, (uint32_t)i);
                            qnode[i] = (uint16_t)NODE_FINAL;
                        } else {
                            const unsigned c2 = nmi[slot];
                            qnode[i] = (uint16_t)c2;
                            if (!finish && sh >= 2) nslot = 4 * (int)c2 + (int)((code >> (sh - 2)) & 3u);
                        }
                    }
                }
            }
            agg_inc(Hn, nslot);
        }
        __syncthreads();
        depth++;
        if (finish) break;               // M, m, cur keep describing the last pass for the final stage
        M = M2; cur ^= 1;
    }
    // after the loop: `M`, `m`, P[], crankBuf[cur] describe the parents of the last pass; crankBuf[cur^1], M2 its children
    const uint16_t *crankPar = crankBase + cur * maxN, *crankKid = crankBase + (cur ^ 1) * maxN;

    // ---------------- leaves that still hold several candidates: keep the max response, first in the reference's
    // candidate order on ties (:381-400)
    const int nStay = M - m;                      // unexpanded parents of the last pass (phase-2 cut)
    const int nMulti = nStay + M2;                // ids: [0,nStay) unexpanded parents, [nStay, nStay+M2) multi-point children
    if (nMulti > maxN) { if (tid == 0) { atomicOr(status, ST_QT_LIMIT); *outCount = 0; } return; }
    __syncthreads();
    for (int i = tid; i < nMulti; i += QT) best[i] = 0ull;
    __syncthreads();
    for (int i = tid; i < n; i += QT) {
        const unsigned nd = qnode[i];
        if (nd == NODE_FINAL) continue;
        const int id = (nd & NODE_STAY) ? (int)P[nd & NODE_MASK] - m : nStay + (int)nd;
        const uint32_t c = cand[i];
        const uint32_t ok = cand_order_key_fast(cand_x(c), cand_y(c), L);
        const unsigned long long v = ((unsigned long long)cand_resp(c) << 32) | (unsigned long long)(0xFFFFFFFFu - ok);
        atomicMax(&best[id], v);
    }
    __syncthreads();
    for (int i = tid; i < n; i += QT) {
        const unsigned nd = qnode[i];
        if (nd == NODE_FINAL) continue;
        const bool stay = (nd & NODE_STAY) != 0;
        const int id = stay ? (int)P[nd & NODE_MASK] - m : nStay + (int)nd;
        const uint32_t c = cand[i];
        const uint32_t ok = cand_order_key_fast(cand_x(c), cand_y(c), L);
        const unsigned long long v = ((unsigned long long)cand_resp(c) << 32) | (unsigned long long)(0xFFFFFFFFu - ok);
        if (best[id] == v) {
            const int pos = atomicAdd(&S.nleaf, 1);
            const uint32_t key = stay ? (((uint32_t)(depth - 1) << 13) | crankPar[nd & NODE_MASK]) : (((uint32_t)depth << 13) | crankKid[nd]);
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
    bitonic_desc<QT>(skey, sval, n2);
    for (int i = tid; i < nleaf; i += QT) sel[i] = cand[sval[i]];
    if (tid == 0) *outCount = nleaf;
}

int launch_quadtree(const PlanDev &hp, const PlanDev *dp, const uint32_t *cand, const int *candCount, const uint32_t *lut,
                    uint32_t *qcode, uint16_t *qnode, uint2 *qleaf, uint32_t *sel, int *selCount, int *status, int B,
                    cudaStream_t st, long *launches)
{
    int maxN = 0;
    for (int l = 0; l < hp.nlevels; l++) if (hp.lv[l].qtMaxN > maxN) maxN = hp.lv[l].qtMaxN;
    const size_t smem = (size_t)maxN * 26;
    if (smem > 48 * 1024) {
        int dev = 0, optin = 0;
        HY_CUDA(cudaGetDevice(&dev));
        HY_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        if ((int)smem + 1024 > optin) { set_error("quadtree needs %zu bytes of shared memory, device allows %d", smem, optin); return HYORB_EUNSUPPORTED; }
        HY_CUDA(cudaFuncSetAttribute(k_quadtree<QT_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HY_CUDA(cudaFuncSetAttribute(k_quadtree<QT_LAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grd(hp.nlevels, B);
    // a handful of images cannot fill the machine anyway: give each (image, level) CTA twice the threads (measured on one C2
    // frame: 0.166 -> 0.112 ms); large batches keep 512 threads per CTA for occupancy
    if (B * hp.nlevels <= 148) HY_CUDA(launch_dependent(k_quadtree<QT_LAT>, grd, dim3(QT_LAT), smem, st, dp, cand, candCount, lut, qcode, qnode, qleaf, sel, selCount, status));
    else HY_CUDA(launch_dependent(k_quadtree<QT_THREADS>, grd, dim3(QT_THREADS), smem, st, dp, cand, candCount, lut, qcode, qnode, qleaf, sel, selCount, status));
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

}  // namespace hyorb
