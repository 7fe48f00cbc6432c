/*
 * quadtree_closed_form.c -- CPU ORACLE helper (test infrastructure, NOT product code).
 *
 * A second, *level-synchronous* restatement of DistributeOctTree (reference
 * src/features/ORBExtractor.cpp:179-403 + DivideNode :121-177) that never builds a linked list.
 * It exists to pin the parallel formulation the CUDA kernel uses (hyslam_b200/csrc/quadtree.cu)
 * against the literal list-based restatement in orb_oracle.c (orc_distribute_octtree) on the CPU,
 * where both can be run on thousands of random inputs.  tests/test_quadtree_closed_form.py does that.
 *
 * Facts used (each follows from the reference code, see DESIGN.md "Quadtree"):
 *  1. Child bounds depend only on the parent's bounds (:123-124), and the x split never looks at y,
 *     so depth-d nodes form a separable, data-independent grid.  A point's path through the tree is
 *     root || interleave(rowbits, colbits) with one (col,row) bit per depth, q = 2*bottom + right
 *     (n1=UL,n2=UR,n3=BL,n4=BR, :151-166).
 *  2. Every pass splits *all* multi-point leaves (phase 1, :257-303) or a prefix of them in sorted
 *     order (phase 2, :324-371), so all splittable leaves always sit at one depth.
 *  3. New children are push_front'ed, so the list is always
 *         [depth D leaves, newest first] [depth D-1 leaves, newest first] ... [root leaves]
 *     and a phase-1 pass visits the multi-point nodes of depth d newest-first.
 *  4. "Creation rank" (crank) of a child = (position of its parent in this pass's visiting order, q);
 *     the canonical tie policy of SURVEY.md A.4 (node address order == creation order) makes the
 *     phase-2 sort key (count, crank).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_API __attribute__((visibility("default")))
enum { ORC_OK = 0, ORC_EINVAL = -1, ORC_ECAPACITY = -2, ORC_EUNSUPPORTED = -3, ORC_ENOMEM = -4 };

#define QT_DMAX 13   /* depth bits per axis: separates all integer points of a < 8192 px span */

/* column / row path bits of an integer lattice coordinate: binary descent with the reference's
 * ceil-half split (:123-124) and strict `<` comparisons (:154-163) */
static uint32_t axis_bits(int v, int lo, int hi)
{
    uint32_t bits = 0;
    for (int d = 0; d < QT_DMAX; d++) {
        const int half = (int)ceilf((float)(hi - lo) / 2);
        const int mid = lo + half;
        bits <<= 1;
        if ((float)v < (float)mid) hi = mid;
        else { bits |= 1u; lo = mid; }
    }
    return bits;
}

static uint32_t interleave(uint32_t row, uint32_t col)   /* 2 bits per depth, MSB = depth 1 */
{
    uint32_t m = 0;
    for (int d = 0; d < QT_DMAX; d++) {
        const int s = QT_DMAX - 1 - d;
        m = (m << 2) | (((row >> s) & 1u) << 1) | ((col >> s) & 1u);
    }
    return m;
}

typedef struct { uint32_t key; int node; } sort_ent;
static int cmp_desc(const void *a, const void *b)
{
    const sort_ent *x = (const sort_ent *)a, *y = (const sort_ent *)b;
    return x->key < y->key ? 1 : (x->key > y->key ? -1 : 0);
}
typedef struct { uint32_t key; int cand; } leaf_ent;
static int cmp_leaf_desc(const void *a, const void *b)
{
    const leaf_ent *x = (const leaf_ent *)a, *y = (const leaf_ent *)b;
    return x->key < y->key ? 1 : (x->key > y->key ? -1 : 0);
}

ORC_API int orc_quadtree_closed_form(const float *px, const float *py, const float *presp, int n,
                                     int minX, int maxX, int minY, int maxY, int N,
                                     int32_t *out_idx, int cap)
{
    if (n == 0) return 0;
    const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));   /* :183 */
    if (nIni < 1 || nIni > 64) return ORC_EUNSUPPORTED;
    const float hX = (float)(maxX - minX) / nIni;                          /* :185 */
    const int H = maxY - minY;

    uint64_t *code = (uint64_t *)malloc(sizeof(uint64_t) * n);  /* root << 2*DMAX | morton */
    int *node = (int *)malloc(sizeof(int) * n);                 /* crank of the point's current node; -1 = in a final leaf */
    int *cnt = (int *)calloc((size_t)n + 64, sizeof(int));      /* per node of the current depth, indexed by crank */
    int *ncnt = (int *)calloc((size_t)n + 64, sizeof(int));
    int *pidx = (int *)malloc(sizeof(int) * ((size_t)n + 64));  /* visiting position of a multi node, -1 otherwise */
    int *slot = (int *)calloc(4 * ((size_t)n + 64), sizeof(int));
    int *slot_new = (int *)malloc(sizeof(int) * 4 * ((size_t)n + 64));
    sort_ent *order = (sort_ent *)malloc(sizeof(sort_ent) * ((size_t)n + 64));
    leaf_ent *leaf = (leaf_ent *)malloc(sizeof(leaf_ent) * ((size_t)n + 64));
    int *best = (int *)malloc(sizeof(int) * ((size_t)n + 64));
    int rc = ORC_OK, nleaf = 0;
    if (!code || !node || !cnt || !ncnt || !pidx || !slot || !slot_new || !order || !leaf || !best) { rc = ORC_ENOMEM; goto done; }

    /* depth 0: roots (:192-225).  crank_0 = rank among non-empty roots counted from the right,
     * so that "visit newest first" == ascending root index, which is what the first pass does. */
    {
        int rootcnt[64]; memset(rootcnt, 0, sizeof(rootcnt));
        int rootcrank[64];
        for (int i = 0; i < n; i++) {
            const int r = (int)(px[i] / hX);
            if (r < 0 || r >= nIni) { rc = ORC_EUNSUPPORTED; goto done; }
            const int ulx = (int)(hX * (float)r), urx = (int)(hX * (float)(r + 1));
            const uint32_t col = axis_bits((int)px[i], ulx, urx);
            const uint32_t row = axis_bits((int)py[i], 0, H);
            code[i] = ((uint64_t)r << (2 * QT_DMAX)) | interleave(row, col);
            rootcnt[r]++;
        }
        int F = 0;
        for (int r = nIni - 1; r >= 0; r--) if (rootcnt[r]) rootcrank[r] = F++;
        for (int r = 0; r < nIni; r++) if (rootcnt[r]) cnt[rootcrank[r]] = rootcnt[r];
        for (int i = 0; i < n; i++) {
            const int r = (int)(code[i] >> (2 * QT_DMAX));
            node[i] = rootcrank[r];
            if (rootcnt[r] == 1) { leaf[nleaf].key = (0u << 20) | (uint32_t)node[i]; leaf[nleaf].cand = i; nleaf++; node[i] = -1; }
        }
        int size = F, depth = 0, phase2 = 0, finish = 0;
        while (!finish) {
            /* multi nodes of this depth and their visiting order */
            int M = 0;
            for (int c = 0; c < F; c++) {
                pidx[c] = -1;
                if (cnt[c] > 1) { order[M].key = ((phase2 ? (uint32_t)cnt[c] : 0u) << 13) | (uint32_t)c; order[M].node = c; M++; }
            }
            if (F > 8192 || n >= (1 << 19)) { rc = ORC_EUNSUPPORTED; goto done; }
            qsort(order, M, sizeof(sort_ent), cmp_desc);     /* (count desc, crank desc); phase 1: crank desc */
            for (int p = 0; p < M; p++) pidx[order[p].node] = p;
            /* children histogram */
            memset(slot, 0, sizeof(int) * 4 * (size_t)(M + 1));
            const int sh = 2 * (QT_DMAX - 1 - depth);
            for (int i = 0; i < n; i++) {
                if (node[i] < 0) continue;
                const int q = (int)((code[i] >> sh) & 3u);
                slot[4 * pidx[node[i]] + q]++;
            }
            /* how many parents get expanded: all (phase 1) or until size >= N (phase 2, :370) */
            int m = M, nsize = size, nToExpand = 0;
            for (int p = 0; p < M; p++) {
                int k = 0;
                for (int q = 0; q < 4; q++) k += slot[4 * p + q] > 0;
                nsize += k - 1;
                if (phase2 && nsize >= N) { m = p + 1; break; }
            }
            /* creation ranks of the children */
            int F2 = 0;
            for (int s = 0; s < 4 * m; s++) {
                slot_new[s] = -1;
                if (slot[s] > 0) { slot_new[s] = F2; ncnt[F2] = slot[s]; if (slot[s] > 1) nToExpand++; F2++; }
            }
            /* move points; freeze singletons; points of unexpanded parents stay (final pass only) */
            for (int c = 0; c < F; c++) best[c] = -1;
            for (int i = 0; i < n; i++) {
                if (node[i] < 0) continue;
                const int p = pidx[node[i]];
                if (p >= m) {          /* unexpanded multi node: a final leaf of depth `depth`; keep max response, first wins */
                    const int c = node[i];
                    if (best[c] < 0 || presp[i] > presp[best[c]]) best[c] = i;
                    node[i] = -1;
                    continue;
                }
                const int q = (int)((code[i] >> sh) & 3u);
                const int s = 4 * p + q;
                node[i] = slot_new[s];
                if (slot[s] == 1) { leaf[nleaf].key = ((uint32_t)(depth + 1) << 20) | (uint32_t)node[i]; leaf[nleaf].cand = i; nleaf++; node[i] = -1; }
            }
            for (int c = 0; c < F; c++) if (best[c] >= 0) { leaf[nleaf].key = ((uint32_t)depth << 20) | (uint32_t)c; leaf[nleaf].cand = best[c]; nleaf++; }
            const int prevSize = size;
            size = nsize; depth++; F = F2;
            { int *t = cnt; cnt = ncnt; ncnt = t; }
            if (size >= N || size == prevSize) finish = 1;                         /* :309 / :374 */
            else if (!phase2 && size + nToExpand * 3 > N) phase2 = 1;               /* :313 */
            if (depth >= QT_DMAX && !finish) { rc = ORC_EUNSUPPORTED; goto done; }
        }
        /* remaining multi-point nodes of the last depth are leaves too */
        for (int c = 0; c < F; c++) best[c] = -1;
        for (int i = 0; i < n; i++) {
            if (node[i] < 0) continue;
            const int c = node[i];
            if (best[c] < 0 || presp[i] > presp[best[c]]) best[c] = i;
        }
        for (int c = 0; c < F; c++) if (best[c] >= 0) { leaf[nleaf].key = ((uint32_t)depth << 20) | (uint32_t)c; leaf[nleaf].cand = best[c]; nleaf++; }
    }
    /* list order: deepest first, newest first within a depth */
    qsort(leaf, nleaf, sizeof(leaf_ent), cmp_leaf_desc);
    if (nleaf > cap) { rc = ORC_ECAPACITY; goto done; }
    for (int i = 0; i < nleaf; i++) out_idx[i] = leaf[i].cand;
    rc = nleaf;
done:
    free(code); free(node); free(cnt); free(ncnt); free(pidx); free(slot); free(slot_new); free(order); free(leaf); free(best);
    return rc;
}
