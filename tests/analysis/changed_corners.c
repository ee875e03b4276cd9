/* ANALYSIS TOOL (test infrastructure, uses the oracle's update_step): how many triangle evaluations of the change-driven
 * PTP sweep have at least one corner whose value changed in the previous iteration? A triangle whose two neighbour values are
 * bit-identical to those of the vertex's previous relaxation returns the same p as then, which is >= the vertex's stored value:
 * it cannot lower the vertex. This simulation counts such triangles and CHECKS the claim (restricted minimum == full minimum,
 * bit for bit, at every relaxation). Build + run: python tests/analysis/changed_corners.py
 * Result (icosphere f = 100 / 200, one source): the claim holds (0 mismatches), but only 7 % of the triangles of the executed
 * relaxations have two unchanged corners; together with the causal skip 63 % of the triangles remain against 69-70 % with the
 * causal skip alone — not worth a per-neighbour "changed" byte in the batched kernel (profiles/README.md). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#define ORC_NIL 0xFFFFFFFFu
static inline uint32_t he_next(uint32_t he) { return 3 * (he / 3) + (he + 1) % 3; }
static inline uint32_t he_prev(uint32_t he) { return 3 * (he / 3) + (he + 2) % 3; }
#define REAL float
#define SUF f32
#define SQRT sqrtf
#define ABS fabsf
#include "../../oracle/ptp_oracle_impl.h"

/* out: 0 relaxations (change-driven), 1 triangles of those, 2 triangles with a changed corner (or vertex new in the window),
 *      3 vertex-updates (window sizes), 4 mismatches of the restricted minimum, 5 triangles passing BOTH the causal test and
 *      the changed-corner test, 6 triangles passing the causal test */
void analyze_f32(uint32_t n_v, const float *GT, const uint32_t *VT, const uint32_t *OT, const uint32_t *EVT, const uint32_t *sources,
                 uint32_t n_sources, const uint32_t *limits, uint32_t n_limits, const uint32_t *sorted, uint64_t *out)
{
    float *d[2] = {malloc(4 * n_v), malloc(4 * n_v)};
    int32_t *chgs[2] = {malloc(4 * n_v), malloc(4 * n_v)}; /* per written buffer (parity of the iteration) */   /* last iteration at which the vertex's written value differed from what the buffer held */
    int32_t *seen = calloc(n_v, 4);   /* how many times the vertex has been in the window */
    for (uint32_t v = 0; v < n_v; v++) { d[0][v] = d[1][v] = INFINITY; chgs[0][v] = chgs[1][v] = -10; }
    for (uint32_t i = 0; i < n_sources; i++) d[0][sources[i]] = d[1][sources[i]] = 0;
    memset(out, 0, 8 * 10);
    uint32_t i = 1, j = 2, dd = 0, iter = 0, max_iter = n_limits << 1;
    int32_t k = 0;
    while (n_limits >= 3 && i < j && iter++ < max_iter) {
        if (i < (j >> 1)) i = j >> 1;
        const uint32_t start = limits[i], end = limits[j], n_cond = limits[i + 1] - start;
        const float *od = d[dd];
        float *nd = d[!dd];
        k++;
        const int32_t *chg = chgs[(k - 1) & 1];
        for (uint32_t vi = start; vi < end; vi++) {
            const uint32_t v = sorted[vi];
            out[3]++;
            /* change-driven: relax iff v or a ring neighbour changed at k-1, or v is new in the window (first two visits) */
            int any = chg[v] == k - 1 || seen[v] < 2;
            const uint32_t stop = EVT[v];
            for (uint32_t he = stop; he != ORC_NIL;) {
                if (chg[VT[he_next(he)]] == k - 1 || chg[VT[he_prev(he)]] == k - 1) any = 1;
                he = OT[he_prev(he)];
                if (he == stop) he = ORC_NIL;
            }
            float full = od[v], restr = od[v];
            if (any) {
                out[0]++;
                for (uint32_t he = stop; he != ORC_NIL;) {
                    const uint32_t a = VT[he_next(he)], b = VT[he_prev(he)];
                    const float p = update_step_f32(GT, VT, od, he);
                    const int changed = seen[v] < 2 || chg[a] == k - 1 || chg[b] == k - 1;
                    const float lo = od[a] < od[b] ? od[a] : od[b];
                    const int causal = !(lo > od[v] * (1.0f + 0x1p-14f) && lo >= 0x1p-60f);
                    {   /* potential of a two-sided causal bound: p >= min(max(t_a, t_b), min_i (t_i + |X_i|)) */
                        float Xa[3], Xb[3];
                        for (int c = 0; c < 3; c++) { Xa[c] = GT[3 * (size_t)a + c] - GT[3 * (size_t)v + c]; Xb[c] = GT[3 * (size_t)b + c] - GT[3 * (size_t)v + c]; }
                        const float ea = od[a] + norm3_f32(Xa), eb = od[b] + norm3_f32(Xb);
                        const float hi = od[a] > od[b] ? od[a] : od[b];
                        const float e = ea < eb ? ea : eb;
                        const float bound = hi < e ? hi : e;
                        static uint64_t dummy; (void)dummy;
                        if (causal && bound > od[v] * (1.0f + 0x1p-14f)) out[8]++;
                        if (causal && bound > od[v] * (1.0f + 0x1p-14f) && !(p >= od[v])) out[9]++; /* bound violated: p would have mattered */
                    }
                    out[1]++;
                    out[2] += changed;
                    out[6] += causal;
                    out[5] += causal && changed;
                    if (p < full) full = p;
                    if (changed && p < restr) restr = p;
                    he = OT[he_prev(he)];
                    if (he == stop) he = ORC_NIL;
                }
                if (memcmp(&full, &restr, 4)) out[4]++;
            }
            /* (skipped: the buffer already holds the result; checked by the GPU parity tests, not here) */
            const float prev = nd[v];
            float nv = any ? full : prev;
            if (!any) { /* verify the skip as well */
                float chk = od[v];
                for (uint32_t he = stop; he != ORC_NIL;) {
                    const float p = update_step_f32(GT, VT, od, he);
                    if (p < chk) chk = p;
                    he = OT[he_prev(he)];
                    if (he == stop) he = ORC_NIL;
                }
                if (memcmp(&chk, &prev, 4)) out[7]++;
                nv = chk;
            }
            if (memcmp(&nv, &prev, 4)) chgs[k & 1][v] = k;
            nd[v] = nv;
            seen[v]++;
        }
        uint32_t count = 0;
        for (uint32_t vi = start; vi < start + n_cond; vi++) {
            const uint32_t v = sorted[vi];
            const float err = fabsf(nd[v] - od[v]) / od[v];
            count += err < 1e-3;
        }
        if (n_cond == count) i++;
        if (j < n_limits - 1) j++;
        dd = !dd;
    }
    free(d[0]); free(d[1]); free(chgs[0]); free(chgs[1]); free(seen);
}
