/* ANALYSIS TOOL (test infrastructure, uses the oracle's update_step): potential of a second, two-sided causal skip in the
 * batched sweep, at lane level and at WARP level (32 consecutive relaxed vertices of the window, triangle slots in the rotated
 * order of the batched kernel: starting at the neighbour of smallest rank).
 * Rule tested for a triangle (v; a, b) with cur = d[v], thr = cur (1 + 2^-14), lo/hi = min/max(t_a, t_b):
 *   old rule:  lo > thr                                                         (both corners above the vertex)
 *   new rule:  hi >= thr, q_ab >= 0 (acute at v), rho^2 <= 8, (hi - cur) >= 2^-9 max(cur - lo, 0),
 *              and the lo corner cannot reach v along its edge: lo >= thr or |X_lo|^2 >= (thr - lo)^2 (1 + 16u)
 * Also checks the claim behind it: whenever the new rule fires, update_step returns p >= cur. */
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

/* out: 0 relaxations, 1 triangle slots, 2 lane-level evaluated (old rule), 3 lane-level evaluated (old + new rule),
 *      4 warp-slots total, 5 warp-slots evaluated (old), 6 warp-slots evaluated (old + new), 7 violations (new rule fired, p < cur) */
static uint64_t extra[4];
void analyze2_extra(uint64_t *o) { for (int i = 0; i < 4; i++) o[i] = extra[i]; }
void analyze2_f32(uint32_t n_v, const float *GT, const uint32_t *VT, const uint32_t *OT, const uint32_t *EVT, const uint32_t *sources,
                  uint32_t n_sources, const uint32_t *limits, uint32_t n_limits, const uint32_t *sorted, const uint32_t *inv, uint64_t *out)
{
    float *d[2] = {malloc(4 * n_v), malloc(4 * n_v)};
    int32_t *chgs[2] = {malloc(4 * n_v), malloc(4 * n_v)};
    int32_t *seen = calloc(n_v, 4);
    for (uint32_t v = 0; v < n_v; v++) { d[0][v] = d[1][v] = INFINITY; chgs[0][v] = chgs[1][v] = -10; }
    for (uint32_t i = 0; i < n_sources; i++) d[0][sources[i]] = d[1][sources[i]] = 0;
    memset(out, 0, 8 * 8);
    uint32_t i = 1, j = 2, dd = 0, iter = 0, max_iter = n_limits << 1;
    int32_t k = 0;
    while (n_limits >= 3 && i < j && iter++ < max_iter) {
        if (i < (j >> 1)) i = j >> 1;
        const uint32_t start = limits[i], end = limits[j], n_cond = limits[i + 1] - start;
        const float *od = d[dd];
        float *nd = d[!dd];
        k++;
        const int32_t *chg = chgs[(k - 1) & 1];
        uint32_t lane = 0;
        uint8_t w_old[8] = {0}, w_new[8] = {0}, w_any[8] = {0};
        for (uint32_t vi = start; vi < end; vi++) {
            const uint32_t v = sorted[vi];
            int any = chg[v] == k - 1 || seen[v] < 2;
            const uint32_t stop = EVT[v];
            uint32_t hes[16], nh = 0;
            for (uint32_t he = stop; he != ORC_NIL;) {
                if (chg[VT[he_next(he)]] == k - 1 || chg[VT[he_prev(he)]] == k - 1) any = 1;
                if (nh < 16) hes[nh++] = he;
                he = OT[he_prev(he)];
                if (he == stop) he = ORC_NIL;
            }
            float nv = nd[v];
            if (any) {
                out[0]++;
                /* rotation: start at the neighbour (he_next corner) of smallest rank */
                uint32_t rot = 0;
                for (uint32_t q = 1; q < nh; q++) if (inv[VT[he_next(hes[q])]] < inv[VT[he_next(hes[rot])]]) rot = q;
                float full = od[v];
                const float cur = od[v], thr = cur * (1.0f + 0x1p-14f);
                for (uint32_t q = 0; q < nh && q < 8; q++) {
                    const uint32_t he = hes[(q + rot) % nh];
                    const uint32_t a = VT[he_next(he)], b = VT[he_prev(he)];
                    const float p = update_step_f32(GT, VT, od, he);
                    if (p < full) full = p;
                    float Xa[3], Xb[3];
                    for (int c = 0; c < 3; c++) { Xa[c] = GT[3 * (size_t)a + c] - GT[3 * (size_t)v + c]; Xb[c] = GT[3 * (size_t)b + c] - GT[3 * (size_t)v + c]; }
                    const float qa = dot3_f32(Xa, Xa), qb = dot3_f32(Xb, Xb), qab = dot3_f32(Xa, Xb);
                    const float lo = od[a] < od[b] ? od[a] : od[b], hi = od[a] < od[b] ? od[b] : od[a];
                    const float qlo = od[a] < od[b] ? qa : qb;
                    const int old_skip = lo > thr && lo >= 0x1p-60f;
                    int new_skip = 0;
                    if (!old_skip && hi >= thr && hi < INFINITY && qab >= 0 && fmaxf(qa, qb) <= 8.0f * fminf(qa, qb)) {
                        const float gap = cur > lo ? cur - lo : 0.0f;
                        const float dlo = thr - lo;
                        const int edge_ok = lo >= thr || qlo >= (dlo * dlo) * (1.0f + 0x1p-20f);
                        new_skip = (hi - cur) >= 0x1p-9f * gap && (hi - cur) >= 0x1p-39f && hi <= 0x1p23f && edge_ok;
                    }
                    if (new_skip && p < cur) out[7]++;
                    if (!old_skip && !new_skip) { extra[0]++; if (p < cur) extra[1]++; if (hi < thr) extra[2]++; }
                    out[1]++;
                    out[2] += !old_skip;
                    out[3] += !old_skip && !new_skip;
                    w_any[q] = 1;
                    w_old[q] |= !old_skip;
                    w_new[q] |= !old_skip && !new_skip;
                }
                nv = full;
                if (++lane == 32 || vi + 1 == end) {
                    for (int q = 0; q < 8; q++) { out[4] += w_any[q]; out[5] += w_old[q]; out[6] += w_new[q]; }
                    memset(w_old, 0, 8); memset(w_new, 0, 8); memset(w_any, 0, 8);
                    lane = 0;
                }
            }
            const float prev = nd[v];
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
