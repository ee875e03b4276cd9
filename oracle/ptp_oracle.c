/* TEST INFRASTRUCTURE — CPU oracle for the PTP hot path (see ptp_oracle.h). NOT product code. */
#include "ptp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* trig/next/prev, src/che.cpp:18-34 */
static inline uint32_t he_next(uint32_t he) { return 3 * (he / 3) + (he + 1) % 3; }
static inline uint32_t he_prev(uint32_t he) { return 3 * (he / 3) + (he + 2) % 3; }

/* che::update_evt_ot_et, src/che.cpp:1295-1362 */
int orc_che_build(uint32_t n_v, uint32_t n_f, const uint32_t *VT, uint32_t *OT, uint32_t *EVT)
{
    const uint32_t n_he = 3 * n_f;
    int manifold = 1;
    memset(EVT, 0xFF, sizeof(uint32_t) * n_v);
    if (!n_f) return manifold;

    /* he_p_vertex: half-edges by origin vertex, in increasing half-edge order (CSR form of the
     * reference's vector<index_t>[n_vertices], :1301-1308) */
    uint32_t *off = (uint32_t *)calloc((size_t)n_v + 1, sizeof(uint32_t));
    uint32_t *lst = (uint32_t *)malloc(sizeof(uint32_t) * n_he);
    for (uint32_t he = 0; he < n_he; he++) {
        EVT[VT[he]] = he;
        off[VT[he] + 1]++;
    }
    for (uint32_t v = 0; v < n_v; v++) off[v + 1] += off[v];
    uint32_t *fill = (uint32_t *)malloc(sizeof(uint32_t) * (n_v ? n_v : 1));
    memcpy(fill, off, sizeof(uint32_t) * n_v);
    for (uint32_t he = 0; he < n_he; he++) lst[fill[VT[he]]++] = he;
    free(fill);

    memset(OT, 0xFF, sizeof(uint32_t) * n_he);                              /* :1311 */
    for (uint32_t he = 0; he < n_he; he++) {                                 /* :1314-1330 */
        if (OT[he] != ORC_NIL) continue;
        const uint32_t a = VT[he];
        for (uint32_t k = off[a]; k < off[a + 1]; k++) {
            const uint32_t h = lst[k];
            if (VT[he_prev(h)] == VT[he_next(he)])
                if (OT[he] == ORC_NIL && OT[he_prev(h)] == ORC_NIL) {
                    OT[he] = he_prev(h);
                    OT[he_prev(h)] = he;
                }
        }
    }

    for (uint32_t he = 0; he < n_he; he++)                                   /* :1343-1352 */
        if (OT[he] == ORC_NIL && EVT[VT[he]] != ORC_NIL) {
            if (OT[EVT[VT[he]]] == ORC_NIL && EVT[VT[he]] != he) {
                manifold = 0;
                EVT[VT[he]] = ORC_NIL;
            } else
                EVT[VT[he]] = he;
        }

    free(off);
    free(lst);
    return manifold;
}

/* che::compute_toplesets, src/che.cpp:546-593; che::link, :102-112 */
uint32_t orc_compute_toplesets(uint32_t n_v, const uint32_t *VT, const uint32_t *OT, const uint32_t *EVT,
                               const uint32_t *sources, uint32_t n_sources, uint32_t k,
                               uint32_t *toplesets, uint32_t *sorted, uint32_t *limits)
{
    if (!n_sources) return 0;                                                /* :548 */
    memset(toplesets, 0xFF, sizeof(uint32_t) * n_v);

    uint32_t level = 0, p = 0, nl = 0;
    for (uint32_t s = 0; s < n_sources; s++) {                               /* :555-561 */
        sorted[p++] = sources[s];
        if (toplesets[sources[s]] == ORC_NIL) toplesets[sources[s]] = level;
    }

    limits[nl++] = 0;
    for (uint32_t i = 0; i < p; i++) {                                       /* :564-589 */
        const uint32_t v = sorted[i];
        if (toplesets[v] > level) {
            level++;
            if (level > k) break;
            limits[nl++] = i;
        }
        /* link(v): for each star half-edge: next(he), and prev(he) when OT[prev(he)] == NIL */
        const uint32_t stop = EVT[v];
        for (uint32_t he = stop; he != ORC_NIL;) {
            uint32_t u = VT[he_next(he)];
            if (toplesets[u] == ORC_NIL) { toplesets[u] = toplesets[v] + 1; sorted[p++] = u; }
            if (OT[he_prev(he)] == ORC_NIL) {
                u = VT[he_prev(he)];
                if (toplesets[u] == ORC_NIL) { toplesets[u] = toplesets[v] + 1; sorted[p++] = u; }
            }
            he = OT[he_prev(he)];
            if (he == stop) he = ORC_NIL;
        }
    }
    limits[nl++] = p;                                                        /* :592 */
    return nl;
}

#define REAL float
#define SUF f32
#define SQRT sqrtf
#define ABS fabsf
#include "ptp_oracle_impl.h"
#undef REAL
#undef SUF
#undef SQRT
#undef ABS

#define REAL double
#define SUF f64
#define SQRT sqrt
#define ABS fabs
#include "ptp_oracle_impl.h"
#undef REAL
#undef SUF
#undef SQRT
#undef ABS
