/* TEST INFRASTRUCTURE — the CPU oracle for the PTP hot path. NOT product code.
 *
 * Plain-C restatement of the reference's (larc/gproshan) CPU algorithm for the path
 * BASELINE.json's north_star names. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the checker.
 *
 * Parity status: PINNED. tests/test_oracle_vs_reference.py compares every function here,
 * bit for bit, with the reference's own unmodified sources compiled into oracle/_ref/
 * (oracle/Makefile, oracle/ref_driver.cpp); tests/golden/ holds vectors generated from that
 * reference build (tests/golden/make_golden.py) so the pin travels to boxes without /root/reference.
 * The reference ships no golden vectors or known-answer tests of its own for this path.
 *
 * file:line citations are relative to /root/reference.
 */
#ifndef PTP_ORACLE_H
#define PTP_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#define ORC_NIL 0xFFFFFFFFu

#ifdef __cplusplus
extern "C" {
#endif

/* che::update_evt_ot_et, src/che.cpp:1295-1362: opposite table OT[he] and extra vertex table EVT[v]
 * from the face list VT (3 vertex ids per face). Returns 1 if manifold, 0 otherwise. */
int orc_che_build(uint32_t n_v, uint32_t n_f, const uint32_t *VT, uint32_t *OT, uint32_t *EVT);

/* che::compute_toplesets, src/che.cpp:546-593 (link: :102-112). limits must hold n_v + 2 entries,
 * sorted n_v + n_sources entries. Returns the number of limits written (0 when n_sources == 0). */
uint32_t orc_compute_toplesets(uint32_t n_v, const uint32_t *VT, const uint32_t *OT, const uint32_t *EVT,
                               const uint32_t *sources, uint32_t n_sources, uint32_t k,
                               uint32_t *toplesets, uint32_t *sorted, uint32_t *limits);

/* stats[0] = iterations executed, stats[1] = vertex-updates (sum of window sizes),
 * stats[2] = largest window, stats[3] = final value of d */
#define ORC_PTP_DECL(SUF, REAL)                                                                      \
    /* parallel_toplesets_propagation_cpu, src/geodesics_ptp.cpp:122-199. Distances follow the CPU  \
     * function exactly (older Jacobi buffer is returned, :193-198). clusters (may be NULL) follow   \
     * the reference GPU rule (src/cuda/geodesics_ptp.cu:257-282) double-buffered with the          \
     * distances, pre-filled with cluster_fill, returned for the same buffer as the distances. */   \
    void orc_ptp_cpu_##SUF(uint32_t n_v, const REAL *GT, const uint32_t *VT, const uint32_t *OT,     \
                           const uint32_t *EVT, const uint32_t *sources, uint32_t n_sources,         \
                           const uint32_t *limits, uint32_t n_limits, const uint32_t *sorted,        \
                           REAL *dist, uint32_t *clusters, uint32_t cluster_fill, uint64_t *stats);  \
    /* update_step, src/geodesics_ptp.cpp:201-262 */                                                 \
    REAL orc_update_step_##SUF(const REAL *GT, const uint32_t *VT, const REAL *dist, uint32_t he);   \
    /* normalize_ptp, src/geodesics_ptp.cpp:264-276 */                                               \
    void orc_normalize_ptp_##SUF(REAL *dist, size_t n);

ORC_PTP_DECL(f32, float)
ORC_PTP_DECL(f64, double)

#ifdef __cplusplus
}
#endif
#endif
