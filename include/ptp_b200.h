/* ptp_b200 — C ABI of the B200-native Parallel Toplesets Propagation (PTP) geodesic solver.
 *
 * Drop-in boundary for the PTP_GPU path of larc/gproshan. Plain pointers and sizes only; every
 * entry point returns PTP_OK (0) or a negative error code, with ptp_last_error() giving the text.
 * `file:line` citations are relative to the reference repository (larc/gproshan); each entry point
 * names the reference interface it replaces. INTEGRATION.md shows the gproshan-side binding.
 *
 * Conventions (reference include/include.h:12-27, include/che.h:41-47):
 *   index_t = uint32_t, NIL = 0xFFFFFFFF, real_t = float (f32 entry points) or double (f64).
 *   Mesh = compact half-edge tables: GT[V][3] vertex positions, VT[H] origin vertex of half-edge,
 *   OT[H] opposite half-edge or NIL, EVT[V] one outgoing half-edge per vertex (border half-edge for
 *   border vertices, NIL for isolated vertices); H = 3 * faces.
 *
 * Parity contract: distances equal parallel_toplesets_propagation_cpu (src/geodesics_ptp.cpp:122-199)
 * bit for bit (stronger than the stated 1e-5 / 1e-10 relative tolerance); toplesets / sorted / limits
 * equal che::compute_toplesets (src/che.cpp:546-593) bit for bit.
 *
 * There is no CPU fallback: every compute entry point needs a CUDA device (sm_100a build).
 * Nothing here calls cudaDeviceReset() (the reference does, src/cuda/geodesics_ptp.cu:22).
 */
#ifndef PTP_B200_H
#define PTP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTP_NIL 0xFFFFFFFFu

enum {
    PTP_OK = 0,
    PTP_ERR_INVALID = -1,   /* bad argument (null pointer, index out of range, sizes) */
    PTP_ERR_CUDA = -2,      /* CUDA runtime / driver failure (text in ptp_last_error) */
    PTP_ERR_CAPACITY = -3,  /* caller-provided output buffer too small */
    PTP_ERR_MESH = -4,      /* mesh tables inconsistent (one-ring walk does not terminate) */
    PTP_ERR_NO_DEVICE = -5  /* no CUDA device / kernel image not loadable on this device */
};

typedef struct ptp_mesh ptp_mesh_t; /* opaque device-resident mesh + solver workspace */

/* per-call statistics; times are device times from CUDA events, milliseconds */
typedef struct ptp_stats {
    uint64_t n_reached;      /* limits.back(): vertices reached from the sources (duplicates counted)  */
    uint64_t n_levels;       /* number of toplesets (= limits.size() - 1)                             */
    uint64_t iterations;     /* PTP iterations executed (src/geodesics_ptp.cpp:145)                   */
    uint64_t vertex_updates; /* sum over iterations of the window size limits[j] - limits[i]          */
    uint64_t max_window;     /* largest window                                                        */
    uint64_t relaxations;    /* one-ring relaxations actually executed (<= vertex_updates: vertices whose  *
                              * inputs did not change since they were last relaxed keep their stored value) */
    uint64_t gpu_launches;   /* kernels this call launched                                            */
    double ms_toplesets;     /* BFS + topleset-order layout (batched: mean per-CTA time over the launch)  */
    double ms_solve;         /* relaxation sweep + scatter back to vertex order (batched: mean per CTA;   *
                              * single solve with device toplesets: start of the BFS kernel to end of sweep team) */
    double ms_total;         /* first launch to last launch, device time (CUDA events)                */
} ptp_stats_t;

/* ------------------------------------------------------------------------------------------------ */

const char *ptp_last_error(void);           /* thread-local text of the last failure               */
int ptp_device_count(void);                 /* number of CUDA devices, <= 0 when none                */
const char *ptp_version(void);

/* Behaviour switches. Every switch has a name, a default and an environment variable PTP_<NAME IN CAPITALS> that
 * replaces the default when set (read once, at the first use of the library); ptp_set_option changes it at run time
 * (takes effect at the next call). ptp_option_name / ptp_option_doc enumerate them (NULL past the end). The switches
 * select among kernel variants that all produce the same bits, except "newest" (which Jacobi buffer is returned). */
int ptp_set_option(const char *name, long value);
long ptp_get_option(const char *name);
const char *ptp_option_name(int index);
const char *ptp_option_doc(int index);

/* pinned host memory helpers (optional; any host pointer is accepted by the entry points below) */
void *ptp_host_alloc(size_t bytes);
void ptp_host_free(void *p);

/* Mesh upload. Replaces CHE::CHE(che*) + cuda_create_CHE (src/che.cpp:36-46, src/cuda/che.cu:29-48),
 * which the reference repeats on every solve; here the mesh stays resident until ptp_mesh_destroy.
 * Host tables are copied, the caller keeps ownership. Builds the per-vertex one-ring table on the
 * device (the for_star order of include/che.h:10). `device` is a CUDA ordinal. OT and EVT may both be
 * NULL: they are then built on the device from VT (see ptp_che_build). */
int ptp_mesh_create_f32(const float *GT, const uint32_t *VT, const uint32_t *OT, const uint32_t *EVT,
                        uint64_t n_vertices, uint64_t n_half_edges, int device, ptp_mesh_t **out);
int ptp_mesh_create_f64(const double *GT, const uint32_t *VT, const uint32_t *OT, const uint32_t *EVT,
                        uint64_t n_vertices, uint64_t n_half_edges, int device, ptp_mesh_t **out);
void ptp_mesh_destroy(ptp_mesh_t *mesh);
/* Replace the vertex positions of a resident mesh (GT[V][3], same connectivity): what a caller does after editing a
 * che in place (noise, smoothing, che::reload of the same topology) instead of destroying and re-creating the handle.
 * The reference uploads the whole CHE on every solve (src/cuda/che.cu:29-48). */
int ptp_mesh_update_positions_f32(ptp_mesh_t *mesh, const float *GT);
int ptp_mesh_update_positions_f64(ptp_mesh_t *mesh, const double *GT);
/* Threading: a ptp_mesh_t owns ONE solver workspace. Calls on the same handle are serialised by the library (a mutex
 * per handle); calls on different handles — other meshes, or the same mesh uploaded to other devices — run
 * concurrently. ptp_last_error() is per thread. */
/* name of the dominant kernel of the last solve on this mesh (measurement: which single-solve variant ran) */
const char *ptp_mesh_last_kernel(const ptp_mesh_t *mesh);

/* CHE tables from a face list, on the device. Replaces che::update_evt_ot_et (src/che.cpp:1295-1362; serial,
 * ~22 s at 10 M vertices) for oriented edge-manifold input, for which OT / EVT equal the reference's bit for
 * bit. VT[H] = 3 vertex ids per face; OT[H], EVT[V] are host outputs. *manifold = 0 when a directed edge occurs
 * twice or a vertex has two border half-edges (the reference pairs such input in half-edge order; here it is
 * reported, and ptp_mesh_create_* refuses it). *ms = device time of the build. ptp_mesh_create_* with
 * OT == EVT == NULL runs the same build before the one-ring table. */
int ptp_che_build(const uint32_t *VT, uint64_t n_vertices, uint64_t n_half_edges, uint32_t *OT, uint32_t *EVT,
                  int device, int *manifold, double *ms);
uint64_t ptp_mesh_n_vertices(const ptp_mesh_t *mesh);
uint64_t ptp_mesh_n_half_edges(const ptp_mesh_t *mesh);
int ptp_mesh_real_size(const ptp_mesh_t *mesh); /* 4 or 8 */
int ptp_mesh_device(const ptp_mesh_t *mesh);
uint64_t ptp_mesh_device_bytes(const ptp_mesh_t *mesh); /* device memory currently held */

/* Topleset construction on the device. Replaces che::compute_toplesets (src/che.cpp:546-593):
 *   sorted[0..S) = sources in the given order (duplicates kept), level 0; BFS in queue order;
 *   limits = [0, S, ..., p] one entry per level start plus the end; k caps the levels (PTP_NIL = no cap).
 * Outputs (host; any may be NULL): toplesets[V] (NIL for unreached), sorted[sorted_capacity]
 * (first limits.back() entries valid; needs V + number of duplicate sources), limits[limits_capacity].
 * *n_limits receives limits.size(). */
int ptp_toplesets(ptp_mesh_t *mesh, const uint32_t *sources, uint32_t n_sources, uint32_t k,
                  uint32_t *toplesets, uint32_t *sorted, uint64_t sorted_capacity,
                  uint32_t *limits, uint64_t limits_capacity, uint32_t *n_limits, ptp_stats_t *stats);

/* Solve with caller-provided toplesets. Replaces parallel_toplesets_propagation_gpu
 * (src/cuda/geodesics_ptp.cu:20-85, declared include/geodesics_ptp.h:36) and
 * parallel_toplesets_propagation_coalescence_gpu (src/cuda/geodesics_ptp_coalescence.cu:21-100,
 * include/geodesics_ptp.h:34): ptp_out_t{dist, clusters} -> dist / clusters, toplesets_t{limits, index}
 * -> limits / sorted. dist[V] receives the distances (INF for unreached). clusters may be NULL; when
 * given it receives, for every reached vertex, 1 + the index of its nearest source (rule of
 * src/cuda/geodesics_ptp.cu:277) and `cluster_fill` elsewhere. limits.size() < 3 performs no sweep
 * (the reference reads limits[2] unconditionally). */
int ptp_solve_f32(ptp_mesh_t *mesh, const uint32_t *sources, uint32_t n_sources,
                  const uint32_t *limits, uint32_t n_limits, const uint32_t *sorted,
                  float *dist, uint32_t *clusters, uint32_t cluster_fill, ptp_stats_t *stats);
int ptp_solve_f64(ptp_mesh_t *mesh, const uint32_t *sources, uint32_t n_sources,
                  const uint32_t *limits, uint32_t n_limits, const uint32_t *sorted,
                  double *dist, uint32_t *clusters, uint32_t cluster_fill, ptp_stats_t *stats);

/* Toplesets + solve in one device pipeline, nothing but the sources going up and the results coming
 * down. Replaces geodesics::run_parallel_toplesets_propagation_gpu (src/geodesics.cpp:225-240), i.e.
 * what `geodesics(mesh, sources, PTP_GPU, ...)` executes. sorted_index (may be NULL) receives the BFS
 * order like geodesics::sorted_index (first n_reached entries, capacity sorted_capacity). */
int ptp_geodesics_f32(ptp_mesh_t *mesh, const uint32_t *sources, uint32_t n_sources,
                      float *dist, uint32_t *clusters, uint32_t cluster_fill,
                      uint32_t *sorted_index, uint64_t sorted_capacity, ptp_stats_t *stats);
int ptp_geodesics_f64(ptp_mesh_t *mesh, const uint32_t *sources, uint32_t n_sources,
                      double *dist, uint32_t *clusters, uint32_t cluster_fill,
                      uint32_t *sorted_index, uint64_t sorted_capacity, ptp_stats_t *stats);

/* Accuracy harness: the error of the distances after each PTP iteration, as the reference's test executable records it
 * (iter_error_run_ptp_gpu, src/cuda/test_geodesics_ptp.cu:164-211; written to <mesh>_error.iter by
 * src/test_geodesics_ptp.cpp:198-214): for every iteration whose window already ends at the last topleset,
 * errors[k] = compute_error(new distances, exact) = 100 / (V - S) * sum over exact > 0 of |dist - exact| / exact
 * (src/test_geodesics_ptp.cpp:361-370) and iters[k] = the iteration number (1-based). The reference copies the whole
 * distance array to the host after every such iteration; here the sums are formed on the device. The schedule is the
 * solver's own (with the j/2 clamp and the iteration cap of src/geodesics_ptp.cpp:137-147, which the reference's harness
 * loop leaves out). exact[V] are the reference distances; dist (may be NULL) receives the final distances.
 * At most `capacity` records are written; *n_out receives their number. */
int ptp_geodesics_error_iter_f32(ptp_mesh_t *mesh, const uint32_t *sources, uint32_t n_sources, const float *exact, float *dist,
                                 uint32_t *iters, float *errors, uint32_t capacity, uint32_t *n_out, ptp_stats_t *stats);
int ptp_geodesics_error_iter_f64(ptp_mesh_t *mesh, const uint32_t *sources, uint32_t n_sources, const double *exact, double *dist,
                                 uint32_t *iters, double *errors, uint32_t capacity, uint32_t *n_out, ptp_stats_t *stats);

/* Batched independent solves (distance-matrix rows; the callers are sampling / key_components style
 * loops such as src/sampling.cpp:23-34). Source set b is sources[offsets[b] .. offsets[b+1]); with
 * offsets == NULL every source is its own single-source solve (n_batch = n_sources).
 * rows receives n_batch rows of V reals (row-major). rows_on_device != 0 means `rows` is a device
 * pointer on the mesh's device (used by the multi-GPU gather); otherwise a host pointer.
 * `stream` is a cudaStream_t (NULL = the legacy default stream). */
int ptp_solve_batched_f32(ptp_mesh_t *mesh, const uint32_t *sources, const uint64_t *offsets,
                          uint32_t n_batch, uint64_t n_sources, float *rows, int rows_on_device,
                          void *stream, ptp_stats_t *stats);
int ptp_solve_batched_f64(ptp_mesh_t *mesh, const uint32_t *sources, const uint64_t *offsets,
                          uint32_t n_batch, uint64_t n_sources, double *rows, int rows_on_device,
                          void *stream, ptp_stats_t *stats);

/* The same batch over several devices of ONE process (gproshan is a single C++ process; the callers are loops such as
 * src/sampling.cpp:23-34). meshes[0 .. n_devices) are handles of the SAME mesh created on different devices
 * (ptp_mesh_create_* with different `device`). Source sets are block-partitioned over the devices (the first
 * n_batch % n_devices devices take one more), one host thread per device runs its shard. rows receives all n_batch rows:
 *   rows_on_device == 0: a host pointer; every device copies its rows straight into place (no collective);
 *   rows_on_device != 0: a device pointer on meshes[0]'s device; the other devices' rows travel there over NVLink with
 *     NCCL (grouped ncclSend / ncclRecv; libnccl.so.2 is loaded at run time), each shard in `gather_chunks` pieces
 *     (ptp_set_option; default: one piece per two waves of CTAs) so that the transfer of a piece overlaps the solving of the next.
 * stats: counters summed over the devices; ms_solve = kernel time of the slowest device, ms_total = WALL time of the call. */
int ptp_solve_batched_multi_f32(ptp_mesh_t *const *meshes, int n_devices, const uint32_t *sources, const uint64_t *offsets,
                                uint32_t n_batch, uint64_t n_sources, float *rows, int rows_on_device, ptp_stats_t *stats);
int ptp_solve_batched_multi_f64(ptp_mesh_t *const *meshes, int n_devices, const uint32_t *sources, const uint64_t *offsets,
                                uint32_t n_batch, uint64_t n_sources, double *rows, int rows_on_device, ptp_stats_t *stats);

/* Farthest-point sampling on the resident mesh. Replaces farthest_point_sampling_ptp_gpu
 * (src/cuda/geodesics_ptp.cu:87-172): starting from samples[0..n_initial), repeatedly solve from all
 * samples so far and append the arg-max vertex (first maximum, like cublasI?amax) until n_total samples
 * or max distance <= radio. samples must hold n_total entries; *n_out receives the final count,
 * *max_dist the last maximum (INF if never read, as the reference). */
int ptp_farthest_point_sampling_f32(ptp_mesh_t *mesh, uint32_t *samples, uint32_t n_initial, uint32_t n_total,
                                    float radio, uint32_t *n_out, float *max_dist, ptp_stats_t *stats);
int ptp_farthest_point_sampling_f64(ptp_mesh_t *mesh, uint32_t *samples, uint32_t n_initial, uint32_t n_total,
                                    double radio, uint32_t *n_out, double *max_dist, ptp_stats_t *stats);

/* Measurement helper, not part of the reference interface: nanoseconds per grid barrier (the fused
 * arrive + reduce + poll barrier of the single-solve kernels) for `ctas` CTAs of `block` threads; < 0 on error.
 * ctas < 0 measures the hardware barrier of ONE thread-block cluster of -ctas CTAs instead (BFS team). */
double ptp_debug_barrier_ns(int ctas, int block, int n);

/* Verification helper, not part of the reference interface: the inverse Gram matrix of update_step
 * (src/geodesics_ptp.cpp:212-231) is computed with ONE reciprocal shared by its three IEEE divisions; this runs that code
 * and three plain IEEE divisions on `n` generated operand sets (Gram matrices of random edge pairs over the whole exponent
 * range, raw bit patterns, zeros / infinities / NaNs / denormals) and counts results that differ in any bit
 * (*mismatches, expected 0) and the sets that took the shared-reciprocal path (*shared_path). real_size = 4 | 8.
 * samples160: NULL, or room for 16 x 10 doubles describing the first offenders (q00, q01, q11, det, then got / want x 3). */
int ptp_debug_inv_gram_check(uint64_t n, uint64_t seed, int real_size, uint64_t *mismatches, uint64_t *shared_path,
                             double *samples160);

/* Verification helper, not part of the reference interface: the batched sweep decides the acceptance condition of
 * update_step (src/geodesics_ptp.cpp:239-253: c0 < 0 and c1 < 0 after 43 rounded operations) from its two-term form
 * e = Q (t - p) wherever |e| exceeds a proven bound on the difference of the two evaluations (csrc/ptp_device.cuh,
 * tri_front). This runs both on `n` generated cases — random triangles within and at the edge of the admitted shapes,
 * distances random or chosen so that e nearly cancels — and counts the cases in which the short form decided and the
 * reference chain disagrees (*disagreements, expected 0), the cases it decided (*decided) and the cases generated on
 * admitted triangles (*flagged). real_size = 4 | 8. */
int ptp_debug_sign_short_check(uint64_t n, uint64_t seed, int real_size, uint64_t *disagreements, uint64_t *decided,
                               uint64_t *flagged);

/* Verification helper, not part of the reference interface: on triangles whose per-mesh flag bounds the squared edge
 * lengths, the float square root of the Dijkstra fallback (src/geodesics_ptp.cpp:254-259) runs the IEEE sequence without
 * its range test; this compares the two forms over EVERY float of [2^-96, 2^96] (*tested of them; *mismatches expected 0). */
int ptp_debug_sqrt_check(uint64_t *mismatches, uint64_t *tested);

/* Verification helper, not part of the reference interface: the batched sweep skips a triangle with ONE corner above the
 * vertex when update_step (src/geodesics_ptp.cpp:201-262) provably cannot return a value below the vertex's (two-sided
 * causal skip, csrc/ptp_device.cuh: two_sided_ok / two_sided_skip). This evaluates the reference chain on `n` generated
 * cases — random triangles, distances random or at the edges of the rule — and counts the cases in which the rule fired and
 * update_step returned less than the vertex's value (*violations, expected 0), the cases in which it fired (*fired) and
 * the cases generated on admitted triangles (*flagged). real_size = 4 | 8. */
int ptp_debug_two_sided_check(uint64_t n, uint64_t seed, int real_size, uint64_t *violations, uint64_t *fired, uint64_t *flagged);

#ifdef __cplusplus
}
#endif
#endif /* PTP_B200_H */
