// Drop-in replacement for gproshan's CUDA PTP translation units
//   src/cuda/geodesics_ptp.cu  and  src/cuda/geodesics_ptp_coalescence.cu
// It is compiled INSIDE a gproshan build (against gproshan's own include/geodesics_ptp.h and include/che.h;
// nothing from gproshan is copied here) and defines, with the reference's exact signatures, the three host
// entry points those files export (include/geodesics_ptp.h:34,36,42). Each forwards to the extern "C" ABI of
// libptp_b200.so (include/ptp_b200.h). See INTEGRATION.md for the CMake change.
//
// Differences a caller can observe, all deliberate:
//   * no cudaDeviceReset() (src/cuda/geodesics_ptp.cu:22,89): the mesh stays resident between calls, keyed by
//     the che* (call gproshan::ptp_b200_release(mesh) before deleting or editing a mesh);
//   * distances are those of parallel_toplesets_propagation_cpu bit for bit (the reference GPU code returns
//     the other Jacobi buffer and contracts FMAs);
//   * clusters of unreached vertices are left untouched only in the sense of the reference's host array: they
//     receive NIL here (the reference uploads and downloads whatever the caller's array held);
//   * CUDA failures are reported on stderr and the call returns -1 seconds instead of being ignored.
#include "geodesics_ptp.h"

#include "ptp_b200.h"

#include <cmath>
#include <cstdio>
#include <map>
#include <mutex>
#include <type_traits>

namespace gproshan {

namespace {

constexpr bool kSingle = std::is_same<real_t, float>::value;

struct Resident { ptp_mesh_t *h; size_t nv, nhe; };
std::map<che *, Resident> g_cache;
std::mutex g_mu;

ptp_mesh_t *resident(che *mesh)
{
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_cache.find(mesh);
    if (it != g_cache.end()) {
        if (it->second.nv == mesh->n_vertices() && it->second.nhe == mesh->n_half_edges()) return it->second.h;
        ptp_mesh_destroy(it->second.h); // same address, different mesh: rebuild
        g_cache.erase(it);
    }
    CHE tables(mesh); // friend view of GT / VT / OT / EVT (include/che.h:129,134-146; src/che.cpp:36-46)
    ptp_mesh_t *h = nullptr;
    int rc;
    if constexpr (kSingle)
        rc = ptp_mesh_create_f32((const float *) tables.GT, tables.VT, tables.OT, tables.EVT, tables.n_vertices, tables.n_half_edges, 0, &h);
    else
        rc = ptp_mesh_create_f64((const double *) tables.GT, tables.VT, tables.OT, tables.EVT, tables.n_vertices, tables.n_half_edges, 0, &h);
    if (rc != PTP_OK) {
        fprintf(stderr, "[ptp_b200] mesh upload failed: %s\n", ptp_last_error());
        return nullptr;
    }
    g_cache[mesh] = {h, mesh->n_vertices(), mesh->n_half_edges()};
    return h;
}

double solve(const ptp_out_t & ptp_out, che * mesh, const std::vector<index_t> & sources, const toplesets_t & toplesets)
{
    ptp_mesh_t *h = resident(mesh);
    if (!h) return -1;
    ptp_stats_t st;
    int rc;
    if constexpr (kSingle)
        rc = ptp_solve_f32(h, sources.data(), (uint32_t) sources.size(), toplesets.limits.data(), (uint32_t) toplesets.limits.size(),
                           toplesets.index, (float *) ptp_out.dist, ptp_out.clusters, NIL, &st);
    else
        rc = ptp_solve_f64(h, sources.data(), (uint32_t) sources.size(), toplesets.limits.data(), (uint32_t) toplesets.limits.size(),
                           toplesets.index, (double *) ptp_out.dist, ptp_out.clusters, NIL, &st);
    if (rc != PTP_OK) {
        fprintf(stderr, "[ptp_b200] solve failed: %s\n", ptp_last_error());
        return -1;
    }
    return st.ms_total / 1000; // seconds, like the reference's cudaEvent timing (src/cuda/geodesics_ptp.cu:77-84)
}

} // namespace

// release the device copy of a mesh (call before the che is deleted or edited)
void ptp_b200_release(che * mesh)
{
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_cache.find(mesh);
    if (it != g_cache.end()) {
        ptp_mesh_destroy(it->second.h);
        g_cache.erase(it);
    }
}

// toplesets + solve on the device: what geodesics::run_parallel_toplesets_propagation_gpu (src/geodesics.cpp:225-240)
// does with a CPU BFS in front. sorted_index may be null.
double geodesics_ptp_b200(che * mesh, const std::vector<index_t> & sources, distance_t * dist, index_t * clusters, index_t * sorted_index)
{
    ptp_mesh_t *h = resident(mesh);
    if (!h) return -1;
    ptp_stats_t st;
    int rc;
    if constexpr (kSingle)
        rc = ptp_geodesics_f32(h, sources.data(), (uint32_t) sources.size(), (float *) dist, clusters, NIL, sorted_index, mesh->n_vertices(), &st);
    else
        rc = ptp_geodesics_f64(h, sources.data(), (uint32_t) sources.size(), (double *) dist, clusters, NIL, sorted_index, mesh->n_vertices(), &st);
    if (rc != PTP_OK && rc != PTP_ERR_CAPACITY) { // CAPACITY: duplicate sources overflow the caller's V-entry sorted_index
        fprintf(stderr, "[ptp_b200] geodesics failed: %s\n", ptp_last_error());
        return -1;
    }
    return st.ms_total / 1000;
}

// include/geodesics_ptp.h:36
double parallel_toplesets_propagation_gpu(const ptp_out_t & ptp_out, che * mesh, const std::vector<index_t> & sources, const toplesets_t & toplesets)
{
    return solve(ptp_out, mesh, sources, toplesets);
}

// include/geodesics_ptp.h:34 — the reference rebuilds a topleset-ordered che on the CPU here (ptp_coalescence,
// ~22 s at 10 M vertices); the library does that re-ordering on the device for every solve, so both entry points
// share one path. set_inf is unused by the reference as well (src/cuda/geodesics_ptp_coalescence.cu:21).
double parallel_toplesets_propagation_coalescence_gpu(const ptp_out_t & ptp_out, che * mesh, const std::vector<index_t> & sources, const toplesets_t & toplesets, const bool & )
{
    return solve(ptp_out, mesh, sources, toplesets);
}

// include/geodesics_ptp.h:42
distance_t farthest_point_sampling_ptp_gpu(che * mesh, std::vector<index_t> & samples, double & time_fps, size_t n, distance_t radio)
{
    time_fps = 0;
    ptp_mesh_t *h = resident(mesh);
    if (!h || samples.empty()) return INFINITY;
    if (n >= mesh->n_vertices()) n = mesh->n_vertices() >> 1; // src/cuda/geodesics_ptp.cu:125
    const size_t n0 = samples.size();
    if (n <= n0) return INFINITY;
    samples.resize(n);
    uint32_t n_out = 0;
    distance_t max_dist = INFINITY;
    ptp_stats_t st;
    int rc;
    if constexpr (kSingle)
        rc = ptp_farthest_point_sampling_f32(h, samples.data(), (uint32_t) n0, (uint32_t) n, (float) radio, &n_out, (float *) &max_dist, &st);
    else
        rc = ptp_farthest_point_sampling_f64(h, samples.data(), (uint32_t) n0, (uint32_t) n, (double) radio, &n_out, (double *) &max_dist, &st);
    if (rc != PTP_OK) {
        fprintf(stderr, "[ptp_b200] farthest point sampling failed: %s\n", ptp_last_error());
        samples.resize(n0);
        return INFINITY;
    }
    samples.resize(n_out);
    time_fps = st.ms_total / 1000;
    return max_dist;
}

} // namespace gproshan
