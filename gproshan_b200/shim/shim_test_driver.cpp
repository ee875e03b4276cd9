// TEST INFRASTRUCTURE: extern "C" handles so tests can call the C++ shim (reference signatures) and the
// reference's own CPU PTP on the SAME gproshan::che object, built by the reference's own constructor.
#include "geodesics.h"
#include "geodesics_ptp.h"

#include <cstring>
#include <vector>

using namespace gproshan;

namespace gproshan {
void ptp_b200_release(che * mesh);
size_t key_components_ptp_b200(che * mesh, const std::vector<index_t> & key_points, real_t radio_fraction, index_t * comp_out);
double distance_rows_ptp_b200(che * mesh, const std::vector<index_t> & points, distance_t * rows, int n_devices);
index_t ** sampling_shape_ptp_b200(std::vector<index_t> & points, size_t *& sizes, vertex *& normals, che * shape, size_t n_points, distance_t radio);
double geodesics_ptp_b200(che * mesh, const std::vector<index_t> & sources, distance_t * dist, index_t * clusters, index_t * sorted_index);
}

extern "C" {

int shim_sizeof_real() { return (int) sizeof(real_t); }

void * shim_che_create(const real_t * xyz, unsigned n_v, const unsigned * faces, unsigned n_f)
{
    return new che((const vertex *) xyz, n_v, faces, n_f);
}

void shim_che_destroy(void * m)
{
    ptp_b200_release((che *) m);
    delete (che *) m;
}

// compute_toplesets (reference, CPU) -> parallel_toplesets_propagation_gpu (shim) and ..._cpu (reference)
double shim_ptp_gpu_vs_cpu(void * m_, const unsigned * sources, unsigned n_sources, int coalescence, real_t * dist_gpu, real_t * dist_cpu,
                           unsigned * clusters_gpu)
{
    che * m = (che *) m_;
    std::vector<index_t> src(sources, sources + n_sources), limits;
    index_t * toplesets = new index_t[m->n_vertices()];
    index_t * sorted = new index_t[m->n_vertices() + n_sources];
    m->compute_toplesets(toplesets, sorted, limits, src);
    const index_t * idx = sorted;
    double secs;
    if(coalescence)
        secs = parallel_toplesets_propagation_coalescence_gpu({dist_gpu, clusters_gpu}, m, src, {limits, idx});
    else
        secs = parallel_toplesets_propagation_gpu({dist_gpu, clusters_gpu}, m, src, {limits, idx});
    if(limits.size() >= 3)
        parallel_toplesets_propagation_cpu(dist_cpu, m, src, {limits, idx});
    delete [] toplesets;
    delete [] sorted;
    return secs;
}

double shim_geodesics(void * m_, const unsigned * sources, unsigned n_sources, real_t * dist, unsigned * clusters, unsigned * sorted_index)
{
    std::vector<index_t> src(sources, sources + n_sources);
    return geodesics_ptp_b200((che *) m_, src, dist, clusters, sorted_index);
}

unsigned shim_fps(void * m_, unsigned * samples, unsigned n0, unsigned n, real_t radio, real_t * max_dist, double * secs)
{
    std::vector<index_t> s(samples, samples + n0);
    *max_dist = farthest_point_sampling_ptp_gpu((che *) m_, s, *secs, n, radio);
    memcpy(samples, s.data(), sizeof(index_t) * s.size());
    return (unsigned) s.size();
}

// the `geodesics` class itself (include/geodesics.h) with option PTP_GPU (1) or PTP_CPU; copies out what its public
// interface exposes: operator[] -> dist, operator() -> sorted_index, clusters, n_sorted_index(); then normalize() and
// operator[] again -> dist_normalized. e_dist != 0 exercises the external-allocation constructor argument.
unsigned shim_geodesics_class(void * m_, const unsigned * sources, unsigned n_sources, int opt, int cluster, int external_dist,
                              real_t * dist, unsigned * sorted_index, unsigned * clusters, real_t * dist_normalized)
{
    che * m = (che *) m_;
    std::vector<index_t> src(sources, sources + n_sources);
    const size_t n = m->n_vertices();
    std::vector<distance_t> ext(external_dist ? n : 0);
    geodesics g(m, src, (geodesics::option_t) opt, external_dist ? ext.data() : nullptr, cluster != 0);
    for(index_t v = 0; v < n; v++)
    {
        dist[v] = g[v];
        sorted_index[v] = g(v);
        if(cluster) clusters[v] = g.clusters[v];
    }
    const unsigned ns = (unsigned) g.n_sorted_index();
    g.normalize();
    for(index_t v = 0; v < n; v++)
        dist_normalized[v] = g[v];
    return ns;
}

// in-place edit of the che (what smoothing / noise tools do through che::get_vertex / set_vertices): the shim must notice
void shim_che_set_vertices(void * m_, const real_t * xyz)
{
    che * m = (che *) m_;
    m->set_vertices((const vertex *) xyz);
}

// both entry points with the caller's arrays pre-filled (the coalescence arm must leave unreached entries untouched)
double shim_ptp_gpu_prefilled(void * m_, const unsigned * sources, unsigned n_sources, int coalescence, real_t * dist_io, unsigned * clusters_io)
{
    che * m = (che *) m_;
    std::vector<index_t> src(sources, sources + n_sources), limits;
    index_t * toplesets = new index_t[m->n_vertices()];
    index_t * sorted = new index_t[m->n_vertices() + n_sources];
    m->compute_toplesets(toplesets, sorted, limits, src);
    const index_t * idx = sorted;
    const double secs = coalescence ? parallel_toplesets_propagation_coalescence_gpu({dist_io, clusters_io}, m, src, {limits, idx})
                                    : parallel_toplesets_propagation_gpu({dist_io, clusters_io}, m, src, {limits, idx});
    delete [] toplesets;
    delete [] sorted;
    return secs;
}

double shim_distance_rows(void * m_, const unsigned * points, unsigned n_points, real_t * rows, int n_devices)
{
    std::vector<index_t> p(points, points + n_points);
    return distance_rows_ptp_b200((che *) m_, p, rows, n_devices);
}

// sampling_shape_ptp_b200 flattened: sizes[n_points], then the patches back to back in `flat` (capacity flat_cap)
unsigned long shim_sampling_shape(void * m_, const unsigned * points, unsigned n_points, real_t radio, unsigned long * sizes_out, unsigned * flat, unsigned long flat_cap)
{
    std::vector<index_t> p(points, points + n_points);
    size_t * sizes = nullptr;
    vertex * normals = nullptr;
    index_t ** idx = sampling_shape_ptp_b200(p, sizes, normals, (che *) m_, n_points, radio);
    unsigned long total = 0;
    for(unsigned i = 0; i < n_points; i++)
    {
        sizes_out[i] = sizes[i];
        for(size_t k = 0; k < sizes[i] && total < flat_cap; k++) flat[total++] = idx[i][k];
        delete [] idx[i];
    }
    delete [] idx;
    delete [] sizes;
    delete [] normals;
    return total;
}

unsigned long shim_key_components(void * m_, const unsigned * kps, unsigned n_kps, real_t radio_fraction, unsigned * comp_out)
{
    std::vector<index_t> k(kps, kps + n_kps);
    return key_components_ptp_b200((che *) m_, k, radio_fraction, comp_out);
}

int shim_option_ptp_gpu() { return (int) geodesics::PTP_GPU; }
int shim_option_ptp_cpu() { return (int) geodesics::PTP_CPU; }
int shim_option_fm() { return (int) geodesics::FM; }

void shim_normalize_ptp(real_t * dist, unsigned n) { normalize_ptp(dist, n); }

} // extern "C"
