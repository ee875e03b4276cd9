// TEST INFRASTRUCTURE: extern "C" handles so tests can call the C++ shim (reference signatures) and the
// reference's own CPU PTP on the SAME gproshan::che object, built by the reference's own constructor.
#include "geodesics_ptp.h"

#include <cstring>
#include <vector>

using namespace gproshan;

namespace gproshan {
void ptp_b200_release(che * mesh);
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

} // extern "C"
