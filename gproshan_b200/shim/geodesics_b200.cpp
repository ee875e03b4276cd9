// Drop-in for the PTP arms of gproshan's `geodesics` class (include/geodesics.h:18-72, src/geodesics.cpp:16-110,
// 213-240). Compiled INSIDE a gproshan build against gproshan's own geodesics.h — same constructor signature, same
// public members and accessors — in place of src/geodesics.cpp when only the PTP options are wanted, or as the model
// for the one-function patch of INTEGRATION.md §3 when the Fast Marching / heat arms (Armadillo, CHOLMOD) stay.
//
//   option_t::PTP_GPU  -> toplesets AND solve on the device (geodesics_ptp_b200: ptp_geodesics_{f32,f64}); `dist`,
//                         `clusters` and `sorted_index` are filled exactly as run_parallel_toplesets_propagation_gpu does
//                         through che::compute_toplesets + parallel_toplesets_propagation[_coalescence]_gpu
//                         (src/geodesics.cpp:225-240): same distances as the reference's CPU PTP bit for bit, same BFS order.
//   option_t::PTP_CPU  -> the reference's own run_parallel_toplesets_propagation_cpu (src/geodesics.cpp:213-223), i.e.
//                         gproshan's parallel_toplesets_propagation_cpu: unchanged code path, kept so that callers that
//                         compare the two arms (src/test_geodesics_ptp.cpp) still link.
//   FM, HEAT_FLOW, HEAT_FLOW_GPU are other algorithms (out of scope) and are absent at compile time from
//   this file: selecting one reports it on stderr and leaves the distances at INFINITY (no assert, no abort).
//
// n_sorted stays 0 after PTP, exactly as in the reference (src/geodesics.cpp:25): radio() / farthest() are therefore
// invalid after PTP there too, and normalize() takes the normalize_ptp branch (src/geodesics.cpp:76-90).
#include "geodesics.h"
#include "geodesics_ptp.h"

#include <cassert>
#include <cstdio>
#include <cstring>

using namespace std;

namespace gproshan {

// gproshan_b200/shim/geodesics_ptp_b200.cpp
double geodesics_ptp_b200(che * mesh, const std::vector<index_t> & sources, distance_t * dist, index_t * clusters, index_t * sorted_index);

geodesics::geodesics(che * mesh, const vector<index_t> & sources, const option_t & opt, distance_t *const & e_dist, const bool & cluster, const size_t & n_iter, const distance_t & radio): n_vertices(mesh->n_vertices())
{
	assert(n_vertices > 0);

	free_dist = e_dist == nullptr;
	dist = free_dist ? new distance_t[n_vertices] : e_dist;
	clusters = cluster ? new index_t[n_vertices] : nullptr;
	sorted_index = new index_t[n_vertices];
	n_sorted = 0;

	memset(sorted_index, -1, n_vertices * sizeof(index_t));
	for(index_t v = 0; v < n_vertices; v++)
		dist[v] = INFINITY;

	assert(sources.size() > 0);
	execute(mesh, sources, n_iter, radio, opt);
}

geodesics::~geodesics()
{
	if(free_dist)		delete [] dist;
	if(sorted_index)	delete [] sorted_index;
	if(clusters)		delete [] clusters;
}

const distance_t & geodesics::operator[](const index_t & i) const
{
	assert(i < n_vertices);
	return dist[i];
}

const index_t & geodesics::operator()(const index_t & i) const
{
	assert(i < n_vertices);
	return sorted_index[i];
}

const distance_t & geodesics::radio() const
{
	assert(n_sorted != 0);
	return dist[farthest()];
}

const index_t & geodesics::farthest() const
{
	assert(n_sorted != 0);
	return sorted_index[n_sorted - 1];
}

const size_t & geodesics::n_sorted_index() const
{
	return n_sorted;
}

void geodesics::copy_sorted_index(index_t * indexes, const size_t & n) const
{
	assert(n <= n_sorted);
	memcpy(indexes, sorted_index, n * sizeof(index_t));
}

void geodesics::normalize()
{
	if(!n_sorted)
	{
		normalize_ptp(dist, n_vertices);	// src/geodesics_ptp.cpp:264-276: divide by the largest finite distance
		return;
	}

	distance_t max = dist[farthest()];

	#pragma omp parallel for
	for(size_t i = 0; i < n_sorted; i++)
		dist[sorted_index[i]] /= max;
}

void geodesics::execute(che * mesh, const vector<index_t> & sources, const size_t & n_iter, const distance_t & radio, const option_t & opt)
{
	switch(opt)
	{
		case PTP_CPU: run_parallel_toplesets_propagation_cpu(mesh, sources, n_iter, radio);
			break;
#ifdef GPROSHAN_CUDA
		case PTP_GPU: run_parallel_toplesets_propagation_gpu(mesh, sources, n_iter, radio);
			break;
#endif // GPROSHAN_CUDA
		default:
			fprintf(stderr, "[ptp_b200] geodesics: option %d is not a PTP option; this build of the class provides PTP_GPU and PTP_CPU only\n", (int) opt);
			break;
	}
}

// src/geodesics.cpp:213-223, unchanged: the reference's CPU arm
void geodesics::run_parallel_toplesets_propagation_cpu(che * mesh, const vector<index_t> & sources, const size_t &, const distance_t &)
{
	index_t * toplesets = new index_t[n_vertices];
	vector<index_t> limits;
	mesh->compute_toplesets(toplesets, sorted_index, limits, sources);

	parallel_toplesets_propagation_cpu({dist, clusters}, mesh, sources, {limits, sorted_index});

	delete [] toplesets;
}

#ifdef GPROSHAN_CUDA
// src/geodesics.cpp:225-240 with the serial che::compute_toplesets (3.5 s at 10 M vertices) moved onto the device
void geodesics::run_parallel_toplesets_propagation_gpu(che * mesh, const vector<index_t> & sources, const size_t &, const distance_t &)
{
	const double time_ptp = geodesics_ptp_b200(mesh, sources, dist, clusters, sorted_index);
	if(time_ptp < 0)
		fprintf(stderr, "[ptp_b200] geodesics(PTP_GPU) failed; distances left at INFINITY\n");
}
#endif // GPROSHAN_CUDA

} // namespace gproshan
