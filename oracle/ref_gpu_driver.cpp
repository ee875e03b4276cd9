// TEST / BENCH INFRASTRUCTURE — not product code. Second baseline: the reference's OWN CUDA implementation of
// the PTP path, compiled UNMODIFIED for sm_100a from /root/reference/src/cuda (oracle/Makefile `refgpu`; the
// output goes to oracle/_ref/, git-ignored). Only bench.py's `reference_gpu` leg and tests/ may load it, and
// only in a process of their own: the reference calls cudaDeviceReset() at the top of every solve
// (src/cuda/geodesics_ptp.cu:22), which would tear down any other CUDA user of the process.
//
// Reference entry points wrapped (file:line relative to /root/reference):
//   parallel_toplesets_propagation_gpu              src/cuda/geodesics_ptp.cu:20-85
//   parallel_toplesets_propagation_coalescence_gpu  src/cuda/geodesics_ptp_coalescence.cu:21-100
//   farthest_point_sampling_ptp_gpu                 src/cuda/geodesics_ptp.cu:87-172
//   iter_error_parallel_toplesets_propagation_gpu   src/cuda/test_geodesics_ptp.cu:20-70 (its harness loop :164-211)
// The mesh handle is the `che *` made by ref_driver.cpp (linked into the same library).

#include "geodesics_ptp.h"
#include "test_geodesics_ptp.h"

#include <cmath>
#include <vector>

using namespace gproshan;

// src/cuda/test_geodesics_ptp.cu calls compute_error, which the reference defines in src/test_geodesics_ptp.cpp:361-370
// — a file that cannot be built here (it pulls in the heat method: CHOLMOD). Its eight lines are restated for the link.
namespace gproshan {
distance_t compute_error(const distance_t * dist, const distance_t * exact, const size_t & n, const size_t & s)
{
	distance_t error = 0;
	#pragma omp parallel for reduction(+: error)
	for(index_t v = 0; v < n; v++)
		if(exact[v] > 0) error += std::abs(dist[v] - exact[v]) / exact[v];
	return error * 100 / (n - s);
}
} // namespace gproshan

extern "C" {

// returns the reference's own timer (seconds, CUDA events around upload + loop + download)
double ref_ptp_gpu(void * m_, const unsigned * sources, unsigned n_sources, const unsigned * limits, unsigned n_limits,
				const unsigned * sorted, real_t * dist, unsigned * clusters)
{
	che * m = (che *) m_;
	std::vector<index_t> src(sources, sources + n_sources);
	std::vector<index_t> lim(limits, limits + n_limits);
	const index_t * idx = sorted;
	return parallel_toplesets_propagation_gpu({dist, clusters}, m, src, {lim, idx});
}

// the single-source arm of geodesics::run_parallel_toplesets_propagation_gpu (src/geodesics.cpp:233-236);
// ptp_coalescence (a new che per solve, serial) runs inside the call but outside the reference's own timer
double ref_ptp_coalescence_gpu(void * m_, const unsigned * sources, unsigned n_sources, const unsigned * limits, unsigned n_limits,
				const unsigned * sorted, real_t * dist, unsigned * clusters)
{
	che * m = (che *) m_;
	std::vector<index_t> src(sources, sources + n_sources);
	std::vector<index_t> lim(limits, limits + n_limits);
	const index_t * idx = sorted;
	return parallel_toplesets_propagation_coalescence_gpu({dist, clusters}, m, src, {lim, idx});
}

// samples_io holds n_in samples on entry and room for n; returns the number of samples on exit
unsigned ref_fps_gpu(void * m_, unsigned * samples_io, unsigned n_in, unsigned n, real_t radio, real_t * max_dist, double * seconds)
{
	che * m = (che *) m_;
	std::vector<index_t> s(samples_io, samples_io + n_in);
	*max_dist = farthest_point_sampling_ptp_gpu(m, s, *seconds, n, radio);
	for(size_t i = 0; i < s.size(); i++) samples_io[i] = s[i];
	return (unsigned) s.size();
}

// per-iteration error of the reference's harness; returns the number of (iteration, error) records written
unsigned ref_iter_error_gpu(void * m_, const unsigned * sources, unsigned n_sources, const unsigned * limits, unsigned n_limits,
				const unsigned * sorted, const real_t * exact, unsigned * iters, real_t * errors, unsigned cap)
{
	che * m = (che *) m_;
	std::vector<index_t> src(sources, sources + n_sources);
	std::vector<index_t> lim(limits, limits + n_limits);
	double t;
	std::vector<std::pair<index_t, distance_t> > r = iter_error_parallel_toplesets_propagation_gpu(m, src, lim, sorted, exact, t);
	unsigned n = 0;
	for(; n < r.size() && n < cap; n++) { iters[n] = r[n].first; errors[n] = r[n].second; }
	return n;
}

} // extern "C"
