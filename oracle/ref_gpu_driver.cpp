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
// The mesh handle is the `che *` made by ref_driver.cpp (linked into the same library).

#include "geodesics_ptp.h"

#include <vector>

using namespace gproshan;

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

} // extern "C"
