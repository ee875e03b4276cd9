// TEST INFRASTRUCTURE — not product code. Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the library built from this file.
//
// Thin extern "C" driver around the UNMODIFIED reference sources, compiled where they lie under
// /root/reference (see oracle/Makefile; outputs go to oracle/_ref/, which is git-ignored).
// It exposes the reference's own CPU implementation of the PTP path so that (1) the plain-C
// restatement in oracle/ptp_oracle.c can be pinned against it and (2) the CPU baseline can be the
// reference itself (cpu_baseline.kind == "reference").
//
// Reference entry points wrapped (file:line relative to /root/reference):
//   che::che(const vertex*, n_v, const index_t*, n_f)        src/che.cpp:84-87 (init :1254-1263)
//   che_off (OFF reader / writer)                             src/che_off.cpp:28-100
//   che::compute_toplesets                                    src/che.cpp:546-593
//   parallel_toplesets_propagation_cpu                        src/geodesics_ptp.cpp:122-199
//   parallel_toplesets_propagation_coalescence_cpu            src/geodesics_ptp.cpp:40-120
//   update_step                                               src/geodesics_ptp.cpp:201-262
//   normalize_ptp                                             src/geodesics_ptp.cpp:264-276
//
// Built twice: default (real_t = double) and with -DSINGLE_P (real_t = float).

#include "geodesics_ptp.h"
#include "che_off.h"

#include <cstring>
#include <vector>

using namespace gproshan;

namespace {

// che with tables injected directly (skips the serial update_evt_ot_et, ~22 s at 10 M vertices).
// Uses only protected members the reference exposes to subclasses (include/che.h:36-47,122-126).
struct che_raw: public che
{
	che_raw(const real_t * xyz, size_t n_v, const index_t * vt, const index_t * ot, const index_t * evt, size_t n_f)
	{
		init(n_v, n_f);
		memcpy(GT, xyz, sizeof(vertex) * n_v);
		memcpy(VT, vt, sizeof(index_t) * 3 * n_f);
		memcpy(OT, ot, sizeof(index_t) * 3 * n_f);
		memcpy(EVT, evt, sizeof(index_t) * n_v);
	}
};

} // namespace

extern "C" {

int ref_sizeof_real() { return (int) sizeof(real_t); }

void * ref_che_create(const real_t * xyz, unsigned n_v, const unsigned * faces, unsigned n_f)
{
	static_assert(sizeof(vertex) == 3 * sizeof(real_t), "vertex must be 3 packed reals");
	return new che((const vertex *) xyz, n_v, faces, n_f);
}

void * ref_che_create_raw(const real_t * xyz, unsigned n_v, const unsigned * vt, const unsigned * ot, const unsigned * evt, unsigned n_f)
{
	return new che_raw(xyz, n_v, vt, ot, evt, n_f);
}

// che_off::che_off(file) — the reference's OFF reader, src/che_off.cpp:16-19,28-80
void * ref_che_read_off(const char * path) { return new che_off(path); }

// che_off::write_file, src/che_off.cpp:82-100 (appends ".off" to the name itself)
void ref_che_write_off(void * m, const char * path_without_ext) { che_off::write_file((che *) m, path_without_ext); }

void ref_che_destroy(void * m) { delete (che *) m; }

unsigned ref_che_n_vertices(void * m) { return (unsigned) ((che *) m)->n_vertices(); }
unsigned ref_che_n_half_edges(void * m) { return (unsigned) ((che *) m)->n_half_edges(); }

// copy out the CHE tables the PTP path consumes (GT, VT, OT, EVT)
void ref_che_tables(void * m_, real_t * gt, unsigned * vt, unsigned * ot, unsigned * evt)
{
	che * m = (che *) m_;
	for(index_t v = 0; v < m->n_vertices(); v++)
	{
		const vertex & p = m->gt(v);
		if(gt) { gt[3 * v] = p.x; gt[3 * v + 1] = p.y; gt[3 * v + 2] = p.z; }
		if(evt) evt[v] = m->evt(v);
	}
	for(index_t he = 0; he < m->n_half_edges(); he++)
	{
		if(vt) vt[he] = m->vt(he);
		if(ot) ot[he] = m->ot(he);
	}
}

// returns number of entries written to limits (0 if sources empty); limits must hold n_v + 2
unsigned ref_compute_toplesets(void * m_, const unsigned * sources, unsigned n_sources, unsigned k,
								unsigned * toplesets, unsigned * sorted, unsigned * limits)
{
	che * m = (che *) m_;
	std::vector<index_t> src(sources, sources + n_sources);
	std::vector<index_t> lim;
	index_t * t = toplesets, * s = sorted;
	m->compute_toplesets(t, s, lim, src, k);
	memcpy(limits, lim.data(), sizeof(index_t) * lim.size());
	return (unsigned) lim.size();
}

void ref_ptp_cpu(void * m_, const unsigned * sources, unsigned n_sources, const unsigned * limits, unsigned n_limits,
				const unsigned * sorted, real_t * dist, unsigned * clusters)
{
	che * m = (che *) m_;
	std::vector<index_t> src(sources, sources + n_sources);
	std::vector<index_t> lim(limits, limits + n_limits);
	const index_t * idx = sorted;
	parallel_toplesets_propagation_cpu({dist, clusters}, m, src, {lim, idx});
}

void ref_ptp_coalescence_cpu(void * m_, const unsigned * sources, unsigned n_sources, const unsigned * limits, unsigned n_limits,
				const unsigned * sorted, real_t * dist, unsigned * clusters)
{
	che * m = (che *) m_;
	std::vector<index_t> src(sources, sources + n_sources);
	std::vector<index_t> lim(limits, limits + n_limits);
	const index_t * idx = sorted;
	parallel_toplesets_propagation_coalescence_cpu({dist, clusters}, m, src, {lim, idx});
}

real_t ref_update_step(void * m_, const real_t * dist, unsigned he)
{
	return update_step((che *) m_, dist, he);
}

void ref_normalize_ptp(real_t * dist, size_t n)
{
	normalize_ptp(dist, n);
}

} // extern "C"
