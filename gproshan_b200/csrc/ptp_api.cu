// Host side + kernels entry points of libptp_b200.so: the extern "C" ABI declared in include/ptp_b200.h.
// Device algorithms live in ptp_device.cuh. Build: gproshan_b200/build.py (nvcc, sm_100a, -lineinfo).
#include "../../include/ptp_b200.h"
#include "ptp_device.cuh"
#ifndef PTP_MAXL1
#define PTP_MAXL1 0
#endif

#include <cuda_profiler_api.h>

#include <algorithm>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <dlfcn.h>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <vector>

using namespace ptp;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) {                                                                              \
            char b_[512];                                                                                     \
            snprintf(b_, sizeof b_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));     \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorNoKernelImageForDevice ||                   \
                                e_ == cudaErrorInsufficientDriver                                             \
                            ? PTP_ERR_NO_DEVICE                                                               \
                            : PTP_ERR_CUDA,                                                                   \
                        b_);                                                                                  \
        }                                                                                                     \
    } while (0)

// Behaviour switches (ptp_set_option / ptp_get_option, include/ptp_b200.h). Each starts from the environment variable
// PTP_<NAME IN CAPITALS> when it is set (read once, at the first use of the library) and can be changed at any time
// through the API; the library reads the environment nowhere else.
struct Option { const char *name; long value; const char *doc; };
Option g_options[] = {
    {"fused", 5, "single solve: 5 BFS-cluster kernel + sweep kernel side by side (falls back to 4), 4 the same two teams in one "
                 "cluster launch, 1 two-team kernel with grid barriers, 0 three launches, 2 debug (the teams one after the other)"},
    {"stage", -1, "window staged in shared memory: -1 per-variant default (on for the three-launch sweep only), 0 off, "
                  "1 also in the cluster kernel, 2 also in the two-team kernel"},
    {"bfs_ctas", 0, "CTAs of the BFS team in the two-team kernel (0 = one per SM)"},
    {"cluster", 8, "CTAs of the BFS thread-block cluster (<= 8 portable, <= 16 non-portable)"},
    {"geo_single", 0, "single solve reads the per-mesh geometry table"},
    {"geo", 0, "batched solves read the per-mesh geometry table"},
    {"elastic", 1, "one-CTA-per-solve batched kernel: idle CTAs execute ticketed chunks of running solves"},
    {"causal", 1, "batched solves skip triangles that provably cannot lower a vertex (both neighbours above it; bit-exact)"},
    {"sign_short", 1, "batched solves (with causal) decide update_step's acceptance condition from its two-term form where provably equal (bit-exact)"},
    {"two_sided", 1, "batched solves (with causal) also skip triangles with one corner above the vertex when the other provably cannot reach it (bit-exact)"},
    {"team", 0, "batched solves: CTAs per solve (0 = 1 when the batch fills the chip, num_sms / batch otherwise; 1 = always one CTA per solve)"},
    {"newest", 0, "return the Jacobi buffer WRITTEN by the last iteration (what the reference's CUDA code copies back, "
                  "src/cuda/geodesics_ptp.cu:60-66) instead of the one it read (the reference's CPU code, src/geodesics_ptp.cpp:193-198)"},
    {"rows_chunk_mb", 8192, "batched solves into HOST rows: size of the device staging buffer; a batch whose rows exceed it is solved in "
                            "several launches"},
    {"stream_rows", 1, "batched solves into HOST rows: copy finished rows to the host while the kernel is still solving the rest"},
    {"gather_chunks", 0, "ptp_solve_batched_multi_*: each device's shard is solved in this many pieces; the NCCL transfer of a "
                         "piece to the root device runs while the next piece is being solved (0 = one piece per two waves of CTAs, "
                         "at least 1: pieces smaller than that cost more in solver efficiency than the overlap returns)"},
    {"profile_range", 0, "bracket every single solve / batched call with cudaProfilerStart/Stop (ncu --replay-mode app-range)"},
    {"debug", 0, "print per-phase device timers to stderr"},
};
constexpr int N_OPTIONS = (int)(sizeof g_options / sizeof g_options[0]);
std::mutex g_opt_mu;

void options_init()
{
    static bool done = [] {
        for (Option &o : g_options) {
            std::string env = "PTP_";
            for (const char *c = o.name; *c; c++) env += (char)toupper((unsigned char)*c);
            if (const char *e = getenv(env.c_str())) o.value = atol(e);
        }
        return true;
    }();
    (void)done;
}

long opt(const char *name)
{
    options_init();
    for (const Option &o : g_options)
        if (!strcmp(o.name, name)) return o.value;
    return 0;
}

constexpr int GRID_BLOCK = 512;    // threads per CTA, cooperative BFS kernel
constexpr int FUSED_BLOCK = 512;   // fused single-solve kernel: 2 CTAs per SM (one BFS-team CTA + one sweep-team CTA)
#ifndef PTP_CLUSTER_BLOCK
#define PTP_CLUSTER_BLOCK 768 // measured on C3 (f64): 768 threads / 80 registers 27.9 ms, 1024 / 64 29.8, 512 / 128 28.0
#endif
constexpr int CLUSTER_BLOCK = PTP_CLUSTER_BLOCK; // cluster single-solve kernel: 1 CTA per SM (BFS cluster + sweep team)
constexpr int SOLVE_BLOCK = 1024;  // threads per CTA, cooperative sweep kernel (one pass per iteration on C3-size windows)
template <class R> struct BatchCfg;                    // threads per CTA, one-solve-per-CTA kernels
#ifndef PTP_BATCH_BLOCK_F32
#define PTP_BATCH_BLOCK_F32 1024
#endif
#ifndef PTP_BATCH_MINBLOCKS_F32
#define PTP_BATCH_MINBLOCKS_F32 1
#endif
template <> struct BatchCfg<float> { static constexpr int BLOCK = PTP_BATCH_BLOCK_F32; static constexpr int MINB = PTP_BATCH_MINBLOCKS_F32; };
template <> struct BatchCfg<double> { static constexpr int BLOCK = 512; static constexpr int MINB = 1; };
constexpr int FLAT_BLOCK = 256;
#ifndef PTP_GRID_MAP
#define PTP_GRID_MAP 4 // lanes per vertex in the whole-GPU sweep: 8 (one triangle per lane) or 4 (two per lane)
#endif

// ------------------------------------------------------------------------------------------------
// kernels

template <class R> __global__ void k_pad_gt(const R *__restrict__ gt, R *__restrict__ gt4, u32 V)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; // one thread per output scalar
    if (i < (size_t)V * 4) {
        const size_t v = i >> 2, c = i & 3;
        gt4[i] = c < 3 ? gt[v * 3 + c] : R(0);
    }
}

__device__ __forceinline__ u32 he_next(u32 he) { return 3 * (he / 3) + (he + 1) % 3; }
__device__ __forceinline__ u32 he_prev(u32 he) { return 3 * (he / 3) + (he + 2) % 3; }

// One-ring rows in for_star order (include/che.h:10): he = EVT[v]; he = OT[prev(he)] until back at EVT[v] or NIL.
// Entry k = VT[next(he_k)]; an open fan appends VT[prev(he_last)] (che::link, src/che.cpp:102-112).
// pass 0: rows with <= 8 entries are final, longer ones get the OVF marker and are counted;
// pass 1: overflow rows allocate from the pool and store their entries.
__global__ void k_ring_build(const u32 *__restrict__ VT, const u32 *__restrict__ OT, const u32 *__restrict__ EVT, u32 V,
                             u32 H, u32 *__restrict__ ring8, u32 *__restrict__ pool, ull *counters, int pass)
{
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    u32 *row = ring8 + (size_t)v * GL;
    if (pass == 1 && row[0] != OVF) return;

    const u32 he0 = EVT[v];
    u32 buf[GL];
    u32 cnt = 0, last = NIL;
    bool open = false, bad = false;
    u32 off = 0;
    if (pass == 1) off = row[1];
    if (he0 != NIL) {
        if (he0 >= H) bad = true;
        u32 he = he0;
        while (!bad) {
            const u32 n = VT[he_next(he)];
            if (n >= V) { bad = true; break; }
            if (cnt < GL) buf[cnt] = n;
            if (pass == 1) pool[off + cnt] = n;
            cnt++;
            const u32 ph = he_prev(he);
            const u32 o = OT[ph];
            if (o == NIL) {
                open = true;
                last = VT[ph];
                if (last >= V) bad = true;
                break;
            }
            if (o >= H || cnt >= (1u << 23)) { bad = true; break; }
            he = o;
            if (he == he0) break;
        }
    }
    if (bad) {
        atomicAdd(counters + 1, 1ull);
        for (u32 k = 0; k < GL; k++) row[k] = NIL;
        return;
    }
    const u32 entries = cnt + (open ? 1u : 0u);
    if (pass == 1) {
        if (open) pool[off + cnt] = last;
        return;
    }
    if (entries <= GL) {
        if (open) buf[cnt] = last;
        for (u32 k = 0; k < GL; k++) row[k] = k < entries ? buf[k] : NIL;
        if (open) row[0] |= OPEN_BIT;
    } else {
        const u32 o2 = (u32)atomicAdd(counters, (ull)entries);
        row[0] = OVF;
        row[1] = o2;
        row[2] = entries;
        row[3] = open ? 1u : 0u;
        for (u32 k = 4; k < GL; k++) row[k] = NIL;
    }
}

// u in ring(v) must imply v in ring(u): the change-driven sweep stamps the ring of a vertex that moved and
// relies on that ring containing every vertex that reads it. counters[2] counts violations.
__global__ void k_ring_check(const u32 *__restrict__ ring8, const u32 *__restrict__ pool, u32 V, ull *counters)
{
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    auto row_len = [&](const u32 *row) { u32 n = 0; while (n < GL && row[n] != NIL) n++; return n; };
    auto get = [&](const u32 *row, bool ovf, u32 k) { return ovf ? pool[row[1] + k] : (k == 0 ? (row[0] & ~OPEN_BIT) : row[k]); };
    const u32 *rv = ring8 + (size_t)v * GL;
    const bool ov = rv[0] == OVF;
    const u32 lv = ov ? rv[2] : row_len(rv);
    u32 bad = 0;
    for (u32 k = 0; k < lv; k++) {
        const u32 n = get(rv, ov, k);
        const u32 *rn = ring8 + (size_t)n * GL;
        const bool on = rn[0] == OVF;
        const u32 ln = on ? rn[2] : row_len(rn);
        bool found = false;
        for (u32 q = 0; q < ln && !found; q++) found = get(rn, on, q) == v;
        if (!found) bad++;
    }
    if (bad) atomicAdd(counters + 2, (ull)bad);
}

// Geometry table (MeshView::geo): for ring slot k of vertex v, the inverse Gram matrix of triangle k = (v, n_k, n_{k+1})
// and |X_k|, computed with the very operations of update_step (src/geodesics_ptp.cpp:208-231, 257-258) so that a
// relaxation reading the table produces the same bits as one recomputing them. One thread per vertex; rows in the
// overflow pool (one-rings longer than 8) are not covered (their relaxation recomputes the geometry).
template <class R>
__global__ void k_geo_build(const typename Ops<R>::vec4 *__restrict__ GT4, const u32 *__restrict__ ring8, u32 V,
                            typename Ops<R>::vec4 *__restrict__ geo)
{
    typedef Ops<R> O;
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const u32 *row = ring8 + (size_t)v * GL;
    R *out = reinterpret_cast<R *>(geo + (size_t)v * GL);
    u32 e[GL];
    for (u32 k = 0; k < GL; k++) e[k] = row[k];
    u32 len = 0;
    bool open = false;
    if (e[0] != OVF && e[0] != NIL) {
        open = (e[0] & OPEN_BIT) != 0;
        e[0] &= ~OPEN_BIT;
        while (len < GL && e[len] != NIL) len++;
    }
    const u32 n_tri = len == 0 ? 0 : (open ? len - 1 : len);
    const P3<R> Ps = load_pos<R>(GT4 + v);
    P3<R> X[GL];
    R q[GL];
    for (u32 k = 0; k < len; k++) {
        const P3<R> Pn = load_pos<R>(GT4 + e[k]);
        X[k] = {O::sub(Pn.x, Ps.x), O::sub(Pn.y, Ps.y), O::sub(Pn.z, Ps.z)};
        q[k] = dot3(X[k], X[k]);
    }
    for (u32 k = 0; k < GL; k++) {
        R rec[4] = {R(0), R(0), R(0), R(0)};
        if (k < len) {
            rec[3] = O::sqrt(q[k]);
            if (k < n_tri) {
                const u32 k1 = k + 1 < len ? k + 1 : 0;
                const TriQ<R> Q = tri_geom<R>(X[k], X[k1], q[k], q[k1]);
                rec[0] = Q.Q00; rec[1] = Q.Q01; rec[2] = Q.Q11;
            }
        }
        for (u32 c = 0; c < 4; c++) out[k * 4 + c] = rec[c];
    }
}

// Causal-safe flags (MeshView::safe8): bit k of safe8[v] = the planar update on triangle k = (v, n_k, n_{k+1}) obeys the
// bound p >= min(t0, t1)(1 - 75 u) (causal_safe in ptp_device.cuh). One thread per vertex; overflow rows get 0.
// safe8[V + v]: the same per triangle for the short sign test of the acceptance condition (sign_short_ok; all 0 when the
// "sign_short" option is off).
template <class R>
__global__ void k_safe_build(const typename Ops<R>::vec4 *__restrict__ GT4, const u32 *__restrict__ ring8, u32 V, unsigned char *__restrict__ safe8,
                             u32 with_sign, u32 with_two)
{
    typedef Ops<R> O;
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const u32 *row = ring8 + (size_t)v * GL;
    u32 e[GL];
    for (u32 k = 0; k < GL; k++) e[k] = row[k];
    u32 len = 0, bits = 0, sbits = 0, tbits = 0;
    bool open = false;
    if (e[0] != OVF && e[0] != NIL) {
        open = (e[0] & OPEN_BIT) != 0;
        e[0] &= ~OPEN_BIT;
        while (len < GL && e[len] != NIL) len++;
    }
    const u32 n_tri = len == 0 ? 0 : (open ? len - 1 : len);
    const P3<R> Ps = load_pos<R>(GT4 + v);
    P3<R> X[GL];
    R q[GL];
    for (u32 k = 0; k < len; k++) {
        const P3<R> Pn = load_pos<R>(GT4 + e[k]);
        X[k] = {O::sub(Pn.x, Ps.x), O::sub(Pn.y, Ps.y), O::sub(Pn.z, Ps.z)};
        q[k] = dot3(X[k], X[k]);
    }
    for (u32 k = 0; k < n_tri; k++) {
        const u32 k1 = k + 1 < len ? k + 1 : 0;
        if (causal_safe<R>(X[k], X[k1], q[k], q[k1])) bits |= 1u << k;
        if (with_sign) {
            const R q01 = dot3(X[k], X[k1]);
            const R det = O::sub(O::mul(q[k], q[k1]), O::mul(q01, q01)); // as update_step computes it
            if (sign_short_ok<R>(q[k], q[k1], det)) sbits |= 1u << k;
        }
        if (with_two && two_sided_ok<R>(X[k], X[k1], q[k], q[k1])) tbits |= 1u << k; // (safe8[2V + v]: two-sided causal skip)
    }
    safe8[v] = (unsigned char)bits;
    safe8[(size_t)V + v] = (unsigned char)sbits;
    safe8[2 * (size_t)V + v] = (unsigned char)tbits;
}

// ------------------------------------------------------------------------------------------------
// CHE construction on the device (reference: che::update_evt_ot_et, src/che.cpp:1295-1362, serial, ~22 s at
// 10 M vertices). Directed edges (a -> b) go into an open-addressing hash table keyed by (a << 32 | b);
// OT[he] is the half-edge stored under (b, a). For edge-manifold input (every directed edge once) this is the
// reference's pairing exactly; duplicates / multiple border half-edges at a vertex raise the non-manifold flag.

__device__ __forceinline__ u32 edge_hash(ull k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (u32)k;
}

__global__ void k_che_insert(const u32 *__restrict__ VT, u32 H, u32 V, ull *keys, u32 *vals, u32 mask, u32 *evt1, ull *flags)
{
    const u32 he = blockIdx.x * blockDim.x + threadIdx.x;
    if (he >= H) return;
    const u32 a = VT[he], b = VT[he_next(he)];
    if (a >= V || b >= V) { atomicAdd(flags + 1, 1ull); return; }
    atomicMax(evt1 + a, he + 1); // EVT[v] = last half-edge leaving v (:1304-1308), stored +1 so 0 means none
    const ull key = ((ull)a << 32) | b;
    u32 slot = edge_hash(key) & mask;
    while (true) {
        const ull prev = atomicCAS(keys + slot, ~0ull, key);
        if (prev == ~0ull) { vals[slot] = he; return; }
        if (prev == key) { atomicAdd(flags, 1ull); return; } // the same directed edge twice: not edge-manifold
        slot = (slot + 1) & mask;
    }
}

__global__ void k_che_pair(const u32 *__restrict__ VT, u32 H, u32 V, const ull *__restrict__ keys, const u32 *__restrict__ vals,
                           u32 mask, u32 *OT, u32 *border_cnt, u32 *border_he)
{
    const u32 he = blockIdx.x * blockDim.x + threadIdx.x;
    if (he >= H) return;
    const u32 a = VT[he], b = VT[he_next(he)];
    if (a >= V || b >= V) { OT[he] = NIL; return; }
    const ull key = ((ull)b << 32) | a;
    u32 slot = edge_hash(key) & mask, opp = NIL;
    while (true) {
        const ull k = keys[slot];
        if (k == key) { opp = vals[slot]; break; }
        if (k == ~0ull) break;
        slot = (slot + 1) & mask;
    }
    OT[he] = opp;
    if (opp == NIL) { // border half-edge: becomes EVT of its origin (:1343-1352)
        atomicAdd(border_cnt + a, 1u);
        border_he[a] = he;
    }
}

__global__ void k_che_evt(u32 V, const u32 *__restrict__ evt1, const u32 *__restrict__ border_cnt, const u32 *__restrict__ border_he,
                          u32 *EVT, ull *flags)
{
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const u32 bc = border_cnt[v];
    if (bc > 1) atomicAdd(flags, 1ull); // two border half-edges at one vertex: non-manifold vertex
    EVT[v] = bc == 1 ? border_he[v] : (evt1[v] ? evt1[v] - 1 : NIL);
}

struct TeamFlat { // plain grid-stride launch, no synchronisation
    __device__ __forceinline__ u32 cta() const { return blockIdx.x; }
    __device__ __forceinline__ u32 nctas() const { return gridDim.x; }
    __device__ __forceinline__ u32 sync(u32 = 0) { return 0; }
    template <class T> static __device__ __forceinline__ T ld(const T *p) { return *p; }
};

template <class R>
__global__ void __launch_bounds__(GRID_BLOCK) k_bfs_grid(MeshView<R> m, Work<R> w, const u32 *sources, u32 S, u32 kcap, u32 sent, ull *bar)
{
    TeamGrid t{bar, 0, 0, gridDim.x};
    t.err = w.ctrl + C_ERROR;
    bfs_run<R, TeamGrid, false>(t, m, w, sources, S, kcap, sent);
}

// DEBUG (PTP_FUSED=2): the two halves of the fused kernel as two launches, to time the streamed sweep alone
template <class R>
__global__ void __launch_bounds__(FUSED_BLOCK, 2) k_dbg_producer(MeshView<R> m, Work<R> w, const u32 *sources, u32 S, u32 sent, ull *bar)
{
    TeamGrid t{bar, 0, 0, gridDim.x};
    t.err = w.ctrl + C_ERROR;
    bfs_run<R, TeamGrid, true>(t, m, w, sources, S, NIL, sent);
}
template <class R>
__global__ void __launch_bounds__(FUSED_BLOCK, 2)
k_dbg_consumer(MeshView<R> m, Work<R> w, const u32 *sources, u32 S, R *dist_out, u32 sent, ull *bar)
{
    TeamGrid t{bar, 0, 0, gridDim.x};
    t.err = w.ctrl + C_ERROR;
    const u32 d = ptp_run<R, TeamGrid, false, PTP_GRID_MAP, true>(t, m, w, sources, S, 0u, 0u, sent, w.tile_sum + 2048, m.ring_symmetric != 0);
    scatter_run<R, TeamGrid, false>(t, m, w, d, dist_out, nullptr, 0u);
}

// DEBUG / verification: Ops<R>::inv_gram (three divisions sharing one reciprocal) against three plain IEEE divisions, bit for
// bit, on (a) Gram matrices of random edge pairs at random scales, (b) raw random bit patterns and specials
// (ptp_debug_inv_gram_check). out[0] = mismatching results, out[1] = cases that took the shared-reciprocal path.
__device__ __forceinline__ ull dbg_mix(ull x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
template <class R> struct DbgBits;
template <> struct DbgBits<float> {
    static __device__ __forceinline__ float from(ull h) { return __uint_as_float((u32)h); }
    static __device__ __forceinline__ bool same(float a, float b) { return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b); }
    static __device__ __forceinline__ float scale(int e) { return __uint_as_float((u32)(127 + e) << 23); }
    static constexpr int ESPAN = 30;
};
template <> struct DbgBits<double> {
    static __device__ __forceinline__ double from(ull h) { return __longlong_as_double((long long)h); }
    static __device__ __forceinline__ bool same(double a, double b) { return __double_as_longlong(a) == __double_as_longlong(b) || (a != a && b != b); }
    static __device__ __forceinline__ double scale(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }
    static constexpr int ESPAN = 500;
};
template <class R>
__global__ void k_dbg_inv_gram(ull n, ull seed, ull *out, double *samples)
{
    typedef Ops<R> O;
    typedef DbgBits<R> B;
    ull bad = 0, fast = 0;
    for (ull i = blockIdx.x * (ull)blockDim.x + threadIdx.x; i < n; i += (ull)gridDim.x * blockDim.x) {
        ull h = dbg_mix(seed ^ (i * 0xD1342543DE82EF95ull));
        R q00, q01, q11, det;
        const u32 kind = (u32)(h & 7u);
        if (kind < 5) { // Gram matrix of two random edges, common scale 2^e, relative length and angle free
            const int e = (int)((h >> 8) % (2u * B::ESPAN + 1u)) - B::ESPAN;
            R x[6];
            for (int k = 0; k < 6; k++) {
                h = dbg_mix(h);
                x[k] = (R)((double)(long long)(h >> 11) * (1.0 / 4503599627370496.0) - 1.0) * B::scale(e);
            }
            if (kind == 3) { x[3] = x[0]; x[4] = x[1]; h = dbg_mix(h); x[5] = x[2] * (R)(1.0 + (double)(h & 0xffff) * 1e-7); } // nearly parallel
            if (kind == 4) { x[3] = -x[1]; x[4] = x[0]; x[2] = 0; x[5] = 0; }                                                // right angle: q01 == 0
            const P3<R> X0 = {x[0], x[1], x[2]}, X1 = {x[3], x[4], x[5]};
            q00 = dot3(X0, X0); q11 = dot3(X1, X1); q01 = dot3(X0, X1);
            det = O::sub(O::mul(q00, q11), O::mul(q01, q01));
        } else { // raw bit patterns: every class of special operand turns up
            q00 = B::from(dbg_mix(h + 1)); q01 = B::from(dbg_mix(h + 2)); q11 = B::from(dbg_mix(h + 3)); det = B::from(dbg_mix(h + 4));
            if (kind == 6) { q00 = O::abs(q00); q11 = O::abs(q11); det = O::abs(det); }
            if (kind == 7) { const u32 w = (u32)(h >> 40) & 7u; const R sp[8] = {R(0), -R(0), O::inf(), -O::inf(), O::sub(O::inf(), O::inf()), R(1), B::scale(-B::ESPAN * 2), B::scale(B::ESPAN * 2)};
                             if (h & 0x100) q00 = sp[w]; if (h & 0x200) q01 = sp[(w + 1) & 7]; if (h & 0x400) q11 = sp[(w + 2) & 7]; if (h & 0x800) det = sp[(w + 3) & 7]; }
        }
        R a, b, c;
        fast += O::inv_gram(q00, q01, q11, det, a, b, c) ? 1u : 0u;
        const R ra = O::div(q11, det), rb = O::div(-q01, det), rc = O::div(q00, det);
        const u32 nb = (B::same(a, ra) ? 0u : 1u) + (B::same(b, rb) ? 0u : 1u) + (B::same(c, rc) ? 0u : 1u);
        if (nb && samples) { // the first few offenders, for diagnosis: operands and both results
            const ull at = atomicAdd(out + 2, 1ull);
            if (at < 16) {
                double *sp = samples + at * 10;
                sp[0] = (double)q00; sp[1] = (double)q01; sp[2] = (double)q11; sp[3] = (double)det;
                sp[4] = (double)a; sp[5] = (double)ra; sp[6] = (double)b; sp[7] = (double)rb; sp[8] = (double)c; sp[9] = (double)rc;
            }
        }
        bad += nb;
    }
    if (bad) atomicAdd(out, bad);
    if (fast) atomicAdd(out + 1, fast);
}

// DEBUG / verification: the short sign test of update_step's acceptance condition (tri_front, sign_short) against the
// reference's 43-operation chain (ptp_debug_sign_short_check). Triangles of random shape and scale (incl. shapes at the
// edge of what sign_short_ok admits), distances around them random or ADVERSARIAL: tp chosen so that one of the two-term
// forms nearly cancels (|e| from 2^-30 to 2^-2 of its terms). out[0] = evaluations in which the short form decided and the
// full chain disagrees (expected 0), out[1] = evaluations decided by the short form, out[2] = evaluations of flagged triangles.
template <class R>
__global__ void k_dbg_sign_short(ull n, ull seed, ull *out)
{
    typedef Ops<R> O;
    typedef DbgBits<R> B;
    ull bad = 0, decided = 0, flagged = 0;
    for (ull i = blockIdx.x * (ull)blockDim.x + threadIdx.x; i < n; i += (ull)gridDim.x * blockDim.x) {
        ull h = dbg_mix(seed ^ (i * 0xD1342543DE82EF95ull));
        auto unif = [&]() -> double { h = dbg_mix(h); return (double)(long long)(h >> 11) * (1.0 / 9007199254740992.0); }; // [0, 1)
        const u32 kind = (u32)(h & 7u);
        const int e = (int)((h >> 8) % 27u) - 13; // edge scale 2^e
        // two edges: length ratio up to 3, angle from 12 to 168 degrees (the flag admits roughly 18..162), random frame
        const double len0 = 1.0 + unif(), len1 = len0 * (kind & 1 ? 1.0 + 2.0 * unif() : 1.0 / (1.0 + 2.0 * unif()));
        const double ang = (12.0 + 156.0 * unif()) * 0.017453292519943295;
        double f0[3] = {unif() - 0.5, unif() - 0.5, unif() - 0.5}, f1[3] = {unif() - 0.5, unif() - 0.5, unif() - 0.5};
        double n0 = sqrt(f0[0] * f0[0] + f0[1] * f0[1] + f0[2] * f0[2]) + 1e-300;
        for (int k = 0; k < 3; k++) f0[k] /= n0;
        double d = f0[0] * f1[0] + f0[1] * f1[1] + f0[2] * f1[2];
        for (int k = 0; k < 3; k++) f1[k] -= d * f0[k];
        double n1 = sqrt(f1[0] * f1[0] + f1[1] * f1[1] + f1[2] * f1[2]) + 1e-300;
        for (int k = 0; k < 3; k++) f1[k] /= n1;
        const double sc = (double)B::scale(e);
        P3<R> X0, X1;
        X0.x = (R)(len0 * f0[0] * sc); X0.y = (R)(len0 * f0[1] * sc); X0.z = (R)(len0 * f0[2] * sc);
        X1.x = (R)(len1 * (cos(ang) * f0[0] + sin(ang) * f1[0]) * sc);
        X1.y = (R)(len1 * (cos(ang) * f0[1] + sin(ang) * f1[1]) * sc);
        X1.z = (R)(len1 * (cos(ang) * f0[2] + sin(ang) * f1[2]) * sc);
        const R q00 = dot3(X0, X0), q11 = dot3(X1, X1), q01 = dot3(X0, X1);
        const R det = O::sub(O::mul(q00, q11), O::mul(q01, q01));
        if (!sign_short_ok<R>(q00, q11, det)) continue;
        flagged++;
        const TriQ<R> Q = tri_geom<R>(X0, X1, q00, q11);
        // distances: base value of random magnitude, difference of the order of the edges
        const double base = sc * (kind < 6 ? 16.0 + 500.0 * unif() : 16.0 + 1e-3 * unif());
        R t0 = (R)(base), t1 = (R)(base + sc * len0 * (2.0 * unif() - 1.0));
        if (kind >= 2 && kind < 6) {
            // adversarial: e0 = 0 is the front running along edge X1 (p = t1 + |X1|, i.e. tp1 = -|X1|, tp0 = Q01 |X1| / Q00),
            // e1 = 0 the same along X0; the planar update only depends on t1 - t0, so the zero sits at
            // t1 - t0 = -|X1| (1 + Q01 / Q00)  resp.  |X0| (1 + Q01 / Q11). Move off it by a relative 2^-2 ... 2^-30, or not at all.
            const bool first = (kind & 1) != 0;
            const u32 k = (u32)((h >> 40) % 30u);
            const double off = k == 29 ? 0.0 : ldexp(1.0, -2 - (int)k) * ((h >> 50) & 1 ? 1.0 : -1.0);
            const double x0 = sqrt((double)q00), x1 = sqrt((double)q11);
            const double dz = first ? -x1 * (1.0 + (double)Q.Q01 / (double)Q.Q00) : x0 * (1.0 + (double)Q.Q01 / (double)Q.Q11);
            t1 = (R)((double)t0 + dz * (1.0 + off));
        }
        if (!(t0 >= R(0)) || !(t1 >= R(0)) || !(t0 < O::inf()) || !(t1 < O::inf())) continue;
        bool fb_full = false, fb_short = false;
        const R pf = tri_front<R>(X0, X1, Q, t0, t1, fb_full, false);
        const R ps = tri_front<R>(X0, X1, Q, t0, t1, fb_short, true);
        // did the short form decide? re-evaluate its predicate (the same expressions)
        bool dec = false;
        {
            const R delta = O::add(O::mul(t0, O::add(Q.Q00, Q.Q01)), O::mul(t1, O::add(Q.Q01, Q.Q11)));
            const R sumQ = O::add(O::add(O::add(Q.Q00, Q.Q01), Q.Q01), Q.Q11);
            const R inner = O::sub(O::add(O::add(O::mul(O::mul(t0, t0), Q.Q00), O::mul(O::mul(t0, t1), O::add(Q.Q01, Q.Q01))), O::mul(O::mul(t1, t1), Q.Q11)), R(1));
            const R dis = O::sub(O::mul(delta, delta), O::mul(sumQ, inner));
            if (!(dis < R(0))) {
                const R p = O::div(O::add(delta, O::sqrt(dis)), sumQ);
                const R tp0 = O::sub(t0, p), tp1 = O::sub(t1, p);
                const R T = O::add(O::abs(tp0), O::abs(tp1));
                const R M = O::mul(O::mul(Q.Q00 > Q.Q11 ? Q.Q00 : Q.Q11, T), SignShort<R>::margin());
                const R e0 = O::add(O::mul(Q.Q00, tp0), O::mul(Q.Q01, tp1)), e1 = O::add(O::mul(Q.Q01, tp0), O::mul(Q.Q11, tp1));
                dec = T >= SignShort<R>::t_min() && T <= SignShort<R>::t_max() && O::abs(e0) > M && O::abs(e1) > M;
            }
        }
        decided += dec ? 1u : 0u;
        if (fb_full != fb_short || !B::same(pf, ps)) bad++;
    }
    if (bad) atomicAdd(out, bad);
    if (decided) atomicAdd(out + 1, decided);
    if (flagged) atomicAdd(out + 2, flagged);
}

// DEBUG / verification: Ops<float>::sqrt_n against __fsqrt_rn over EVERY float in [2^-96, 2^96] (ptp_debug_sqrt_check)
__global__ void k_dbg_sqrt(ull *out)
{
    ull bad = 0, n = 0;
    const u32 lo = (127u - 96u) << 23, hi = (127u + 96u) << 23; // bit patterns of 2^-96 and 2^96
    for (ull b = (ull)lo + blockIdx.x * (ull)blockDim.x + threadIdx.x; b <= (ull)hi; b += (ull)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((u32)b);
        bad += __float_as_uint(Ops<float>::sqrt_n(x)) != __float_as_uint(__fsqrt_rn(x)) ? 1u : 0u;
        n++;
    }
    if (bad) atomicAdd(out, bad);
    atomicAdd(out + 1, n);
}

// DEBUG / verification: the two-sided causal skip (two_sided_ok + two_sided_skip) against update_step
// (ptp_debug_two_sided_check). Triangles of random shape and scale, and around each (cur, lo, hi) random or ADVERSARIAL:
// hi - cur at the edge of what (G) admits, thr - lo within rounding of the edge length (the boundary of (E)), hi = thr
// exactly, cur = 0, lo = 0. Whenever the rule fires the reference chain is evaluated (sign_short off):
// out[0] = cases with p < cur (expected 0), out[1] = cases in which the rule fired, out[2] = cases on admitted triangles.
template <class R>
__global__ void k_dbg_two_sided(ull n, ull seed, ull *out)
{
    typedef Ops<R> O;
    typedef DbgBits<R> B;
    ull bad = 0, fired = 0, flagged = 0;
    for (ull i = blockIdx.x * (ull)blockDim.x + threadIdx.x; i < n; i += (ull)gridDim.x * blockDim.x) {
        ull h = dbg_mix(seed ^ (i * 0xD1342543DE82EF95ull));
        auto unif = [&]() -> double { h = dbg_mix(h); return (double)(long long)(h >> 11) * (1.0 / 9007199254740992.0); };
        const u32 kind = (u32)(h & 15u);
        const int e = (int)((h >> 8) % 25u) - 12;
        const double len0 = 1.0 + unif(), len1 = len0 * (kind & 1 ? 1.0 + 1.9 * unif() : 1.0 / (1.0 + 1.9 * unif()));
        const double ang = (15.0 + 78.0 * unif()) * 0.017453292519943295; // up to 93 degrees: the flag must reject the obtuse ones
        double f0[3] = {unif() - 0.5, unif() - 0.5, unif() - 0.5}, f1[3] = {unif() - 0.5, unif() - 0.5, unif() - 0.5};
        const double n0 = sqrt(f0[0] * f0[0] + f0[1] * f0[1] + f0[2] * f0[2]) + 1e-300;
        for (int k = 0; k < 3; k++) f0[k] /= n0;
        const double dt = f0[0] * f1[0] + f0[1] * f1[1] + f0[2] * f1[2];
        for (int k = 0; k < 3; k++) f1[k] -= dt * f0[k];
        const double n1 = sqrt(f1[0] * f1[0] + f1[1] * f1[1] + f1[2] * f1[2]) + 1e-300;
        for (int k = 0; k < 3; k++) f1[k] /= n1;
        const double sc = (double)B::scale(e);
        P3<R> X0, X1;
        X0.x = (R)(len0 * f0[0] * sc); X0.y = (R)(len0 * f0[1] * sc); X0.z = (R)(len0 * f0[2] * sc);
        X1.x = (R)(len1 * (cos(ang) * f0[0] + sin(ang) * f1[0]) * sc);
        X1.y = (R)(len1 * (cos(ang) * f0[1] + sin(ang) * f1[1]) * sc);
        X1.z = (R)(len1 * (cos(ang) * f0[2] + sin(ang) * f1[2]) * sc);
        const R q00 = dot3(X0, X0), q11 = dot3(X1, X1);
        if (!two_sided_ok<R>(X0, X1, q00, q11)) continue;
        flagged++;
        const bool lo_is_0 = ((h >> 20) & 1) != 0;                      // which corner is the upstream one
        const double x_lo = sqrt((double)(lo_is_0 ? q00 : q11));
        // lo: far from / near / at the sources; cur above it by a fraction of the edge; hi above cur
        double lo = sc * (kind < 12 ? 4.0 + 800.0 * unif() : (kind < 14 ? 2.0 * unif() : 0.0));
        double cur = lo + x_lo * 1.3 * unif();
        if (kind == 15) { cur = 0.0; lo = 0.0; }
        const double gap = cur - lo;
        double hi = cur + sc * len0 * 1.5 * unif();
        const u32 adv = (u32)((h >> 24) & 7u);
        if (adv == 1) hi = cur + gap * ldexp(1.0, -9) * (1.0 + ldexp(unif() - 0.5, -18));      // (G) at its edge
        if (adv == 2) hi = cur * (1.0 + ldexp(1.0, -14)) * (1.0 + ldexp(unif() - 0.5, -20));   // hi ~ thr
        if (adv == 3) cur = (lo + x_lo) * (1.0 - ldexp(1.0, -14)) * (1.0 + ldexp(unif() - 0.5, -16)), hi = cur + sc * len0 * unif(); // (E) at its edge
        if (adv == 4) hi = cur + ldexp(1.0, -39) * (1.0 + unif());                             // g ~ g_min
        const R rcur = (R)cur, rlo = (R)lo, rhi = (R)hi;
        const R thr = O::mul(rcur, Causal<R>::up());
        const R t0 = lo_is_0 ? rlo : rhi, t1 = lo_is_0 ? rhi : rlo;
        // (the arguments exactly as the ring walk forms them: corner 0 = current neighbour, corner 1 = next)
        if (!two_sided_skip<R>(rcur, thr, t1 < t0 ? t1 : t0, t1 < t0 ? t0 : t1, t1 < t0 ? q11 : q00)) continue;
        fired++;
        const R p = update_tri<R>(X0, X1, q00, q11, t0, t1, false);
        if (p < rcur) bad++;
    }
    if (bad) atomicAdd(out, bad);
    if (fired) atomicAdd(out + 1, fired);
    if (flagged) atomicAdd(out + 2, flagged);
}

// DEBUG / measurement: n grid barriers and nothing else (ptp_debug_barrier_ns)
// mode 0: barriers only. mode 1: every iteration each CTA writes a word, barrier, every thread reads the word its
// neighbour CTA wrote (plain load: the acquire barrier flushed L1) and the value feeds the next iteration — the
// barrier + one dependent L2 round trip of a PTP iteration. mode 2: same with a poll that does not acquire (no
// CCTL.IVALL) and an ld.cg read.
__global__ void k_dbg_barriers(ull *bar, u32 n, u32 *sink, u32 mode)
{
    TeamGrid t{bar, 0, 0, gridDim.x};
    t.relaxed_poll = mode == 2u ? 1u : 0u;
    u32 *box = reinterpret_cast<u32 *>(bar + 16); // [gridDim.x] words after the barrier words (bar is 4 KB)
    u32 acc = 0;
    for (u32 i = 0; i < n; i++) {
        if (mode) {
            if (threadIdx.x == 0) box[blockIdx.x] = i + acc;
        }
        acc += t.sync(i & 1u);
        if (mode) {
            const u32 *p = box + (blockIdx.x + 1u) % gridDim.x;
            acc += (mode == 2u ? __ldcg(p) : *(const volatile u32 *)p) & 1u;
        }
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}

// DEBUG / measurement: n hardware cluster barriers and nothing else (PTP_CLUSTER_BARRIER in ptp_debug_barrier_ns)
__global__ void k_dbg_cluster_barriers(u32 n, u32 *sink)
{
    cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
    u32 acc = 0;
    for (u32 i = 0; i < n; i++) { cl.sync(); acc += i; }
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}

template <class R> __global__ void k_inv_init(MeshView<R> m, Work<R> w)
{
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < m.V) w.inv[v] = NIL;
}
template <class R> __global__ void k_inv_fill(MeshView<R> m, Work<R> w, u32 p)
{
    const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < p) {
        const u32 v = w.sorted[r];
        if (v < m.V) atomicMin(&w.inv[v], r);
        else atomicAdd(w.ctrl + C_ABORT, 1ull); // out-of-range entry: reported by the host before anything uses the table
    }
}

// exact distances into rank order (measurement mode of the sweep: Work::exactS)
template <class R> __global__ void k_gather_exact(Work<R> w, const R *__restrict__ exact, R *__restrict__ exactS)
{
    const u32 p = (u32)w.ctrl[C_REACHED];
    for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < p; r += gridDim.x * blockDim.x) exactS[r] = exact[w.sorted[r]];
}

template <class R> __global__ void __launch_bounds__(FLAT_BLOCK) k_layout(MeshView<R> m, Work<R> w, u32 sent)
{
    TeamFlat t;
    layout_run<R, TeamFlat>(t, m, w, (u32)w.ctrl[C_REACHED], sent);
}

template <class R, bool CL>
__global__ void __launch_bounds__(SOLVE_BLOCK)
k_solve_grid(MeshView<R> m, Work<R> w, const u32 *sources, u32 S, R *dist_out, u32 *cl_out, u32 cl_fill, u32 sent, ull *bar, u32 staged)
{
    extern __shared__ __align__(16) unsigned char ptp_dyn_smem[];
    TeamGrid t{bar, 0, 0, gridDim.x};
    t.err = w.ctrl + C_ERROR;
    const u32 nl = (u32)w.ctrl[C_NLIMITS], p = (u32)w.ctrl[C_REACHED];
    const u32 d = ptp_run<R, TeamGrid, CL, PTP_GRID_MAP, false>(t, m, w, sources, S, nl, p, sent, w.tile_sum + 2048, m.ring_symmetric != 0,
                                                               staged ? ptp_dyn_smem : nullptr);
    scatter_run<R, TeamGrid, CL>(t, m, w, d, dist_out, cl_out, cl_fill);
}

// Single solve, one cooperative launch, two teams: CTAs [0, nb) build the toplesets and lay out the rows
// (producer), CTAs [nb, gridDim) sweep behind them (consumer). The BFS (~#levels dependent steps) and the sweep
// (~#levels dependent iterations) overlap instead of adding up.
template <class R, bool CL>
__global__ void __launch_bounds__(FUSED_BLOCK, 2)
k_geodesics_fused(MeshView<R> m, Work<R> w, const u32 *sources, u32 S, R *dist_out, u32 *cl_out, u32 cl_fill, u32 sent, ull *bar, u32 nb,
                  u32 staged)
{
    extern __shared__ __align__(16) unsigned char ptp_dyn_smem[];
    if (blockIdx.x < nb) {
        TeamGrid t{bar, 0, 0, nb};
        t.err = w.ctrl + C_ERROR;
        if (blockIdx.x == 0 && threadIdx.x == 0) w.ctrl[C_TSTART] = global_timer();
        bfs_run<R, TeamGrid, true>(t, m, w, sources, S, NIL, sent);
        if (blockIdx.x == 0 && threadIdx.x == 0) w.ctrl[C_TBFS] = global_timer();
    } else {
        TeamGrid t{bar + 64, 0, nb, gridDim.x - nb};
        const u32 d = ptp_run<R, TeamGrid, CL, PTP_GRID_MAP, true>(t, m, w, sources, S, 0u, 0u, sent, w.tile_sum + 2048, m.ring_symmetric != 0,
                                                                  staged ? ptp_dyn_smem : nullptr);
        scatter_run<R, TeamGrid, CL>(t, m, w, d, dist_out, cl_out, cl_fill);
        if (blockIdx.x == nb && threadIdx.x == 0) w.ctrl[C_TEND] = global_timer();
    }
}

// Single solve, one launch of thread-block clusters: cluster 0 (CTAs [0, nb)) builds the toplesets with hardware cluster
// barriers (bfs_run_cluster), every other CTA belongs to the sweep team (one CTA per SM, window staged in shared
// memory) that lays out and relaxes the levels as they appear. The sweep team also presets the BFS tables (all-ones:
// key = ~0, inv = NIL) while the BFS cluster waits for C_FILLED, so the whole solve stays one launch.
template <class R, bool CL, bool GEO>
__global__ void __launch_bounds__(CLUSTER_BLOCK, 1)
k_geodesics_cluster(MeshView<R> m, Work<R> w, const u32 *sources, u32 S, R *dist_out, u32 *cl_out, u32 cl_fill, u32 sent, ull *bar, u32 nb,
                    u32 staged)
{
    extern __shared__ __align__(16) unsigned char ptp_dyn_smem[];
    if (blockIdx.x < nb) {
        if (blockIdx.x == 0 && threadIdx.x == 0) w.ctrl[C_TSTART] = global_timer();
        bfs_run_cluster<R, true>(m, w, sources, S);
        if (blockIdx.x == 0 && threadIdx.x == 0) w.ctrl[C_TBFS] = global_timer();
    } else {
        TeamGrid t{bar + 64, 0, nb, gridDim.x - nb};
        t.err = w.ctrl + C_ERROR;
        const u32 tid = t.cta() * blockDim.x + threadIdx.x, nth = t.nctas() * blockDim.x;
        uint4 *key4 = reinterpret_cast<uint4 *>(w.key); // V * 8 bytes, 16-byte aligned (cudaMalloc)
        const uint4 ones = make_uint4(NIL, NIL, NIL, NIL);
        for (u32 i = tid; i < m.V / 2; i += nth) key4[i] = ones;
        if (tid == 0 && (m.V & 1u)) w.key[m.V - 1] = ~0ull;
        for (u32 v = tid; v < m.V; v += nth) w.inv[v] = NIL;
        t.sync();
        if (tid == 0) flag_store(w.ctrl + C_FILLED, 1ull);
        const u32 d = ptp_run<R, TeamGrid, CL, PTP_GRID_MAP, true, GEO>(t, m, w, sources, S, 0u, 0u, sent, w.tile_sum + 2048,
                                                                               m.ring_symmetric != 0, staged ? ptp_dyn_smem : nullptr);
        scatter_run<R, TeamGrid, CL>(t, m, w, d, dist_out, cl_out, cl_fill);
        if (blockIdx.x == nb && threadIdx.x == 0) w.ctrl[C_TEND] = global_timer();
    }
}

// The same two teams as TWO launches that run side by side (PTP_FUSED=5): the BFS needs one cluster, the sweep team
// needs no clusters at all, and a single cluster launch leaves the SMs that do not fill a GPC's last cluster idle
// (15 clusters of 8 = 120 of 148 SMs). 8 + 140 CTAs, one per SM; C3: 28.3 -> 24.0 ms.
template <class R>
__global__ void __launch_bounds__(CLUSTER_BLOCK, 1) k_toplesets_cluster(MeshView<R> m, Work<R> w, const u32 *sources, u32 S)
{
    // programmatic dependent launch: the sweep kernel (next in the stream, launched with programmatic stream
    // serialisation) may start as soon as every CTA of this grid is resident and has got here — which is exactly the
    // guarantee needed: the cluster has its eight SMs of one GPC before the sweep CTAs take whatever is left
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (blockIdx.x == 0 && threadIdx.x == 0) w.ctrl[C_TSTART] = global_timer();
    bfs_run_cluster<R, true>(m, w, sources, S);
    if (blockIdx.x == 0 && threadIdx.x == 0) w.ctrl[C_TBFS] = global_timer();
}

template <class R, bool CL, bool GEO>
__global__ void __launch_bounds__(CLUSTER_BLOCK, 1)
k_sweep_streamed(MeshView<R> m, Work<R> w, const u32 *sources, u32 S, R *dist_out, u32 *cl_out, u32 cl_fill, u32 sent, ull *bar, u32 staged)
{
    extern __shared__ __align__(16) unsigned char ptp_dyn_smem[];
    TeamGrid t{bar + 64, 0, 0, gridDim.x};
    t.err = w.ctrl + C_ERROR;
    const u32 tid = t.cta() * blockDim.x + threadIdx.x, nth = t.nctas() * blockDim.x;
    uint4 *key4 = reinterpret_cast<uint4 *>(w.key);
    const uint4 ones = make_uint4(NIL, NIL, NIL, NIL);
    for (u32 i = tid; i < m.V / 2; i += nth) key4[i] = ones;
    if (tid == 0 && (m.V & 1u)) w.key[m.V - 1] = ~0ull;
    for (u32 v = tid; v < m.V; v += nth) w.inv[v] = NIL;
    t.sync();
    if (tid == 0) flag_store(w.ctrl + C_FILLED, 1ull);
    const u32 d = ptp_run<R, TeamGrid, CL, PTP_GRID_MAP, true, GEO>(t, m, w, sources, S, 0u, 0u, sent, w.tile_sum + 2048, m.ring_symmetric != 0,
                                                                   staged ? ptp_dyn_smem : nullptr);
    scatter_run<R, TeamGrid, CL>(t, m, w, d, dist_out, cl_out, cl_fill);
    if (blockIdx.x == 0 && threadIdx.x == 0) w.ctrl[C_TEND] = global_timer();
}

// one CTA per solve, CTAs pull source sets from a queue
template <class R, bool GEO, bool CAUSAL>
__global__ void __launch_bounds__(BatchCfg<R>::BLOCK, BatchCfg<R>::MINB)
k_batched(MeshView<R> m, const Work<R> *works, const u32 *sources, const ull *offsets, u32 first, u32 B, R *rows, u32 sent,
          ull *queue, ull *totals, HelpDesc *descs, u32 *counters /* [0] idle CTAs, [1] solves done */, u32 n_slots,
          unsigned char *row_done /* optional: [B] set once row b is complete (host copies it out meanwhile) */)
{
    __shared__ u32 s_b;
    __shared__ u32 s_wl[2];
    TeamCta t;
    if (blockIdx.x >= n_slots) { // a CTA without a workspace of its own (more CTAs than slots): helper from the start
        if (descs) help_loop<R, GEO, CAUSAL>(m.geo, works, descs, n_slots, counters, counters + 1, B);
        return;
    }
    const Work<R> w = works[blockIdx.x];
    HelpDesc *help = descs ? descs + blockIdx.x : nullptr;
    while (true) {
        if (threadIdx.x == 0) s_b = (u32)atomicAdd(queue, 1ull);
        __syncthreads();
        const u32 b = s_b;
        __syncthreads();
        if (b >= B) {
            // no solve left for this CTA: lend its threads to the relax passes of the solves still running
            if (descs) help_loop<R, GEO, CAUSAL>(m.geo, works, descs, n_slots, counters, counters + 1, B);
            break;
        }
        const ull o0 = offsets ? offsets[first + b] : (ull)(first + b);
        const u32 S = offsets ? (u32)(offsets[first + b + 1] - o0) : 1u;
        const u32 *src = sources + o0;
        if (threadIdx.x == 0) { w.ctrl[C_OVFALLOC] = 0; w.ctrl[C_RELAXED] = 0; }
        const ull t0 = global_timer();
        bfs_run_cta<R>(m, w, src, S);
        const ull t1 = global_timer();
        const u32 nl = (u32)w.ctrl[C_NLIMITS], p = (u32)w.ctrl[C_REACHED];
        layout_rows_thread<R, CAUSAL>(m, w, 0u, p, threadIdx.x, blockDim.x, sent, [](const u32 *q) { return *q; });
        __syncthreads();
        const ull t2 = global_timer();
        const u32 d = ptp_run<R, TeamCta, false, 1, false, GEO, CAUSAL>(t, m, w, src, S, nl, p, sent, s_wl, m.ring_symmetric != 0, nullptr, help, counters);
        scatter_run<R, TeamCta, false>(t, m, w, d, rows + (size_t)b * m.V, nullptr, 0u);
        __syncthreads();
        if (threadIdx.x == 0) {
            const ull t3 = global_timer();
            atomicAdd(totals + 6, t1 - t0); // per-phase device time summed over solves (ns)
            atomicAdd(totals + 7, t2 - t1);
            atomicAdd(totals + 8, t3 - t2);
            if (descs) { __threadfence(); atomicAdd(counters + 1, 1u); }
            if (row_done) { __threadfence(); *(volatile unsigned char *)(row_done + b) = 1; } // after the barrier behind the scatter
            atomicAdd(totals + 0, w.ctrl[C_ITER]);
            atomicAdd(totals + 1, w.ctrl[C_UPDATES]);
            atomicMax(totals + 2, w.ctrl[C_MAXWIN]);
            atomicAdd(totals + 3, (ull)(nl ? nl - 1 : 0));
            atomicAdd(totals + 4, (ull)p);
            atomicAdd(totals + 5, w.ctrl[C_RELAXED]);
            if (w.ctrl[C_ERROR]) atomicMax(totals + 9, w.ctrl[C_ERROR]); // device watchdog (elastic chunks): the host must know
        }
    }
}

// Batched solves, a TEAM of CTAs per solve. With one CTA per solve 148 solves are in flight and their windows
// (rows + two distance buffers of ~10^5 vertices each) add up to ~1.5 GB: every Jacobi iteration re-streams its window
// from HBM and every dependent gather pays DRAM latency. With teams of `team_size` CTAs only gridDim / team_size solves
// are in flight, each finishing team_size times sooner, so the windows of all of them together stay resident in the
// 126 MB L2 for the 30-60 iterations a vertex spends in a window. The price is a grid barrier among the team's CTAs
// per iteration (1 us against ~20 us of relaxation work per iteration). Every phase is the whole-GPU code of the
// single solve (bfs_run, layout_run, ptp_run, scatter_run on a TeamGrid); teams take solves from one queue, the index
// travels in the barrier word so that every CTA of the team sees the same one.
template <class R, bool GEO, bool CAUSAL>
__global__ void __launch_bounds__(BatchCfg<R>::BLOCK)
k_batched_teams(MeshView<R> m, const Work<R> *works, const u32 *sources, const ull *offsets, u32 first, u32 B, R *rows, u32 sent,
                ull *queue, ull *totals, ull *bars, u32 team_size, unsigned char *row_done)
{
    const u32 team_id = blockIdx.x / team_size;
    TeamGrid t{bars + (size_t)team_id * 16, 0, team_id * team_size, team_size};
    const Work<R> w = works[team_id];
    t.err = w.ctrl + C_ERROR;
    const bool lead = blockIdx.x == t.cta0 && threadIdx.x == 0;
    while (!t.dead) {
        const ull word = t.sync_full(0u, lead ? atomicAdd(queue, 1ull) : 0ull);
        const u32 b = (u32)(word >> 24);
        if (b >= B) break;
        const ull o0 = offsets ? offsets[first + b] : (ull)(first + b);
        const u32 S = offsets ? (u32)(offsets[first + b + 1] - o0) : 1u;
        const u32 *src = sources + o0;
        if (lead) { w.ctrl[C_OVFALLOC] = 0; w.ctrl[C_RELAXED] = 0; }
        const ull t0 = global_timer();
        bfs_run<R, TeamGrid, false>(t, m, w, src, S, NIL, sent);
        const u32 nl = (u32)TeamGrid::ld_sync(w.ctrl + C_NLIMITS), p = (u32)TeamGrid::ld_sync(w.ctrl + C_REACHED);
        const ull t1 = global_timer();
        layout_rows_thread<R, CAUSAL>(m, w, 0u, p, t.cta() * blockDim.x + threadIdx.x, t.nctas() * blockDim.x, sent,
                                      [](const u32 *q) { return TeamGrid::ld(q); });
        t.sync();
        const ull t2 = global_timer();
        const u32 d = ptp_run<R, TeamGrid, false, 1, false, GEO, CAUSAL>(t, m, w, src, S, nl, p, sent, w.tile_sum + 2048, m.ring_symmetric != 0);
        scatter_run<R, TeamGrid, false>(t, m, w, d, rows + (size_t)b * m.V, nullptr, 0u);
        t.sync(); // every CTA's share of C_RELAXED is in; the workspace may be reused
        if (lead) {
            const ull t3 = global_timer();
            if (row_done) { __threadfence(); *(volatile unsigned char *)(row_done + b) = 1; } // behind the team barrier after the scatter
            atomicAdd(totals + 6, t1 - t0); // per-phase device time summed over solves (ns)
            atomicAdd(totals + 7, t2 - t1);
            atomicAdd(totals + 8, t3 - t2);
            atomicAdd(totals + 0, w.ctrl[C_ITER]);
            atomicAdd(totals + 1, w.ctrl[C_UPDATES]);
            atomicMax(totals + 2, w.ctrl[C_MAXWIN]);
            atomicAdd(totals + 3, (ull)(nl ? nl - 1 : 0));
            atomicAdd(totals + 4, (ull)p);
            atomicAdd(totals + 5, *(volatile ull *)(w.ctrl + C_RELAXED));
        }
    }
    if (lead && *(volatile ull *)(w.ctrl + C_ERROR)) atomicMax(totals + 9, *(volatile ull *)(w.ctrl + C_ERROR));
}

// arg-max of |x| with the smallest index on ties (cublasI?amax semantics, src/cuda/geodesics_ptp.cu:139-141);
// appends the winner to the sample list. Single CTA: the array is read once, bandwidth-trivial next to a solve.
template <class R> __global__ void __launch_bounds__(1024)
k_argmax_append(const R *__restrict__ x, u32 V, u32 *samples, u32 n, R *maxval, const ull *ctrl, ull *sticky)
{
    // the solve that produced x ran just before in this stream: carry its watchdog word over (ctrl is cleared per solve)
    if (threadIdx.x == 0 && ctrl[C_ERROR]) *sticky = ctrl[C_ERROR];
    __shared__ R s_v[32];
    __shared__ u32 s_i[32];
    R bv = R(-1);
    u32 bi = NIL;
    for (u32 i = threadIdx.x; i < V; i += blockDim.x) {
        const R a = Ops<R>::abs(x[i]);
        if (a > bv) { bv = a; bi = i; }   // strided scan keeps the smallest index per thread
    }
    for (u32 o = 16; o; o >>= 1) {
        const R ov = __shfl_xor_sync(0xFFFFFFFFu, bv, o);
        const u32 oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const u32 nw = blockDim.x >> 5;
        bv = threadIdx.x < nw ? s_v[threadIdx.x] : R(-1);
        bi = threadIdx.x < nw ? s_i[threadIdx.x] : NIL;
        for (u32 o = 16; o; o >>= 1) {
            const R ov = __shfl_xor_sync(0xFFFFFFFFu, bv, o);
            const u32 oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (threadIdx.x == 0) { samples[n] = bi; *maxval = x[bi]; }
    }
}

// ------------------------------------------------------------------------------------------------
// host structures

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

} // namespace

struct ptp_mesh {
    std::mutex mu; // one workspace per mesh: calls on the same handle are serialised (different handles run concurrently)
    int device = 0;
    int real_size = 0;
    u64 V = 0, H = 0;
    int num_sms = 0;
    void *GT4 = nullptr;
    bool ring_symmetric = true;
    u32 *ring8 = nullptr;
    u32 *ovf = nullptr;
    u64 ovf_total = 0;
    unsigned char *safe8 = nullptr; // [3V] causal-safe / short-sign-test / two-sided-skip triangle flags, built at the first batched call (k_safe_build)
    int safe_sign = -1;             // options safe8[V..3V) was built with: bit 0 "sign_short", bit 1 "two_sided"
    void *geo = nullptr; // geometry table, built at the first batched call (k_geo_build)
    bool geo_failed = false; // the table did not fit: do not try again
    bool two_failed = false; // the two-launch single solve did not get both kernels resident once: use one launch from now on
    bool last_two = false;   // the last single solve ran as two launches
    const char *last_kernel = ""; // dominant kernel of the last call on this mesh (ptp_mesh_last_kernel)
    u64 bytes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};

    // single-solve workspace (lazy)
    u64 ws_scap = 0; // source capacity the workspace was sized for (0 = not allocated)
    bool ws_cl = false, ws_top = false;
    std::vector<std::pair<void *, size_t>> ws_allocs; // (pointer, bytes): `bytes` below is what is currently held
    void *w_key = nullptr, *w_sorted = nullptr, *w_inv = nullptr, *w_limits = nullptr, *w_tile = nullptr, *w_posS = nullptr,
         *w_ringS = nullptr, *w_ovfS = nullptr, *w_dist[2] = {nullptr, nullptr}, *w_cl[2] = {nullptr, nullptr},
         *w_top = nullptr, *w_ctrl = nullptr, *w_bar = nullptr, *w_src = nullptr, *w_out = nullptr, *w_clout = nullptr,
         *w_maxval = nullptr, *w_wl = nullptr, *w_dirty[2] = {nullptr, nullptr};
    void *h_ctrl = nullptr; // pinned

    // batched workspace (lazy)
    u32 bt_slots = 0;  // per-solve workspaces allocated
    u32 bt_team = 0;   // CTAs per solve they were laid out for (1 = one CTA per solve)
    u32 bt_grid = 0;   // CTAs of a batched launch
    u64 bt_scap = 0;
    std::vector<std::pair<void *, size_t>> bt_allocs;
    void *bt_works = nullptr, *bt_queue = nullptr, *bt_src = nullptr, *bt_off = nullptr, *bt_rows = nullptr, *bt_help = nullptr,
         *bt_bars = nullptr, *bt_ctrl = nullptr;
    u64 bt_src_cap = 0, bt_off_cap = 0, bt_rows_cap = 0;

    // multi-device batched solves: this device's shard of the rows before it travels to the root device, comm stream
    void *bt_done = nullptr, *bt_hdone = nullptr; // per-row completion flags (device) and their pinned host copy
    u64 bt_done_cap = 0;
    void *mg_rows = nullptr;
    u64 mg_rows_cap = 0;
    cudaStream_t mg_stream = nullptr;
    cudaEvent_t mg_ev = nullptr;
};

namespace {

int dev_alloc(ptp_mesh *m, void **p, size_t bytes, std::vector<std::pair<void *, size_t>> *track)
{
    *p = nullptr;
    if (bytes == 0) bytes = 16;
    CK(cudaMalloc(p, bytes));
    m->bytes += bytes;
    if (track) track->push_back({*p, bytes});
    return PTP_OK;
}

void free_list(ptp_mesh *m, std::vector<std::pair<void *, size_t>> &l)
{
    for (auto &a : l) {
        cudaFree(a.first);
        m->bytes -= std::min<u64>(m->bytes, a.second);
    }
    l.clear();
}

template <class R> MeshView<R> mesh_view(const ptp_mesh *m)
{
    MeshView<R> v;
    v.V = (u32)m->V;
    v.ring_symmetric = m->ring_symmetric ? 1u : 0u;
    v.newest = opt("newest") ? 1u : 0u;
    v.GT4 = (const typename Ops<R>::vec4 *)m->GT4;
    v.ring8 = m->ring8;
    v.ovf = m->ovf;
    v.geo = (const typename Ops<R>::vec4 *)m->geo;
    v.safe8 = m->safe8;
    return v;
}

template <class R> Work<R> work_view(const ptp_mesh *m)
{
    Work<R> w;
    w.key = (ull *)m->w_key;
    w.sorted = (u32 *)m->w_sorted;
    w.inv = (u32 *)m->w_inv;
    w.limits = (u32 *)m->w_limits;
    w.tile_sum = (u32 *)m->w_tile;
    w.posS = (typename Ops<R>::vec4 *)m->w_posS;
    w.ringS = (u32 *)m->w_ringS;
    w.ovfS = (u32 *)m->w_ovfS;
    w.dist[0] = (R *)m->w_dist[0];
    w.dist[1] = (R *)m->w_dist[1];
    w.cl[0] = (u32 *)m->w_cl[0];
    w.cl[1] = (u32 *)m->w_cl[1];
    w.wl = (u32 *)m->w_wl;
    w.dirty[0] = (unsigned char *)m->w_dirty[0];
    w.dirty[1] = (unsigned char *)m->w_dirty[1];
    w.toplesets = nullptr;
    w.ctrl = (ull *)m->w_ctrl;
    return w;
}

// (re)allocate the single-solve workspace for up to `scap` sources
template <class R> int ensure_workspace(ptp_mesh *m, u64 S, bool need_cl, bool need_top)
{
    if (m->ws_scap >= S && m->ws_scap != 0 && (!need_cl || m->ws_cl) && (!need_top || m->ws_top)) return PTP_OK;
    const u64 scap = std::max<u64>(std::max<u64>(S, m->ws_scap), 1024);
    const bool cl = need_cl || m->ws_cl, top = need_top || m->ws_top;
    free_list(m, m->ws_allocs);
    m->ws_scap = 0;
    const u64 V = m->V, N = V + scap;
    int rc;
#define WS(ptr, bytes)                                               \
    if ((rc = dev_alloc(m, &(ptr), (bytes), &m->ws_allocs)) != PTP_OK) return rc;
    WS(m->w_key, 8 * V)
    WS(m->w_sorted, 4 * N)
    WS(m->w_inv, 4 * V)
    WS(m->w_limits, 4 * (V + 2))
    WS(m->w_tile, 4 * 4096)
    WS(m->w_posS, 4 * sizeof(R) * (N + 1))
    WS(m->w_ringS, 4 * GL * N)
    WS(m->w_ovfS, 4 * std::max<u64>(m->ovf_total, 4))
    WS(m->w_dist[0], sizeof(R) * (N + 1))
    WS(m->w_dist[1], sizeof(R) * (N + 1))
    WS(m->w_wl, 4 * N)
    WS(m->w_dirty[0], N + 1)
    WS(m->w_dirty[1], N + 1)
    if (cl) {
        WS(m->w_cl[0], 4 * (N + 1))
        WS(m->w_cl[1], 4 * (N + 1))
        WS(m->w_clout, 4 * V)
    } else {
        m->w_cl[0] = m->w_cl[1] = m->w_clout = nullptr;
    }
    if (top) { WS(m->w_top, 4 * V) } else m->w_top = nullptr;
    WS(m->w_ctrl, 8 * C_COUNT)
    WS(m->w_bar, 1024)
    WS(m->w_src, 4 * scap)
    WS(m->w_out, sizeof(R) * V)
    WS(m->w_maxval, 16)
#undef WS
    m->ws_scap = scap;
    m->ws_cl = cl;
    m->ws_top = top;
    return PTP_OK;
}

int check_sources(const ptp_mesh *m, const u32 *sources, u64 n)
{
    if (!sources || n == 0) return fail(PTP_ERR_INVALID, "sources must be non-empty");
    for (u64 i = 0; i < n; i++)
        if (sources[i] >= m->V) return fail(PTP_ERR_INVALID, "source index out of range");
    return PTP_OK;
}

template <class K> int coop_grid(K kernel, int block, const ptp_mesh *m, int *grid)
{
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0));
    if (per_sm < 1) return fail(PTP_ERR_CUDA, "cooperative kernel does not fit on an SM");
    *grid = m->num_sms; // one CTA per SM: the fewest barrier participants that still cover the chip
    return PTP_OK;
}

float ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

// launch BFS (device toplesets). Leaves limits/sorted/inv/ctrl on the device.
template <class R> int launch_bfs(ptp_mesh *m, u32 S, u32 kcap, bool want_top)
{
    MeshView<R> mv = mesh_view<R>(m);
    Work<R> w = work_view<R>(m);
    w.toplesets = want_top ? (u32 *)m->w_top : nullptr;
    int grid;
    int rc = coop_grid(k_bfs_grid<R>, GRID_BLOCK, m, &grid);
    if (rc) return rc;
    const u32 *src = (const u32 *)m->w_src;
    ull *bar = (ull *)m->w_bar;
    u32 sent = (u32)(m->V + m->ws_scap);
    void *args[] = {&mv, &w, &src, &S, &kcap, &sent, &bar};
    CK(cudaMemsetAsync(m->w_bar, 0, 1024, m->stream));
    CK(cudaLaunchCooperativeKernel((void *)k_bfs_grid<R>, dim3(grid), dim3(GRID_BLOCK), args, 0, m->stream));
    return PTP_OK;
}

template <class R> int launch_layout(ptp_mesh *m)
{
    MeshView<R> mv = mesh_view<R>(m);
    Work<R> w = work_view<R>(m);
    const int grid = m->num_sms * 8;
    k_layout<R><<<grid, FLAT_BLOCK, 0, m->stream>>>(mv, w, (u32)(m->V + m->ws_scap));
    CK(cudaGetLastError());
    return PTP_OK;
}

bool use_staging() { return opt("stage") != 0 && PTP_GRID_MAP == 4; }

template <class R> int launch_solve(ptp_mesh *m, u32 S, bool cl, u32 cl_fill, const R *exactS = nullptr, double *iter_err = nullptr, u32 iter_cap = 0)
{
    MeshView<R> mv = mesh_view<R>(m);
    Work<R> w = work_view<R>(m);
    w.exactS = exactS;
    w.iter_err = iter_err;
    w.iter_cap = iter_cap;
    int grid;
    void *fn = cl ? (void *)k_solve_grid<R, true> : (void *)k_solve_grid<R, false>;
    int rc = cl ? coop_grid(k_solve_grid<R, true>, SOLVE_BLOCK, m, &grid) : coop_grid(k_solve_grid<R, false>, SOLVE_BLOCK, m, &grid);
    if (rc) return rc;
    const u32 *src = (const u32 *)m->w_src;
    R *out = (R *)m->w_out;
    u32 *clo = (u32 *)m->w_clout;
    ull *bar = (ull *)m->w_bar;
    u32 sent = (u32)(m->V + m->ws_scap);
    u32 staged = use_staging() ? 1u : 0u;
    const size_t smem = staged ? SOLVE_BLOCK * Stage4<R>::bytes_per_thread() : 0;
    if (staged) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {&mv, &w, &src, &S, &out, &clo, &cl_fill, &sent, &bar, &staged};
    CK(cudaMemsetAsync(m->w_bar, 0, 1024, m->stream));
    CK(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(SOLVE_BLOCK), args, smem, m->stream));
    m->last_kernel = sizeof(R) == 8 ? "k_solve_grid<double>" : "k_solve_grid<float>";
    return PTP_OK;
}

// number of CTAs given to the BFS/layout team of the fused single-solve kernel (PTP_BFS_CTAS overrides)
int bfs_ctas(const ptp_mesh *m)
{
    const int env = (int)opt("bfs_ctas");
    int nb = env > 0 ? env : m->num_sms; // default: one BFS CTA and one sweep CTA per SM
    return std::max(1, std::min(nb, 2 * m->num_sms - 1));
}

bool use_fused() { return opt("fused") != 0; }

template <class R> int launch_fused(ptp_mesh *m, u32 S, bool cl, u32 cl_fill)
{
    MeshView<R> mv = mesh_view<R>(m);
    Work<R> w = work_view<R>(m);
    if (!cl) w.cl[0] = w.cl[1] = nullptr;
    void *fn = cl ? (void *)k_geodesics_fused<R, true> : (void *)k_geodesics_fused<R, false>;
    // Staging the window in shared memory pays in the stand-alone sweep (1 CTA per SM); with two CTAs per SM it
    // takes 160 KB of the SM's 228 KB L1/shared array away from both teams and measured slower (C3: 32.8 vs
    // 29.6 ms), so the fused kernel runs unstaged unless PTP_STAGE=2.
    const bool stage_fused = opt("stage") == 2;
    u32 staged = (stage_fused && PTP_GRID_MAP == 4) ? 1u : 0u;
    size_t smem = staged ? FUSED_BLOCK * Stage4<R>::bytes_per_thread() : 0;
    int per_sm = 0;
    if (staged) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, FUSED_BLOCK, smem));
    if (per_sm < 2 && staged) { // not enough shared memory for two staged CTAs per SM: run unstaged
        staged = 0;
        smem = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, FUSED_BLOCK, 0));
    }
    if (per_sm < 2) return fail(PTP_ERR_CUDA, "fused kernel needs two resident CTAs per SM");
    const int grid = 2 * m->num_sms;
    const u32 *src = (const u32 *)m->w_src;
    R *out = (R *)m->w_out;
    u32 *clo = (u32 *)m->w_clout;
    ull *bar = (ull *)m->w_bar;
    u32 sent = (u32)(m->V + m->ws_scap);
    u32 nb = (u32)bfs_ctas(m);
    void *args[] = {&mv, &w, &src, &S, &out, &clo, &cl_fill, &sent, &bar, &nb, &staged};
    CK(cudaMemsetAsync(m->w_bar, 0, 1024, m->stream));
    CK(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(FUSED_BLOCK), args, smem, m->stream));
    m->last_kernel = sizeof(R) == 8 ? "k_geodesics_fused<double>" : "k_geodesics_fused<float>";
    return PTP_OK;
}

// PTP_FUSED=4 (default when the device can co-schedule the clusters): BFS on one thread-block cluster + sweep team.
// Returns PTP_OK and sets *launched = false when the configuration is not available (caller falls back).
int cluster_size()
{
    return std::max(1, std::min((int)opt("cluster"), 16));
}

// Geometry table (MeshView::geo), built once per mesh on first use. Not fatal when it does not fit: the kernels then
// recompute the geometry (`*ok` = false).
template <class R> int ensure_geo(ptp_mesh *m, cudaStream_t stream, bool *ok)
{
    *ok = m->geo != nullptr;
    if (m->geo || m->geo_failed) return PTP_OK;
    const size_t bytes = 4 * sizeof(R) * GL * m->V;
    if (cudaMalloc(&m->geo, bytes) != cudaSuccess) {
        cudaGetLastError();
        m->geo = nullptr;
        m->geo_failed = true;
        return PTP_OK;
    }
    m->bytes += bytes;
    k_geo_build<R><<<(unsigned)((m->V + 127) / 128), 128, 0, stream>>>((const typename Ops<R>::vec4 *)m->GT4, m->ring8, (u32)m->V,
                                                                      (typename Ops<R>::vec4 *)m->geo);
    CK(cudaGetLastError());
    *ok = true;
    return PTP_OK;
}

template <class R> int ensure_safe(ptp_mesh *m, cudaStream_t stream)
{
    const int with_sign = ((PTP_SIGN_SHORT && opt("sign_short") != 0) ? 1 : 0) | ((PTP_SIGN_SHORT && PTP_TWO_SIDED && opt("two_sided") != 0) ? 2 : 0);
    if (m->safe8 && m->safe_sign == with_sign) return PTP_OK;
    int rc;
    if (!m->safe8 && (rc = dev_alloc(m, (void **)&m->safe8, 3 * m->V, nullptr))) return rc;
    k_safe_build<R><<<(unsigned)((m->V + 127) / 128), 128, 0, stream>>>((const typename Ops<R>::vec4 *)m->GT4, m->ring8, (u32)m->V, m->safe8,
                                                                     (u32)(with_sign & 1), (u32)(with_sign >> 1));
    CK(cudaGetLastError());
    m->safe_sign = with_sign;
    return PTP_OK;
}

template <class R> int launch_cluster(ptp_mesh *m, u32 S, bool cl, u32 cl_fill, bool *launched)
{
    *launched = false;
    // Single solve with the geometry table (PTP_GEO_SINGLE=1, off): takes 3 divisions + 2 square roots per triangle off
    // the dependent FP chain of an iteration (thread-0 stamps on C3: compute 3.2 -> 2.1 us per iteration) but the two
    // extra 32-byte records per lane lengthen the gather phase by as much (1.05 -> 2.0 us): 30.1 vs 28.5 ms per solve.
    const bool want_geo = opt("geo_single") != 0;
    bool geo = false;
    int rc;
    if (want_geo && (rc = ensure_geo<R>(m, m->stream, &geo))) return rc;
    MeshView<R> mv = mesh_view<R>(m);
    if (!geo) mv.geo = nullptr;
    Work<R> w = work_view<R>(m);
    if (!cl) w.cl[0] = w.cl[1] = nullptr;
    void *fn = geo ? (cl ? (void *)k_geodesics_cluster<R, true, true> : (void *)k_geodesics_cluster<R, false, true>)
                   : (cl ? (void *)k_geodesics_cluster<R, true, false> : (void *)k_geodesics_cluster<R, false, false>);
    // the staged window needs (window + entering topleset) <= groups of the sweep team; with ~110 sweep CTAs the widest
    // C3 windows do not fit and the streamed sweep measured faster unstaged (27.9 vs 29.0 ms): PTP_STAGE=1 turns it on
    const bool want_stage = opt("stage") == 1;
    u32 staged = (want_stage && PTP_GRID_MAP == 4) ? 1u : 0u;
    size_t smem = staged ? CLUSTER_BLOCK * Stage4<R>::bytes_per_thread() : 0;
    const int csize = cluster_size();
    if (csize > 8 && cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        return PTP_OK;
    }
    if (staged) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = (unsigned)csize;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    attrs[1].id = cudaLaunchAttributeCooperative;
    attrs[1].val.cooperative = 1;
    cfg.blockDim = dim3(CLUSTER_BLOCK);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = m->stream;
    cfg.attrs = attrs;
    cfg.numAttrs = 2;
    cfg.gridDim = dim3((unsigned)(m->num_sms / csize * csize));
    int n_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&n_clusters, fn, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return PTP_OK;
    }
    // every CTA must be resident at once (software grid barrier of the sweep team + the BFS cluster it waits for)
    const int grid = std::min(n_clusters, m->num_sms / csize) * csize;
    if (grid < csize + 16) return PTP_OK;
    cfg.gridDim = dim3((unsigned)grid);
    const u32 *src = (const u32 *)m->w_src;
    R *out = (R *)m->w_out;
    u32 *clo = (u32 *)m->w_clout;
    ull *bar = (ull *)m->w_bar;
    u32 sent = (u32)(m->V + m->ws_scap);
    u32 nb = (u32)csize;
    void *args[] = {&mv, &w, &src, &S, &out, &clo, &cl_fill, &sent, &bar, &nb, &staged};
    CK(cudaMemsetAsync(m->w_bar, 0, 1024, m->stream));
    cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (opt("debug")) fprintf(stderr, "[ptp] cluster launch refused (%s): falling back to the two-team kernel\n", cudaGetErrorString(e));
        return PTP_OK;
    }
    if (opt("debug")) fprintf(stderr, "[ptp] cluster kernel: grid %d, cluster %d, staged %u, smem %zu\n", grid, csize, staged, smem);
    *launched = true;
    m->last_kernel = sizeof(R) == 8 ? "k_geodesics_cluster<double>" : "k_geodesics_cluster<float>";
    return PTP_OK;
}

// PTP_FUSED=5: BFS cluster kernel + sweep kernel side by side (see k_toplesets_cluster): same stream, the second one a
// programmatic dependent of the first, so that it starts once the BFS cluster is resident instead of when it has finished.
template <class R> int launch_two_kernels(ptp_mesh *m, u32 S, bool cl, u32 cl_fill, bool *launched)
{
    *launched = false;
    m->last_two = false;
    const int csize = cluster_size();
    if (m->two_failed || m->num_sms < csize + 32) return PTP_OK;
    MeshView<R> mv = mesh_view<R>(m);
    mv.geo = nullptr;
    Work<R> w = work_view<R>(m);
    if (!cl) w.cl[0] = w.cl[1] = nullptr;
    const u32 *src = (const u32 *)m->w_src;
    R *out = (R *)m->w_out;
    u32 *clo = (u32 *)m->w_clout;
    ull *bar = (ull *)m->w_bar;
    u32 sent = (u32)(m->V + m->ws_scap);
    void *fb = (void *)k_toplesets_cluster<R>;
    void *fs = cl ? (void *)k_sweep_streamed<R, true, false> : (void *)k_sweep_streamed<R, false, false>;
    if (csize > 8 && cudaFuncSetAttribute(fb, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        return PTP_OK;
    }
    // both kernels must be loaded before either runs: with lazy module loading the first launch of the second one would
    // otherwise wait for the device to drain, i.e. for the first one, which is waiting for it
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, fb));
    CK(cudaFuncGetAttributes(&fa, fs));
    CK(cudaMemsetAsync(m->w_bar, 0, 1024, m->stream));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)csize;
    at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)csize);
    cfg.blockDim = dim3(CLUSTER_BLOCK);
    cfg.stream = m->stream;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    void *a1[] = {&mv, &w, &src, &S};
    if (cudaLaunchKernelExC(&cfg, fb, a1) != cudaSuccess) {
        cudaGetLastError();
        return PTP_OK;
    }
    // (a plain launch, not a cooperative one: that would be held back until the device is idle; with one CTA per
    // remaining SM every CTA is resident, and the watchdog covers the rest)
    cudaLaunchConfig_t cfg2 = {};
    cudaLaunchAttribute at2[1];
    at2[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at2[0].val.programmaticStreamSerializationAllowed = 1;
    // window staged in shared memory ("stage" = 1): measured on C3, see profiles/README.md
    u32 staged = (opt("stage") == 1 && PTP_GRID_MAP == 4) ? 1u : 0u;
    const size_t smem2 = staged ? CLUSTER_BLOCK * Stage4<R>::bytes_per_thread() : 0;
    if (staged && cudaFuncSetAttribute(fs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2) != cudaSuccess) {
        cudaGetLastError();
        staged = 0;
    }
    cfg2.gridDim = dim3((unsigned)(m->num_sms - csize));
    cfg2.blockDim = dim3(CLUSTER_BLOCK);
    cfg2.dynamicSmemBytes = staged ? smem2 : 0;
    cfg2.stream = m->stream;
    cfg2.attrs = at2;
    cfg2.numAttrs = 1;
    void *a2[] = {&mv, &w, &src, &S, &out, &clo, &cl_fill, &sent, &bar, &staged};
    if (cudaLaunchKernelExC(&cfg2, fs, a2) != cudaSuccess) {
        // the dependent launch was refused: the BFS kernel already in the stream gives up through its watchdog; start
        // over with a clean control block and let the caller use the one-launch kernel, now and from now on
        cudaGetLastError();
        m->two_failed = true;
        CK(cudaStreamSynchronize(m->stream));
        CK(cudaMemsetAsync(m->w_ctrl, 0, 8 * C_COUNT, m->stream));
        return PTP_OK;
    }
    m->last_kernel = sizeof(R) == 8 ? "k_sweep_streamed<double> + k_toplesets_cluster<double>" : "k_sweep_streamed<float> + k_toplesets_cluster<float>";
    m->last_two = true;
    *launched = true;
    return PTP_OK;
}

void fill_stats(const ptp_mesh *m, ptp_stats_t *st, u64 launches, double ms_top, double ms_solve, double ms_total)
{
    if (!st) return;
    const ull *c = (const ull *)m->h_ctrl;
    st->n_reached = c[C_REACHED];
    st->n_levels = c[C_NLIMITS] ? c[C_NLIMITS] - 1 : 0;
    st->iterations = c[C_ITER];
    st->vertex_updates = c[C_UPDATES];
    st->max_window = c[C_MAXWIN];
    st->relaxations = c[C_RELAXED];
    st->gpu_launches = launches;
    st->ms_toplesets = ms_top;
    st->ms_solve = ms_solve;
    st->ms_total = ms_total;
}

int che_build_device(const u32 *d_vt, u64 V, u64 H, u32 *d_ot, u32 *d_evt, cudaStream_t stream, bool *manifold)
{
    u64 cap = 1;
    while (cap < 2 * H) cap <<= 1;
    ull *keys = nullptr, *flags = nullptr;
    u32 *vals = nullptr, *evt1 = nullptr, *bcnt = nullptr, *bhe = nullptr;
    auto cleanup = [&]() { cudaFree(keys); cudaFree(vals); cudaFree(evt1); cudaFree(bcnt); cudaFree(bhe); cudaFree(flags); };
    auto run = [&]() -> int {
        CK(cudaMalloc(&keys, 8 * cap));
        CK(cudaMalloc(&vals, 4 * cap));
        CK(cudaMalloc(&evt1, 4 * V));
        CK(cudaMalloc(&bcnt, 4 * V));
        CK(cudaMalloc(&bhe, 4 * V));
        CK(cudaMalloc(&flags, 16));
        CK(cudaMemsetAsync(keys, 0xFF, 8 * cap, stream));
        CK(cudaMemsetAsync(evt1, 0, 4 * V, stream));
        CK(cudaMemsetAsync(bcnt, 0, 4 * V, stream));
        CK(cudaMemsetAsync(flags, 0, 16, stream));
        const unsigned gh = (unsigned)((H + 255) / 256), gv = (unsigned)((V + 255) / 256);
        k_che_insert<<<gh, 256, 0, stream>>>(d_vt, (u32)H, (u32)V, keys, vals, (u32)(cap - 1), evt1, flags);
        k_che_pair<<<gh, 256, 0, stream>>>(d_vt, (u32)H, (u32)V, keys, vals, (u32)(cap - 1), d_ot, bcnt, bhe);
        k_che_evt<<<gv, 256, 0, stream>>>((u32)V, evt1, bcnt, bhe, d_evt, flags);
        CK(cudaGetLastError());
        ull hf[2] = {0, 0};
        CK(cudaMemcpyAsync(hf, flags, 16, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (hf[1]) return fail(PTP_ERR_INVALID, "face list references a vertex >= n_vertices");
        *manifold = hf[0] == 0;
        return PTP_OK;
    };
    const int rc = run();
    cleanup();
    return rc;
}

template <class R>
int mesh_create(const R *GT, const u32 *VT, const u32 *OT, const u32 *EVT, u64 V, u64 H, int device, ptp_mesh_t **out)
{
    if (!out) return fail(PTP_ERR_INVALID, "out is null");
    *out = nullptr;
    if (!GT || !VT) return fail(PTP_ERR_INVALID, "null mesh table");
    if ((OT == nullptr) != (EVT == nullptr)) return fail(PTP_ERR_INVALID, "pass both OT and EVT, or neither (built on the device)");
    if (V == 0 || H == 0 || H % 3 != 0) return fail(PTP_ERR_INVALID, "need V > 0 and H = 3 * faces > 0");
    if (V >= 0x7FFFFFF0ull || H >= 0xFFFFFFF0ull) return fail(PTP_ERR_INVALID, "mesh too large for 31-bit vertex ranks");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(PTP_ERR_NO_DEVICE, "no such CUDA device");
    CK(cudaSetDevice(device));

    ptp_mesh *m = new (std::nothrow) ptp_mesh();
    if (!m) return fail(PTP_ERR_INVALID, "out of host memory");
    m->device = device;
    m->real_size = (int)sizeof(R);
    m->V = V;
    m->H = H;
    int rc = PTP_OK;
    auto bail = [&](int code) {
        ptp_mesh_destroy(m);
        return code;
    };
#define CKM(call)                                     \
    do {                                              \
        rc = [&]() -> int { CK(call); return PTP_OK; }(); \
        if (rc) return bail(rc);                      \
    } while (0)

    CKM(cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, device));
    CKM(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    for (auto &e : m->ev) CKM(cudaEventCreate(&e));
    CKM(cudaHostAlloc(&m->h_ctrl, 8 * C_COUNT, cudaHostAllocDefault));

    void *d_gt = nullptr, *d_vt = nullptr, *d_ot = nullptr, *d_evt = nullptr, *d_cnt = nullptr;
    CKM(cudaMalloc(&d_gt, sizeof(R) * 3 * V));
    CKM(cudaMalloc(&d_vt, 4 * H));
    CKM(cudaMalloc(&d_ot, 4 * H));
    CKM(cudaMalloc(&d_evt, 4 * V));
    CKM(cudaMalloc(&d_cnt, 32));
    auto free_tmp = [&]() {
        cudaFree(d_gt); cudaFree(d_vt); cudaFree(d_ot); cudaFree(d_evt); cudaFree(d_cnt);
    };
#define CKT(call)                                     \
    do {                                              \
        rc = [&]() -> int { CK(call); return PTP_OK; }(); \
        if (rc) { free_tmp(); return bail(rc); }      \
    } while (0)
    CKT(cudaMemcpyAsync(d_gt, GT, sizeof(R) * 3 * V, cudaMemcpyHostToDevice, m->stream));
    CKT(cudaMemcpyAsync(d_vt, VT, 4 * H, cudaMemcpyHostToDevice, m->stream));
    if (OT) {
        CKT(cudaMemcpyAsync(d_ot, OT, 4 * H, cudaMemcpyHostToDevice, m->stream));
        CKT(cudaMemcpyAsync(d_evt, EVT, 4 * V, cudaMemcpyHostToDevice, m->stream));
    } else {
        bool manifold = true;
        if ((rc = che_build_device((const u32 *)d_vt, V, H, (u32 *)d_ot, (u32 *)d_evt, m->stream, &manifold))) { free_tmp(); return bail(rc); }
        if (!manifold) { free_tmp(); bail(0); return fail(PTP_ERR_MESH, "face list is not an oriented edge-manifold mesh"); }
    }
    CKT(cudaMemsetAsync(d_cnt, 0, 32, m->stream));

    if ((rc = dev_alloc(m, &m->GT4, sizeof(R) * 4 * V, nullptr))) { free_tmp(); return bail(rc); }
    if ((rc = dev_alloc(m, (void **)&m->ring8, 4 * GL * V, nullptr))) { free_tmp(); return bail(rc); }

    k_pad_gt<R><<<(unsigned)((V * 4 + 255) / 256), 256, 0, m->stream>>>((const R *)d_gt, (R *)m->GT4, (u32)V);
    k_ring_build<<<(unsigned)((V + 127) / 128), 128, 0, m->stream>>>((const u32 *)d_vt, (const u32 *)d_ot, (const u32 *)d_evt,
                                                                      (u32)V, (u32)H, m->ring8, nullptr, (ull *)d_cnt, 0);
    CKT(cudaGetLastError());
    ull cnt[2] = {0, 0};
    CKT(cudaMemcpyAsync(cnt, d_cnt, 16, cudaMemcpyDeviceToHost, m->stream));
    CKT(cudaStreamSynchronize(m->stream));
    if (cnt[1]) { free_tmp(); bail(0); return fail(PTP_ERR_MESH, "inconsistent CHE tables: a one-ring walk left the mesh or did not terminate"); }
    m->ovf_total = cnt[0];
    if (m->ovf_total >= 0xFFFFFFF0ull) { free_tmp(); bail(0); return fail(PTP_ERR_INVALID, "overflow pool too large"); }
    if (m->ovf_total) {
        if ((rc = dev_alloc(m, (void **)&m->ovf, 4 * m->ovf_total, nullptr))) { free_tmp(); return bail(rc); }
        // pass 1 reuses the offsets stored in the rows by pass 0
        k_ring_build<<<(unsigned)((V + 127) / 128), 128, 0, m->stream>>>((const u32 *)d_vt, (const u32 *)d_ot, (const u32 *)d_evt,
                                                                          (u32)V, (u32)H, m->ring8, m->ovf, (ull *)d_cnt, 1);
        CKT(cudaGetLastError());
        CKT(cudaStreamSynchronize(m->stream));
    }
    k_ring_check<<<(unsigned)((V + 127) / 128), 128, 0, m->stream>>>(m->ring8, m->ovf, (u32)V, (ull *)d_cnt);
    CKT(cudaGetLastError());
    ull asym[4] = {0, 0, 0, 0};
    CKT(cudaMemcpyAsync(asym, d_cnt, 32, cudaMemcpyDeviceToHost, m->stream));
    CKT(cudaStreamSynchronize(m->stream));
    m->ring_symmetric = asym[2] == 0;
    free_tmp();
#undef CKT
#undef CKM
    *out = m;
    return PTP_OK;
}

template <class R> int upload_sources(ptp_mesh *m, const u32 *sources, u32 S)
{
    CK(cudaMemcpyAsync(m->w_src, sources, 4ull * S, cudaMemcpyHostToDevice, m->stream));
    CK(cudaMemsetAsync(m->w_ctrl, 0, 8 * C_COUNT, m->stream));
    return PTP_OK;
}

int fetch_ctrl(ptp_mesh *m)
{
    CK(cudaMemcpyAsync(m->h_ctrl, m->w_ctrl, 8 * C_COUNT, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    const ull wd = ((const ull *)m->h_ctrl)[C_ERROR];
    if (wd) {
        static const char *what[] = {"", "grid barrier", "sweep team waiting for toplesets / rows", "layout warp waiting for the BFS",
                                     "layout warp waiting for its turn to publish", "BFS cluster waiting for its tables", "elastic relax chunks"};
        return fail(PTP_ERR_CUDA, std::string("device watchdog: a wait did not complete (") + (wd < 7 ? what[wd] : "?") + "); results discarded");
    }
    return PTP_OK;
}

template <class R>
int toplesets_impl(ptp_mesh *m, const u32 *sources, u32 S, u32 k, u32 *toplesets, u32 *sorted, u64 scap, u32 *limits, u64 lcap,
                   u32 *n_limits, ptp_stats_t *st)
{
    int rc;
    CK(cudaSetDevice(m->device));
    if ((rc = check_sources(m, sources, S))) return rc;
    if ((rc = ensure_workspace<R>(m, S, false, toplesets != nullptr))) return rc;
    if ((rc = upload_sources<R>(m, sources, S))) return rc;
    CK(cudaEventRecord(m->ev[0], m->stream));
    if ((rc = launch_bfs<R>(m, S, k, toplesets != nullptr))) return rc;
    CK(cudaEventRecord(m->ev[1], m->stream));
    if ((rc = fetch_ctrl(m))) return rc;
    const ull *c = (const ull *)m->h_ctrl;
    const u64 nl = c[C_NLIMITS], p = c[C_REACHED];
    if (n_limits) *n_limits = (u32)nl;
    if (limits) {
        if (lcap < nl) return fail(PTP_ERR_CAPACITY, "limits buffer too small");
        CK(cudaMemcpyAsync(limits, m->w_limits, 4 * nl, cudaMemcpyDeviceToHost, m->stream));
    }
    if (sorted) {
        if (scap < p) return fail(PTP_ERR_CAPACITY, "sorted buffer too small (needs V + duplicate sources)");
        CK(cudaMemcpyAsync(sorted, m->w_sorted, 4 * p, cudaMemcpyDeviceToHost, m->stream));
    }
    if (toplesets) CK(cudaMemcpyAsync(toplesets, m->w_top, 4 * m->V, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    const double ms = ev_ms(m->ev[0], m->ev[1]);
    fill_stats(m, st, 1, ms, 0, ms);
    return PTP_OK;
}

template <class R>
int solve_impl(ptp_mesh *m, const u32 *sources, u32 S, const u32 *limits, u32 nl, const u32 *sorted, R *dist, u32 *clusters,
               u32 cl_fill, ptp_stats_t *st)
{
    int rc;
    CK(cudaSetDevice(m->device));
    if ((rc = check_sources(m, sources, S))) return rc;
    if (!limits || !sorted || !dist || nl < 2) return fail(PTP_ERR_INVALID, "limits (>= 2 entries), sorted and dist are required");
    if (nl > m->V + 2) return fail(PTP_ERR_INVALID, "limits longer than V + 2");
    const u64 p = limits[nl - 1];
    for (u32 i = 1; i < nl; i++)
        if (limits[i] < limits[i - 1]) return fail(PTP_ERR_INVALID, "limits must be non-decreasing");
    if (p > m->V + S) return fail(PTP_ERR_INVALID, "limits.back() exceeds V + n_sources");
    if ((rc = ensure_workspace<R>(m, S, clusters != nullptr, false))) return rc;
    if ((rc = upload_sources<R>(m, sources, S))) return rc;
    CK(cudaMemcpyAsync(m->w_sorted, sorted, 4 * p, cudaMemcpyHostToDevice, m->stream));
    CK(cudaMemcpyAsync(m->w_limits, limits, 4ull * nl, cudaMemcpyHostToDevice, m->stream));
    ull hc[2] = {nl, p};
    CK(cudaMemcpyAsync(m->w_ctrl, hc, 16, cudaMemcpyHostToDevice, m->stream));
    CK(cudaEventRecord(m->ev[0], m->stream));
    MeshView<R> mv = mesh_view<R>(m);
    Work<R> w = work_view<R>(m);
    k_inv_init<R><<<(unsigned)((m->V + 255) / 256), 256, 0, m->stream>>>(mv, w);
    k_inv_fill<R><<<(unsigned)((p + 255) / 256), 256, 0, m->stream>>>(mv, w, (u32)p);
    CK(cudaGetLastError());
    // the caller's `sorted` is only trusted after this check: an entry >= V (NIL padding, which the reference's kernels
    // tolerate through `if(v < n_vertices)`) would index the mesh tables out of bounds in the layout pass
    CK(cudaMemcpyAsync((ull *)m->h_ctrl + C_ABORT, (ull *)m->w_ctrl + C_ABORT, 8, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    if (((const ull *)m->h_ctrl)[C_ABORT]) return fail(PTP_ERR_INVALID, "sorted[0 .. limits.back()) holds a vertex index >= n_vertices");
    if ((rc = launch_layout<R>(m))) return rc;
    CK(cudaEventRecord(m->ev[1], m->stream));
    if ((rc = launch_solve<R>(m, S, clusters != nullptr, cl_fill))) return rc;
    CK(cudaEventRecord(m->ev[2], m->stream));
    CK(cudaMemcpyAsync(dist, m->w_out, sizeof(R) * m->V, cudaMemcpyDeviceToHost, m->stream));
    if (clusters) CK(cudaMemcpyAsync(clusters, m->w_clout, 4 * m->V, cudaMemcpyDeviceToHost, m->stream));
    if ((rc = fetch_ctrl(m))) return rc;
    fill_stats(m, st, 4, ev_ms(m->ev[0], m->ev[1]), ev_ms(m->ev[1], m->ev[2]), ev_ms(m->ev[0], m->ev[2]));
    return PTP_OK;
}

// device pipeline: sources already in w_src, ctrl zeroed. Records ev[0..2].
template <class R> int pipeline(ptp_mesh *m, u32 S, bool cl, u32 cl_fill)
{
    int rc;
    CK(cudaEventRecord(m->ev[0], m->stream));
    // PTP_FUSED: 5 (default) BFS cluster kernel + sweep kernel side by side, 4 the same two teams in ONE cluster launch,
    // 1 two-team kernel (grid barriers only), 0 three launches, 2 debug (the two teams one after the other)
    const int dbg = (int)opt("fused");
    if (dbg == 2 && !cl) {
        MeshView<R> mv = mesh_view<R>(m);
        Work<R> w = work_view<R>(m);
        w.cl[0] = w.cl[1] = nullptr;
        const u32 *src = (const u32 *)m->w_src;
        R *out = (R *)m->w_out;
        ull *bar = (ull *)m->w_bar;
        u32 sent = (u32)(m->V + m->ws_scap);
        void *a1[] = {&mv, &w, &src, &S, &sent, &bar};
        void *a2[] = {&mv, &w, &src, &S, &out, &sent, &bar};
        CK(cudaMemsetAsync(m->w_bar, 0, 1024, m->stream));
        CK(cudaLaunchCooperativeKernel((void *)k_dbg_producer<R>, dim3(m->num_sms), dim3(FUSED_BLOCK), a1, 0, m->stream));
        CK(cudaEventRecord(m->ev[1], m->stream));
        CK(cudaMemsetAsync(m->w_bar, 0, 1024, m->stream));
        CK(cudaLaunchCooperativeKernel((void *)k_dbg_consumer<R>, dim3(m->num_sms), dim3(FUSED_BLOCK), a2, 0, m->stream));
        CK(cudaEventRecord(m->ev[2], m->stream));
        return PTP_OK;
    }
    if (use_fused()) {
        bool launched = false;
        m->last_two = false;
        if (dbg >= 5 && (rc = launch_two_kernels<R>(m, S, cl, cl_fill, &launched))) return rc;
        if (!launched && dbg != 1 && (rc = launch_cluster<R>(m, S, cl, cl_fill, &launched))) return rc; // PTP_FUSED=1: the two-team kernel
        if (!launched && (rc = launch_fused<R>(m, S, cl, cl_fill))) return rc;
        CK(cudaEventRecord(m->ev[1], m->stream));
        CK(cudaEventRecord(m->ev[2], m->stream));
        return PTP_OK;
    }
    if ((rc = launch_bfs<R>(m, S, NIL, false))) return rc;
    if ((rc = launch_layout<R>(m))) return rc;
    CK(cudaEventRecord(m->ev[1], m->stream));
    if ((rc = launch_solve<R>(m, S, cl, cl_fill))) return rc;
    CK(cudaEventRecord(m->ev[2], m->stream));
    return PTP_OK;
}

template <class R>
int geodesics_impl(ptp_mesh *m, const u32 *sources, u32 S, R *dist, u32 *clusters, u32 cl_fill, u32 *sorted_index, u64 scap,
                   ptp_stats_t *st)
{
    int rc;
    CK(cudaSetDevice(m->device));
    if ((rc = check_sources(m, sources, S))) return rc;
    if (!dist) return fail(PTP_ERR_INVALID, "dist is null");
    if ((rc = ensure_workspace<R>(m, S, clusters != nullptr, false))) return rc;
    for (int attempt = 0;; attempt++) {
        if ((rc = upload_sources<R>(m, sources, S))) return rc;
        // "profile_range": the solve as ONE profiler range, so that `ncu --replay-mode range / app-range` measures the two
        // kernels of the default path while they run side by side (kernel replay serialises them)
        const bool prof = opt("profile_range") != 0;
        if (prof) { CK(cudaStreamSynchronize(m->stream)); cudaProfilerStart(); }
        rc = pipeline<R>(m, S, clusters != nullptr, cl_fill);
        if (prof) { cudaStreamSynchronize(m->stream); cudaProfilerStop(); }
        if (rc) return rc;
        CK(cudaMemcpyAsync(dist, m->w_out, sizeof(R) * m->V, cudaMemcpyDeviceToHost, m->stream));
        if (clusters) CK(cudaMemcpyAsync(clusters, m->w_clout, 4 * m->V, cudaMemcpyDeviceToHost, m->stream));
        rc = fetch_ctrl(m);
        if (rc && m->last_two && attempt == 0 && ((const ull *)m->h_ctrl)[C_ERROR]) {
            // the two kernels did not become resident together (the watchdog ended them): one launch from now on
            m->two_failed = true;
            continue;
        }
        if (rc) return rc;
        break;
    }
    if (opt("fused") == 2)
        fill_stats(m, st, 2, ev_ms(m->ev[0], m->ev[1]), ev_ms(m->ev[1], m->ev[2]), ev_ms(m->ev[0], m->ev[2]));
    else if (use_fused()) {
        // one launch: the producer / consumer split comes from %globaltimer stamps written by the kernel
        const ull *c = (const ull *)m->h_ctrl;
        const double t_bfs = c[C_TBFS] > c[C_TSTART] ? (c[C_TBFS] - c[C_TSTART]) * 1e-6 : 0.0;
        const double t_all = ev_ms(m->ev[0], m->ev[2]);
        if (opt("debug")) {
            fprintf(stderr, "[ptp] producer polls by the sweep team: %llu\n", c[C_ARGMAX]);
            if (c[C_TPHASE + 6]) fprintf(stderr, "[ptp] sweep thread-0 ms: relax %.2f | wait for producers %.2f | barrier %.2f | post-barrier %.2f\n",
                    c[C_TPHASE + 6] * 1e-6, c[C_TPHASE + 7] * 1e-6, c[C_TPHASE + 8] * 1e-6, c[C_TPHASE + 9] * 1e-6);
#ifdef PTP_PHASE_TIMERS
            {
                ull t[16];
                cudaMemcpyFromSymbol(t, g_dbg_t, sizeof t);
                fprintf(stderr, "[ptp] sweep CTA-0 thread-0 cumulative ms: top %.2f | row %.2f | gathers %.2f | compute %.2f | min+commit %.2f | stamps %.2f | rest %.2f | publish %.2f | barrier %.2f | post %.2f\n",
                        t[0] * 1e-6, t[1] * 1e-6, t[2] * 1e-6, t[3] * 1e-6, t[4] * 1e-6, t[5] * 1e-6, t[6] * 1e-6, t[7] * 1e-6, t[8] * 1e-6, t[9] * 1e-6);
                ull z[16] = {0};
                cudaMemcpyToSymbol(g_dbg_t, z, sizeof z);
            }
#endif
            if (c[C_TPHASE])
                fprintf(stderr, "[ptp] cluster BFS thread-0 ms: claim %.2f | barrier1 %.2f | publish %.2f | own+scan %.2f | barrier2 %.2f | place+barrier3 %.2f\n",
                        c[C_TPHASE] * 1e-6, c[C_TPHASE + 1] * 1e-6, c[C_TPHASE + 5] * 1e-6, c[C_TPHASE + 2] * 1e-6, c[C_TPHASE + 3] * 1e-6, c[C_TPHASE + 4] * 1e-6);
        }
        fill_stats(m, st, m->last_two ? 2 : 1, t_bfs, c[C_TEND] > c[C_TSTART] ? (c[C_TEND] - c[C_TSTART]) * 1e-6 : t_all, t_all);
    } else
        fill_stats(m, st, 3, ev_ms(m->ev[0], m->ev[1]), ev_ms(m->ev[1], m->ev[2]), ev_ms(m->ev[0], m->ev[2]));
    if (sorted_index) {
        // exactly the limits.back() entries che::compute_toplesets writes (src/che.cpp:555-592); the rest of the caller's
        // array is left as it was (geodesics::geodesics presets it to NIL, src/geodesics.cpp:26)
        const u64 p = ((const ull *)m->h_ctrl)[C_REACHED];
        CK(cudaMemcpyAsync(sorted_index, m->w_sorted, 4 * std::min<u64>(scap, p), cudaMemcpyDeviceToHost, m->stream));
        CK(cudaStreamSynchronize(m->stream));
        if (scap < p) return fail(PTP_ERR_CAPACITY, "sorted_index buffer too small (needs V + duplicate sources)");
    }
    return PTP_OK;
}

// CTAs per solve of the batched path: the "team" option, or (0) the default for this mesh
// Default (0): one CTA per solve when the batch can occupy the chip (measured on C5, 296 sources: 1 CTA per solve 263,
// teams of 2 / 4 / 8 / 16 / 37: 204 / 202 / 186 / 157 / 109 sources/s — the team barrier and the per-level BFS barriers cost
// more than L2 residency of the windows returns), a team of num_sms / B CTAs (at most 37) per solve when it cannot
// (8 sources on the 2 M-vertex sphere: 60 ms with teams of 18, 172 ms with one CTA per solve + elastic helpers).
int batch_team(const ptp_mesh *m, u32 B)
{
    long t = opt("team");
    if (t <= 0) t = (u64)B * 2 <= (u64)m->num_sms ? std::min<long>(37, m->num_sms / std::max<u32>(B, 1u)) : 1;
    return (int)std::max<long>(1, std::min<long>(t, m->num_sms));
}

// Per-iteration error of a solve (src/cuda/test_geodesics_ptp.cu:164-211, written to `<mesh>_error.iter` by
// src/test_geodesics_ptp.cpp:198-214): the three-launch path (BFS, layout, stand-alone sweep) with the sweep in
// measurement mode. errors[k] = 100 / (V - S) * sum over exact > 0 of |dist - exact| / exact after iteration iters[k],
// for every iteration whose window ends at the last topleset.
template <class R>
int error_iter_impl(ptp_mesh *m, const u32 *sources, u32 S, const R *exact, R *dist, u32 *iters, R *errors, u32 cap, u32 *n_out,
                    ptp_stats_t *st)
{
    int rc;
    CK(cudaSetDevice(m->device));
    if ((rc = check_sources(m, sources, S))) return rc;
    if (!exact || !iters || !errors || !n_out || cap == 0) return fail(PTP_ERR_INVALID, "exact, iters, errors (capacity > 0) and n_out are required");
    if ((rc = ensure_workspace<R>(m, S, false, false))) return rc;
    R *d_exact = nullptr, *d_exactS = nullptr;
    double *d_rec = nullptr;
    std::vector<double> rec(2 * (size_t)cap);
    auto run = [&]() -> int {
        CK(cudaMalloc(&d_exact, sizeof(R) * m->V));
        CK(cudaMalloc(&d_exactS, sizeof(R) * (m->V + m->ws_scap + 1)));
        CK(cudaMalloc(&d_rec, 16 * (size_t)cap));
        CK(cudaMemcpyAsync(d_exact, exact, sizeof(R) * m->V, cudaMemcpyHostToDevice, m->stream));
        CK(cudaMemsetAsync(d_rec, 0, 16 * (size_t)cap, m->stream));
        int r2;
        if ((r2 = upload_sources<R>(m, sources, S))) return r2;
        CK(cudaEventRecord(m->ev[0], m->stream));
        if ((r2 = launch_bfs<R>(m, S, NIL, false))) return r2;
        if ((r2 = launch_layout<R>(m))) return r2;
        k_gather_exact<R><<<m->num_sms * 4, 256, 0, m->stream>>>(work_view<R>(m), d_exact, d_exactS);
        CK(cudaGetLastError());
        CK(cudaEventRecord(m->ev[1], m->stream));
        if ((r2 = launch_solve<R>(m, S, false, 0, d_exactS, d_rec, cap))) return r2;
        CK(cudaEventRecord(m->ev[2], m->stream));
        if (dist) CK(cudaMemcpyAsync(dist, m->w_out, sizeof(R) * m->V, cudaMemcpyDeviceToHost, m->stream));
        CK(cudaMemcpyAsync(rec.data(), d_rec, 16 * (size_t)cap, cudaMemcpyDeviceToHost, m->stream));
        return fetch_ctrl(m);
    };
    rc = run();
    cudaFree(d_exact); cudaFree(d_exactS); cudaFree(d_rec);
    if (rc) return rc;
    const ull *c = (const ull *)m->h_ctrl;
    const u32 n = (u32)std::min<ull>(c[C_NITERR], cap);
    const bool all_reached = c[C_REACHED] >= m->V; // an unreached vertex with exact > 0 makes the reference's sum infinite
    for (u32 k = 0; k < n; k++) {
        iters[k] = (u32)rec[2 * k];
        errors[k] = all_reached ? (R)(rec[2 * k + 1] * 100.0 / (double)(m->V - S)) : (R)INFINITY;
    }
    *n_out = n;
    fill_stats(m, st, 4, ev_ms(m->ev[0], m->ev[1]), ev_ms(m->ev[1], m->ev[2]), ev_ms(m->ev[0], m->ev[2]));
    return PTP_OK;
}

template <class R> int ensure_batch(ptp_mesh *m, u64 max_s, u64 n_src, u64 n_off, u64 rows_elems, u32 B)
{
    int rc;
    const u32 team = (u32)batch_team(m, B);
    int per_sm = 0;
    if (team > 1) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_batched_teams<R, false, true>, BatchCfg<R>::BLOCK, 0));
    else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_batched<R, false, true>, BatchCfg<R>::BLOCK, 0));
    if (per_sm < 1) return fail(PTP_ERR_CUDA, "batched kernel does not fit on an SM");
    // one CTA per solve: every resident CTA may own a solve; teams: one CTA per SM, num_sms / team solves in flight
    const u32 max_slots = team > 1 ? (u32)m->num_sms / team : (u32)(m->num_sms * per_sm);
    const u32 want = std::max<u32>(1u, std::min<u32>(max_slots, B));
    if (m->bt_slots < want || m->bt_scap < max_s || m->bt_team != team) {
        free_list(m, m->bt_allocs);
        m->bt_slots = 0;
        m->bt_src = m->bt_off = m->bt_rows = nullptr;
        m->bt_src_cap = m->bt_off_cap = m->bt_rows_cap = 0;
        const u64 V = m->V, scap = std::max<u64>(std::max<u64>(max_s, m->bt_scap), 16), N = V + scap;
        const u64 ovfn = std::max<u64>(m->ovf_total, 4), qb = (N + 1 + 15) / 16 * 16, tileb = 4 * 4096;
        const u64 per_slot = 8 * V + 4 * N + 4 * V + 4 * (V + 2) + tileb + 4 * sizeof(R) * (N + 1) + 4 * GL * N + 4 * ovfn +
                             2 * sizeof(R) * (N + 1) + 8 * C_COUNT + 4 * N + 2 * qb + 4096;
        // as many workspaces as the batch can use and the device can hold (the rows staging buffer comes on top)
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        const u64 usable = free_b > sizeof(R) * rows_elems + (256ull << 20) ? free_b - sizeof(R) * rows_elems - (256ull << 20) : 0;
        const u32 fit = (u32)std::min<u64>(usable / per_slot, 1u << 20);
        if (fit < 1) return fail(PTP_ERR_CUDA, "not enough device memory for one batched-solve workspace");
        const u32 slots = std::min(want, fit);
        std::vector<Work<R>> hw(slots);
        auto &tr = m->bt_allocs;
        char *b_key, *b_sorted, *b_inv, *b_limits, *b_tile, *b_pos, *b_ring, *b_ovf, *b_d0, *b_d1, *b_ctrl, *b_wl, *b_q0, *b_q1;
#define BT(ptr, per)                                                                     \
    if ((rc = dev_alloc(m, (void **)&(ptr), (u64)(per) * slots, &tr)) != PTP_OK) return rc;
        BT(b_key, 8 * V)
        BT(b_sorted, 4 * N)
        BT(b_inv, 4 * V)
        BT(b_limits, 4 * (V + 2))
        BT(b_tile, tileb)
        BT(b_pos, 4 * sizeof(R) * (N + 1))
        BT(b_ring, 4 * GL * N)
        BT(b_ovf, 4 * ovfn)
        BT(b_d0, sizeof(R) * (N + 1))
        BT(b_d1, sizeof(R) * (N + 1))
        BT(b_ctrl, 8 * C_COUNT)
        BT(b_wl, 4 * N)
        BT(b_q0, qb)
        BT(b_q1, qb)
#undef BT
        for (u32 s = 0; s < slots; s++) {
            Work<R> &w = hw[s];
            w.key = (ull *)(b_key + (u64)s * 8 * V);
            w.sorted = (u32 *)(b_sorted + (u64)s * 4 * N);
            w.inv = (u32 *)(b_inv + (u64)s * 4 * V);
            w.limits = (u32 *)(b_limits + (u64)s * 4 * (V + 2));
            w.tile_sum = (u32 *)(b_tile + (u64)s * tileb);
            w.posS = (typename Ops<R>::vec4 *)(b_pos + (u64)s * 4 * sizeof(R) * (N + 1));
            w.ringS = (u32 *)(b_ring + (u64)s * 4 * GL * N);
            w.ovfS = (u32 *)(b_ovf + (u64)s * 4 * ovfn);
            w.dist[0] = (R *)(b_d0 + (u64)s * sizeof(R) * (N + 1));
            w.dist[1] = (R *)(b_d1 + (u64)s * sizeof(R) * (N + 1));
            w.cl[0] = w.cl[1] = nullptr;
            w.toplesets = nullptr;
            w.ctrl = (ull *)(b_ctrl + (u64)s * 8 * C_COUNT);
            w.wl = (u32 *)(b_wl + (u64)s * 4 * N);
            w.dirty[0] = (unsigned char *)(b_q0 + (u64)s * qb);
            w.dirty[1] = (unsigned char *)(b_q1 + (u64)s * qb);
        }
        if ((rc = dev_alloc(m, &m->bt_works, sizeof(Work<R>) * slots, &tr))) return rc;
        if ((rc = dev_alloc(m, &m->bt_queue, 128, &tr))) return rc;
        if ((rc = dev_alloc(m, &m->bt_help, sizeof(HelpDesc) * slots + 64, &tr))) return rc;
        if ((rc = dev_alloc(m, &m->bt_bars, 128 * (u64)slots + 128, &tr))) return rc;
        CK(cudaMemcpy(m->bt_works, hw.data(), sizeof(Work<R>) * slots, cudaMemcpyHostToDevice));
        m->bt_ctrl = b_ctrl;
        m->bt_slots = slots;
        m->bt_team = team;
        m->bt_grid = team > 1 ? slots * team : (u32)(m->num_sms * per_sm);
        m->bt_scap = scap;
    }
    if (m->bt_src_cap < n_src) {
        if ((rc = dev_alloc(m, &m->bt_src, 4 * n_src, &m->bt_allocs))) return rc;
        m->bt_src_cap = n_src;
    }
    if (m->bt_off_cap < n_off) {
        if ((rc = dev_alloc(m, &m->bt_off, 8 * n_off, &m->bt_allocs))) return rc;
        m->bt_off_cap = n_off;
    }
    if (m->bt_rows_cap < rows_elems) {
        if ((rc = dev_alloc(m, &m->bt_rows, sizeof(R) * rows_elems, &m->bt_allocs))) return rc;
        m->bt_rows_cap = rows_elems;
    }
    return PTP_OK;
}

template <class R>
int batched_impl(ptp_mesh *m, const u32 *sources, const u64 *offsets, u32 B, u64 n_src, R *rows, int on_device, void *stream_,
                 ptp_stats_t *st)
{
    int rc;
    CK(cudaSetDevice(m->device));
    if (B == 0) return fail(PTP_ERR_INVALID, "empty batch");
    if (!rows) return fail(PTP_ERR_INVALID, "rows is null");
    if (!offsets && n_src != B) return fail(PTP_ERR_INVALID, "without offsets n_sources must equal n_batch");
    if ((rc = check_sources(m, sources, n_src))) return rc;
    u64 max_s = 1;
    if (offsets) {
        if (offsets[0] != 0 || offsets[B] != n_src) return fail(PTP_ERR_INVALID, "offsets must start at 0 and end at n_sources");
        for (u32 b = 0; b < B; b++) {
            if (offsets[b + 1] <= offsets[b]) return fail(PTP_ERR_INVALID, "every source set must be non-empty");
            max_s = std::max<u64>(max_s, offsets[b + 1] - offsets[b]);
        }
    }
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : m->stream;
    // host rows: stage through a device buffer of at most ~8 GiB, chunk by chunk
    u64 chunk = B;
    if (!on_device) {
        const u64 budget = (u64)std::max<long>(1, opt("rows_chunk_mb")) << 20;
        chunk = std::max<u64>(1, std::min<u64>(B, budget / (sizeof(R) * m->V)));
    }
    if ((rc = ensure_batch<R>(m, max_s, n_src, offsets ? B + 1 : 0, on_device ? 0 : chunk * m->V, B))) return rc;
    // Geometry table ("geo" option): the mesh-constant half of update_step (3 of 4 divisions, 2 of 3 square roots) read from
    // a table shared by every solve instead of recomputed. Bit-exact, 25 % fewer instructions per relaxation, and yet
    // measured SLOWER in float (C5: 229-234 vs 239-241 sources/s: +128 B of DRAM traffic per relaxation on a kernel that
    // is bound by memory latency, not issue slots) and only +4 % in double, so it stays opt-in.
    const bool use_geo = opt("geo") != 0;
    bool geo_ok = false;
    if (use_geo && (rc = ensure_geo<R>(m, stream, &geo_ok))) return rc;
    CK(cudaMemcpyAsync(m->bt_src, sources, 4 * n_src, cudaMemcpyHostToDevice, stream));
    if (offsets) CK(cudaMemcpyAsync(m->bt_off, offsets, 8 * (u64)(B + 1), cudaMemcpyHostToDevice, stream));
    ull *queue = (ull *)m->bt_queue;
    CK(cudaMemsetAsync(queue, 0, 128, stream));
    CK(cudaMemsetAsync(m->bt_ctrl, 0, 8 * C_COUNT * (u64)m->bt_slots, stream));
    CK(cudaEventRecord(m->ev[0], stream));
    // causal skip ("causal" option): needs the per-mesh safe flags and 28-bit ranks (three flag bits per ring entry, 32-bit record offsets); not combined with the geometry table
    // (whose records are indexed by the un-rotated ring slots)
    const bool causal = opt("causal") != 0 && !use_geo && m->V + m->bt_scap + 2 < (1ull << 28);
    if (causal && (rc = ensure_safe<R>(m, stream))) return rc;
    MeshView<R> mv = mesh_view<R>(m);
    if (!use_geo) mv.geo = nullptr; // (the single-solve path may have built the table; the batched kernel uses it on request only)
    u64 launches = 0;
    const u32 team = m->bt_team;
    const bool prof = opt("profile_range") != 0; // the whole batch as one profiler range (ncu --replay-mode app-range)
    if (prof) { CK(cudaStreamSynchronize(stream)); cudaProfilerStart(); }
    for (u64 first = 0; first < B; first += chunk) {
        const u32 nb = (u32)std::min<u64>(chunk, B - first);
        R *dst = on_device ? rows + first * m->V : (R *)m->bt_rows;
        CK(cudaMemsetAsync(queue, 0, 8, stream));
        const Work<R> *works = (const Work<R> *)m->bt_works;
        const u32 *d_src = (const u32 *)m->bt_src;
        const ull *d_off = offsets ? (const ull *)m->bt_off : nullptr;
        u32 first32 = (u32)first, nb32 = nb, sent = (u32)(m->V + m->bt_scap);
        ull *totals = queue + 1;
        unsigned char *row_done = nullptr;
        if (!on_device && opt("stream_rows") != 0 && nb >= 32) {
            if (m->bt_done_cap < nb) {
                cudaFree(m->bt_done);
                if (m->bt_hdone) cudaFreeHost(m->bt_hdone);
                m->bt_done = m->bt_hdone = nullptr;
                m->bt_done_cap = 0;
                CK(cudaMalloc(&m->bt_done, chunk));
                CK(cudaHostAlloc(&m->bt_hdone, chunk, cudaHostAllocDefault));
                m->bt_done_cap = chunk;
            }
            if (!m->mg_stream) CK(cudaStreamCreateWithFlags(&m->mg_stream, cudaStreamNonBlocking));
            row_done = (unsigned char *)m->bt_done;
            CK(cudaMemsetAsync(row_done, 0, nb, stream));
        }
        if (team > 1) {
            // a team of CTAs per solve; the teams synchronise with software grid barriers, so every CTA must be resident:
            // cooperative launch (one CTA per SM)
            ull *bars = (ull *)m->bt_bars;
            u32 tsz = team;
            const u32 teams = std::min<u32>(m->bt_slots, nb);
            CK(cudaMemsetAsync(bars, 0, 128 * (u64)m->bt_slots + 128, stream));
            void *fn = mv.geo ? (void *)k_batched_teams<R, true, false>
                              : (causal ? (void *)k_batched_teams<R, false, true> : (void *)k_batched_teams<R, false, false>);
            void *args[] = {&mv, &works, &d_src, &d_off, &first32, &nb32, &dst, &sent, &queue, &totals, &bars, &tsz, &row_done};
            CK(cudaLaunchCooperativeKernel(fn, dim3(teams * team), dim3(BatchCfg<R>::BLOCK), args, 0, stream));
            m->last_kernel = sizeof(R) == 8 ? "k_batched_teams<double>" : "k_batched_teams<float>";
        } else {
            // elastic mode ("elastic" option): every resident CTA is launched; those without a solve help from the start
            const bool elastic = opt("elastic") != 0;
            HelpDesc *descs = elastic ? (HelpDesc *)m->bt_help : nullptr;
            u32 *counters = (u32 *)((char *)m->bt_help + sizeof(HelpDesc) * m->bt_slots);
            if (elastic) CK(cudaMemsetAsync(m->bt_help, 0, sizeof(HelpDesc) * m->bt_slots + 64, stream));
            const u32 grid = elastic ? m->bt_grid : std::min<u32>(m->bt_slots, nb);
            auto kern = mv.geo ? k_batched<R, true, false> : (causal ? k_batched<R, false, true> : k_batched<R, false, false>);
#if PTP_MAXL1
            // the kernel uses < 1 KB of shared memory and lives on L1 hits of its gathers: ask for the largest L1 split
            cudaFuncSetAttribute((const void *)kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
#endif
            kern<<<grid, BatchCfg<R>::BLOCK, 0, stream>>>(mv, works, d_src, d_off, first32, nb32, dst, sent, queue, totals, descs, counters,
                                                         m->bt_slots, row_done);
            CK(cudaGetLastError());
            m->last_kernel = sizeof(R) == 8 ? "k_batched<double>" : "k_batched<float>";
        }
        launches++;
        if (!on_device && row_done) {
            // Rows leave for the host while the kernel is still solving the rest: a second stream polls the completion flags
            // and copies every finished prefix of the batch (solves finish roughly in queue order), so that only the last
            // wave's rows are copied after the kernel has ended.
            unsigned char *h_done = (unsigned char *)m->bt_hdone;
            u64 copied = 0;
            bool kernel_done = false;
            while (copied < nb) {
                if (!kernel_done) {
                    if (cudaStreamQuery(stream) == cudaSuccess) kernel_done = true;
                    else cudaGetLastError(); // cudaErrorNotReady is not an error: do not leave it behind for a later check
                }
                u64 upto = nb;
                if (!kernel_done) {
                    CK(cudaMemcpyAsync(h_done, row_done, nb, cudaMemcpyDeviceToHost, m->mg_stream));
                    CK(cudaStreamSynchronize(m->mg_stream));
                    upto = copied;
                    while (upto < nb && h_done[upto]) upto++;
                }
                // (copies of at least 16 rows, or the rest: each copy costs a launch)
                if (upto > copied && (kernel_done || upto - copied >= 16 || upto == nb)) {
                    CK(cudaMemcpyAsync(rows + (first + copied) * m->V, (R *)m->bt_rows + copied * m->V, sizeof(R) * (upto - copied) * m->V,
                                       cudaMemcpyDeviceToHost, m->mg_stream));
                    copied = upto;
                } else if (!kernel_done) {
                    std::this_thread::sleep_for(std::chrono::microseconds(300));
                }
            }
            CK(cudaStreamSynchronize(stream));      // a launch failure surfaces here
            CK(cudaStreamSynchronize(m->mg_stream)); // the staging buffer is reused by the next chunk
        } else if (!on_device) {
            CK(cudaMemcpyAsync(rows + first * m->V, m->bt_rows, sizeof(R) * (u64)nb * m->V, cudaMemcpyDeviceToHost, stream));
        }
    }
    CK(cudaEventRecord(m->ev[1], stream));
    ull tot[16];
    CK(cudaMemcpyAsync(tot, queue, 128, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (prof) cudaProfilerStop();
    if (tot[10]) {
        static const char *what[] = {"", "grid barrier", "sweep team waiting for toplesets / rows", "layout warp waiting for the BFS",
                                     "layout warp waiting for its turn to publish", "BFS cluster waiting for its tables", "elastic relax chunks"};
        return fail(PTP_ERR_CUDA, std::string("device watchdog in a batched solve: a wait did not complete (") + (tot[10] < 7 ? what[tot[10]] : "?") +
                                      "); rows discarded");
    }
#ifdef PTP_COUNT_TRI
    {
        ull c[4];
        cudaMemcpyFromSymbol(c, g_tri_cnt, sizeof c);
        fprintf(stderr, "[ptp] causal: relaxations %llu, triangles %llu, evaluated %llu (%.1f %%), warp-level lane-slots %llu (%.1f %% of triangles)\n",
                c[3], c[0], c[1], 100.0 * c[1] / (double)c[0], c[2], 100.0 * c[2] / (double)c[0]);
        ull z[4] = {0, 0, 0, 0};
        cudaMemcpyToSymbol(g_tri_cnt, z, sizeof z);
    }
#endif
    if (st) {
        st->iterations = tot[1];
        st->vertex_updates = tot[2];
        st->max_window = tot[3];
        st->n_levels = tot[4];
        st->n_reached = tot[5];
        st->relaxations = tot[6];
        st->gpu_launches = launches;
        // CTA-time spent in BFS + layout, and in the sweep, averaged over the CTAs that ran (device timers)
        const double ctas = (double)std::min<u64>(m->bt_slots, B); // (teams: per team)
        if (opt("debug"))
            fprintf(stderr, "[ptp] batched mean per-CTA ms: bfs %.1f layout %.1f sweep+scatter %.1f\n", tot[7] * 1e-6 / ctas,
                    tot[8] * 1e-6 / ctas, tot[9] * 1e-6 / ctas);
        st->ms_toplesets = (double)(tot[7] + tot[8]) * 1e-6 / ctas;
        st->ms_solve = (double)tot[9] * 1e-6 / ctas;
        st->ms_total = ev_ms(m->ev[0], m->ev[1]);
    }
    return PTP_OK;
}

template <class R>
int fps_impl(ptp_mesh *m, u32 *samples, u32 n_initial, u32 n_total, R radio, u32 *n_out, R *max_dist, ptp_stats_t *st)
{
    int rc;
    CK(cudaSetDevice(m->device));
    if (!samples || n_initial == 0) return fail(PTP_ERR_INVALID, "need at least one initial sample");
    if ((rc = check_sources(m, samples, n_initial))) return rc;
    // src/cuda/geodesics_ptp.cu:125: n >= n_vertices is clamped to n_vertices / 2
    u64 want = n_total;
    if (want >= m->V) want = m->V >> 1;
    if ((rc = ensure_workspace<R>(m, std::max<u64>(want, n_initial), false, false))) return rc;
    u32 n = n_initial;
    R maxd = (R)INFINITY;
    u64 launches = 0;
    for (int attempt = 0;; attempt++) {
        n = n_initial;
        maxd = (R)INFINITY;
        CK(cudaMemcpyAsync(m->w_src, samples, 4ull * n, cudaMemcpyHostToDevice, m->stream));
        CK(cudaMemsetAsync(m->w_maxval, 0, 16, m->stream)); // [0] last maximum, [1] sticky watchdog word of the whole run
        CK(cudaEventRecord(m->ev[3], m->stream));
        // the reference loops `n -= samples.size(); while(n-- && max_dist > radio)` (:127-148)
        while (n < want && maxd > radio) {
            CK(cudaMemsetAsync(m->w_ctrl, 0, 8 * C_COUNT, m->stream));
            if ((rc = pipeline<R>(m, n, false, 0))) return rc;
            k_argmax_append<R><<<1, 1024, 0, m->stream>>>((const R *)m->w_out, (u32)m->V, (u32 *)m->w_src, n, (R *)m->w_maxval,
                                                          (const ull *)m->w_ctrl, (ull *)m->w_maxval + 1);
            CK(cudaGetLastError());
            launches += m->last_two ? 3 : 2;
            n++;
            const bool last = n >= want;
            if (radio > 0 || last) {
                // the reference reads the maximum back only when it needs it (:143-144); the watchdog word comes with it
                ull hv[2] = {0, 0};
                CK(cudaMemcpyAsync(hv, m->w_maxval, 16, cudaMemcpyDeviceToHost, m->stream));
                CK(cudaStreamSynchronize(m->stream));
                memcpy(&maxd, hv, sizeof(R));
                if (hv[1]) break; // a solve of this run was ended by the watchdog: its arg-max is void
            }
        }
        CK(cudaEventRecord(m->ev[2], m->stream));
        ull hv[2] = {0, 0};
        CK(cudaMemcpyAsync(hv, m->w_maxval, 16, cudaMemcpyDeviceToHost, m->stream));
        CK(cudaStreamSynchronize(m->stream));
        if (hv[1]) {
            if (m->last_two && attempt == 0) { // the two-kernel solve did not get both kernels resident: one launch from now on
                m->two_failed = true;
                continue;
            }
            return fail(PTP_ERR_CUDA, "device watchdog: a wait inside a farthest-point-sampling solve did not complete; samples discarded");
        }
        break;
    }
    CK(cudaMemcpyAsync(samples, m->w_src, 4ull * n, cudaMemcpyDeviceToHost, m->stream));
    if ((rc = fetch_ctrl(m))) return rc;
    if (n_out) *n_out = n;
    if (max_dist) *max_dist = maxd;
    fill_stats(m, st, launches, 0, 0, ev_ms(m->ev[3], m->ev[2]));
    if (st) st->ms_solve = st->ms_total;
    return PTP_OK;
}

template <class R> int update_positions(ptp_mesh *m, const R *GT)
{
    CK(cudaSetDevice(m->device));
    if (!GT) return fail(PTP_ERR_INVALID, "GT is null");
    void *d_gt = nullptr;
    CK(cudaMalloc(&d_gt, sizeof(R) * 3 * m->V));
    auto run = [&]() -> int {
        CK(cudaMemcpyAsync(d_gt, GT, sizeof(R) * 3 * m->V, cudaMemcpyHostToDevice, m->stream));
        k_pad_gt<R><<<(unsigned)((m->V * 4 + 255) / 256), 256, 0, m->stream>>>((const R *)d_gt, (R *)m->GT4, (u32)m->V);
        CK(cudaGetLastError());
        if (m->safe8) {
            k_safe_build<R><<<(unsigned)((m->V + 127) / 128), 128, 0, m->stream>>>((const typename Ops<R>::vec4 *)m->GT4, m->ring8, (u32)m->V, m->safe8,
                                                                                 (u32)(m->safe_sign & 1), (u32)(m->safe_sign >> 1));
            CK(cudaGetLastError());
        }
        if (m->geo) { // the geometry table depends on the positions
            k_geo_build<R><<<(unsigned)((m->V + 127) / 128), 128, 0, m->stream>>>((const typename Ops<R>::vec4 *)m->GT4, m->ring8, (u32)m->V,
                                                                              (typename Ops<R>::vec4 *)m->geo);
            CK(cudaGetLastError());
        }
        CK(cudaStreamSynchronize(m->stream));
        return PTP_OK;
    };
    const int rc = run();
    cudaFree(d_gt);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// Batched solves over several devices of one process (ptp_solve_batched_multi_*). gproshan is a single C++ process
// (the caller shape is src/sampling.cpp:23-34): one host thread per device drives that device's shard through
// batched_impl; with host rows every device copies its rows straight into the caller's matrix, with device rows the
// shards travel to the root device over NVLink with NCCL (grouped ncclSend / ncclRecv), piece by piece, while the next
// piece is being solved. NCCL is loaded at run time (dlopen: the process may already hold one, e.g. PyTorch's).

struct Nccl {
    typedef void *comm_t;
    int (*CommInitAll)(comm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};

Nccl &nccl()
{
    static Nccl n = [] {
        Nccl x;
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { x.why = std::string("libnccl.so.2 cannot be loaded: ") + dlerror(); return x; }
        auto sym = [&](const char *name) { return dlsym(h, name); };
        x.CommInitAll = (decltype(x.CommInitAll))sym("ncclCommInitAll");
        x.CommDestroy = (decltype(x.CommDestroy))sym("ncclCommDestroy");
        x.Send = (decltype(x.Send))sym("ncclSend");
        x.Recv = (decltype(x.Recv))sym("ncclRecv");
        x.GroupStart = (decltype(x.GroupStart))sym("ncclGroupStart");
        x.GroupEnd = (decltype(x.GroupEnd))sym("ncclGroupEnd");
        x.GetErrorString = (decltype(x.GetErrorString))sym("ncclGetErrorString");
        x.ok = x.CommInitAll && x.CommDestroy && x.Send && x.Recv && x.GroupStart && x.GroupEnd && x.GetErrorString;
        if (!x.ok) x.why = "libnccl.so.2 lacks ncclCommInitAll / ncclSend / ncclRecv";
        return x;
    }();
    return n;
}

// communicators are kept per device list (creating them costs ~100 ms)
struct CommSet { std::vector<int> devs; std::vector<Nccl::comm_t> comms; };
std::mutex g_comm_mu;
std::vector<CommSet> g_comm_sets;

int get_comms(const std::vector<int> &devs, std::vector<Nccl::comm_t> *out)
{
    Nccl &n = nccl();
    if (!n.ok) return fail(PTP_ERR_CUDA, n.why);
    std::lock_guard<std::mutex> lock(g_comm_mu);
    for (CommSet &c : g_comm_sets)
        if (c.devs == devs) { *out = c.comms; return PTP_OK; }
    CommSet c;
    c.devs = devs;
    c.comms.resize(devs.size());
    const int rc = n.CommInitAll(c.comms.data(), (int)devs.size(), devs.data());
    if (rc != 0) return fail(PTP_ERR_CUDA, std::string("ncclCommInitAll: ") + n.GetErrorString(rc));
    g_comm_sets.push_back(c);
    *out = c.comms;
    return PTP_OK;
}

struct Rendezvous { // reusable barrier for the device threads + the coordinator
    std::mutex mu;
    std::condition_variable cv;
    int n, waiting = 0, generation = 0;
    explicit Rendezvous(int n_) : n(n_) {}
    void wait()
    {
        std::unique_lock<std::mutex> lk(mu);
        const int g = generation;
        if (++waiting == n) { waiting = 0; generation++; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != g; });
    }
};

template <class R>
int batched_multi_impl(ptp_mesh *const *ms, int G, const u32 *sources, const u64 *offsets, u32 B, u64 n_src, R *rows, int on_device,
                       ptp_stats_t *st)
{
    if (!ms || G < 1) return fail(PTP_ERR_INVALID, "need at least one mesh handle");
    if (B == 0 || !rows || !sources) return fail(PTP_ERR_INVALID, "empty batch or null pointer");
    if (!offsets && n_src != B) return fail(PTP_ERR_INVALID, "without offsets n_sources must equal n_batch");
    std::vector<int> devs(G);
    for (int g = 0; g < G; g++) {
        if (!ms[g]) return fail(PTP_ERR_INVALID, "null mesh handle");
        if (ms[g]->real_size != (int)sizeof(R)) return fail(PTP_ERR_INVALID, "mesh precision does not match this entry point");
        if (ms[g]->V != ms[0]->V || ms[g]->H != ms[0]->H) return fail(PTP_ERR_INVALID, "the handles must hold the same mesh");
        devs[g] = ms[g]->device;
        for (int k = 0; k < g; k++)
            if (devs[k] == devs[g]) return fail(PTP_ERR_INVALID, "one handle per device: two handles share a device");
    }
    const u64 V = ms[0]->V;
    const bool gather = on_device && G > 1;
    std::vector<Nccl::comm_t> comms;
    int rc;
    if (gather && (rc = get_comms(devs, &comms))) return rc;
    // measured on 8 x B200, 128 sources per GPU: 2 pieces of 64 -> 1 343 sources/s, 1 piece -> as fast as host rows (1 754)
    long want_pieces = opt("gather_chunks");
    if (want_pieces <= 0) want_pieces = (long)((u64)B / (u64)G / (2ull * (u64)std::max(1, ms[0]->num_sms)));
    const int pieces = gather ? (int)std::max<long>(1, std::min<long>(want_pieces, 64)) : 1;

    // contiguous block partition of the batch; the first B % G devices get one more
    std::vector<u64> lo(G + 1, 0);
    for (int g = 0; g < G; g++) lo[g + 1] = lo[g] + B / G + ((u64)g < B % G ? 1 : 0);

    std::vector<int> codes(G, PTP_OK);
    std::vector<std::string> errs(G);
    std::vector<ptp_stats_t> stats(G);
    Rendezvous meet(G + 1);
    const auto t_begin = std::chrono::steady_clock::now();

    auto worker = [&](int g) {
        ptp_mesh *m = ms[g];
        std::lock_guard<std::mutex> lock(m->mu);
        memset(&stats[g], 0, sizeof(ptp_stats_t));
        auto step = [&](int code) { if (code != PTP_OK && codes[g] == PTP_OK) { codes[g] = code; errs[g] = g_err; } return code; };
        const u64 nb = lo[g + 1] - lo[g];
        R *dst_base = nullptr;
        if (cudaSetDevice(m->device) != cudaSuccess) step(fail(PTP_ERR_CUDA, "cudaSetDevice failed"));
        if (codes[g] == PTP_OK && nb) {
            if (!on_device) dst_base = rows + lo[g] * V;          // host matrix: this device's rows land in place
            else if (g == 0) dst_base = rows + lo[g] * V;          // root device: in place in the assembled matrix
            else {                                                  // peer: staged here, then sent to the root
                if (m->mg_rows_cap < nb * V) {
                    cudaFree(m->mg_rows);
                    m->mg_rows = nullptr;
                    m->mg_rows_cap = 0;
                    if (cudaMalloc(&m->mg_rows, sizeof(R) * nb * V) != cudaSuccess) { cudaGetLastError(); step(fail(PTP_ERR_CUDA, "cudaMalloc of the row staging buffer failed")); }
                    else m->mg_rows_cap = nb * V;
                }
                dst_base = (R *)m->mg_rows;
            }
        }
        if (gather && codes[g] == PTP_OK) {
            if (!m->mg_stream && cudaStreamCreateWithFlags(&m->mg_stream, cudaStreamNonBlocking) != cudaSuccess) step(fail(PTP_ERR_CUDA, "stream creation failed"));
            if (!m->mg_ev && cudaEventCreateWithFlags(&m->mg_ev, cudaEventDisableTiming) != cudaSuccess) step(fail(PTP_ERR_CUDA, "event creation failed"));
        }
        for (int c = 0; c < pieces; c++) {
            const u64 a = lo[g] + nb * c / pieces, b = lo[g] + nb * (c + 1) / pieces;
            if (codes[g] == PTP_OK && b > a) {
                ptp_stats_t s1;
                memset(&s1, 0, sizeof s1);
                std::vector<u64> off;
                const u32 *src = sources + a;
                u64 ns = b - a;
                if (offsets) {
                    off.resize(b - a + 1);
                    for (u64 k = a; k <= b; k++) off[k - a] = offsets[k] - offsets[a];
                    src = sources + offsets[a];
                    ns = offsets[b] - offsets[a];
                }
                step(batched_impl<R>(m, src, offsets ? off.data() : nullptr, (u32)(b - a), ns, dst_base + (a - lo[g]) * V, on_device, nullptr, &s1));
                stats[g].iterations += s1.iterations;
                stats[g].vertex_updates += s1.vertex_updates;
                stats[g].relaxations += s1.relaxations;
                stats[g].n_levels += s1.n_levels;
                stats[g].n_reached += s1.n_reached;
                stats[g].gpu_launches += s1.gpu_launches;
                stats[g].max_window = std::max(stats[g].max_window, s1.max_window);
                stats[g].ms_toplesets += s1.ms_toplesets;
                stats[g].ms_solve += s1.ms_solve;
                stats[g].ms_total += s1.ms_total;
            }
            if (gather) {
                meet.wait(); // piece c is solved on every device (batched_impl returns after its stream has drained)
                meet.wait(); // the coordinator has enqueued the transfer of piece c; go on with piece c + 1
            }
        }
    };

    std::vector<std::thread> threads;
    for (int g = 0; g < G; g++) threads.emplace_back(worker, g);
    int comm_rc = PTP_OK;
    std::string comm_err;
    if (gather) {
        Nccl &n = nccl();
        const int dt = sizeof(R) == 4 ? 7 : 8; // ncclFloat32 / ncclFloat64
        for (int c = 0; c < pieces; c++) {
            meet.wait();
            bool all_ok = comm_rc == PTP_OK;
            for (int g = 0; g < G; g++) all_ok = all_ok && codes[g] == PTP_OK;
            if (all_ok) {
                int r = n.GroupStart();
                for (int g = 1; g < G && r == 0; g++) {
                    const u64 nb = lo[g + 1] - lo[g];
                    const u64 a = nb * c / pieces, b = nb * (c + 1) / pieces;
                    if (b == a) continue;
                    r = n.Send((const R *)ms[g]->mg_rows + a * V, (b - a) * V, dt, 0, comms[g], ms[g]->mg_stream);
                    if (r == 0) r = n.Recv(rows + (lo[g] + a) * V, (b - a) * V, dt, g, comms[0], ms[0]->mg_stream);
                }
                const int r2 = n.GroupEnd();
                if (r == 0) r = r2;
                if (r != 0) { comm_rc = PTP_ERR_CUDA; comm_err = std::string("NCCL send/recv: ") + n.GetErrorString(r); }
            }
            meet.wait();
        }
    }
    for (std::thread &t : threads) t.join();
    if (gather && comm_rc == PTP_OK) {
        for (int g = 0; g < G; g++) {
            cudaSetDevice(ms[g]->device);
            if (ms[g]->mg_stream && cudaStreamSynchronize(ms[g]->mg_stream) != cudaSuccess) {
                comm_rc = PTP_ERR_CUDA;
                comm_err = "the row transfer to the root device failed";
            }
        }
        cudaSetDevice(ms[0]->device);
    }
    for (int g = 0; g < G; g++)
        if (codes[g] != PTP_OK) return fail(codes[g], "device " + std::to_string(devs[g]) + ": " + errs[g]);
    if (comm_rc != PTP_OK) return fail(comm_rc, comm_err);
    if (st) {
        memset(st, 0, sizeof *st);
        for (int g = 0; g < G; g++) {
            st->iterations += stats[g].iterations;
            st->vertex_updates += stats[g].vertex_updates;
            st->relaxations += stats[g].relaxations;
            st->n_levels += stats[g].n_levels;
            st->n_reached += stats[g].n_reached;
            st->gpu_launches += stats[g].gpu_launches;
            st->max_window = std::max(st->max_window, stats[g].max_window);
            st->ms_toplesets = std::max(st->ms_toplesets, stats[g].ms_toplesets);
            st->ms_solve = std::max(st->ms_solve, stats[g].ms_total); // slowest device: kernel time of its shard
        }
        st->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    }
    return PTP_OK;
}

} // namespace

// ------------------------------------------------------------------------------------------------
// extern "C"

extern "C" {

const char *ptp_last_error(void) { return g_err.c_str(); }

const char *ptp_version(void) { return "ptp_b200 0.1 (sm_100a)"; }

int ptp_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void *ptp_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) {
        g_err = "cudaHostAlloc failed";
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void ptp_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int ptp_mesh_create_f32(const float *GT, const uint32_t *VT, const uint32_t *OT, const uint32_t *EVT, uint64_t V, uint64_t H,
                        int device, ptp_mesh_t **out)
{
    return mesh_create<float>(GT, VT, OT, EVT, V, H, device, out);
}
int ptp_mesh_create_f64(const double *GT, const uint32_t *VT, const uint32_t *OT, const uint32_t *EVT, uint64_t V, uint64_t H,
                        int device, ptp_mesh_t **out)
{
    return mesh_create<double>(GT, VT, OT, EVT, V, H, device, out);
}

int ptp_che_build(const uint32_t *VT, uint64_t V, uint64_t H, uint32_t *OT, uint32_t *EVT, int device, int *manifold, double *ms)
{
    if (!VT || !OT || !EVT) return fail(PTP_ERR_INVALID, "null table");
    if (V == 0 || H == 0 || H % 3 != 0 || V >= 0x7FFFFFF0ull || H >= 0x7FFFFFF0ull) return fail(PTP_ERR_INVALID, "bad sizes");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(PTP_ERR_NO_DEVICE, "no such CUDA device");
    CK(cudaSetDevice(device));
    u32 *d_vt = nullptr, *d_ot = nullptr, *d_evt = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto run = [&]() -> int {
        CK(cudaMalloc(&d_vt, 4 * H));
        CK(cudaMalloc(&d_ot, 4 * H));
        CK(cudaMalloc(&d_evt, 4 * V));
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaMemcpy(d_vt, VT, 4 * H, cudaMemcpyHostToDevice));
        CK(cudaEventRecord(e0, nullptr));
        bool mf = true;
        int rc = che_build_device(d_vt, V, H, d_ot, d_evt, nullptr, &mf);
        if (rc) return rc;
        CK(cudaEventRecord(e1, nullptr));
        CK(cudaMemcpy(OT, d_ot, 4 * H, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(EVT, d_evt, 4 * V, cudaMemcpyDeviceToHost));
        if (manifold) *manifold = mf ? 1 : 0;
        if (ms) { float t = 0; cudaEventElapsedTime(&t, e0, e1); *ms = t; }
        return PTP_OK;
    };
    const int rc = run();
    cudaFree(d_vt); cudaFree(d_ot); cudaFree(d_evt);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

// verification helper (not part of the reference interface): see k_dbg_inv_gram
int ptp_debug_inv_gram_check(uint64_t n, uint64_t seed, int real_size, uint64_t *mismatches, uint64_t *shared_path, double *samples160)
{
    ull *out = nullptr;
    double *smp = nullptr;
    if (real_size != 4 && real_size != 8) return fail(PTP_ERR_INVALID, "real_size must be 4 or 8");
    if (cudaMalloc(&out, 24) != cudaSuccess) { cudaGetLastError(); return fail(PTP_ERR_NO_DEVICE, "no CUDA device"); }
    if (samples160 && cudaMalloc(&smp, 160 * sizeof(double)) != cudaSuccess) { cudaFree(out); return fail(PTP_ERR_CUDA, "cudaMalloc"); }
    cudaMemset(out, 0, 24);
    if (smp) cudaMemset(smp, 0, 160 * sizeof(double));
    if (real_size == 4) k_dbg_inv_gram<float><<<1184, 256>>>((ull)n, (ull)seed, out, smp);
    else k_dbg_inv_gram<double><<<1184, 256>>>((ull)n, (ull)seed, out, smp);
    ull h[3] = {0, 0, 0};
    cudaError_t e = cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && smp) e = cudaMemcpy(samples160, smp, 160 * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(out);
    cudaFree(smp);
    if (e != cudaSuccess) return fail(PTP_ERR_CUDA, cudaGetErrorString(e));
    if (mismatches) *mismatches = h[0];
    if (shared_path) *shared_path = h[1];
    return PTP_OK;
}

// verification helper (not part of the reference interface): see k_dbg_sign_short
int ptp_debug_sign_short_check(uint64_t n, uint64_t seed, int real_size, uint64_t *disagreements, uint64_t *decided, uint64_t *flagged)
{
    ull *out = nullptr;
    if (real_size != 4 && real_size != 8) return fail(PTP_ERR_INVALID, "real_size must be 4 or 8");
    if (cudaMalloc(&out, 24) != cudaSuccess) { cudaGetLastError(); return fail(PTP_ERR_NO_DEVICE, "no CUDA device"); }
    cudaMemset(out, 0, 24);
    if (real_size == 4) k_dbg_sign_short<float><<<1184, 256>>>((ull)n, (ull)seed, out);
    else k_dbg_sign_short<double><<<1184, 256>>>((ull)n, (ull)seed, out);
    ull h[3] = {0, 0, 0};
    const cudaError_t e = cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
    cudaFree(out);
    if (e != cudaSuccess) return fail(PTP_ERR_CUDA, cudaGetErrorString(e));
    if (disagreements) *disagreements = h[0];
    if (decided) *decided = h[1];
    if (flagged) *flagged = h[2];
    return PTP_OK;
}

// verification helper (not part of the reference interface): see k_dbg_sqrt
int ptp_debug_sqrt_check(uint64_t *mismatches, uint64_t *tested)
{
    ull *out = nullptr;
    if (cudaMalloc(&out, 16) != cudaSuccess) { cudaGetLastError(); return fail(PTP_ERR_NO_DEVICE, "no CUDA device"); }
    cudaMemset(out, 0, 16);
    k_dbg_sqrt<<<1184, 256>>>(out);
    ull h[2] = {0, 0};
    const cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    cudaFree(out);
    if (e != cudaSuccess) return fail(PTP_ERR_CUDA, cudaGetErrorString(e));
    if (mismatches) *mismatches = h[0];
    if (tested) *tested = h[1];
    return PTP_OK;
}

// verification helper (not part of the reference interface): see k_dbg_two_sided
int ptp_debug_two_sided_check(uint64_t n, uint64_t seed, int real_size, uint64_t *violations, uint64_t *fired, uint64_t *flagged)
{
    ull *out = nullptr;
    if (real_size != 4 && real_size != 8) return fail(PTP_ERR_INVALID, "real_size must be 4 or 8");
    if (cudaMalloc(&out, 24) != cudaSuccess) { cudaGetLastError(); return fail(PTP_ERR_NO_DEVICE, "no CUDA device"); }
    cudaMemset(out, 0, 24);
    if (real_size == 4) k_dbg_two_sided<float><<<1184, 256>>>((ull)n, (ull)seed, out);
    else k_dbg_two_sided<double><<<1184, 256>>>((ull)n, (ull)seed, out);
    ull h[3] = {0, 0, 0};
    const cudaError_t e = cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
    cudaFree(out);
    if (e != cudaSuccess) return fail(PTP_ERR_CUDA, cudaGetErrorString(e));
    if (violations) *violations = h[0];
    if (fired) *fired = h[1];
    if (flagged) *flagged = h[2];
    return PTP_OK;
}

// measurement helper (not part of the reference interface): nanoseconds per grid barrier of `ctas` CTAs x `block` threads
double ptp_debug_barrier_ns(int ctas, int block, int n)
{
    ull *bar = nullptr;
    u32 *sink = nullptr;
    cudaEvent_t e0, e1;
    // block >= 100000 selects a measurement mode: block = mode * 100000 + threads (see k_dbg_barriers)
    u32 mode = (u32)(block / 100000);
    block %= 100000;
    if (cudaMalloc(&bar, 4096) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) return -1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    u32 nn = (u32)n;
    void *args[] = {&bar, &nn, &sink, &mode};
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaMemset(bar, 0, 4096);
        cudaEventRecord(e0);
        if (ctas < 0) { // hardware barrier of ONE thread-block cluster of -ctas CTAs
            cudaLaunchConfig_t cfg = {};
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)-ctas;
            at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
            cfg.gridDim = dim3((unsigned)-ctas);
            cfg.blockDim = dim3(block);
            cfg.attrs = at;
            cfg.numAttrs = 1;
            if (-ctas > 8) cudaFuncSetAttribute((void *)k_dbg_cluster_barriers, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            void *cargs[] = {&nn, &sink};
            if (cudaLaunchKernelExC(&cfg, (void *)k_dbg_cluster_barriers, cargs) != cudaSuccess) { cudaGetLastError(); best = -1; break; }
        } else
        if (cudaLaunchCooperativeKernel((void *)k_dbg_barriers, dim3(ctas), dim3(block), args, 0, nullptr) != cudaSuccess) { best = -1; break; }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    cudaFree(bar); cudaFree(sink); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best < 0 ? -1 : (double)best * 1e6 / n;
}

void ptp_mesh_destroy(ptp_mesh_t *m)
{
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    free_list(m, m->ws_allocs);
    free_list(m, m->bt_allocs);
    cudaFree(m->GT4);
    cudaFree(m->ring8);
    cudaFree(m->ovf);
    cudaFree(m->geo);
    cudaFree(m->safe8);
    cudaFree(m->mg_rows);
    cudaFree(m->bt_done);
    if (m->bt_hdone) cudaFreeHost(m->bt_hdone);
    if (m->mg_stream) cudaStreamDestroy(m->mg_stream);
    if (m->mg_ev) cudaEventDestroy(m->mg_ev);
    if (m->h_ctrl) cudaFreeHost(m->h_ctrl);
    for (auto &e : m->ev)
        if (e) cudaEventDestroy(e);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

const char *ptp_mesh_last_kernel(const ptp_mesh_t *m) { return m ? m->last_kernel : ""; }
uint64_t ptp_mesh_n_vertices(const ptp_mesh_t *m) { return m ? m->V : 0; }
uint64_t ptp_mesh_n_half_edges(const ptp_mesh_t *m) { return m ? m->H : 0; }
int ptp_mesh_real_size(const ptp_mesh_t *m) { return m ? m->real_size : 0; }
int ptp_mesh_device(const ptp_mesh_t *m) { return m ? m->device : -1; }
uint64_t ptp_mesh_device_bytes(const ptp_mesh_t *m) { return m ? m->bytes : 0; }

#define NEED(m, rs)                                                                     \
    if (!(m)) return fail(PTP_ERR_INVALID, "mesh is null");                             \
    if ((m)->real_size != (rs)) return fail(PTP_ERR_INVALID, "mesh precision does not match this entry point"); \
    std::lock_guard<std::mutex> lock_((m)->mu);

int ptp_set_option(const char *name, long value)
{
    if (!name) return fail(PTP_ERR_INVALID, "option name is null");
    options_init();
    std::lock_guard<std::mutex> lock(g_opt_mu);
    for (Option &o : g_options)
        if (!strcmp(o.name, name)) { o.value = value; return PTP_OK; }
    return fail(PTP_ERR_INVALID, std::string("unknown option: ") + name);
}

long ptp_get_option(const char *name)
{
    return name ? opt(name) : 0;
}

const char *ptp_option_name(int index) { return index >= 0 && index < N_OPTIONS ? g_options[index].name : nullptr; }
const char *ptp_option_doc(int index) { return index >= 0 && index < N_OPTIONS ? g_options[index].doc : nullptr; }

int ptp_mesh_update_positions_f32(ptp_mesh_t *m, const float *GT)
{
    NEED(m, 4) return update_positions<float>(m, GT);
}
int ptp_mesh_update_positions_f64(ptp_mesh_t *m, const double *GT)
{
    NEED(m, 8) return update_positions<double>(m, GT);
}

int ptp_toplesets(ptp_mesh_t *m, const uint32_t *sources, uint32_t S, uint32_t k, uint32_t *toplesets, uint32_t *sorted,
                  uint64_t scap, uint32_t *limits, uint64_t lcap, uint32_t *n_limits, ptp_stats_t *st)
{
    if (!m) return fail(PTP_ERR_INVALID, "mesh is null");
    std::lock_guard<std::mutex> lock_(m->mu);
    return m->real_size == 4 ? toplesets_impl<float>(m, sources, S, k, toplesets, sorted, scap, limits, lcap, n_limits, st)
                             : toplesets_impl<double>(m, sources, S, k, toplesets, sorted, scap, limits, lcap, n_limits, st);
}

int ptp_solve_f32(ptp_mesh_t *m, const uint32_t *sources, uint32_t S, const uint32_t *limits, uint32_t nl, const uint32_t *sorted,
                  float *dist, uint32_t *clusters, uint32_t fill, ptp_stats_t *st)
{
    NEED(m, 4) return solve_impl<float>(m, sources, S, limits, nl, sorted, dist, clusters, fill, st);
}
int ptp_solve_f64(ptp_mesh_t *m, const uint32_t *sources, uint32_t S, const uint32_t *limits, uint32_t nl, const uint32_t *sorted,
                  double *dist, uint32_t *clusters, uint32_t fill, ptp_stats_t *st)
{
    NEED(m, 8) return solve_impl<double>(m, sources, S, limits, nl, sorted, dist, clusters, fill, st);
}

int ptp_geodesics_f32(ptp_mesh_t *m, const uint32_t *sources, uint32_t S, float *dist, uint32_t *clusters, uint32_t fill,
                      uint32_t *sorted_index, uint64_t scap, ptp_stats_t *st)
{
    NEED(m, 4) return geodesics_impl<float>(m, sources, S, dist, clusters, fill, sorted_index, scap, st);
}
int ptp_geodesics_f64(ptp_mesh_t *m, const uint32_t *sources, uint32_t S, double *dist, uint32_t *clusters, uint32_t fill,
                      uint32_t *sorted_index, uint64_t scap, ptp_stats_t *st)
{
    NEED(m, 8) return geodesics_impl<double>(m, sources, S, dist, clusters, fill, sorted_index, scap, st);
}

int ptp_geodesics_error_iter_f32(ptp_mesh_t *m, const uint32_t *sources, uint32_t S, const float *exact, float *dist, uint32_t *iters,
                                 float *errors, uint32_t cap, uint32_t *n_out, ptp_stats_t *st)
{
    NEED(m, 4) return error_iter_impl<float>(m, sources, S, exact, dist, iters, errors, cap, n_out, st);
}
int ptp_geodesics_error_iter_f64(ptp_mesh_t *m, const uint32_t *sources, uint32_t S, const double *exact, double *dist, uint32_t *iters,
                                 double *errors, uint32_t cap, uint32_t *n_out, ptp_stats_t *st)
{
    NEED(m, 8) return error_iter_impl<double>(m, sources, S, exact, dist, iters, errors, cap, n_out, st);
}

int ptp_solve_batched_f32(ptp_mesh_t *m, const uint32_t *sources, const uint64_t *offsets, uint32_t B, uint64_t n_src, float *rows,
                          int on_device, void *stream, ptp_stats_t *st)
{
    NEED(m, 4) return batched_impl<float>(m, sources, offsets, B, n_src, rows, on_device, stream, st);
}
int ptp_solve_batched_f64(ptp_mesh_t *m, const uint32_t *sources, const uint64_t *offsets, uint32_t B, uint64_t n_src, double *rows,
                          int on_device, void *stream, ptp_stats_t *st)
{
    NEED(m, 8) return batched_impl<double>(m, sources, offsets, B, n_src, rows, on_device, stream, st);
}

int ptp_solve_batched_multi_f32(ptp_mesh_t *const *meshes, int n_devices, const uint32_t *sources, const uint64_t *offsets, uint32_t B,
                                uint64_t n_src, float *rows, int on_device, ptp_stats_t *st)
{
    return batched_multi_impl<float>(meshes, n_devices, sources, offsets, B, n_src, rows, on_device, st);
}
int ptp_solve_batched_multi_f64(ptp_mesh_t *const *meshes, int n_devices, const uint32_t *sources, const uint64_t *offsets, uint32_t B,
                                uint64_t n_src, double *rows, int on_device, ptp_stats_t *st)
{
    return batched_multi_impl<double>(meshes, n_devices, sources, offsets, B, n_src, rows, on_device, st);
}

int ptp_farthest_point_sampling_f32(ptp_mesh_t *m, uint32_t *samples, uint32_t n_initial, uint32_t n_total, float radio,
                                    uint32_t *n_out, float *max_dist, ptp_stats_t *st)
{
    NEED(m, 4) return fps_impl<float>(m, samples, n_initial, n_total, radio, n_out, max_dist, st);
}
int ptp_farthest_point_sampling_f64(ptp_mesh_t *m, uint32_t *samples, uint32_t n_initial, uint32_t n_total, double radio,
                                    uint32_t *n_out, double *max_dist, ptp_stats_t *st)
{
    NEED(m, 8) return fps_impl<double>(m, samples, n_initial, n_total, radio, n_out, max_dist, st);
}

} // extern "C"
