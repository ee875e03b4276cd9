/* Synthetic triangle meshes (the inputs BASELINE.json's configs name) and a host-side builder for the
 * compact-half-edge (CHE) tables OT / EVT from a face list. Host helpers for tests, bench.py and
 * callers that do not already hold a gproshan `che`; nothing here is on the GPU hot path.
 *
 * CHE conventions follow the reference (file:line relative to larc/gproshan):
 *   VT[he]  origin vertex of half-edge he; face f owns half-edges 3f, 3f+1, 3f+2   include/che.h:41-47
 *   OT[he]  opposite half-edge or NIL on a border                                   src/che.cpp:1311-1330
 *   EVT[v]  last half-edge leaving v in VT order, overridden by the border half-edge
 *           for border vertices; NIL for isolated vertices                          src/che.cpp:1304-1308,1343-1352
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NIL 0xFFFFFFFFu

/* ---------------------------------------------------------------- random numbers (MT19937 family) */

void mg_mt19937(uint32_t seed, size_t n, uint32_t *out)
{
    uint32_t mt[624];
    int idx = 624;
    mt[0] = seed;
    for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    for (size_t k = 0; k < n; k++) {
        if (idx >= 624) {
            for (int i = 0; i < 624; i++) {
                uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        out[k] = y;
    }
}

typedef struct { uint64_t mt[312]; int idx; } mt64_t;

static void mt64_seed(mt64_t *s, uint64_t seed)
{
    s->mt[0] = seed;
    for (int i = 1; i < 312; i++) s->mt[i] = 6364136223846793005ull * (s->mt[i - 1] ^ (s->mt[i - 1] >> 62)) + (uint64_t)i;
    s->idx = 312;
}

static uint64_t mt64_next(mt64_t *s)
{
    if (s->idx >= 312) {
        for (int i = 0; i < 312; i++) {
            uint64_t x = (s->mt[i] & 0xFFFFFFFF80000000ull) | (s->mt[(i + 1) % 312] & 0x7FFFFFFFull);
            s->mt[i] = s->mt[(i + 156) % 312] ^ (x >> 1) ^ ((x & 1ull) ? 0xB5026F5AA96619E9ull : 0ull);
        }
        s->idx = 0;
    }
    uint64_t x = s->mt[s->idx++];
    x ^= (x >> 29) & 0x5555555555555555ull;
    x ^= (x << 17) & 0x71D67FFFEDA60000ull;
    x ^= (x << 37) & 0xFFF7EEE000000000ull;
    x ^= x >> 43;
    return x;
}

/* scale every vertex radially by 1 + sigma*u, u ~ U(-1,1) from mt19937_64(seed): u = 2*(x>>11)*2^-53 - 1 */
void mg_radial_noise(double *xyz, size_t n_v, double sigma, uint64_t seed)
{
    mt64_t s;
    mt64_seed(&s, seed);
    for (size_t v = 0; v < n_v; v++) {
        const double u = 2.0 * ((double)(mt64_next(&s) >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
        const double k = 1.0 + sigma * u;
        xyz[3 * v] *= k; xyz[3 * v + 1] *= k; xyz[3 * v + 2] *= k;
    }
}

/* ---------------------------------------------------------------- grid */

/* nx*ny vertices on z=0, x=i/(nx-1), y=j/(ny-1), id=j*nx+i; each cell split along (i,j)-(i+1,j+1):
 * faces (a,b,d),(a,d,c) with a=(i,j) b=(i+1,j) c=(i,j+1) d=(i+1,j+1). */
void mg_grid(uint32_t nx, uint32_t ny, double *xyz, uint32_t *faces)
{
    for (uint32_t j = 0; j < ny; j++)
        for (uint32_t i = 0; i < nx; i++) {
            size_t v = (size_t)j * nx + i;
            xyz[3 * v] = (double)i / (double)(nx - 1);
            xyz[3 * v + 1] = (double)j / (double)(ny - 1);
            xyz[3 * v + 2] = 0.0;
        }
    size_t f = 0;
    for (uint32_t j = 0; j + 1 < ny; j++)
        for (uint32_t i = 0; i + 1 < nx; i++) {
            uint32_t a = j * nx + i, b = a + 1, c = a + nx, d = c + 1;
            faces[3 * f] = a; faces[3 * f + 1] = b; faces[3 * f + 2] = d; f++;
            faces[3 * f] = a; faces[3 * f + 1] = d; faces[3 * f + 2] = c; f++;
        }
}

/* ---------------------------------------------------------------- torus */

/* nu*nv vertices, id=a*nv+b, theta=2pi a/nu (major), phi=2pi b/nv (minor); 2*nu*nv faces */
void mg_torus(uint32_t nu, uint32_t nv, double R, double r, double *xyz, uint32_t *faces)
{
    const double two_pi = 6.283185307179586476925286766559;
    for (uint32_t a = 0; a < nu; a++)
        for (uint32_t b = 0; b < nv; b++) {
            const double th = two_pi * (double)a / (double)nu, ph = two_pi * (double)b / (double)nv;
            size_t v = (size_t)a * nv + b;
            xyz[3 * v] = (R + r * cos(ph)) * cos(th);
            xyz[3 * v + 1] = (R + r * cos(ph)) * sin(th);
            xyz[3 * v + 2] = r * sin(ph);
        }
    size_t f = 0;
    for (uint32_t a = 0; a < nu; a++)
        for (uint32_t b = 0; b < nv; b++) {
            uint32_t a1 = (a + 1) % nu, b1 = (b + 1) % nv;
            uint32_t v00 = a * nv + b, v10 = a1 * nv + b, v01 = a * nv + b1, v11 = a1 * nv + b1;
            faces[3 * f] = v00; faces[3 * f + 1] = v10; faces[3 * f + 2] = v11; f++;
            faces[3 * f] = v00; faces[3 * f + 1] = v11; faces[3 * f + 2] = v01; f++;
        }
}

/* ---------------------------------------------------------------- class-I icosphere */

static const double ICO_T = 1.6180339887498948482045868343656;
static const int ICO_F[20][3] = {
    {0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11},
    {1, 5, 9}, {5, 11, 4}, {11, 10, 2}, {10, 7, 6}, {7, 1, 8},
    {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8}, {3, 8, 9},
    {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};

typedef struct { uint32_t f; int edge_id[12][12]; } ico_t;

static uint32_t ico_edge_vertex(const ico_t *ic, int u, int v, uint32_t t)
{
    /* vertex t steps from corner u towards corner v on edge (u,v), 0 < t < f */
    if (u > v) { int w = u; u = v; v = w; t = ic->f - t; }
    return 12u + (uint32_t)ic->edge_id[u][v] * (ic->f - 1) + (t - 1);
}

static uint32_t ico_vid(const ico_t *ic, int F, uint32_t i, uint32_t j)
{
    const uint32_t f = ic->f;
    const int A = ICO_F[F][0], B = ICO_F[F][1], C = ICO_F[F][2];
    if (i == 0 && j == 0) return (uint32_t)A;
    if (i == f && j == 0) return (uint32_t)B;
    if (i == 0 && j == f) return (uint32_t)C;
    if (j == 0) return ico_edge_vertex(ic, A, B, i);
    if (i == 0) return ico_edge_vertex(ic, A, C, j);
    if (i + j == f) return ico_edge_vertex(ic, B, C, j);
    const uint32_t per_face = (f - 1) * (f - 2) / 2;
    /* interior rows j = 1..f-2, row j holds i = 1..f-1-j */
    const uint32_t before = (j - 1) * (f - 1) - (j - 1) * j / 2;
    return 12u + 30u * (f - 1) + (uint32_t)F * per_face + before + (i - 1);
}

void mg_icosphere_counts(uint32_t f, uint64_t *n_v, uint64_t *n_f)
{
    *n_v = 10ull * f * f + 2;
    *n_f = 20ull * f * f;
}

static void ico_put(double *xyz, size_t v, double x, double y, double z)
{
    const double n = sqrt(x * x + y * y + z * z);
    xyz[3 * v] = x / n; xyz[3 * v + 1] = y / n; xyz[3 * v + 2] = z / n;
}

/* frequency-f subdivision of the icosahedron projected on the unit sphere: 10f^2+2 vertices, 20f^2 faces.
 * Numbering: 12 corners, then edge vertices (30 edges in first-seen order), then face interiors. */
void mg_icosphere(uint32_t f, double *xyz, uint32_t *faces)
{
    const double t = ICO_T;
    const double P[12][3] = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
                             {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
    ico_t ic;
    ic.f = f;
    memset(ic.edge_id, -1, sizeof(ic.edge_id));
    int n_edges = 0;
    for (int F = 0; F < 20; F++)
        for (int k = 0; k < 3; k++) {
            int u = ICO_F[F][k], v = ICO_F[F][(k + 1) % 3];
            if (u > v) { int w = u; u = v; v = w; }
            if (ic.edge_id[u][v] < 0) ic.edge_id[u][v] = n_edges++;
        }

    for (int c = 0; c < 12; c++) ico_put(xyz, (size_t)c, P[c][0], P[c][1], P[c][2]);
    for (int u = 0; u < 12; u++)
        for (int v = u + 1; v < 12; v++) {
            if (ic.edge_id[u][v] < 0) continue;
            for (uint32_t s = 1; s < f; s++) {
                const double a = (double)(f - s), b = (double)s;
                ico_put(xyz, ico_edge_vertex(&ic, u, v, s),
                        (P[u][0] * a + P[v][0] * b) / f, (P[u][1] * a + P[v][1] * b) / f, (P[u][2] * a + P[v][2] * b) / f);
            }
        }
    #pragma omp parallel for schedule(dynamic, 1)
    for (int F = 0; F < 20; F++) {
        const double *A = P[ICO_F[F][0]], *B = P[ICO_F[F][1]], *C = P[ICO_F[F][2]];
        for (uint32_t j = 1; j + 1 < f; j++)
            for (uint32_t i = 1; i + j < f; i++) {
                const double a = (double)(f - i - j), b = (double)i, c = (double)j;
                ico_put(xyz, ico_vid(&ic, F, i, j),
                        (A[0] * a + B[0] * b + C[0] * c) / f, (A[1] * a + B[1] * b + C[1] * c) / f, (A[2] * a + B[2] * b + C[2] * c) / f);
            }
    }
    #pragma omp parallel for schedule(dynamic, 1)
    for (int F = 0; F < 20; F++) {
        size_t k = (size_t)F * f * f;
        for (uint32_t j = 0; j < f; j++)
            for (uint32_t i = 0; i + j < f; i++) {
                faces[3 * k] = ico_vid(&ic, F, i, j);
                faces[3 * k + 1] = ico_vid(&ic, F, i + 1, j);
                faces[3 * k + 2] = ico_vid(&ic, F, i, j + 1);
                k++;
                if (i + j + 1 < f) {
                    faces[3 * k] = ico_vid(&ic, F, i + 1, j);
                    faces[3 * k + 1] = ico_vid(&ic, F, i + 1, j + 1);
                    faces[3 * k + 2] = ico_vid(&ic, F, i, j + 1);
                    k++;
                }
            }
    }
}

/* ---------------------------------------------------------------- CHE tables on the host */

static inline uint32_t he_next(uint32_t he) { return 3 * (he / 3) + (he + 1) % 3; }

/* OT / EVT for an oriented triangle soup. For edge-manifold input (every directed edge a->b appears
 * once, so it has at most one opposite b->a) this produces exactly the reference's tables; inputs that
 * are not edge-manifold are rejected (returns 0) rather than paired in an order-dependent way.
 * Returns 1 on success. */
int mg_che_build(uint32_t n_v, uint32_t n_f, const uint32_t *VT, uint32_t *OT, uint32_t *EVT)
{
    const uint32_t n_he = 3 * n_f;
    uint32_t *off = (uint32_t *)calloc((size_t)n_v + 2, sizeof(uint32_t));
    uint32_t *lst = (uint32_t *)malloc(sizeof(uint32_t) * (n_he ? n_he : 1));
    if (!off || !lst) { free(off); free(lst); return 0; }
    for (uint32_t he = 0; he < n_he; he++) off[VT[he] + 1]++;
    for (uint32_t v = 0; v < n_v; v++) off[v + 1] += off[v];
    uint32_t *fill = (uint32_t *)malloc(sizeof(uint32_t) * (n_v ? n_v : 1));
    memcpy(fill, off, sizeof(uint32_t) * n_v);
    for (uint32_t he = 0; he < n_he; he++) lst[fill[VT[he]]++] = he;
    free(fill);

    int ok = 1;
    #pragma omp parallel for schedule(static) reduction(&: ok)
    for (uint32_t he = 0; he < n_he; he++) {
        const uint32_t a = VT[he], b = VT[he_next(he)];
        uint32_t opp = NIL, same = 0;
        for (uint32_t k = off[b]; k < off[b + 1]; k++)          /* half-edges leaving b */
            if (VT[he_next(lst[k])] == a) { if (opp != NIL) ok = 0; opp = lst[k]; }
        for (uint32_t k = off[a]; k < off[a + 1]; k++)          /* duplicates of a->b */
            if (VT[he_next(lst[k])] == b) same++;
        if (same != 1) ok = 0;
        OT[he] = opp;
    }

    #pragma omp parallel for schedule(static) reduction(&: ok)
    for (uint32_t v = 0; v < n_v; v++) {
        uint32_t e = NIL, borders = 0;
        if (off[v + 1] > off[v]) e = lst[off[v + 1] - 1];
        for (uint32_t k = off[v]; k < off[v + 1]; k++)
            if (OT[lst[k]] == NIL) { e = lst[k]; borders++; }
        if (borders > 1) ok = 0;                                  /* non-manifold vertex */
        EVT[v] = e;
    }
    free(off);
    free(lst);
    return ok;
}
