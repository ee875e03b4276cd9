/* TEST INFRASTRUCTURE — included twice by ptp_oracle.c with REAL = float / double.
 * See ptp_oracle.h for the contract. Arithmetic is written one operation per reference
 * operation, in the reference's order; the file is compiled with -ffp-contract=off and no -march
 * so no FMA can be formed. */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* vertex::operator, (dot)  src/vertex.cpp:41-44 : x*v.x + y*v.y + z*v.z, left to right */
static inline REAL FN(dot3_)(const REAL *a, const REAL *b)
{
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

/* vertex::operator* () (norm)  src/vertex.cpp:36-39 */
static inline REAL FN(norm3_)(const REAL *a)
{
    return SQRT(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
}

/* update_step, src/geodesics_ptp.cpp:201-262 */
static inline REAL FN(update_step_)(const REAL *GT, const uint32_t *VT, const REAL *dist, uint32_t he)
{
    uint32_t x[3];
    x[0] = VT[he_next(he)];
    x[1] = VT[he_prev(he)];
    x[2] = VT[he];

    REAL X[2][3];
    for (int c = 0; c < 3; c++) {
        X[0][c] = GT[3 * (size_t)x[0] + c] - GT[3 * (size_t)x[2] + c];
        X[1][c] = GT[3 * (size_t)x[1] + c] - GT[3 * (size_t)x[2] + c];
    }

    REAL t[2];
    t[0] = dist[x[0]];
    t[1] = dist[x[1]];

    REAL q[2][2];
    q[0][0] = FN(dot3_)(X[0], X[0]);
    q[0][1] = FN(dot3_)(X[0], X[1]);
    q[1][0] = FN(dot3_)(X[1], X[0]);
    q[1][1] = FN(dot3_)(X[1], X[1]);

    REAL det = q[0][0] * q[1][1] - q[0][1] * q[1][0];
    REAL Q[2][2];
    Q[0][0] = q[1][1] / det;
    Q[0][1] = -q[0][1] / det;
    Q[1][0] = -q[1][0] / det;
    Q[1][1] = q[0][0] / det;

    REAL delta = t[0] * (Q[0][0] + Q[1][0]) + t[1] * (Q[0][1] + Q[1][1]);
    REAL dis = delta * delta -
               (Q[0][0] + Q[0][1] + Q[1][0] + Q[1][1]) *
               (t[0] * t[0] * Q[0][0] + t[0] * t[1] * (Q[1][0] + Q[0][1]) + t[1] * t[1] * Q[1][1] - 1);

    REAL p = (delta + SQRT(dis)) / (Q[0][0] + Q[0][1] + Q[1][0] + Q[1][1]);

    REAL tp[2];
    tp[0] = t[0] - p;
    tp[1] = t[1] - p;

    REAL n[3];
    for (int c = 0; c < 3; c++)
        n[c] = tp[0] * (X[0][c] * Q[0][0] + X[1][c] * Q[1][0]) + tp[1] * (X[0][c] * Q[0][1] + X[1][c] * Q[1][1]);

    REAL cond[2];
    cond[0] = FN(dot3_)(X[0], n);
    cond[1] = FN(dot3_)(X[1], n);

    REAL c[2];
    c[0] = cond[0] * Q[0][0] + cond[1] * Q[0][1];
    c[1] = cond[0] * Q[1][0] + cond[1] * Q[1][1];

    if (t[0] == (REAL)INFINITY || t[1] == (REAL)INFINITY || dis < 0 || c[0] >= 0 || c[1] >= 0) {
        REAL dp[2];
        dp[0] = dist[x[0]] + FN(norm3_)(X[0]);
        dp[1] = dist[x[1]] + FN(norm3_)(X[1]);
        p = dp[dp[1] < dp[0]];
    }
    return p;
}

REAL FN(orc_update_step_)(const REAL *GT, const uint32_t *VT, const REAL *dist, uint32_t he)
{
    return FN(update_step_)(GT, VT, dist, he);
}

/* parallel_toplesets_propagation_cpu, src/geodesics_ptp.cpp:122-199 */
void FN(orc_ptp_cpu_)(uint32_t n_v, const REAL *GT, const uint32_t *VT, const uint32_t *OT,
                      const uint32_t *EVT, const uint32_t *sources, uint32_t n_sources,
                      const uint32_t *limits, uint32_t n_limits, const uint32_t *sorted,
                      REAL *dist, uint32_t *clusters, uint32_t cluster_fill, uint64_t *stats)
{
    REAL *pdist[2] = {dist, (REAL *)malloc(sizeof(REAL) * (n_v ? n_v : 1))};
    uint32_t *pcl[2] = {NULL, NULL};

    #pragma omp parallel for
    for (uint32_t v = 0; v < n_v; v++)
        pdist[0][v] = pdist[1][v] = (REAL)INFINITY;                       /* :127-129 */

    if (clusters) {
        pcl[0] = clusters;
        pcl[1] = (uint32_t *)malloc(sizeof(uint32_t) * (n_v ? n_v : 1));
        #pragma omp parallel for
        for (uint32_t v = 0; v < n_v; v++)
            pcl[0][v] = pcl[1][v] = cluster_fill;
    }

    for (uint32_t i = 0; i < n_sources; i++) {                            /* :131-135 */
        pdist[0][sources[i]] = pdist[1][sources[i]] = 0;
        if (clusters) pcl[0][sources[i]] = pcl[1][sources[i]] = i + 1;     /* src/cuda/geodesics_ptp.cu:191-195 */
    }

    uint32_t d = 0;
    uint32_t start, end, n_cond, count;
    uint32_t i = 1, j = 2;
    uint32_t iter = 0;
    uint32_t max_iter = n_limits << 1;                                     /* :143 */
    uint64_t updates = 0, max_window = 0, iterations = 0;

    /* the reference indexes limits[2] unconditionally (UB when n_limits < 3); the oracle defines
     * that case as "no iteration": sources 0, everything else INF */
    while (n_limits >= 3 && i < j && iter++ < max_iter) {                  /* :145 */
        if (i < (j >> 1)) i = (j >> 1);                                    /* :147 */

        start = limits[i];
        end = limits[j];
        n_cond = limits[i + 1] - start;

        const REAL *old_d = pdist[d];
        REAL *new_d = pdist[!d];
        const uint32_t *old_c = pcl[d];
        uint32_t *new_c = pcl[!d];

        #pragma omp parallel for
        for (uint32_t vi = start; vi < end; vi++) {                        /* :153-171 */
            const uint32_t v = sorted[vi];
            new_d[v] = old_d[v];
            if (clusters) new_c[v] = old_c[v];                             /* geodesics_ptp.cu:268 */

            /* for_star, include/che.h:10 */
            const uint32_t stop = EVT[v];
            for (uint32_t he = stop; he != ORC_NIL;) {
                REAL p = FN(update_step_)(GT, VT, old_d, he);
                if (p < new_d[v]) {
                    new_d[v] = p;
                    if (clusters) {                                        /* geodesics_ptp.cu:277 */
                        const uint32_t xp = VT[he_prev(he)], xn = VT[he_next(he)];
                        new_c[v] = old_d[xp] < old_d[xn] ? old_c[xp] : old_c[xn];
                    }
                }
                he = OT[he_prev(he)];
                if (he == stop) he = ORC_NIL;
            }
        }

        count = 0;
        #pragma omp parallel for reduction(+: count)
        for (uint32_t vi = start; vi < start + n_cond; vi++) {             /* :173-185 */
            const uint32_t v = sorted[vi];
            const REAL err = ABS(new_d[v] - old_d[v]) / old_d[v];
            count += err < 1e-3;                                           /* PTP_TOL, include/geodesics_ptp.h:11; double compare */
        }

        if (n_cond == count) i++;                                          /* :186 */
        if (j < n_limits - 1) j++;                                         /* :187 */
        d = !d;                                                            /* :189 */

        iterations++;
        updates += end - start;
        if (end - start > max_window) max_window = end - start;
    }

    /* :193-198 — the function hands back pdist[!d], the OLDER buffer */
    if (dist != pdist[!d]) memcpy(dist, pdist[!d], sizeof(REAL) * n_v);
    free(pdist[1]);
    if (clusters) {
        if (clusters != pcl[!d]) memcpy(clusters, pcl[!d], sizeof(uint32_t) * n_v);
        free(pcl[1]);
    }
    if (stats) {
        stats[0] = iterations;
        stats[1] = updates;
        stats[2] = max_window;
        stats[3] = d;
    }
}

/* normalize_ptp, src/geodesics_ptp.cpp:264-276 */
void FN(orc_normalize_ptp_)(REAL *dist, size_t n)
{
    REAL max_d = 0;
    #pragma omp parallel for reduction(max: max_d)
    for (size_t v = 0; v < n; v++)
        if (dist[v] < (REAL)INFINITY)
            max_d = dist[v] > max_d ? dist[v] : max_d;

    #pragma omp parallel for
    for (size_t v = 0; v < n; v++)
        dist[v] /= max_d;
}

#undef FN
#undef CAT
#undef CAT_
