/*
 * oracle/seeds_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the first consumer of the Frangi outputs,
 * SeedExtractor::extractSeeds (pnr-vaa3d/seed.cpp:556-791), used by the parity
 * tests as the "downstream seed set" check named by BASELINE.json:north_star:
 * the seeds extracted from (GPU J8, GPU Vx/Vy/Vz) must match the seeds extracted
 * from (reference J8, reference Vx/Vy/Vz).
 *
 * The reference routine is itself a per-z-layer port of ImageJ's MaximumFinder
 * (public algorithm): 8-neighbour local maxima of the layer (border pixels and
 * pixels equal to the layer minimum excluded), visited from the highest down;
 * each one flood-fills the connected set within `tolerance` below it; the
 * maximum is dropped if the fill meets a higher pixel, an already processed
 * pixel or the image border; otherwise one seed is emitted at the member of the
 * equal-height plateau nearest to the plateau's centroid.
 *
 * Parity status: PINNED against the compiled reference: tests/test_oracle_golden.py
 * (test_case_b_outputs_and_seeds: the seed list of tests/golden/case_b_seeds.npz, generated
 * from oracle/_ref by tools/make_golden.py; test_port_equals_reference_where_built: against
 * oracle/_ref directly); bit-exact seed positions are required.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* pixel flags, seed.cpp:1053-1059 */
enum { F_MAXIMUM = 1, F_LISTED = 2, F_PROCESSED = 4, F_MAX_AREA = 8, F_EQUAL = 16, F_MAX_POINT = 32 };

/* neighbour order, seed.cpp:1051-1052 (N, NE, E, SE, S, SW, W, NW) */
static const int NBR_DX[8] = { 0, 1, 1, 1, 0, -1, -1, -1 };
static const int NBR_DY[8] = { -1, -1, 0, 1, 1, 1, 0, -1 };

static inline int nbr_inside(int x, int y, int d, int w, int h)
{
    int x2 = x + NBR_DX[d], y2 = y + NBR_DY[d]; /* equivalent to isWithin, seed.cpp:1027-1049 */
    return x2 >= 0 && x2 < w && y2 >= 0 && y2 < h;
}

static inline int on_border(int x, int y, int w, int h)
{
    return x == 0 || x == w - 1 || y == 0 || y == h - 1;
}

static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

/* The per-layer pre-pass of extractSeeds (seed.cpp:574-632): layer range, 8-neighbour candidate maxima
 * (flags[p] = F_MAXIMUM), and the candidates ranked by height -- value in the upper 32 bits, pixel offset in
 * the lower, sorted ascending.  flags must be zeroed by the caller.  Returns the number of candidates. */
static int layer_candidates(const uint8_t *layer, int w, int h, uint8_t *flags, int64_t *ranked, float *lo_out,
                            float *hi_out)
{
    const int64_t plane = (int64_t)w * h;
    /* layer range, seed.cpp:578-586 */
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (int64_t p = 0; p < plane; ++p) {
        float v = (float)(int)layer[p];
        if (lo > v) lo = v;
        if (hi < v) hi = v;
    }
    /* candidate maxima, seed.cpp:590-614 */
    int n_max = 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float v = layer[(int64_t)y * w + x];
            if (v == lo) continue;
            if (on_border(x, y, w, h)) continue;
            int is_max = 1;
            for (int d = 0; d < 8; ++d) {
                float vn = layer[(int64_t)(y + NBR_DY[d]) * w + (x + NBR_DX[d])];
                if (vn > v) { is_max = 0; break; }
            }
            if (is_max) { flags[(int64_t)y * w + x] = F_MAXIMUM; ++n_max; }
        }
    /* rank by height: value in the upper 32 bits, pixel offset below, seed.cpp:616-632 */
    float to_int = (float)(2e9 / (hi - lo));
    int k = 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int p = x + y * w;
            if (flags[p] == F_MAXIMUM) {
                float fv = layer[p];
                int iv = (int)((fv - lo) * to_int);
                ranked[k++] = ((int64_t)iv << 32) | (int64_t)p;
            }
        }
    qsort(ranked, (size_t)n_max, sizeof(int64_t), cmp_i64);
    *lo_out = lo; *hi_out = hi;
    return n_max;
}

/* The pre-pass alone, for the parity test of frangi_gpu_seed_candidates: per layer the range, the number of
 * candidates and their ranked keys (all layers concatenated).  Returns the total number of keys (at most `cap`
 * are written), -1 on allocation failure. */
ORACLE_API long oracle_seed_candidates(const uint8_t *J8, int w, int h, int l, uint8_t *layer_min, uint8_t *layer_max,
                                       int *n_max_out, int64_t *keys, long cap)
{
    const int64_t plane = (int64_t)w * h;
    uint8_t *flags = (uint8_t *)malloc((size_t)plane);
    int64_t *ranked = (int64_t *)malloc(sizeof(int64_t) * (size_t)plane);
    long total = 0;
    if (!flags || !ranked) { free(flags); free(ranked); return -1; }
    for (int z = 0; z < l; ++z) {
        float lo, hi;
        memset(flags, 0, (size_t)plane);
        int n = layer_candidates(J8 + (int64_t)z * plane, w, h, flags, ranked, &lo, &hi);
        layer_min[z] = (uint8_t)lo; layer_max[z] = (uint8_t)hi; n_max_out[z] = n;
        for (int k = 0; k < n; ++k, ++total)
            if (total < cap) keys[total] = ranked[k];
    }
    free(flags); free(ranked);
    return total;
}

/* Returns the number of seeds; writes at most `cap` rows (x,y,z,vx,vy,vz). */
ORACLE_API long oracle_extract_seeds(double tolerance, const uint8_t *J8, int w, int h, int l,
                                     const uint8_t *Vx, const uint8_t *Vy, const uint8_t *Vz,
                                     float *out, long cap)
{
    const int64_t plane = (int64_t)w * h;
    uint8_t *flags = (uint8_t *)malloc((size_t)plane);
    int *fill = (int *)malloc(sizeof(int) * (size_t)plane);
    int64_t *ranked = (int64_t *)malloc(sizeof(int64_t) * (size_t)plane);
    long count = 0;
    if (!flags || !fill || !ranked) { free(flags); free(fill); free(ranked); return -1; }

    for (int z = 0; z < l; ++z) {
        const uint8_t *layer = J8 + (int64_t)z * plane;
        memset(flags, 0, (size_t)plane);
        float lo, hi;
        int n_max = layer_candidates(layer, w, h, flags, ranked, &lo, &hi);

        /* analyse from the highest maximum down, seed.cpp:643-782 */
        for (int im = n_max - 1; im >= 0; --im) {
            int start = (int)ranked[im];
            if (flags[start] & F_PROCESSED) continue;
            int x0 = start % w, y0 = start / w;
            float v0 = layer[start];
            int retry;
            do {
                fill[0] = start;
                flags[start] |= (F_EQUAL | F_LISTED);
                int n_fill = 1, cur = 0;
                int edge_max = on_border(x0, y0, w, h);
                int possible = 1;
                double sum_x = x0, sum_y = y0;
                int n_equal = 1;
                retry = 0;
                do {
                    int p = fill[cur];
                    int x = p % w, y = p / w;
                    for (int d = 0; d < 8; ++d) {
                        if (!nbr_inside(x, y, d, w, h)) continue;
                        int p2 = p + NBR_DY[d] * w + NBR_DX[d];
                        if (flags[p2] & F_LISTED) continue;
                        if (flags[p2] & F_PROCESSED) { possible = 0; break; }
                        int x2 = x + NBR_DX[d], y2 = y + NBR_DY[d];
                        float v2 = layer[p2];
                        if (v2 > v0 + 0.0f) { possible = 0; break; }
                        if (v2 >= v0 - (float)tolerance) {
                            if (v2 > v0) { retry = 1; start = p2; v0 = v2; x0 = x2; y0 = y2; }
                            fill[n_fill++] = p2;
                            flags[p2] |= F_LISTED;
                            if (on_border(x2, y2, w, h)) { edge_max = 1; possible = 0; break; }
                            if (v2 == v0) {
                                flags[p2] |= F_EQUAL;
                                sum_x += x2; sum_y += y2; ++n_equal;
                            }
                        }
                    }
                    ++cur;
                } while (cur < n_fill);

                if (retry) {
                    for (int q = 0; q < n_fill; ++q) flags[fill[q]] = 0;
                } else {
                    int keep = ~(possible ? F_LISTED : (F_LISTED | F_EQUAL));
                    double cx = sum_x / n_equal, cy = sum_y / n_equal;
                    double best = 1e20;
                    int nearest = 0;
                    for (int q = 0; q < n_fill; ++q) {
                        int p = fill[q];
                        int x = p % w, y = p / w;
                        flags[p] &= (uint8_t)keep;
                        flags[p] |= F_PROCESSED;
                        if (possible) {
                            flags[p] |= F_MAX_AREA;
                            if (flags[p] & F_EQUAL) {
                                double d2 = (cx - x) * (cx - x) + (cy - y) * (cy - y);
                                if (d2 < best) { best = d2; nearest = q; }
                            }
                        }
                    }
                    if (possible) {
                        int p = fill[nearest];
                        flags[p] |= F_MAX_POINT;
                        if (!edge_max) {
                            int x = p % w, y = p / w;
                            int64_t si = (int64_t)z * plane + (int64_t)y * w + x;
                            /* direction decode, seed.cpp:767-771 */
                            float ux = (((float)Vx[si] / 255) * 2) - 1;
                            float uy = (((float)Vy[si] / 255) * 2) - 1;
                            float uz = (((float)Vz[si] / 255) * 2) - 1;
                            float un = (float)sqrt((double)ux * ux + (double)uy * uy + (double)uz * uz);
                            if (count < cap) {
                                float *row = out + 6 * count;
                                row[0] = (float)x; row[1] = (float)y; row[2] = (float)z;
                                row[3] = ux / un; row[4] = uy / un; row[5] = uz / un;
                            }
                            ++count;
                        }
                    }
                }
            } while (retry);
        }
    }
    free(flags); free(fill); free(ranked);
    return count;
}
