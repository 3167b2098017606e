/* CPU oracle (C restatement) for the CMDA event-representation path.
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs, never by the product library.
 *
 * Every function restates the reference's algorithm (file:line cited) in plain C:
 * float32 arithmetic, one rounding per operation, left to right, no FMA contraction
 * (build with -ffp-contract=off, see oracle/Makefile).  The voxel accumulation runs in
 * the deterministic reference's order: corner pass major, event index minor
 * (SURVEY.md Q4), which makes the raw grid bit-identical to the single-thread
 * reference and to oracle/cmda_oracle.py; tests/test_oracle_c.py checks this file
 * against the golden fixtures produced by the reference itself.
 *
 * OpenMP is used ONLY across independent windows / images (the unit the reference
 * processes in separate DataLoader workers), never inside one accumulation.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define INT_MIN64 (-2147483648LL)

/* tensor.int() on x86 (dsec.py:41-43): truncation; NaN/inf/|v|>=2^31 -> INT_MIN. */
static inline int64_t trunc_i(float v) {
    if (!(fabsf(v) < 2147483648.0f)) return INT_MIN64; /* also catches NaN */
    return (int64_t)v;
}

/* mmseg/datasets/dsec.py:26-58, normalize_flag=False. grid is [B,H,W], zeroed here. */
void oracle_voxel_grid(const float* time, const float* x, const float* y, const float* pol, int64_t n,
                       int W, int H, int B, float* grid) {
    memset(grid, 0, sizeof(float) * (size_t)B * H * W);                   /* dsec.py:31 */
    if (n <= 0) return;
    const float t_first = time[0];
    const float den = time[n - 1] - time[0];                              /* dsec.py:39 */
    const float cm1 = (float)(B - 1);
    for (int pass = 0; pass < 8; ++pass) {                                /* dsec.py:47-49 */
        const int dx = (pass >> 2) & 1, dy = (pass >> 1) & 1, dt = pass & 1;
        for (int64_t i = 0; i < n; ++i) {
            const float a = time[i] - t_first;
            const float b = cm1 * a;
            const float tn = b / den;                                     /* dsec.py:38-39 */
            const int64_t xl = trunc_i(x[i]) + dx;                        /* dsec.py:41 */
            const int64_t yl = trunc_i(y[i]) + dy;                        /* dsec.py:42 */
            const int64_t tl = trunc_i(tn) + dt;                          /* dsec.py:43 */
            if (!(xl < W && xl >= 0 && yl < H && yl >= 0 && tl >= 0 && tl < B)) continue; /* :50 */
            const float value = 2.0f * pol[i] - 1.0f;                     /* dsec.py:45 */
            const float wx = 1.0f - fabsf((float)xl - x[i]);
            const float wy = 1.0f - fabsf((float)yl - y[i]);
            const float wt = 1.0f - fabsf((float)tl - tn);
            float w = value * wx;                                         /* dsec.py:51-52 */
            w = w * wy;
            w = w * wt;
            grid[(int64_t)H * W * tl + (int64_t)W * yl + xl] += w;        /* dsec.py:54-58 */
        }
    }
}

static void normalize_to_range(float* v, int64_t n, float min_val, float max_val) {
    /* dsec.py:73-77 */
    if (n <= 0) return;
    float mn = v[0], mx = v[0];
    for (int64_t i = 1; i < n; ++i) { if (v[i] < mn) mn = v[i]; if (v[i] > mx) mx = v[i]; }
    const float den = (mx - mn) + 1e-8f;
    const float span = max_val - min_val;
    for (int64_t i = 0; i < n; ++i) {
        float r = (v[i] - mn) / den;
        r = r * span;
        v[i] = r + min_val;
    }
}

/* mmseg/datasets/dsec.py:80-121, numeric clip_range. In place on ev[n]; scratch[n]. */
void oracle_events_norm(float* ev, int64_t n, float clip, float final_range, int enforce, float* scratch) {
    int64_t nz = 0;
    double s = 0.0, s2 = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        if (ev[i] != 0.0f) ++nz;                                          /* dsec.py:88-89 */
        s += ev[i];
        const float sq = ev[i] * ev[i];
        s2 += sq;
    }
    if (nz > 0) {                                                         /* dsec.py:90 */
        const float mean = (float)s / (float)nz;                          /* dsec.py:91 */
        const float var = (float)s2 / (float)nz - mean * mean;
        const float sd = sqrtf(var);                                      /* dsec.py:92 */
        const float den = sd + 1e-8f;
        for (int64_t i = 0; i < n; ++i) {
            const float m = ev[i] != 0.0f ? 1.0f : 0.0f;
            ev[i] = (m * (ev[i] - mean)) / den;                           /* dsec.py:93-94 */
        }
    }
    if (enforce) {                                                        /* dsec.py:106-117 */
        for (int64_t i = 0; i < n; ++i) {
            const float e = ev[i];
            float neg = e > 0.0f ? 0.0f : e;
            neg = neg < -clip ? -clip : (neg > 0.0f ? 0.0f : neg);
            scratch[i] = neg;
            float pos = e < 0.0f ? 0.0f : e;
            pos = pos < 0.0f ? 0.0f : (pos > clip ? clip : pos);
            ev[i] = pos;
        }
        normalize_to_range(ev, n, 0.0f, final_range);
        normalize_to_range(scratch, n, -final_range, 0.0f);
        for (int64_t i = 0; i < n; ++i) ev[i] = ev[i] + scratch[i];
    } else {                                                              /* dsec.py:119-120 */
        for (int64_t i = 0; i < n; ++i) {
            float e = ev[i];
            e = e < -clip ? -clip : (e > clip ? clip : e);
            e = e * final_range;
            e = e / clip;
            ev[i] = e * final_range;
        }
    }
}

/* DSECDataset.get_events_vg (dsec.py:341-366) for S windows [start, finish] (inclusive) of
 * one SoA event store; windows run in parallel (one per thread), each exactly as the
 * reference would run it in a DataLoader worker.  clip[s] < 0 -> default of dsec.py:362.
 * out is [S,B,H,W]; raw (optional) receives the un-normalised grids. */
int oracle_get_events_vg_batch(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p,
                               const int64_t* start, const int64_t* finish, int S, const float* rectify_map,
                               int W, int H, int B, const double* clip, float* out, float* raw, int nthreads) {
    int err = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int s = 0; s < S; ++s) {
        const int64_t n = finish[s] - start[s] + 1;
        const size_t gsz = (size_t)B * H * W;
        float* grid = out + (size_t)s * gsz;
        if (n <= 0) { memset(grid, 0, gsz * sizeof(float)); continue; }
        float* tf = (float*)malloc(sizeof(float) * n);
        float* xf = (float*)malloc(sizeof(float) * n);
        float* yf = (float*)malloc(sizeof(float) * n);
        float* pf = (float*)malloc(sizeof(float) * n);
        float* scratch = (float*)malloc(sizeof(float) * gsz);
        if (!tf || !xf || !yf || !pf || !scratch) { err = 1; free(tf); free(xf); free(yf); free(pf); free(scratch); continue; }
        const int64_t o = start[s];
        const uint32_t t_first = t[o];
        const float t_last = (float)(uint32_t)(t[o + n - 1] - t_first);
        for (int64_t i = 0; i < n; ++i) {
            const float dtf = (float)(uint32_t)(t[o + i] - t_first);      /* dsec.py:347 */
            tf[i] = dtf / t_last;                                         /* dsec.py:348 */
            pf[i] = (float)p[o + i];                                      /* dsec.py:349 */
            if (rectify_map) {
                const float* m = rectify_map + ((size_t)y[o + i] * W + x[o + i]) * 2;  /* dsec.py:351 */
                xf[i] = m[0]; yf[i] = m[1];                               /* dsec.py:352-353 */
            } else { xf[i] = (float)x[o + i]; yf[i] = (float)y[o + i]; }
        }
        oracle_voxel_grid(tf, xf, yf, pf, n, W, H, B, grid);              /* dsec.py:356 */
        if (raw) memcpy(raw + (size_t)s * gsz, grid, gsz * sizeof(float));
        const double c = clip && clip[s] >= 0.0 ? clip[s] : (double)(finish[s] - start[s]) / 500000 * 1.5; /* :362 */
        oracle_events_norm(grid, (int64_t)gsz, (float)c, 1.0f, 1, scratch); /* dsec.py:365 */
        free(tf); free(xf); free(yf); free(pf); free(scratch);
    }
    return err;
}

/* Tolerance inputs of the parity tests for the same windows: per voxel the sum of |w| over its contributions
 * (float64) and their number -- the weights are the float32 products of dsec.py:51-52, the corners those of
 * dsec.py:47-50.  Test infrastructure for the full-size comparisons (numpy takes minutes at 80 M events). */
int oracle_voxel_aux_batch(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p,
                           const int64_t* start, const int64_t* finish, int S, const float* rectify_map,
                           int W, int H, int B, double* abs_w, int32_t* n_contrib, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int s = 0; s < S; ++s) {
        const int64_t n = finish[s] - start[s] + 1;
        const size_t gsz = (size_t)B * H * W;
        double* aw = abs_w + (size_t)s * gsz;
        int32_t* nc = n_contrib + (size_t)s * gsz;
        memset(aw, 0, gsz * sizeof(double));
        memset(nc, 0, gsz * sizeof(int32_t));
        if (n <= 0) continue;
        const int64_t o = start[s];
        const uint32_t t_first = t[o];
        const float t_last = (float)(uint32_t)(t[o + n - 1] - t_first);
        const float t01_first = 0.0f / t_last;
        const float den = t_last / t_last - t01_first;
        const float cm1 = (float)(B - 1);
        for (int64_t i = 0; i < n; ++i) {
            const float t01 = (float)(uint32_t)(t[o + i] - t_first) / t_last;
            const float tn = (cm1 * (t01 - t01_first)) / den;
            float xf, yf;
            if (rectify_map) {
                const float* m = rectify_map + ((size_t)y[o + i] * W + x[o + i]) * 2;
                xf = m[0]; yf = m[1];
            } else { xf = (float)x[o + i]; yf = (float)y[o + i]; }
            const float value = 2.0f * (float)p[o + i] - 1.0f;
            for (int pass = 0; pass < 8; ++pass) {
                const int64_t xl = trunc_i(xf) + ((pass >> 2) & 1), yl = trunc_i(yf) + ((pass >> 1) & 1), tl = trunc_i(tn) + (pass & 1);
                if (!(xl < W && xl >= 0 && yl < H && yl >= 0 && tl >= 0 && tl < B)) continue;
                float w = value * (1.0f - fabsf((float)xl - xf));
                w = w * (1.0f - fabsf((float)yl - yf));
                w = w * (1.0f - fabsf((float)tl - tn));
                const int64_t idx = (int64_t)H * W * tl + (int64_t)W * yl + xl;
                aw[idx] += fabs((double)w);
                nc[idx] += 1;
            }
        }
    }
    return 0;
}

/* utils.py:95-104 == create_cityscapes_image_change.py:22-31 on d[n] (in place). */
static void dead_zone_split_norm(float* d, int64_t n, float thr, float clip, float* scratch) {
    for (int64_t i = 0; i < n; ++i) {
        float v = d[i];
        if (fabsf(v) <= thr) v = 0.0f;
        float neg = v > 0.0f ? 0.0f : v;
        neg = neg < -clip ? -clip : (neg > 0.0f ? 0.0f : neg);
        scratch[i] = neg;
        float pos = v < 0.0f ? 0.0f : v;
        pos = pos < 0.0f ? 0.0f : (pos > clip ? clip : pos);
        d[i] = pos;
    }
    normalize_to_range(d, n, 0.0f, 1.0f);
    normalize_to_range(scratch, n, -1.0f, 0.0f);
    for (int64_t i = 0; i < n; ++i) d[i] = d[i] + scratch[i];
}

static inline uint8_t shifted_px(const uint8_t* g, int H, int W, int r, int c, int s, int dir) {
    /* utils.py:129-132: dir 0=left 1=right 2=up 3=down; unshifted border strip */
    switch (dir) {
        case 0: return c < W - s ? g[(size_t)r * W + c + s] : g[(size_t)r * W + c];
        case 1: return c >= s ? g[(size_t)r * W + c - s] : g[(size_t)r * W + c];
        case 2: return r < H - s ? g[(size_t)(r + s) * W + c] : g[(size_t)r * W + c];
        default: return r >= s ? g[(size_t)(r - s) * W + c] : g[(size_t)r * W + c];
    }
}

/* get_image_change_from_pil (utils.py:108-152) for S gray images [S,H,W]; lut[256] is
 * np.log(g/255*(v1-v0)+v0) computed by numpy on the caller's side; dir_mode: 0 rightdown,
 * 1 rightup, 2 leftdown, 3 leftup, 4 all.  out is [S,H,W] float32. */
int oracle_isr_batch(const uint8_t* gray, int S, int H, int W, int shift, int dir_mode, const float* lut,
                     float thr, float clip, float* out, int nthreads) {
    int err = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int s = 0; s < S; ++s) {
        const size_t n = (size_t)H * W;
        const uint8_t* g = gray + s * n;
        float* o = out + s * n;
        float* term = (float*)malloc(sizeof(float) * n);
        float* scratch = (float*)malloc(sizeof(float) * n);
        if (!term || !scratch) { err = 1; free(term); free(scratch); continue; }
        int dirs[4], nd;
        if (dir_mode == 4) { dirs[0] = 2; dirs[1] = 0; dirs[2] = 3; dirs[3] = 1; nd = 4; }   /* up,left,down,right :133-137 */
        else { dirs[0] = (dir_mode & 2) ? 0 : 1; dirs[1] = (dir_mode & 1) ? 2 : 3; nd = 2; } /* row term, col term :139-151 */
        const float div = nd == 4 ? 4.0f : 2.0f;
        for (int k = 0; k < nd; ++k) {
            for (int r = 0; r < H; ++r)
                for (int c = 0; c < W; ++c)
                    term[(size_t)r * W + c] = lut[shifted_px(g, H, W, r, c, shift, dirs[k])] - lut[g[(size_t)r * W + c]]; /* :92 */
            dead_zone_split_norm(term, (int64_t)n, thr, clip, scratch);
            if (k == 0) for (size_t i = 0; i < n; ++i) o[i] = term[i] / div;
            else for (size_t i = 0; i < n; ++i) o[i] = o[i] + term[i] / div;
        }
        free(term); free(scratch);
    }
    return err;
}

/* get_image_change (create_cityscapes_image_change.py:16-35) for S pairs; lut = np.log(g+log_add).
 * out_f32 and/or out_u8 may be NULL. */
int oracle_image_change_batch(const uint8_t* now, const uint8_t* front, int S, int H, int W, const float* lut,
                              float thr, float clip, float* out_f32, uint8_t* out_u8, int nthreads) {
    int err = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int s = 0; s < S; ++s) {
        const size_t n = (size_t)H * W;
        float* d = (float*)malloc(sizeof(float) * n);
        float* scratch = (float*)malloc(sizeof(float) * n);
        if (!d || !scratch) { err = 1; free(d); free(scratch); continue; }
        for (size_t i = 0; i < n; ++i) d[i] = lut[now[s * n + i]] - lut[front[s * n + i]];   /* :17-21 */
        dead_zone_split_norm(d, (int64_t)n, thr, clip, scratch);                           /* :22-31 */
        if (out_f32) memcpy(out_f32 + s * n, d, n * sizeof(float));
        if (out_u8) for (size_t i = 0; i < n; ++i) {
            float v = d[i] + 1.0f;
            v = v / 2.0f;
            v = v * 255.0f;
            out_u8[s * n + i] = (uint8_t)nearbyintf(v);                   /* np.around -> half to even; :33 */
        }
        free(d); free(scratch);
    }
    return err;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
