/*
 * scda_oracle.c — CPU restatement of the SCDA operator hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under scda_b200/ may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker and as the
 * timed CPU baseline.
 *
 * Every function restates the arithmetic of one reference CUDA kernel (path
 * and line range cited, relative to /root/reference/), one output element per
 * loop iteration, with the implicit float/double promotions of the original
 * expressions written out.  Where nvcc 12.9 contracts a multiply-add of the
 * reference source into one fma (read from `nvcc -ptx` of the unmodified
 * file), the same fma is written explicitly so that this file compiled with
 * -ffp-contract=off produces the bits the reference kernel produces.
 *
 * Parity pin: the reference ships no golden vectors, so these
 * functions are pinned on the GPU box against the reference's own .cu files
 * compiled unmodified (oracle/_ref/libscda_ref.so, tests/test_ref_parity.py)
 * and here against the reference's real cython_bbox (oracle/_ref) and
 * committed fixtures under tests/golden/.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC (oracle/build.py).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* Thread count of the OpenMP loops below (test / bench infrastructure: torchrun exports
 * OMP_NUM_THREADS=1, which would otherwise time the baseline on one core). */
ORACLE_API int oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* ------------------------------------------------------------------ */
/* RoI max pooling                                                      */
/* ------------------------------------------------------------------ */

/* One scaled+rounded RoI, as ROIPoolForward derives it
 * (extensions/_roi_pooling/src/roi_pooling_kernel.cu:45-56). */
typedef struct {
    int batch, x0, y0, x1, y1;
    float bin_h, bin_w;
} pool_roi_t;

static pool_roi_t pool_roi(const float *r, float scale, int ph, int pw)
{
    pool_roi_t q;
    q.batch = (int)r[0];
    q.x0 = (int)roundf(r[1] * scale);
    q.y0 = (int)roundf(r[2] * scale);
    q.x1 = (int)roundf(r[3] * scale);
    q.y1 = (int)roundf(r[4] * scale);
    int rw = (int)fmaxf((float)(q.x1 - q.x0 + 1), 1.0f);
    int rh = (int)fmaxf((float)(q.y1 - q.y0 + 1), 1.0f);
    q.bin_h = (float)rh / (float)ph;
    q.bin_w = (float)rw / (float)pw;
    return q;
}

static int clampi(int v, int lo, int hi)
{
    /* fminf(fmaxf(v, lo), hi) on ints promoted to float, cast back */
    return (int)fminf(fmaxf((float)v, (float)lo), (float)hi);
}

/* roi_pooling_kernel.cu:24-93 (ROIPoolForward). argmax may be NULL. */
ORACLE_API void oracle_roi_pool_forward(const float *feat, float scale, int num_rois,
                                        int H, int W, int C, int PH, int PW,
                                        const float *rois, float *out, int *argmax)
{
#pragma omp parallel for schedule(static)
    for (int n = 0; n < num_rois; ++n) {
        pool_roi_t q = pool_roi(rois + 5 * n, scale, PH, PW);
        for (int c = 0; c < C; ++c) {
            const int plane = (q.batch * C + c) * H * W;
            for (int ph = 0; ph < PH; ++ph) {
                int hs = (int)floorf((float)ph * q.bin_h);
                int he = (int)ceilf((float)(ph + 1) * q.bin_h);
                hs = clampi(hs + q.y0, 0, H);
                he = clampi(he + q.y0, 0, H);
                for (int pw = 0; pw < PW; ++pw) {
                    int ws = (int)floorf((float)pw * q.bin_w);
                    int we = (int)ceilf((float)(pw + 1) * q.bin_w);
                    ws = clampi(ws + q.x0, 0, W);
                    we = clampi(we + q.x0, 0, W);
                    int empty = (he <= hs) || (we <= ws);
                    float best = empty ? 0.0f : -FLT_MAX;
                    int where = -1;
                    for (int h = hs; h < he; ++h)
                        for (int w = ws; w < we; ++w) {
                            float v = feat[plane + h * W + w];
                            if (v > best) { best = v; where = plane + h * W + w; }
                        }
                    size_t o = (((size_t)n * C + c) * PH + ph) * PW + pw;
                    out[o] = best;
                    if (argmax) argmax[o] = where;
                }
            }
        }
    }
}

/* roi_pooling_kernel.cu:128-203 (ROIPoolBackward): gather form, one input
 * element at a time, RoIs in ascending order — the summation order here is
 * the reference's. */
ORACLE_API void oracle_roi_pool_backward(const float *top_diff, float scale, int batch,
                                         int num_rois, int H, int W, int C, int PH, int PW,
                                         const float *rois, float *bottom_diff,
                                         const int *argmax)
{
    pool_roi_t *q = (pool_roi_t *)malloc(sizeof(pool_roi_t) * (size_t)(num_rois > 0 ? num_rois : 1));
    for (int n = 0; n < num_rois; ++n) q[n] = pool_roi(rois + 5 * n, scale, PH, PW);
    const int total = batch * C * H * W;
#pragma omp parallel for schedule(static)
    for (int index = 0; index < total; ++index) {
        int w = index % W, h = (index / W) % H, c = (index / W / H) % C, b = index / W / H / C;
        float g = 0.0f;
        for (int n = 0; n < num_rois; ++n) {
            if (q[n].batch != b) continue;
            if (!(w >= q[n].x0 && w <= q[n].x1 && h >= q[n].y0 && h <= q[n].y1)) continue;
            int p0 = (int)floorf((float)(h - q[n].y0) / q[n].bin_h);
            int p1 = (int)ceilf((float)(h - q[n].y0 + 1) / q[n].bin_h);
            int r0 = (int)floorf((float)(w - q[n].x0) / q[n].bin_w);
            int r1 = (int)ceilf((float)(w - q[n].x0 + 1) / q[n].bin_w);
            p0 = clampi(p0, 0, PH); p1 = clampi(p1, 0, PH);
            r0 = clampi(r0, 0, PW); r1 = clampi(r1, 0, PW);
            size_t base = ((size_t)n * C + c) * PH * PW;
            for (int ph = p0; ph < p1; ++ph)
                for (int pw = r0; pw < r1; ++pw)
                    if (argmax[base + ph * PW + pw] == index)
                        g += top_diff[base + ph * PW + pw];
        }
        bottom_diff[index] = g;
    }
    free(q);
}

/* ------------------------------------------------------------------ */
/* RoIAlign (single bilinear sample per grid point)                     */
/* ------------------------------------------------------------------ */

typedef struct {
    int ok;          /* sample inside [0,H) x [0,W) */
    int ul;          /* offset of the up-left tap inside one channel plane */
    float hr, wr;    /* fractional parts */
} align_tap_t;

/* roi_align_kernel.cu:33-53.  nvcc folds `x + 1.` / `fmaxf(., 0.)` back to
 * float, keeps the two bin-size divisions in double, and contracts
 * ph*bin + start into one float fma. */
static align_tap_t align_tap(const float *r, float scale, int H, int W, int AH, int AW,
                             int ph, int pw)
{
    float x0 = r[1] * scale, y0 = r[2] * scale, x1 = r[3] * scale, y1 = r[4] * scale;
    float rw = fmaxf((x1 - x0) + 1.0f, 0.0f);
    float rh = fmaxf((y1 - y0) + 1.0f, 0.0f);
    float bin_h = (float)((double)rh / ((double)AH - 1.0));
    float bin_w = (float)((double)rw / ((double)AW - 1.0));
    float h = fmaf((float)ph, bin_h, y0);
    float w = fmaf((float)pw, bin_w, x0);
    align_tap_t t;
    t.ok = !(h < 0 || h >= H || w < 0 || w >= W);
    int hs = (int)fminf(floorf(h), (float)(H - 2));
    int ws = (int)fminf(floorf(w), (float)(W - 2));
    t.hr = h - (float)hs;
    t.wr = w - (float)ws;
    t.ul = hs * W + ws;
    return t;
}

/* roi_align_kernel.cu:15-70 (ROIAlignForward). */
ORACLE_API void oracle_roi_align_forward(const float *feat, float scale, int num_rois,
                                         int H, int W, int C, int AH, int AW,
                                         const float *rois, float *out)
{
#pragma omp parallel for schedule(static)
    for (int n = 0; n < num_rois; ++n) {
        const float *r = rois + 5 * n;
        const int img = (int)(r[0] * (float)C * (float)H * (float)W);
        for (int ph = 0; ph < AH; ++ph)
            for (int pw = 0; pw < AW; ++pw) {
                align_tap_t t = align_tap(r, scale, H, W, AH, AW, ph, pw);
                for (int c = 0; c < C; ++c) {
                    size_t o = (((size_t)n * C + c) * AH + ph) * AW + pw;
                    if (!t.ok) { out[o] = 0.0f; continue; }
                    const float *p = feat + img + c * H * W + t.ul;
                    double omh = 1.0 - (double)t.hr, omw = 1.0 - (double)t.wr;
                    /* (ul*(1-h))*(1-w) + ((ur*(1-h))*w) fused, then the two
                     * float-product terms, as the PTX orders them */
                    double acc = fma(omw, omh * (double)p[0], (omh * (double)p[1]) * (double)t.wr);
                    acc = fma(omw, (double)(t.hr * p[W]), acc);
                    acc = acc + (double)(t.wr * (t.hr * p[W + 1]));
                    out[o] = (float)acc;
                }
            }
    }
}

/* roi_align_kernel.cu:94-143 (ROIAlignBackward).  atomicAdd becomes +=, in
 * output-index order; the reference's order is scheduling dependent. */
ORACLE_API void oracle_roi_align_backward(const float *top_diff, float scale, int batch,
                                          int num_rois, int H, int W, int C, int AH, int AW,
                                          const float *rois, float *bottom_diff)
{
    (void)batch;
    /* parallel over channels: each channel plane is private to one thread */
#pragma omp parallel for schedule(static)
    for (int c = 0; c < C; ++c)
        for (int n = 0; n < num_rois; ++n) {
            const float *r = rois + 5 * n;
            const int img = (int)(r[0] * (float)C * (float)H * (float)W);
            for (int ph = 0; ph < AH; ++ph)
                for (int pw = 0; pw < AW; ++pw) {
                    align_tap_t t = align_tap(r, scale, H, W, AH, AW, ph, pw);
                    if (!t.ok) continue;
                    float g = top_diff[(((size_t)n * C + c) * AH + ph) * AW + pw];
                    float *p = bottom_diff + img + c * H * W + t.ul;
                    double omh = 1.0 - (double)t.hr;
                    float omw = 1.0f - t.wr;
                    p[0]     += (float)(((double)g * omh) * (double)omw);
                    p[1]     += (float)(((double)g * omh) * (double)t.wr);
                    p[W]     += (g * t.hr) * omw;
                    p[W + 1] += (g * t.hr) * t.wr;
                }
        }
}

/* ------------------------------------------------------------------ */
/* NMS                                                                  */
/* ------------------------------------------------------------------ */

/* devIoU (extensions/_nms/src/cuda/nms_kernel.cu:16-24) with the +1 box
 * convention.  nvcc contracts Sa + Sb into fma(wb, hb, Sa) and keeps the
 * rounded product for interS. */
static float nms_iou(const float *a, const float *b)
{
    float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    float w = fmaxf((right - left) + 1.0f, 0.0f), h = fmaxf((bottom - top) + 1.0f, 0.0f);
    float inter = w * h;
    float Sa = ((a[2] - a[0]) + 1.0f) * ((a[3] - a[1]) + 1.0f);
    float den = fmaf((b[2] - b[0]) + 1.0f, (b[3] - b[1]) + 1.0f, Sa) - inter;
    return inter / den;
}

/* nms_kernel.cu:26-70: the N x ceil(N/64) suppression bitmask. */
ORACLE_API void oracle_nms_mask(int n, const float *boxes5, float thresh, uint64_t *mask)
{
    const int cb = (n + 63) / 64;
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < n; ++i)
        for (int cblk = 0; cblk < cb; ++cblk) {
            uint64_t t = 0;
            int cols = n - cblk * 64 < 64 ? n - cblk * 64 : 64;
            int start = (i / 64 == cblk) ? (i % 64) + 1 : 0;
            for (int j = start; j < cols; ++j)
                if (nms_iou(boxes5 + 5 * (size_t)i, boxes5 + 5 * (size_t)(cblk * 64 + j)) > thresh)
                    t |= 1ULL << j;
            mask[(size_t)i * cb + cblk] = t;
        }
}

/* extensions/_nms/src/nms_cuda.c:41-58: sequential scan of the bitmask. */
ORACLE_API long oracle_nms_scan(int n, const uint64_t *mask, int64_t *keep)
{
    const int cb = (n + 63) / 64;
    uint64_t *remv = (uint64_t *)calloc((size_t)(cb > 0 ? cb : 1), sizeof(uint64_t));
    long kept = 0;
    for (int i = 0; i < n; ++i) {
        int blk = i / 64, bit = i % 64;
        if (!(remv[blk] & (1ULL << bit))) {
            keep[kept++] = i;
            const uint64_t *row = mask + (size_t)i * cb;
            for (int j = blk; j < cb; ++j) remv[j] |= row[j];
        }
    }
    free(remv);
    return kept;
}

/* gpu_nms end to end (nms_cuda.c:17-67) without materialising the mask:
 * greedy NMS over pre-sorted boxes, suppress iff IoU > thresh (strict).
 * Equivalent to oracle_nms_mask + oracle_nms_scan; used where N is large. */
ORACLE_API long oracle_nms(int n, const float *boxes5, float thresh, int64_t *keep)
{
    unsigned char *dead = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
    long kept = 0;
    for (int i = 0; i < n; ++i) {
        if (dead[i]) continue;
        keep[kept++] = i;
        const float *a = boxes5 + 5 * (size_t)i;
        for (int j = i + 1; j < n; ++j)
            if (!dead[j] && nms_iou(a, boxes5 + 5 * (size_t)j) > thresh) dead[j] = 1;
    }
    free(dead);
    return kept;
}

/* cpu_nms (extensions/_nms/src/nms.c:4-68): explicit order/areas, suppress
 * iff IoU >= thresh.  Unreachable from the reference's Python, kept as the
 * config-5 host comparator. */
ORACLE_API long oracle_cpu_nms(int n, const float *boxes5, const int64_t *order,
                               const float *areas, float thresh, int64_t *keep)
{
    unsigned char *dead = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
    long kept = 0;
    for (int _i = 0; _i < n; ++_i) {
        int64_t i = order[_i];
        if (dead[i]) continue;
        keep[kept++] = i;
        const float *a = boxes5 + 5 * i;
        for (int _j = _i + 1; _j < n; ++_j) {
            int64_t j = order[_j];
            if (dead[j]) continue;
            const float *b = boxes5 + 5 * j;
            float xx1 = fmaxf(a[0], b[0]), yy1 = fmaxf(a[1], b[1]);
            float xx2 = fminf(a[2], b[2]), yy2 = fminf(a[3], b[3]);
            float w = fmaxf(0.0f, xx2 - xx1 + 1), h = fmaxf(0.0f, yy2 - yy1 + 1);
            float inter = w * h;
            float ovr = inter / (areas[i] + areas[j] - inter);
            if (ovr >= thresh) dead[j] = 1;
        }
    }
    free(dead);
    return kept;
}

/* ------------------------------------------------------------------ */
/* IoU matrices                                                         */
/* ------------------------------------------------------------------ */

/* extensions/_cython_bbox/cython_bbox.pyx:44-72: no +1, zero unless both
 * overlaps are > 0, no clamp; plain float ops (gcc x86-64, no fma). */
ORACLE_API void oracle_bbox_overlaps(int N, const float *boxes, int K, const float *query,
                                     float *out)
{
    memset(out, 0, sizeof(float) * (size_t)N * (size_t)K);
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n) {
        const float *b = boxes + 4 * (size_t)n;
        for (int k = 0; k < K; ++k) {
            const float *q = query + 4 * (size_t)k;
            float qa = (q[2] - q[0]) * (q[3] - q[1]);
            float iw = fminf(b[2], q[2]) - fmaxf(b[0], q[0]);
            if (iw > 0) {
                float ih = fminf(b[3], q[3]) - fmaxf(b[1], q[1]);
                if (ih > 0) {
                    float ua = ((b[2] - b[0]) * (b[3] - b[1]) + qa) - iw * ih;
                    out[(size_t)n * K + k] = iw * ih / ua;
                }
            }
        }
    }
}

/* extensions/_bbox_helper/src/cuda/iou_overlap_kernel.cu:33-65: no +1,
 * union clamped to >= 1; nvcc contracts area1 + area2 into
 * fma(w1, h1, area2). `stride` is the row length of both inputs. */
ORACLE_API void oracle_iou_overlap(const float *b1, const float *b2, int stride, int n1, int n2,
                                   float *out)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n1; ++i) {
        const float *a = b1 + (size_t)i * stride;
        for (int j = 0; j < n2; ++j) {
            const float *b = b2 + (size_t)j * stride;
            float area2 = (b[2] - b[0]) * (b[3] - b[1]);
            float w = fmaxf(fminf(a[2], b[2]) - fmaxf(a[0], b[0]), 0.0f);
            float h = fmaxf(fminf(a[3], b[3]) - fmaxf(a[1], b[1]), 0.0f);
            float inter = w * h;
            float uni = fmaxf(fmaf(a[2] - a[0], a[3] - a[1], area2) - inter, 1.0f);
            out[(size_t)i * n2 + j] = inter / uni;
        }
    }
}

/* ------------------------------------------------------------------ */
/* Focal losses                                                         */
/* ------------------------------------------------------------------ */

/* -x*[x>=0] - log(1 + exp(x - 2x*[x>=0])), the double/float mix of
 * focal_loss_sigmoid_kernel.cu:38-41. */
static double log1m_sigmoid(float x)
{
    double pos = (double)(x >= 0);
    float e = expf((float)((double)x - 2.0 * (double)x * pos));
    return -1.0 * (double)x * pos - (double)logf((float)(1.0 + (double)e));
}

/* focal_loss_sigmoid_kernel.cu:12-47.  n = M * num_classes elements. */
ORACLE_API void oracle_sigmoid_focal_forward(int n, const float *logits, const int *targets,
                                             float weight_pos, float gamma, float alpha,
                                             int num_classes, float *losses)
{
    float Np = (float)fmax((double)weight_pos, 1.0);
    float zn = (float)((1.0 - (double)alpha) / (double)Np);
    float zp = alpha / Np;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        int d = i % num_classes, t = targets[i / num_classes];
        float c1 = (float)(t == d + 1);
        float c2 = (float)((t != -1) & (t != d + 1));
        float x = logits[i];
        float p = (float)(1.0 / (1.0 + (double)expf(-x)));
        float term1 = powf((float)(1.0 - (double)p), gamma) * logf(fmaxf(p, FLT_MIN));
        float term2 = (float)((double)powf(p, gamma) * log1m_sigmoid(x));
        float l = 0.0f;
        l += -c1 * term1 * zp;
        l += -c2 * term2 * zn;
        losses[i] = l;
    }
}

/* focal_loss_sigmoid_kernel.cu:49-81. */
ORACLE_API void oracle_sigmoid_focal_backward(int n, const float *logits, const int *targets,
                                              float *dX, float weight_pos, float gamma,
                                              float alpha, int num_classes)
{
    float Np = (float)fmax((double)weight_pos, 1.0);
    float zn = (float)((1.0 - (double)alpha) / (double)Np);
    float zp = alpha / Np;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        int d = i % num_classes, t = targets[i / num_classes];
        float c1 = (float)(t == d + 1);
        float c2 = (float)((t != -1) & (t != d + 1));
        float x = logits[i];
        float p = (float)(1.0 / (1.0 + (double)expf(-x)));
        float term1 = (float)((double)powf((float)(1.0 - (double)p), gamma) *
                              (1.0 - (double)p - (double)(p * gamma * logf(fmaxf(p, FLT_MIN)))));
        float term2 = (float)((double)powf(p, gamma) *
                              (log1m_sigmoid(x) * (1.0 - (double)p) * (double)gamma - (double)p));
        float g = 0.0f;
        g += -c1 * zp * term1;
        g += -c2 * zn * term2;
        dX[i] = g;
    }
}

/* focal_loss_softmax_kernel.cu:12-57 (SpatialSoftmaxKernel +
 * SoftmaxFocalLossKernel).  n = M * num_classes; losses has M entries,
 * priors n entries. */
ORACLE_API void oracle_softmax_focal_forward(int n, const float *logits, const int *targets,
                                             float weight_pos, float gamma, float alpha,
                                             int num_classes, float *losses, float *priors)
{
    const int M = n / num_classes;
    float Np = (float)fmax((double)weight_pos, 1.0);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < M; ++i) {
        const float *x = logits + (size_t)i * num_classes;
        float *P = priors + (size_t)i * num_classes;
        float mx = -FLT_MAX;
        for (int c = 0; c < num_classes; ++c) mx = fmaxf(mx, x[c]);
        float sum = 0.0f;
        for (int c = 0; c < num_classes; ++c) { P[c] = expf(x[c] - mx); sum += P[c]; }
        for (int c = 0; c < num_classes; ++c) P[c] /= sum;
        int label = targets[i];
        float z = (float)(label == 0) * (1 - alpha) / Np + (float)(label >= 1) * alpha / Np;
        losses[i] = 0.0f;
        if (label >= 0) {
            float pl = P[label];
            /* powf(1.0 - p, gamma): double subtraction narrowed to float;
             * log() of a float argument resolves to logf in CUDA C++ */
            losses[i] = -(powf((float)(1.0 - (double)pl), gamma) * logf(fmaxf(pl, FLT_MIN))) * z;
        }
    }
}

/* focal_loss_softmax_kernel.cu:59-100 (GradientWeight + Gradient kernels). */
ORACLE_API void oracle_softmax_focal_backward(int n, const float *logits, const int *targets,
                                              float *dX, float weight_pos, float gamma,
                                              float alpha, int num_classes, const float *priors,
                                              float *buff)
{
    (void)logits;
    const int M = n / num_classes;
    float Np = (float)fmax((double)weight_pos, 1.0);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < M; ++i) {
        int label = targets[i];
        float z = (float)(label == 0) * (1 - alpha) / Np + (float)(label >= 1) * alpha / Np;
        buff[i] = 0.0f;
        if (label >= 0) {
            float p = priors[(size_t)i * num_classes + label];
            float onemp = (float)(1.0 - (double)p);
            buff[i] = (-powf(onemp, gamma) +
                       gamma * powf(onemp, gamma - 1) * p * logf(fmaxf(p, FLT_MIN))) * z;
        }
        for (int c = 0; c < num_classes; ++c) {
            float c1 = (float)(label >= 0), c2 = (float)(label == c);
            dX[(size_t)i * num_classes + c] = c1 * buff[i] * (c2 - priors[(size_t)i * num_classes + c]);
        }
    }
}

ORACLE_API int oracle_abi_version(void) { return 1; }
