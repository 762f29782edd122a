/* ORACLE (test infrastructure, never shipped or measured as the product): plain-C restatement of the reference's detector
 * post-processing, statement by statement, with the same types at every step (float / double / int as the C++ has them).
 *
 *   retina_anchors()      RetinaFace::create_anchor_retinaface   /root/reference/src/retinaface.cpp:210-240
 *   retina_postprocess()  RetinaFace::postprocessing (:154-208)  + RetinaFace::nms (:248-271) + m_cmp (:242-246)
 *                         scales as in the constructor (:21-22)
 *
 * Deviations, all deliberate and listed in DESIGN.md:
 *   - std::sort (:204) is not stable and its order for equal scores is unspecified; this restatement orders equal scores by
 *     anchor index (a stable sort), which is one of the orders std::sort may produce.
 *   - landmark decode is NOT reference code (the deployed model has no landmark head); it applies the reference's own centre
 *     decode (:166-167, variance 0.1) to each of the five points and undoes the letterbox in float.
 * Build: oracle/build_native.py -> oracle/_native/libretina_post.so (gcc -O2 -ffp-contract=off).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int x1, y1, x2, y2;
    float score;
} Bbox; /* src/common.h:13-16 */

typedef struct {
    float cx, cy, sx, sy;
} anchorBox; /* src/retinaface.h:10-15 */

#define CLIP(a, lo, hi) ((a) < (lo) ? (lo) : ((a) > (hi) ? (hi) : (a))) /* MAX(MIN(a, max), min), src/retinaface.h:8 */

/* returns the number of anchors; writes min(count, cap) of them */
int retina_anchors(int w, int h, float *out, int cap) {
    const float steps[3] = {8, 16, 32};
    const int min_sizes[3][2] = {{10, 20}, {32, 64}, {128, 256}};
    int n = 0;
    for (int k = 0; k < 3; ++k) {
        const int fm_h = (int)ceilf(h / steps[k]); /* ceil(h / steps[i]) on float, :215 */
        const int fm_w = (int)ceilf(w / steps[k]);
        for (int i = 0; i < fm_h; ++i)
            for (int j = 0; j < fm_w; ++j)
                for (int l = 0; l < 2; ++l) {
                    const float s_kx = min_sizes[k][l] * 1.0 / w; /* double division, rounded to float, :230 */
                    const float s_ky = min_sizes[k][l] * 1.0 / h;
                    const float cx = (j + 0.5) * steps[k] / w;
                    const float cy = (i + 0.5) * steps[k] / h;
                    if (n < cap) {
                        out[4 * n + 0] = cx;
                        out[4 * n + 1] = cy;
                        out[4 * n + 2] = s_kx;
                        out[4 * n + 3] = s_ky;
                    }
                    ++n;
                }
    }
    return n;
}

typedef struct {
    Bbox b;
    int id;
} Cand;

static void merge_sort(Cand *a, Cand *tmp, int n) { /* stable, descending score */
    if (n < 2) return;
    const int m = n / 2;
    merge_sort(a, tmp, m);
    merge_sort(a + m, tmp, n - m);
    int i = 0, j = m, k = 0;
    while (i < m && j < n) tmp[k++] = (a[j].b.score > a[i].b.score) ? a[j++] : a[i++];
    while (i < m) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(Cand) * n);
}

/* bbox [A][4], conf [A][2], landm [A][10] or NULL. out: up to cap boxes (after NMS and the max_faces cap), out_landm [.][10]
 * (x = column, y = row, frame pixels, float), out_ids = anchor index of every returned box. Returns the count. */
int retina_postprocess(const float *bbox, const float *conf, const float *landm, int n_anchor, int net_w, int net_h, int frame_w,
                       int frame_h, float nms_thr, float bbox_thr, int max_faces, Bbox *out, float *out_landm, int *out_ids, int cap) {
    const int m_INPUT_W = net_w, m_INPUT_H = net_h, m_frameWidth = frame_w, m_frameHeight = frame_h;
    const float m_scale_h = (float)m_INPUT_H / m_frameHeight; /* :21 */
    const float m_scale_w = (float)m_INPUT_W / m_frameWidth;  /* :22 */
    anchorBox *anchor = (anchorBox *)malloc(sizeof(anchorBox) * (size_t)(n_anchor > 0 ? n_anchor : 1));
    Cand *cand = (Cand *)malloc(sizeof(Cand) * (size_t)(n_anchor > 0 ? n_anchor : 1) * 2);
    if (!anchor || !cand) return -1;
    if (retina_anchors(m_INPUT_W, m_INPUT_H, (float *)anchor, n_anchor) != n_anchor) {
        free(anchor);
        free(cand);
        return -2;
    }
    int n = 0;
    for (int i = 0; i < n_anchor; ++i) {
        if (*(conf + 1) > bbox_thr) { /* strict >, :160 */
            anchorBox tmp = anchor[i];
            anchorBox tmp1;
            Bbox result;
            /* decode bbox (y - W; x - H), :166-169 : double intermediates, stored as float */
            tmp1.cx = tmp.cx + *bbox * 0.1 * tmp.sx;
            tmp1.cy = tmp.cy + *(bbox + 1) * 0.1 * tmp.sy;
            tmp1.sx = tmp.sx * exp(*(bbox + 2) * 0.2);
            tmp1.sy = tmp.sy * exp(*(bbox + 3) * 0.2);
            /* float arithmetic, truncated to int, :171-174 */
            result.y1 = (tmp1.cx - tmp1.sx / 2) * m_INPUT_W;
            result.x1 = (tmp1.cy - tmp1.sy / 2) * m_INPUT_H;
            result.y2 = (tmp1.cx + tmp1.sx / 2) * m_INPUT_W;
            result.x2 = (tmp1.cy + tmp1.sy / 2) * m_INPUT_H;
            /* rescale to original size, second truncation, :177-187 */
            if (m_scale_h > m_scale_w) {
                result.y1 = result.y1 / m_scale_w;
                result.y2 = result.y2 / m_scale_w;
                result.x1 = (result.x1 - (m_INPUT_H - m_scale_w * m_frameHeight) / 2) / m_scale_w;
                result.x2 = (result.x2 - (m_INPUT_H - m_scale_w * m_frameHeight) / 2) / m_scale_w;
            } else {
                result.y1 = (result.y1 - (m_INPUT_W - m_scale_h * m_frameWidth) / 2) / m_scale_h;
                result.y2 = (result.y2 - (m_INPUT_W - m_scale_h * m_frameWidth) / 2) / m_scale_h;
                result.x1 = result.x1 / m_scale_h;
                result.x2 = result.x2 / m_scale_h;
            }
            /* :190-193 */
            result.y1 = CLIP(result.y1, 0, m_frameWidth - 1);
            result.x1 = CLIP(result.x1, 0, m_frameHeight - 1);
            result.y2 = CLIP(result.y2, 0, m_frameWidth - 1);
            result.x2 = CLIP(result.x2, 0, m_frameHeight - 1);
            result.score = *(conf + 1);
            cand[n].b = result;
            cand[n].id = i;
            ++n;
        }
        bbox += 4;
        conf += 2;
    }
    merge_sort(cand, cand + n_anchor, n); /* std::sort(..., m_cmp), :204 */
    /* nms, :248-271 — erase() replaced by a keep flag; the visiting order and the arithmetic are unchanged */
    char *dead = (char *)calloc((size_t)(n > 0 ? n : 1), 1);
    float *vArea = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) vArea[i] = (cand[i].b.x2 - cand[i].b.x1 + 1) * (cand[i].b.y2 - cand[i].b.y1 + 1);
    for (int i = 0; i < n; ++i) {
        if (dead[i]) continue;
        for (int j = i + 1; j < n; ++j) {
            if (dead[j]) continue;
            float xx1 = cand[i].b.x1 > cand[j].b.x1 ? cand[i].b.x1 : cand[j].b.x1;
            float yy1 = cand[i].b.y1 > cand[j].b.y1 ? cand[i].b.y1 : cand[j].b.y1;
            float xx2 = cand[i].b.x2 < cand[j].b.x2 ? cand[i].b.x2 : cand[j].b.x2;
            float yy2 = cand[i].b.y2 < cand[j].b.y2 ? cand[i].b.y2 : cand[j].b.y2;
            float w = 0.f > xx2 - xx1 + 1 ? 0.f : xx2 - xx1 + 1;
            float h = 0.f > yy2 - yy1 + 1 ? 0.f : yy2 - yy1 + 1;
            float inter = w * h;
            float ovr = inter / (vArea[i] + vArea[j] - inter);
            if (ovr >= nms_thr) dead[j] = 1;
        }
    }
    int kept = 0;
    for (int i = 0; i < n && kept < cap; ++i) {
        if (dead[i]) continue;
        if (kept >= max_faces) break; /* resize(m_maxFacesPerScene), :206-207 */
        out[kept] = cand[i].b;
        if (out_ids) out_ids[kept] = cand[i].id;
        if (out_landm) {
            float *lm = out_landm + 10 * kept;
            if (landm) {
                const anchorBox a = anchor[cand[i].id];
                const float *l = landm + 10 * (size_t)cand[i].id;
                for (int p = 0; p < 5; ++p) {
                    const float lx = a.cx + l[2 * p] * 0.1 * a.sx;
                    const float ly = a.cy + l[2 * p + 1] * 0.1 * a.sy;
                    float px = lx * m_INPUT_W, py = ly * m_INPUT_H;
                    if (m_scale_h > m_scale_w) {
                        px = px / m_scale_w;
                        py = (py - (m_INPUT_H - m_scale_w * m_frameHeight) / 2) / m_scale_w;
                    } else {
                        px = (px - (m_INPUT_W - m_scale_h * m_frameWidth) / 2) / m_scale_h;
                        py = py / m_scale_h;
                    }
                    lm[2 * p] = px;
                    lm[2 * p + 1] = py;
                }
            } else {
                memset(lm, 0, sizeof(float) * 10);
            }
        }
        ++kept;
    }
    free(dead);
    free(vArea);
    free(anchor);
    free(cand);
    return kept;
}
