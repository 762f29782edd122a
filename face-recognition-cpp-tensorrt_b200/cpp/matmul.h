// Drop-in replacement for the reference's MatMul (/root/reference/src/matmul.h:6-38, src/matmul.cpp): C = embeds x gallery^T.
#ifndef MATMUL_H
#define MATMUL_H

#include "common.h"

class MatMul {
    /*
    A: m x k row-major (gallery), B: n x k row-major (new embeddings), C: n x m row-major with C[i*m + j] = <B_i, A_j>
    (src/matmul.h:8-15; exact fp32 like CUBLAS_COMPUTE_32F).
    */
  public:
    MatMul() {}
    ~MatMul() { fr_gallery_destroy(gallery); }
    MatMul(const MatMul &) = delete;
    MatMul &operator=(const MatMul &) = delete;

    void init(float *knownEmbeds, int numRow, int numCol) {
        fr_gallery_destroy(gallery);  // the reference leaks the previous device copy on /reload (src/matmul.cpp:17); we free it
        gallery = nullptr;
        m = numRow;
        k = numCol;
        frCheck(fr_gallery_create(knownEmbeds, numRow, numCol, device, 0, &gallery));
    }
    void calculate(float *embeds, int embedCount, float *outputs) { frCheck(fr_gallery_sims(gallery, embeds, embedCount, outputs)); }

    // fast path (not in the reference): top-k without materialising C; order (score desc, row asc) = std::max_element for k = 1
    void search(float *embeds, int embedCount, int topk, float *scores, int64_t *rows) {
        frCheck(fr_gallery_topk(gallery, embeds, embedCount, topk, scores, rows));
    }
    // gallery lifecycle without a re-upload (not in the reference, which re-stages and re-uploads everything per enrolment,
    // src/app.cpp:131-217,354-365): append rows after the last one / delete a row by moving the last row into its slot
    void append(float *embeds, int count) {
        if (!gallery) frCheck(fr_gallery_create(nullptr, 0, k ? k : 512, device, 0, &gallery));
        frCheck(fr_gallery_append(gallery, embeds, count));
        m += count;
    }
    int64_t remove(int row) {
        int64_t moved = -1;
        frCheck(fr_gallery_remove(gallery, row, &moved));
        m -= 1;
        return moved;
    }
    FrGallery *handle() const { return gallery; }
    int device = 0;

  private:
    FrGallery *gallery = nullptr;
    int m = 0, k = 0;
};

#endif  // MATMUL_H
