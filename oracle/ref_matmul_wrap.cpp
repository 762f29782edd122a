// extern "C" handle around the reference's own MatMul class (compiled verbatim from /root/reference/src/matmul.cpp)
// so that tests and bench.py can drive it through ctypes. ORACLE / BASELINE ONLY — never linked into libfr_b200.
#include "matmul.h"

extern "C" {
void *ref_matmul_new() {
    try {
        return new MatMul();
    } catch (...) {
        return nullptr;
    }
}
int ref_matmul_init(void *h, float *known, int rows, int cols) {
    try {
        static_cast<MatMul *>(h)->init(known, rows, cols);
        return 0;
    } catch (...) {
        return -1;
    }
}
int ref_matmul_calculate(void *h, float *embeds, int count, float *outputs) {
    try {
        static_cast<MatMul *>(h)->calculate(embeds, count, outputs);
        return 0;
    } catch (...) {
        return -1;
    }
}
void ref_matmul_free(void *h) {
    try {
        delete static_cast<MatMul *>(h);
    } catch (...) {
    }
}
}
