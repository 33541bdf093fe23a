// Detection post-processing (declarations).  See nms.cu.
#pragma once
#include "common.cuh"

namespace ydst {

struct NmsCand {
    float x1, y1, x2, y2, score, cls;
    int key;   // row * nc + class : position in the reference's row-major (box, class) expansion
    int pad;
};

struct Nms {
    int cap = 0, max_det = 0, words = 0;
    NmsCand* cand = nullptr;
    NmsCand* sorted = nullptr;
    unsigned long long* mask = nullptr;
    int* counters = nullptr;   // [0] candidates, [1] kept (<= max_det), [2] overflow flag, [3] tracker inputs m
    float* dets = nullptr;     // [max_det][6]  x1,y1,x2,y2,conf,cls  (score-descending)
    void init(int cap, int max_det);
    void destroy();
    void run(const float* pred, int rows, int nf, float conf, float iou, cudaStream_t st);
    void to_tracker_inputs(float rw, float rh, const int* class_mask_dev, int n_mask, float* tlwh, float* conf, float* cls,
                           cudaStream_t st);
};

}  // namespace ydst
