// Detection post-processing (declarations).  See nms.cu.
#pragma once
#include "common.cuh"

namespace ydst {

struct NmsCand {
    float x1, y1, x2, y2, score, cls;
    int key;   // row * nc + class : position in the reference's row-major (box, class) expansion
    int pad;
};

// All buffers carry a leading image dimension (`batch` images per launch set: blockIdx.y / .x selects the image), so a
// micro-batch of frames costs one set of launches instead of one per frame.  Image b's results sit at counters + 8*b,
// dets + b*max_det*6 (image 0 first, so single-image callers index them as before).
struct NmsRatios { float rw[8], rh[8]; };
// soft_non_max_suppression's keyword options (yolo3/utils/model_build.py:52-53); all zero = the video path
struct NmsOptions {
    int p1p2 = 0;                              // boxes are corners already (is_p1p2)
    int merge = 0;                             // the reference's "Merge NMS" block, as it actually executes (nms.cu)
    int agnostic = 0;                          // no class offset
    int use_classes = 0;                       // keep only the classes whose bit is set
    unsigned long long class_bits[4] = {0, 0, 0, 0};
};
struct Nms {
    int cap = 0, max_det = 0, words = 0, batch = 1;
    NmsCand* cand = nullptr;   // [batch][cap]
    NmsCand* sorted = nullptr; // [batch][cap]
    unsigned long long* mask = nullptr;   // [batch][cap][words]
    int* counters = nullptr;   // [batch][8]: [0] candidates, [1] kept (<= max_det), [2] overflow flag, [3] tracker inputs m
    float* dets = nullptr;     // [batch][max_det][6]  x1,y1,x2,y2,conf,cls  (score-descending)
    void init(int cap, int max_det, int batch = 1);
    void destroy();
    // pred: [nb][rows][nf]
    void run(const float* pred, int rows, int nf, float conf, float iou, cudaStream_t st, int nb = 1, const NmsOptions* options = nullptr);
    // tlwh/conf/cls: [nb][max_det](x4)
    void to_tracker_inputs(float rw, float rh, const int* class_mask_dev, int n_mask, float* tlwh, float* conf, float* cls,
                           cudaStream_t st);
    void to_tracker_inputs_batch(const NmsRatios& r, int nb, const int* class_mask_dev, int n_mask, float* tlwh, float* conf, float* cls,
                                 cudaStream_t st);
};

}  // namespace ydst
