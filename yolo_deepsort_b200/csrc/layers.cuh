// HBM-bound layer kernels (declarations).  See layers.cu.
#pragma once
#include "common.cuh"

namespace ydst {

// u8 HWC frame -> fp32 HWC in [0,1] (x / 255, yolo3/detect/img_detect.py:71-72)
void launch_u8_to_f32(const uint8_t* src, float* dst, long long n, cudaStream_t st);
// NCHW (fp32 or fp16) -> fp32 NHWC, for the Darknet.forward(x) API
void launch_nchw_to_nhwc(const void* src, int src_is_half, float* dst, int N, int C, int H, int W, cudaStream_t st);

// First-layer direct convolution: fp32 NHWC input with 3 channels (unpadded), 3x3 pad 1, stride 1|2,
// fused scale/bias/activation, fp16 padded-NHWC output.  w is fp32 [27][cout] (tap-major, then cin).
// w_hilo (optional): [2][cout][32] fp16 hi/lo split of the same weights; with it, stride 1 and cout in {16,32,48,64} the layer runs
// on the tensor cores (three tcgen05 products hi*hi + lo*hi + hi*lo of a [128 x 32] im2col tile built in shared memory).
void launch_conv_first(const float* in, int N, int H, int W, const float* w, const __half* w_hilo, const float* scale, const float* bias,
                       int cout, int stride, int act, const Act& out, cudaStream_t st);

// MaxPool2d(k, s, padding=(k-1)//2) with -inf padding; zero_pad_br=1 reproduces the reference's
// ZeroPad2d((0,1,0,1)) + MaxPool2d(2,1) (yolo3/models/models.py:58-64).
void launch_maxpool(const Act& in, const Act& out, int k, int stride, int zero_pad_br, cudaStream_t st);
// first layer (Cin = 3, stride 1, tensor cores) fused with MaxPool2d(3, 2, 1): the ReID stem without the full-resolution round trip
bool conv_first_pool_supported(int H, int W, int cout, int stride, int act, const __half* w_hilo);
void launch_conv_first_pool(const float* in, int N, int H, int W, const __half* w_hilo, const float* scale, const float* bias, int cout, int act,
                            const Act& out, cudaStream_t st);
// nearest-neighbour upsample by `s` (UpsampleExpand, yolo3/models/models.py:118-133)
void launch_upsample(const Act& in, const Act& out, int s, cudaStream_t st);
// out = a + b (unfused shortcut, yolo3/models/models.py:304-306)
void launch_add(const Act& a, const Act& b, const Act& out, cudaStream_t st);
// channel-slice copy (route/concat fallback when a producer cannot write in place)
void launch_copy(const Act& in, const Act& out, cudaStream_t st);

// dense NHWC fp16 <-> flat-padded view; fp32 padded -> dense (boundary / test helpers)
void launch_pack(const __half* dense, const Act& padded, cudaStream_t st);
void launch_unpack(const Act& padded, __half* dense, cudaStream_t st);
void launch_unpack_f32(const float* padded, int cstride, int N, int H, int W, int C, float* dense, cudaStream_t st);

// YOLO decode (YOLOLayer.forward inference branch, yolo3/models/models.py:185-224).
// head: fp32 [N*(g+2)*(g+2)][cstride] padded-pixel layout; writes rows [row0, row0 + 3*gy*gx) of pred[N][rows_total][5+nc].
void launch_yolo_decode(const float* head, int cstride, int N, int gy, int gx, int na, const float* anchors_wh, int nc,
                        int img_h, int img_w, float* pred, int rows_total, int row0, cudaStream_t st);

// ReID tail: AvgPool2d((8,4),1) + view + x / ||x||_2 (deep_sort/deep/model.py:86-92): in (B,8,4,512) -> out fp32 [B][512]
void launch_avgpool_l2(const Act& in, float* out, cudaStream_t st);

// Crop + cv2-exact fixed-point bilinear resize to 128x64 + /255 + ImageNet mean/std (deep_sort/deep_sort.py:116-141,
// deep_sort/deep/feature_extractor.py:34-51).  frame: u8 HWC RGB; boxes: tlwh fp32 [m][4]; out fp32 NHWC [m][128][64][3].
// err_flag (device int) is set to 1 if a box yields an empty crop (the reference raises there).
// cv2.resize(u8 HWC, INTER_LINEAR) of a whole frame, optionally swapping R and B (BGR capture -> RGB); same size = (swapping) copy
void launch_resize_u8(const uint8_t* src, int sh, int sw, uint8_t* dst, int dh, int dw, int swap_rb, cudaStream_t st, long long pitch = 0);
void launch_window_boxes(float* pred, int tiles, int rows, int nf, const float* geo_dev, cudaStream_t st);   // geo: [tiles][4] = rw, rh, ox, oy
void launch_crop_resize(const uint8_t* frame, int H, int W, const float* tlwh, int m, float* out, int* err_flag, cudaStream_t st);
// crops of up to 8 frames in one launch: segment f holds crops [start[f], start[f+1]) of the launch, cut from frame[f]
struct CropBatch {
    const uint8_t* frame[8];
    const float* tlwh[8];
    int* err[8];
    int H[8], W[8];
    int start[9];
    int n;
};
void launch_crop_resize_multi(const CropBatch& cb, float* out, cudaStream_t st);

}  // namespace ydst
