// tcgen05 implicit-GEMM convolution (declarations).  See conv_tc.cu.
#pragma once
#include "common.cuh"

namespace ydst {

struct ConvTcParams {
    // --- GEMM tiling ---
    int mode;          // 0: flat (stride 1; M tile = 128 consecutive padded pixels)
                       // 1: patch (stride 2; M tile = TH x TW output pixels of one image)
    int R, S;          // filter size (1x1 or 3x3)
    int block_k;       // 16 / 32 / 64 input channels per k-block (row bytes 32 / 64 / 128)
    int block_n;       // output channels per tile (multiple of 16, <= 256)
    int cin_blocks;    // Cin / block_k
    int cin;           // Cin (K offset of a tap in the packed weights = tap * cin)
    int cout;          // output channels rounded up to 16
    // --- geometry ---
    int N, Ho, Wo;     // logical output dims (padded dims are +2)
    int in_Wp;         // padded input width  (flat mode: == Wo + 2)
    int in_Hp_half;    // patch mode: padded input height / 2
    long long P_total; // flat mode: N * (Ho+2) * (Wo+2)
    int TW, TH, tiles_x, tiles_y;   // patch mode
    int TN;            // patch mode: whole images per tile (> 1 when TW x TH covers a small image)
    int pad_shift;     // patch mode: 1 - pad  (3x3: 0, 1x1: 1)
    // --- epilogue ---
    const float* scale;   // [cout rounded up to block_n]  BN gamma/sqrt(var+eps)   (1 if no BN)
    const float* bias;    // [..]                         BN beta - mean*scale     (conv bias if no BN)
    int act;
    int res_mode;         // 0 none, 1 add after activation (darknet shortcut), 2 add before activation (ReID block)
    const __half* res; int res_ctot, res_coff;
    __half* out; int out_ctot, out_coff;
    float* out_f32;       // when non-null: write fp32 [pixels][cout] instead of fp16 (YOLO head convs)
    // --- halo kernel (conv_tc2_kernel: stride 1, 64-channel blocks) ---
    int v2;               // 1: the A chunk (128 output rows + halo) is loaded ONCE per channel block and all taps read it
    int ksplit;           // K split over channel blocks (grid.z); partial sums meet in `ws`
    int cbs_per_split;    // channel blocks per split
    int halo;             // chunk rows before the tile's first pixel (3x3: Wp + 1, 1x1: 0)
    int a_box_rows, a_boxes;   // the chunk is fetched as a_boxes TMA boxes of a_box_rows rows
    int a_stages, b_stages;
    int b_resident;       // persistent only: all weight stages of an N tile fit in shared memory and stay there across its M tiles
    int persistent;       // 1: 1-D grid of at most one CTA per SM, each looping over output tiles with two TMEM accumulators
    int m_tiles, n_tiles;
    int mpair;            // 128-row accumulators per CTA tile (1 or 2): a pair of M tiles shares every weight stage
    int tpb;              // 64-channel K slices per weight stage (3x3: filter taps, 1 | 3 | 9; 1x1: channel blocks)
    int store_tma;        // 1: epilogue stages 64-column groups in swizzled smem and writes them with TMA bulk stores
    int cta2;             // 1: clusters of two CTAs along M share one tcgen05.mma.cta_group::2 stream; each stages half of the weight tile
    int nteams;           // persistent: epilogue teams of four warps that alternate tiles (2; 1 for wide tiles, whose MMAs outlast an epilogue)
    int nbuf;             // persistent: 16 KB staging buffers per epilogue team (3 when a residual tile is prefetched into them, else 2 or 3)
    int trace_tiles;      // debug: the persistent loop also records per-tile milestones of CTA 0 (YDST_CONV_TRACE=1 only)
    float* ws;            // split-K partials [tile][split][chunk][128][16] fp32
    int* tickets;         // per output tile arrival counter (self-resetting)
    unsigned long long* trace;   // debug: CTA (0,0,0) records clock64 at its pipeline milestones (YDST_CONV_TRACE=1)
    const void* pf_ptr;   // next convolution's packed weights: each CTA asks L2 to prefetch its slice once its own loads are queued
    unsigned pf_bytes;
    int pdl;              // launched with programmatic stream serialization (the producer then prefetches all weight stages first)
    int gemm;             // 1: plain row-major matrices (every row is a valid output row: no padded-pixel geometry); conv_tc_plan_gemm
    int bo_mode;          // UMMA descriptor base-offset mode for row-shifted starts (validated on hardware)
};

// split-K scratch shared by all convs of one owner (Detector / Reid): kernels of one owner run on one stream, in order
struct ConvWorkspace {
    float* partial = nullptr;
    size_t partial_bytes = 0;
    int* tickets = nullptr;
    int n_tickets = 0;
};

struct ConvTcLaunch {
    CUtensorMap tmA[4];   // flat: [0] only.  patch: one per input parity (py*2 + px)
    CUtensorMap tmB;
    ConvTcParams p;
    dim3 grid;
    int smem_bytes;
    int stages;
    const void* w_ptr;    // this convolution's packed weights (what a predecessor prefetches)
    unsigned w_bytes;
};

// Host: encode tensor maps + pick tiling.  `w_packed` is fp16 [cout16][R*S*cin] (K-major).
void conv_tc_plan(ConvTcLaunch& L, const Act& in, const Act& out, const __half* w_packed, int R, int S, int stride,
                  const float* scale, const float* bias, int act, int res_mode, const Act* res, float* out_f32, int cout_real,
                  const ConvWorkspace* ws = nullptr);
// the tiling the planner would pick for a stride-1 conv on the halo kernel (pure host function, no CUDA): for the planner test
struct ConvTiling { int bn, ksplit, cbs_per_split, tpb, a_stages, b_stages, occupancy, smem_bytes, persistent, mpair, b_resident, nbuf, cta2, nteams; double model_us; };
ConvTiling conv_tc_choose_tiling(int m_tiles, int cout16, int taps, int cin_blocks, int halo, size_t ws_bytes, int max_tickets, int res_mode = 0,
                                 bool allow_pers = true);
// Plain GEMM on the same kernel: out[rows][ncols] (fp32, row stride ncols) = A[rows][K] * B[ncols][K]^T * scale[col] + bias[col], A and B
// fp16 row-major, K a multiple of 64, rows a multiple of 128, ncols a multiple of 16.  Used by the appearance cost (cosine_tc.cu).
void conv_tc_plan_gemm(ConvTcLaunch& L, const __half* A, int rows, int K, const __half* Bw, int ncols, float* out_f32, const float* scale,
                       const float* bias, const ConvWorkspace* ws);
void conv_tc_run(const ConvTcLaunch& L, cudaStream_t stream);
void conv_tc_trace_dump();   // YDST_CONV_TRACE=2: print the per-launch timeline collected so far (debug aid)
double conv_tc_flops(const ConvTcLaunch& L);   // useful 2*M*N*K (logical, unpadded)

}  // namespace ydst
