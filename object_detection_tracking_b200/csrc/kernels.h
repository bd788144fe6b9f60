// Parameter blocks and launchers of the non-GEMM kernels (proposals, ROIAlign, detection post-process,
// appearance cost).  All launches are asynchronous on the given stream; nothing syncs with the host.
#pragma once
#include "common.h"

namespace b2 {

// stem.cu
int f32_to_planes(const float* src, __half* hi, __half* lo, size_t n, cudaStream_t s);
int planes_to_f32(const __half* hi, const __half* lo, float* dst, size_t n, cudaStream_t s);
// norm_mode 0: detector (x * (1/255) - mean) / std with BGR constants (models.py:345-355);
// norm_mode 1: torchvision ToTensor + Normalize (x / 255 - mean) / std with RGB constants (feature_extractor.py:190-196)
int resize_u8_launch(const uint8_t* src, int B, int sh, int sw, float* dst, int dh, int dw, cudaStream_t s);
int stem_pack_launch(const void* img, int is_u8, int B, int H, int W, __half* out_hi, __half* out_lo, int Hu, int Wu,
                     int norm_mode, cudaStream_t s);
// compact operand of the detector stem: [B][Hu][Wv = Wu + 3][16] (stem.cu)
int stem_pack16_launch(const void* img, int is_u8, int B, int H, int W, __half* out_hi, __half* out_lo, int Hu, int Wv,
                       cudaStream_t s);
int maxpool_launch(const __half* in_hi, const __half* in_lo, int B, int H, int W, int C, __half* out_hi,
                   __half* out_lo, int Ho, int Wo, cudaStream_t s);

// rpn.cu
struct RpnParams {
  const float* logits[5];   // per level [B * h*w][16]: cols 0..2 objectness, 3..14 deltas (anchor-major)
  int h[5], w[5];
  float stride[5];
  float cell[5][3][4];      // cell anchors x1,y1,x2,y2 (generate_anchors.py semantics, before the +1)
  int B, topk;
  float img_h, img_w, decode_clip, min_size, nms_thr;
  int multi;                // batch-graph semantics: no min-size filter, zero-padded merge, zero-area drop
  float* lvl_boxes;         // [B][5][topk][4]
  float* lvl_scores;        // [B][5][topk]
  int* lvl_count;           // [B][5]
  float* prop_boxes;        // [B][topk][4]
  float* prop_scores;       // [B][topk]
  int* prop_count;          // [B]
};
int rpn_proposals_launch(const RpnParams& p, cudaStream_t s);

// roialign.cu
struct RoiAlignParams {
  const __half* feat_hi[4];   // p2..p5, NHWC
  const __half* feat_lo[4];   // nullptr in fp16 precision
  int H[4], W[4];             // cropped view the ROIAlign sees (models.py:382-390)
  int pitch_H[4], pitch_W[4]; // buffer dims
  float inv_stride[4];
  int C;                      // 256
  int B, rois_per_image;      // boxes [B][rois_per_image][4], count[B]
  const float* boxes;
  const int* count;
  __half* out_hi;             // [B*rois][7][7][C]   (fc6 operand order), or
  __half* out_lo;
  float* out_nchw;            // [B*rois][C][7][7]   fp32 (fpn_box_feat contract)
  int out_res;                // 0 / 7: 7x7 bins from a 14x14 crop; 14: 14x14 bins from a 28x28 crop (mask head,
                              // models.py:936-937), planes output [B*rois][14][14][C]
};
int roialign_launch(const RoiAlignParams& p, cudaStream_t s);

// head.cu
struct HeadPostParams {
  const float* logits;        // [B*rois][ld]: cols 0..nc-1 class logits, nc + c*4 .. box logits of class c
  int ld, num_class, B, rois_per_image;
  int class_agnostic;         // box logits shared by all classes (use_frcnn_class_agnostic)
  const float* rois;          // [B][rois][4]
  const int* roi_count;       // [B]
  float reg_w[4];
  float decode_clip, img_h, img_w, score_thresh, nms_thr;
  int max_per_class, max_total;
  float* probs;               // [B][rois][num_class]
  float* dec_boxes;           // [B][rois][num_class-1][4]
  int* cls_keep;              // [B][num_class-1][max_per_class] roi indices
  int* cls_count;             // [B][num_class-1]
  float* final_boxes;         // [B][max_total][4]
  float* final_probs;         // [B][max_total]
  int* final_labels;          // [B][max_total]
  int* final_count;           // [B]
};
int head_post_launch(const HeadPostParams& p, cudaStream_t s);

// reid.cu
int dwconv3x3_launch(const __half* in_hi, const __half* in_lo, int B, int H, int W, int C, const float* w,
                     const float* bias, __half* out_hi, __half* out_lo, cudaStream_t s);
int gap_launch(const __half* in_hi, const __half* in_lo, int B, int HW, int C, float* out, int out_ld, cudaStream_t s);
int gate_mlp_launch(const float* g, int rows, int C, int Creal, int Cr, const float* w1, const float* b1,
                    const float* w2, const float* b2, float* gates, cudaStream_t s);
int gated_sum4_launch(const __half* const hi[4], const __half* const lo[4], const float* gates, int B, int HW, int C,
                      __half* out_hi, __half* out_lo, cudaStream_t s);
int avgpool2_launch(const __half* in_hi, const __half* in_lo, int B, int H, int W, int C, __half* out_hi,
                    __half* out_lo, cudaStream_t s);

// effdet.cu -- EfficientDet feature network, heads' depthwise halves, post-processing
struct BifpnInput {
  const __half* hi;
  const __half* lo;
  int H, W;               // source map size
  int mode;               // 0 same size, 1 max-pool 3x3/2 SAME from a 2x finer map, 2 nearest 2x upsampling
  int pad_t, pad_l;       // SAME padding of mode 1
  float weight;           // fast-attention edge weight relu(w_i)
};
struct BifpnCombineParams {
  BifpnInput in[3];
  int n_in;
  int B, Ho, Wo, C;
  int weighted;           // 1: node = sum_i x_i * w_i / denom (fastattn); 0: plain sum
  float denom;            // sum_i relu(w_i) + 1e-4
  int swish;              // apply x * sigmoid(x) to the combined node (op_after_combine input)
  __half* out_hi;
  __half* out_lo;
};
int bifpn_combine_launch(const BifpnCombineParams& p, cudaStream_t s);
int dw3x3_plain_launch(const __half* in_hi, const __half* in_lo, int B, int H, int W, int C, const float* w, __half* out_hi,
                       __half* out_lo, cudaStream_t s);

struct EffdetPostParams {
  const float* logits[5];   // per level [h*w][ld]: anchor-major, class-minor (A * num_classes real columns)
  const float* boxes[5];    // per level [h*w][ldb]: A * 4 (ty, tx, th, tw)
  int ld[5], ldb[5], w[5];
  unsigned long long level_off[6];   // flat logit index where each level starts
  unsigned long long total;
  int n_levels, anchors, num_classes, min_level;
  double stride_y[5], stride_x[5], half_y[5][9], half_x[5][9];   // anchor grid (anchors.py:216-257), float64 like numpy
  int k, max_out;
  float score_thresh, nms_thr;
  unsigned int* hist;       // [256]
  uint32_t* state;          // [8]: prefix, mask, need, n_selected (> k-th), n_ties (== k-th), n_candidates
  unsigned long long* cand; // [k]
  unsigned long long* ties; // [k] lowest-index ties of the k-th value
  unsigned int* tie_cnt;    // [effdet_topk_blocks()] ties per block (blocks own contiguous index ranges)
  unsigned int* tie_off;    // exclusive prefix of tie_cnt
  float4* cand_box;         // [k] (ymin, xmin, ymax, xmax), score order
  float* cand_score;
  int* cand_cls;
  int* cand_lvl;
  unsigned long long* mask; // [k][ceil(k/64)]
  float4* out_boxes;        // [max_out] x1 y1 x2 y2, scaled back to the original frame
  float* out_scores;
  int* out_classes;
  int* out_levels;
  int* out_count;
};
int effdet_topk_blocks();
int effdet_post_launch(const EffdetPostParams& p, const float* image_scale_dev, cudaStream_t s);

// effnet.cu -- EfficientNet backbone pieces + EfficientDet input pre-processing
int effnet_preprocess_launch(const uint8_t* img_bgr, int h, int w, int sh, int sw, float* out, int H, int W, cudaStream_t s);
int stem_im2col_launch(const float* img, int H, int W, int Ho, int Wo, int pad_t, int pad_l, __half* out_hi, __half* out_lo,
                       cudaStream_t s);
struct DwConvParams {
  const __half* in_hi;
  const __half* in_lo;
  int H, W, C;              // input map (batch 1), C multiple of 8
  int K, stride, pad_t, pad_l;
  int Ho, Wo;
  const float* w;           // [K*K][C], BatchNorm scale folded in
  const float* bias;        // [C], BatchNorm shift
  __half* out_hi;
  __half* out_lo;
};
int dwconv_bn_swish_launch(const DwConvParams& p, cudaStream_t s);
int dwconv_launch(const DwConvParams& p, int act, cudaStream_t s);   // act 0 none / 2 swish; p.bias may be null
int se_chunks(int HW);      // rows of the partial-sum buffer se_gate_launch needs
int se_fc1_parts();         // rows of the r_part buffer (row stride se_max_nr())
int se_max_nr();
int se_gate_launch(const __half* in_hi, const __half* in_lo, int HW, int Cpad, int C, int nr, float* partial, float* r_part,
                   const float* w1, const float* b1, const float* w2, const float* b2, float* gate, const float* w, int rows,
                   __half* w_hi, __half* w_lo, cudaStream_t s);

// roialign.cu -- ROIAlign 7x7 on the detection's own level + mean over the 49 bins (efficientdet_wrapper.py:265-301)
struct LevelRoiFeatParams {
  const __half* feat_hi[5];
  const __half* feat_lo[5];
  int H[5], W[5];
  float inv_stride[5];
  int C, Creal, min_level, max_out;
  const float4* boxes;      // x1 y1 x2 y2
  const int* levels;
  const int* count;
  float* out;               // [max_out][Creal]
};
int level_roi_feat_launch(const LevelRoiFeatParams& p, cudaStream_t s);

// cosine.cu
int cosine_normalize_rows(const float* src, int rows, int D, __half* hi, __half* lo, int ld, cudaStream_t s);
int rows_to_planes(const float* src, int rows, int D, __half* hi, __half* lo, int ld, float* sqnorm, cudaStream_t s);
int distance_finish(const float* dots, int ld, int na, int nb, int metric, const float* na2, const float* nb2, float* out,
                    cudaStream_t s);
int pair_segmin(const float* dots, int ld, const float* na2, const float* nb2, const int* seg_a, int N, const int* seg_b,
                int M, const unsigned char* gate, float fill, float* out, cudaStream_t s);
int cosine_segmin(const float* dots, int ld, const int* seg_offsets, int T, int N, float* cost, cudaStream_t s);

}  // namespace b2
