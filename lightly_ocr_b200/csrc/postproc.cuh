// Score-map post-processing on the GPU: thresholds, 4-connected component labelling with raster-ordered label ids,
// per-component statistics, and per-component box extraction (dilation extents -> convex hull -> float32 rotating
// calipers -> boxPoints -> diamond rule -> roll -> scaled integer rect).  Replaces the cv2 / numpy host code of
// reference ocr/tools/det_utils.py:35-94, :259-265 and ocr/net.py:82-98.  Compiled with --fmad=false: the float32
// arithmetic must match OpenCV's non-contracted x86 code bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace locr {

struct PostprocParams {
    int B, H, W;            // score maps [B][H][W][2] (text, link interleaved)
    float low_text, link_threshold, text_threshold;
    double scale_x, scale_y;  // ratio_w * 2, ratio_h * 2 (python floats)
    int max_boxes;          // per-image capacity of the outputs
};

// Workspace sizes (bytes) for a batch of B maps of H x W.
size_t postproc_workspace_bytes(int B, int H, int W);

// Outputs (device): boxes fp32 [B][max_boxes][8], rects int32 [B][max_boxes][4] (min_y, min_x, max_y, max_x),
// box_label int32 [B][max_boxes] (cv2 label id of each box), counts int32 [B][2] (boxes kept, components found).
// labels_out (optional): int32 [B][H][W] raster-ordered component ids like cv2.connectedComponents.
// Returns the number of kernel launches issued.
int launch_postproc(const float* score, const PostprocParams& p, void* workspace, float* boxes, int32_t* rects,
                    int32_t* box_label, int32_t* counts, int32_t* labels_out, cudaStream_t s);

}  // namespace locr
