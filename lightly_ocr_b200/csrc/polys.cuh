// The polygon path of getDetBoxes (reference ocr/tools/det_utils.py:97-245 `poly_core`): see polys.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace locr {

// boxes fp32 [B][max_boxes][8], box_label int32 [B][max_boxes], counts int32 [B][2] (as written by launch_postproc),
// labels int32 [B][H][W] (raster-ordered component ids).  Outputs: polys fp64 [B][max_boxes][14][2] (score-map
// coordinates), valid int32 [B][max_boxes] (1 = polygon, 0 = the reference's None) for the first counts[b][0] boxes.
void launch_polys(const float* boxes, const int32_t* box_label, const int32_t* counts, const int32_t* labels, int B,
                  int H, int W, int max_boxes, double* polys, int32_t* valid, cudaStream_t s);

}  // namespace locr
