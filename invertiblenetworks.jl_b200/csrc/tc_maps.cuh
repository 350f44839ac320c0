// tc_maps.cuh - host-side TMA tensor-map builders shared by the tensor-core kernels (conv_tc.cu).
#pragma once
#include <cuda.h>
#include "conv_tc.cuh"

namespace inb {
// pixel-major activation planes [B][D][H][W][pitch] bf16, box (ck channels, tile box), swizzle = ck*2 bytes
CUtensorMap make_act_map(const __nv_bfloat16* base, int pitch, const Geo& g, int B, int ck, const TileBox& tb);
// weight planes [rows][ktot] bf16 (K-major rows), box (ck, rows)
CUtensorMap make_w_map(const __nv_bfloat16* base, int ktot, int rows, int ck);
// generic [rows][pitch] bf16 planes, box (box_c channels, box_rows rows), swizzle = box_c*2 bytes
CUtensorMap make_rows_map(const __nv_bfloat16* base, int pitch, long long rows, int box_c, int box_rows);
CUtensorMap make_rows_map_f32(const float* base, int pitch, long long rows, int box_c, int box_rows);
CUtensorMap make_rows_map_f32_dense(const float* base, int pitch, long long rows, int box_c, int box_rows);
// byte planes [rows][pitch] (the 1-byte lo planes), box (box_c bytes, box_rows rows); swizzle64: SWIZZLE_64B (box_c == 64)
CUtensorMap make_rows_map_u8(const void* base, int pitch, long long rows, int box_c, int box_rows, bool swizzle64);
// fp32 planes [nplanes][rows][cols] (dense, cols * 4 a multiple of 16), box (cols, box_rows, 1), no swizzle
CUtensorMap make_planes_map_f32_dense(const float* base, int cols, long long rows, int nplanes, int box_rows);
}  // namespace inb
