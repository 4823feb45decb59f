// Device-side data layout shared by the step and render kernels.
//
// HBM layout (all arrays owned by swb_engine, struct-of-arrays):
//   live state   pos_x, pos_y            [E*S] f64   (float32-valued where pos_f32)
//                cursor, step_count      [E]   i32
//                reset_next              [E]   u8
//   scene pool   x0,y0,m00..m11,vx,vy    [E*K*S] f64
//                member                  [E*K*S] u32
//                shape,pos_f32           [E*K*S] u8
//                rgb                     [E*K*S] u32 (r | g<<8 | b<<16)
//                factors                 [E*K*S*5] f32 (scale, angle, c0, c1, c2)
//   index of (env e, ring slot k, sprite slot s) = (e*K + k)*S + s
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/spriteworld_b200.h"

namespace swb {

struct DevState {
  int E, S, K;
  double *pos_x, *pos_y;
  int32_t *cursor, *step_count;
  uint8_t *reset_next;
  uint8_t *render_status;  // [E] SWB_ENV_SPAN_OVERFLOW from the last render
  double *p_x, *p_y, *p_m00, *p_m01, *p_m10, *p_m11, *p_vx, *p_vy;
  uint32_t *p_member;
  uint8_t *p_shape, *p_pos_f32;
  uint32_t *p_rgb;
  float *p_factors;
  const double *shape_verts;  // [13][32][2]
  const int32_t *shape_n;     // [13]
};

// action space + episode + task tree; every engine keeps its own copy in device memory and
// hands the step kernel a pointer to it.
struct StepCfg {
  int32_t action_kind;
  double action_scale;
  double motion_cost;
  int32_t keep_in_frame;
  int32_t max_episode_length;
  int32_t n_nodes;
  swb_task_node nodes[SWB_MAX_NODES];
};

// Resampling tables for one axis (Pillow precompute_coeffs / normalize_coeffs_8bpc).
struct AxisTables {
  const int16_t *win_min;    // [n_out]  first input index of the tap window
  const uint8_t *win_len;    // [n_out]  taps (<= 32)
  const uint8_t *win_cls;    // [n_out]  class = distinct tap vector
  const int32_t *prefix;     // [n_cls][33] prefix sums of the tap vector, prefix[len] = total
  const int32_t *program;    // [n_cls][PROG_STRIDE] paired-tap program (see swb_tables.h)
  const int16_t *first_out;  // [n_in] first output whose window contains input index i
  const int16_t *last_out;   // [n_in] last  output whose window contains input index i
};

struct RasterDev {
  int W, H, aa, CW, CH;
  uint32_t bg;  // r | g<<8 | b<<16
  int band_rows;  // output rows per CTA
  int n_bands;
  int max_spans;  // spans kept per (sprite, canvas row)
  int ncls_x, ncls_y;  // distinct tap vectors per axis
  int ny_cap;  // output rows per render tile: (ny_cap-1)*aa + 32 canvas rows fit the H buffer
  // vertical-pass fast path: class id and the 13 distinct coefficients of the interior tap
  // vector of a 5x reduction (taps k and 28-k equal; 4,9,19,24,29 zero; 14 the centre)
  int a5_cls;
  int32_t a5_coef[13];
  AxisTables ax, ay;
};

constexpr int PROG_STRIDE = 2 + 2 * 16 + 2 * 32;  // header, <=16 pairs, <=32 singles

}  // namespace swb
