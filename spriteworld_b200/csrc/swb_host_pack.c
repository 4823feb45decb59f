/* Host-side packing of sampled sprite tables into the scene arrays swb_upload_scenes takes.
 *
 * spriteworld_b200/scene.py arrays_from_layout does this with NumPy (a dozen passes per table:
 * colour map, shape ids, type conversions, one gather per field).  At C2's reset rate the
 * batched environment wants about a million scenes per second from the host, so the common
 * case -- one table of plain numeric factor columns, shape names from a Discrete, the HSV
 * colour map of the reference's renderers/color_maps.py:26-28 -- is done here in one pass per
 * sprite.  Everything else stays on the NumPy path, which is also the specification: the two are
 * compared value by value in tests/test_host_api.py.
 *
 * Arithmetic follows NumPy's elementwise evaluation of the same expressions (every operation
 * rounded to the column's float type, no contraction: build with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

enum { SWB_COL_F32 = 0, SWB_COL_F64 = 1, SWB_COL_I64 = 2 };

static inline double load_f64(const void *p, int code, int64_t r) {
  switch (code) {
    case SWB_COL_F32: return (double)((const float *)p)[r];
    case SWB_COL_F64: return ((const double *)p)[r];
    default: return (double)((const int64_t *)p)[r];
  }
}

/* floor modulus of NumPy's int64 % */
static inline int64_t mod6(int64_t i) {
  if ((uint64_t)i < 6) return i; /* hue in [0, 1): no division */
  int64_t m = i % 6;
  return m < 0 ? m + 6 : m;
}

#define HSV_TO_RGB(T, NAME)                                                              \
  static inline void NAME(T h, T s, T v, uint8_t *out) {                                 \
    const T one = (T)1.0, six = (T)6.0, scale = (T)255;                                  \
    const T h6 = h * six;                                                                \
    const int64_t it = (int64_t)h6; /* int() truncates */                                \
    const T f = h6 - (T)it;                                                              \
    const T sf = s * f;                                                                  \
    const T p = v * (one - s);                                                           \
    const T q = v * (one - sf);                                                          \
    const T omf = one - f;                                                               \
    const T somf = s * omf;                                                              \
    const T t = v * (one - somf);                                                        \
    T r, g, b;                                                                           \
    switch (mod6(it)) {                                                                  \
      case 0: r = v; g = t; b = p; break;                                                \
      case 1: r = q; g = v; b = p; break;                                                \
      case 2: r = p; g = v; b = t; break;                                                \
      case 3: r = p; g = q; b = v; break;                                                \
      case 4: r = t; g = p; b = v; break;                                                \
      default: r = v; g = p; b = q; break;                                               \
    }                                                                                    \
    if (s == (T)0) { r = v; g = v; b = v; }                                              \
    out[0] = (uint8_t)(scale * r);                                                       \
    out[1] = (uint8_t)(scale * g);                                                       \
    out[2] = (uint8_t)(scale * b);                                                       \
  }

HSV_TO_RGB(float, hsv_f32)
HSV_TO_RGB(double, hsv_f64)

/* Distinct values of an array of pointers (an object column's PyObject*), at most `cap`:
 * writes them to `distinct` and the index of a first occurrence of each to `first`.
 * Returns their number, or -1 if there are more than cap. */
int64_t swb_distinct_pointers(const uintptr_t *ptrs, int64_t n, int64_t cap, uintptr_t *distinct,
                              int64_t *first) {
  int64_t k = 0;
  uintptr_t last = 0;
  int have_last = 0;
  for (int64_t i = 0; i < n; ++i) {
    const uintptr_t p = ptrs[i];
    if (have_last && p == last) continue;
    int64_t j = 0;
    while (j < k && distinct[j] != p) ++j;
    if (j == k) {
      if (k == cap) return -1;
      distinct[k] = p;
      first[k] = i;
      ++k;
    }
    last = p;
    have_last = 1;
  }
  return k;
}

typedef struct {
  /* table columns, `code` per column (SWB_COL_*) */
  const void *x, *y, *scale, *angle, *c0, *c1, *c2, *vx, *vy;
  int32_t x_code, y_code, scale_code, angle_code, c_code /* c0, c1, c2 share it */, vx_code, vy_code;
  int32_t pos_f32;            /* both x and y are float32 columns */
  int32_t rgb_mode;           /* 0: colours are RGB already (cast), 1: HSV -> RGB */
  const double *transform;    /* [rows][4] m00, m01, m10, m11; or one row if transform_stride == 0 */
  int64_t transform_stride;   /* 4, or 0: every sprite has the same (scale, angle) */
  const uint32_t *member;     /* [rows] */
  /* shape: either an int64 column of ids, or PyObject* pointers + a table */
  const int64_t *shape_ids;   /* NULL if by pointer */
  const uintptr_t *shape_ptrs;
  const uintptr_t *shape_key; /* [n_shape_keys] */
  const uint8_t *shape_val;   /* [n_shape_keys] */
  int64_t n_shape_keys;
  /* output arrays, flat over (scene, slot) */
  double *o_x, *o_y, *o_m00, *o_m01, *o_m10, *o_m11, *o_vx, *o_vy;
  uint32_t *o_member;
  uint8_t *o_shape, *o_pos_f32, *o_rgb;
  float *o_factors;
} swb_pack_args;

/* out[d(t)] = (double)col[rows[t]], one tight loop per column type */
static void gather_f64(double *out, const void *col, int code, int64_t n, const int64_t *rows,
                       const int64_t *dst) {
#define GATHER(T)                                                            \
  do {                                                                       \
    const T *c = (const T *)col;                                             \
    if (dst) for (int64_t t = 0; t < n; ++t) out[dst[t]] = (double)c[rows[t]]; \
    else for (int64_t t = 0; t < n; ++t) out[t] = (double)c[rows[t]];        \
  } while (0)
  if (code == SWB_COL_F32) GATHER(float);
  else if (code == SWB_COL_F64) GATHER(double);
  else GATHER(int64_t);
#undef GATHER
}

/* items: for t in [0, n): table row rows[t] goes to flat slot dst[t] (dst == NULL: slot t).
 * Returns 0, or 1 if a shape pointer is not in the table. */
int swb_pack_scenes(const swb_pack_args *a, int64_t n, const int64_t *rows, const int64_t *dst) {
  gather_f64(a->o_x, a->x, a->x_code, n, rows, dst);
  gather_f64(a->o_y, a->y, a->y_code, n, rows, dst);
  gather_f64(a->o_vx, a->vx, a->vx_code, n, rows, dst);
  gather_f64(a->o_vy, a->vy, a->vy_code, n, rows, dst);
  for (int64_t t = 0; t < n; ++t) {
    const int64_t d = dst ? dst[t] : t;
    const double *m = a->transform + a->transform_stride * rows[t];
    a->o_m00[d] = m[0]; a->o_m01[d] = m[1]; a->o_m10[d] = m[2]; a->o_m11[d] = m[3];
    a->o_member[d] = a->member[rows[t]];
    a->o_pos_f32[d] = (uint8_t)a->pos_f32;
  }
  if (a->shape_ids) {
    for (int64_t t = 0; t < n; ++t) a->o_shape[dst ? dst[t] : t] = (uint8_t)a->shape_ids[rows[t]];
  } else {
    uintptr_t last = 0;
    uint8_t last_val = 0;
    int have_last = 0;
    for (int64_t t = 0; t < n; ++t) {
      const uintptr_t p = a->shape_ptrs[rows[t]];
      if (!have_last || p != last) {
        int64_t j = 0;
        while (j < a->n_shape_keys && a->shape_key[j] != p) ++j;
        if (j == a->n_shape_keys) return 1;
        last = p;
        last_val = a->shape_val[j];
        have_last = 1;
      }
      a->o_shape[dst ? dst[t] : t] = last_val;
    }
  }
  /* colours and the factor observations (scale, angle, c0, c1, c2 as float32) */
  for (int64_t t = 0; t < n; ++t) {
    const int64_t r = rows[t];
    const int64_t d = dst ? dst[t] : t;
    uint8_t *rgb = a->o_rgb + 3 * d;
    float *f = a->o_factors + 5 * d;
    f[0] = (float)load_f64(a->scale, a->scale_code, r);
    f[1] = (float)load_f64(a->angle, a->angle_code, r);
    if (a->c_code == SWB_COL_F32) {
      const float c0 = ((const float *)a->c0)[r], c1 = ((const float *)a->c1)[r], c2 = ((const float *)a->c2)[r];
      if (a->rgb_mode == 1) hsv_f32(c0, c1, c2, rgb);
      else { rgb[0] = (uint8_t)(int64_t)c0; rgb[1] = (uint8_t)(int64_t)c1; rgb[2] = (uint8_t)(int64_t)c2; }
      f[2] = c0; f[3] = c1; f[4] = c2;
    } else {
      const double c0 = load_f64(a->c0, a->c_code, r), c1 = load_f64(a->c1, a->c_code, r),
                   c2 = load_f64(a->c2, a->c_code, r);
      if (a->rgb_mode == 1) hsv_f64(c0, c1, c2, rgb);
      else { rgb[0] = (uint8_t)(int64_t)c0; rgb[1] = (uint8_t)(int64_t)c1; rgb[2] = (uint8_t)(int64_t)c2; }
      f[2] = (float)c0; f[3] = (float)c1; f[4] = (float)c2;
    }
  }
  return 0;
}

int swb_host_pack_version(void) { return 1; }
