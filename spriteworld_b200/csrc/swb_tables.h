// Host-side construction of the resampling tables the render kernel consumes.
//
// Restates Pillow's Resample.c precompute_coeffs + normalize_coeffs_8bpc for the LANCZOS
// filter (the arithmetic behind `canvas.resize(image_size, Image.ANTIALIAS)`,
// renderers/pil_renderer.py:84): support 3*scale, window [int(c-s+.5), int(c+s+.5)),
// taps normalised in double and quantised to 22-bit fixed point.  On top of the raw taps
// the kernel wants: a class id per output (outputs with identical tap vectors share one
// table), prefix sums per class (horizontal pass over piecewise-constant rows), a
// paired-tap program per class (vertical pass: equal coefficients share a multiply) and
// the inverse maps "input index -> first/last output whose window contains it".
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "swb_device.cuh"

namespace swb {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Pillow PRECISION_BITS

struct AxisHost {
  int n_in = 0, n_out = 0, n_cls = 0, max_len = 0;
  std::vector<int16_t> win_min;
  std::vector<uint8_t> win_len, win_cls;
  std::vector<int32_t> prefix;   // [n_cls][33]
  std::vector<int32_t> program;  // [n_cls][PROG_STRIDE]
  std::vector<int16_t> first_out, last_out;
  std::vector<std::vector<int32_t>> taps;  // per output, for tests
};

inline double sinc_filter(double x) {
  if (x == 0.0) return 1.0;
  x = x * M_PI;
  return std::sin(x) / x;
}
inline double lanczos_filter(double x) {
  if (-3.0 <= x && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3);
  return 0.0;
}

inline bool build_axis(int n_in, int n_out, AxisHost *ax, std::string *err) {
  ax->n_in = n_in;
  ax->n_out = n_out;
  ax->win_min.assign(n_out, 0);
  ax->win_len.assign(n_out, 0);
  ax->win_cls.assign(n_out, 0);
  ax->taps.assign(n_out, {});
  if (n_in == n_out) {
    // Image.resize returns a plain copy when the size does not change (anti_aliasing=1)
    for (int xx = 0; xx < n_out; ++xx) {
      ax->win_min[xx] = (int16_t)xx;
      ax->win_len[xx] = 1;
      ax->taps[xx] = {1 << kPrecisionBits};
    }
  } else {
    const double scale = (double)n_in / n_out;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 3.0 * filterscale;
    const int ksize = (int)std::ceil(support) * 2 + 1;
    std::vector<double> k(ksize);
    for (int xx = 0; xx < n_out; ++xx) {
      const double center = (xx + 0.5) * scale;
      double ww = 0.0;
      const double ss = 1.0 / filterscale;
      int xmin = (int)(center - support + 0.5);
      if (xmin < 0) xmin = 0;
      int xmax = (int)(center + support + 0.5);
      if (xmax > n_in) xmax = n_in;
      xmax -= xmin;
      for (int x = 0; x < xmax; ++x) {
        const double w = lanczos_filter((x + xmin - center + 0.5) * ss);
        k[x] = w;
        ww += w;
      }
      for (int x = 0; x < xmax; ++x)
        if (ww != 0.0) k[x] /= ww;
      if (xmax > 32) {
        *err = "tap window of " + std::to_string(xmax) +
               " > 32 (anti_aliasing > 5 is not supported by the render kernel)";
        return false;
      }
      ax->win_min[xx] = (int16_t)xmin;
      ax->win_len[xx] = (uint8_t)xmax;
      ax->taps[xx].resize(xmax);
      for (int x = 0; x < xmax; ++x) {
        const double v = k[x] * (1 << kPrecisionBits);
        ax->taps[xx][x] = (v < 0) ? (int)(-0.5 + v) : (int)(0.5 + v);
      }
    }
  }
  // classes
  std::map<std::vector<int32_t>, int> cls;
  for (int xx = 0; xx < n_out; ++xx) {
    auto it = cls.find(ax->taps[xx]);
    int id;
    if (it == cls.end()) {
      id = (int)cls.size();
      cls.emplace(ax->taps[xx], id);
    } else {
      id = it->second;
    }
    if (id > 255) {
      *err = "more than 256 distinct tap vectors on one axis";
      return false;
    }
    ax->win_cls[xx] = (uint8_t)id;
    ax->max_len = std::max<int>(ax->max_len, ax->win_len[xx]);
  }
  ax->n_cls = (int)cls.size();
  ax->prefix.assign((size_t)ax->n_cls * 33, 0);
  ax->program.assign((size_t)ax->n_cls * PROG_STRIDE, 0);
  for (const auto &kv : cls) {
    const std::vector<int32_t> &t = kv.first;
    int32_t *P = ax->prefix.data() + (size_t)kv.second * 33;
    int32_t run = 0;
    for (int i = 0; i <= 32; ++i) {
      P[i] = run;
      if (i < (int)t.size()) run += t[i];
    }
    // paired-tap program: taps with equal non-zero coefficient are added before the multiply.
    // Pairs are emitted in ascending order of their first tap.
    int32_t *prog = ax->program.data() + (size_t)kv.second * PROG_STRIDE;
    std::map<int32_t, std::vector<int>> by_coef;
    for (int i = 0; i < (int)t.size(); ++i)
      if (t[i] != 0) by_coef[t[i]].push_back(i);
    std::vector<std::array<int32_t, 3>> pair_list;  // a, b, coefficient
    std::vector<std::array<int32_t, 2>> single_list;
    for (const auto &bc : by_coef) {
      const std::vector<int> &idx = bc.second;
      size_t i = 0;
      for (; i + 1 < idx.size() && pair_list.size() < 16; i += 2) pair_list.push_back({idx[i], idx[i + 1], bc.first});
      for (; i < idx.size(); ++i) single_list.push_back({idx[i], bc.first});
    }
    std::sort(pair_list.begin(), pair_list.end());
    std::sort(single_list.begin(), single_list.end());
    int32_t *pairs = prog + 2, *singles = prog + 2 + 2 * 16;
    for (size_t i = 0; i < pair_list.size(); ++i) {
      pairs[2 * i] = pair_list[i][0] | (pair_list[i][1] << 8);
      pairs[2 * i + 1] = pair_list[i][2];
    }
    for (size_t i = 0; i < single_list.size(); ++i) {
      singles[2 * i] = single_list[i][0];
      singles[2 * i + 1] = single_list[i][1];
    }
    // the interior vector of a 5x LANCZOS reduction: taps k and 28-k are
    // equal, taps 4, 9, 19, 24, 29 are zero, tap 14 is the centre.  The kernel has an
    // unrolled path with immediate offsets for exactly this shape (flag in bit 16 of ns).
    static const int kA5[12] = {0, 1, 2, 3, 5, 6, 7, 8, 10, 11, 12, 13};
    bool a5 = pair_list.size() == 12 && single_list.size() == 1 && single_list[0][0] == 14;
    for (int i = 0; a5 && i < 12; ++i) a5 = pair_list[i][0] == kA5[i] && pair_list[i][1] == 28 - kA5[i];
    prog[0] = (int32_t)pair_list.size();
    prog[1] = (int32_t)single_list.size() | (a5 ? (1 << 16) : 0);
  }
  // inverse maps
  ax->first_out.assign(n_in, (int16_t)(n_out - 1));
  ax->last_out.assign(n_in, 0);
  std::vector<char> seen(n_in, 0);
  for (int xx = 0; xx < n_out; ++xx) {
    for (int i = ax->win_min[xx]; i < ax->win_min[xx] + ax->win_len[xx]; ++i) {
      if (!seen[i]) {
        seen[i] = 1;
        ax->first_out[i] = (int16_t)xx;
        ax->last_out[i] = (int16_t)xx;
      } else {
        ax->first_out[i] = std::min<int16_t>(ax->first_out[i], (int16_t)xx);
        ax->last_out[i] = std::max<int16_t>(ax->last_out[i], (int16_t)xx);
      }
    }
  }
  // an input index no window touches contributes to no output: map it to an empty range
  for (int i = 0; i < n_in; ++i)
    if (!seen[i]) {
      ax->first_out[i] = 1;
      ax->last_out[i] = 0;
    }
  return true;
}

}  // namespace swb
