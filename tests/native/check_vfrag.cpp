// Host-only check of the vertical-pass tensor-core tables (swb_tables.h build_vfrag).
//
// For every block of eight output rows the fragments are expanded back through the register
// layout of mma.sync.m16n8k32 (lane g = lane/4 holds output row g, k = 32 ks + 16 half + 4 (lane%4)
// + byte), multiplied with a random uint8 column of H values, the three int8 limbs recombined as
// d0 + 2^8 d1 + 2^16 d2 in 32-bit arithmetic -- exactly what the render kernel does -- and the
// result compared with the direct convolution with Pillow's 22-bit taps.  Prints one line per
// (n_in, n_out) pair; exit code 1 on any mismatch.
#include <cstdio>
#include <cstdlib>

#include "../../spriteworld_b200/csrc/swb_tables.h"

using namespace swb;

int main() {
  const int cfgs[][2] = {{320, 64}, {640, 128}, {64, 64}, {128, 64}, {192, 64}, {256, 64},
                         {300, 60}, {250, 50}, {80, 16}, {480, 96}, {35, 7}};
  long total_bad = 0;
  srand(12345);
  for (const auto &c : cfgs) {
    AxisHost ay;
    std::string err;
    if (!build_axis(c[0], c[1], &ay, &err)) { printf("axis %d->%d: %s\n", c[0], c[1], err.c_str()); return 1; }
    VFragHost vf;
    if (!build_vfrag(ay, &vf, &err)) { printf("vfrag %d->%d: %s\n", c[0], c[1], err.c_str()); return 1; }
    const int n_blk = (c[1] + 7) / 8;
    long bad = 0;
    for (int b = 0; b < n_blk; ++b) {
      const int o = ay.win_min[8 * b] & ~3;
      std::vector<int> h(c[0] + 32 * vf.nks + 8);
      for (auto &v : h) v = rand() & 255;
      const uint32_t *f = vf.frag.data() + (size_t)vf.blk_cls[b] * vf.nks * 3 * 32 * 2;
      for (int g = 0; g < 8; ++g) {
        const int yo = 8 * b + g;
        if (yo >= c[1]) continue;
        int32_t acc[3] = {0, 0, 0};
        for (int ks = 0; ks < vf.nks; ++ks)
          for (int t = 0; t < 4; ++t)
            for (int half = 0; half < 2; ++half)
              for (int i = 0; i < 4; ++i) {
                const int row = o + 32 * ks + 16 * half + 4 * t + i;
                const int hv = row < (int)h.size() ? h[row] : 0;
                for (int limb = 0; limb < 3; ++limb) {
                  const uint32_t w = f[(((size_t)ks * 3 + limb) * 32 + (g * 4 + t)) * 2 + half];
                  acc[limb] += hv * (int)(int8_t)((w >> (8 * i)) & 255u);
                }
              }
        const int32_t got = (int32_t)((uint32_t)acc[0] + ((uint32_t)acc[1] << 8) + ((uint32_t)acc[2] << 16));
        int64_t want = 0;
        for (int i = 0; i < ay.win_len[yo]; ++i) want += (int64_t)ay.taps[yo][i] * h[ay.win_min[yo] + i];
        if (want != got) ++bad;
      }
    }
    printf("%d -> %d: k-steps %d, classes %d, blocks %d, mismatches %ld\n", c[0], c[1], vf.nks, vf.n_cls,
           n_blk, bad);
    total_bad += bad;
  }
  return total_bad ? 1 : 0;
}
