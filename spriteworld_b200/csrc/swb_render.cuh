// Render kernel: PILRenderer.render for E envs (renderers/pil_renderer.py:67-91).
//
// One CTA renders one band of one frame.  Nothing of the aa-times supersampled canvas is
// ever materialised; the CTA works on what Pillow's two stages are functions of:
//
//   A  vertices   int(canvas * (R.S.v + pos)) in fp64 (sprite.py:128-133, pil_renderer.py:81,
//                 Pillow's (int) truncation), then Pillow's edge records (float32 dx) and,
//                 per edge, the value Pillow's corner-joining refinement would overwrite
//                 its crossing with on its first row / on the polygon's last row.
//   B  spans      Pillow's scan conversion.  B1 is edge-parallel: every edge drops its
//                 float32 crossing (twice where it ends on an interior row) into the
//                 crossing list of each canvas row it spans.  B2 is row-parallel: per
//                 sprite, front to back, sort the list, pair it up with Pillow's
//                 ROUND_UP/ROUND_DOWN rule, add horizontal edges, and fold the spans into
//                 the row's list of VISIBLE segments (x range + sprite), i.e. the painter's
//                 algorithm resolved once per canvas row.
//   C  per sprite region (the outputs whose 2-D tap window can see the sprite), in tiles
//      sized to a 20 KB buffer of H values:
//        H  horizontal LANCZOS pass of the canvas rows the tile needs.  A canvas row is piecewise constant, so an output is
//           bg*K + sum_segments (colour-bg) * (P[b]-P[a]) with P the prefix sums of the
//           22-bit tap vector and [a, b) the segment clamped to the tap window
//           (branch-free).  clip8'ed like Pillow's uint8 intermediate.
//        V  vertical pass over the tile's H values (paired taps: equal coefficients share
//           one multiply), clip8, written into the frame staged in shared memory.
//   D  the staged frame (background + tiles) goes to HBM as 128-bit stores, rows flipped
//      (np.flipud, pil_renderer.py:90).
//
// All pixel arithmetic is integer and associative, so the result is bit-identical to
// Pillow's; the float32 edge arithmetic uses explicit _rn intrinsics (no FMA).
#pragma once
#include <climits>

#include "swb_device.cuh"

namespace swb {

constexpr int R_THREADS = 256;
constexpr int H_NC = 2;           // output columns one H-pass thread owns
constexpr int TILE_X_MAX = 20;    // output columns per tile = row stride of the H buffer
constexpr int HT_ROWS = 112;      // canvas rows a tile's H buffer holds
constexpr int HT_ITEMS = HT_ROWS * TILE_X_MAX;
constexpr int MAX_ROW_SPANS = 12;
constexpr int EV = SWB_MAX_VERTS;  // edge slots per sprite
static_assert(EV == 32, "phase A maps one lane to one vertex / edge");

// Where a frame goes.  Normally one buffer; with the frame gather fused into the kernel
// (swb_step_render_gather) one buffer per rank, each an [n_ranks * E] frame array reached
// over NVLink peer memory, written at env index env_offset + e.
struct RenderTargets {
  uint8_t *dst[SWB_MAX_PEERS];
  int n;
  int env_offset;
  int self;  // index of this rank's own buffer in dst
};

struct RenderLayout {
  int S, rows, M, band_rows, W, aa, ncx, ncy, cap;
  int off_meta, off_edge_i, off_edge_f, off_edge_yr, off_hl, off_region;
  int off_nseg, off_segs, off_prefix, off_xwin, off_ywin, off_scratch, off_frame, total;
  int scratch_bytes, segcap;
  __host__ __device__ RenderLayout(int S_, int rows_, int M_, int band_rows_, int W_, int aa_,
                                   int ncx_, int ncy_)
      : S(S_), rows(rows_), M(M_), band_rows(band_rows_), W(W_), aa(aa_), ncx(ncx_), ncy(ncy_) {
    int o = 0;
    auto take = [&](int bytes) { int r = o; o += (bytes + 15) & ~15; return r; };
    off_meta = take(S * (12 + 8) * 4);     // 12 plan ints + 8 active-edge masks per sprite
    off_edge_i = take(S * EV * 2 * 4);     // x0, y0
    off_edge_f = take(S * EV * 3 * 4);     // dx, ovs (override on the first row), ove (on the last row)
    off_edge_yr = take(S * EV * 4);        // ymin | ymax<<16 of non-horizontal edges, empty otherwise
    off_hl = take(S * EV * 3 * 2);         // horizontal edges: y, xmin, xmax (int16)
    off_region = take(S * 4 * 2);
    // visible segments kept per canvas row: n one-span sprites leave at most 2n-1 pieces
    segcap = M > 1 ? 16 : (2 * S < 4 ? 4 : (2 * S > 16 ? 16 : 2 * S));
    off_nseg = take(rows);
    off_segs = take(rows * segcap * 4);
    off_prefix = take(ncx * 33 * 4);
    off_xwin = take(W * 4);                // per output column: win_min | len<<16 | cls<<24
    off_ywin = take(band_rows * 4);
    cap = (M > 1) ? 16 : 8;                // crossings kept per (sprite, row)
    // scratch = H tile + staged frame; phase B aliases it with the per-row crossing lists and
    // spans of a chunk of sprites, so it must hold at least one sprite spanning every row
    const int frame_bytes = band_rows * W * 3;
    const int need_b = (cap + M) * 4 * rows;
    int ht_bytes = HT_ITEMS * 4;           // one H value = r | g<<10 | b<<20
    if (ht_bytes + ((frame_bytes + 15) & ~15) < need_b) ht_bytes = need_b - frame_bytes;
    off_scratch = take(ht_bytes);
    off_frame = take(frame_bytes);
    scratch_bytes = o - off_scratch;
    total = o;
  }
};

__device__ __forceinline__ int round_up_f(float f) {
  return f >= 0.0f ? (int)floorf(__fadd_rn(f, 0.5f)) : -(int)floorf(__fadd_rn(fabsf(f), 0.5f));
}
__device__ __forceinline__ int round_down_f(float f) {
  return f >= 0.0f ? (int)ceilf(__fsub_rn(f, 0.5f)) : -(int)ceilf(__fsub_rn(fabsf(f), 0.5f));
}
__device__ __forceinline__ float edge_x_at(int y, int y0, float dx, int x0) {
  return __fadd_rn(__fmul_rn((float)(y - y0), dx), (float)x0);
}
__device__ __forceinline__ uint32_t clip8_q22(int v) {  // Pillow clip8: (v >> 22) clamped to 0..255
  uint32_t r;
  asm("cvt.sat.u8.s32 %0, %1;" : "=r"(r) : "r"(v >> 22));
  return r;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" : : "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int lds_s32(uint32_t addr) {
  int v;
  // volatile: must not be scheduled across the barriers that publish the tables it reads
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// merges [xs, xe] into a small unsorted list of disjoint, non-adjacent spans
__device__ __forceinline__ void add_span(int *lxs, int *lxe, int &n, int xs, int xe, bool &ovf) {
  for (int i = 0; i < n;) {
    if (xs <= lxe[i] + 1 && lxs[i] <= xe + 1) {
      xs = min(xs, lxs[i]);
      xe = max(xe, lxe[i]);
      lxs[i] = lxs[n - 1];
      lxe[i] = lxe[n - 1];
      --n;
      i = 0;
    } else {
      ++i;
    }
  }
  if (n < MAX_ROW_SPANS) {
    lxs[n] = xs;
    lxe[n] = xe;
    ++n;
  } else {
    ovf = true;
  }
}

// c_inv20[d] = 2^20 / d + 1: x / d == (x * c_inv20[d]) >> 20 for 0 <= x < 2^20 / d (d <= 64)
__constant__ uint32_t c_inv20[65] = {
    0u, 1048577u, 524289u, 349526u, 262145u, 209716u, 174763u, 149797u,
    131073u, 116509u, 104858u, 95326u, 87382u, 80660u, 74899u, 69906u,
    65537u, 61681u, 58255u, 55189u, 52429u, 49933u, 47663u, 45591u,
    43691u, 41944u, 40330u, 38837u, 37450u, 36158u, 34953u, 33826u,
    32769u, 31776u, 30841u, 29960u, 29128u, 28340u, 27595u, 26887u,
    26215u, 25576u, 24967u, 24386u, 23832u, 23302u, 22796u, 22311u,
    21846u, 21400u, 20972u, 20561u, 20165u, 19785u, 19419u, 19066u,
    18725u, 18397u, 18079u, 17773u, 17477u, 17190u, 16913u, 16645u,
    16385u};
__device__ __forceinline__ int div20(int x, int d) { return (int)(((uint32_t)x * c_inv20[d]) >> 20); }

// Debug-only phase timers (nvcc -DSWB_PHASE_CLOCKS): thread 0 of every CTA adds the cycles
// between consecutive marks to g_phase_clk[id].  Not part of the shipped library.
#ifdef SWB_PHASE_CLOCKS
__device__ unsigned long long g_phase_clk[16];
#define SWB_MARK(id)                                                              \
  do {                                                                            \
    if (tid == 0) {                                                               \
      const long long now_ = clock64();                                           \
      atomicAdd(&g_phase_clk[id], (unsigned long long)(now_ - mark_));            \
      mark_ = now_;                                                               \
    }                                                                             \
  } while (0)
#else
#define SWB_MARK(id) do { } while (0)
#endif

template <bool kPeers>  // kPeers: also store the frame into the other ranks' buffers
__global__ void __launch_bounds__(R_THREADS, 5)
render_kernel(DevState st, RasterDev rd, const RenderLayout L, const RenderTargets targets,
              int env_base) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int e = env_base + blockIdx.x;  // frames is indexed by the absolute env id
  const int band = blockIdx.y;
  const int tid = threadIdx.x;
  const int S = st.S;

  // per-sprite plan (ints): vertex count, last polygon row, first row / row count in this
  // band, horizontal-edge count, bucket shift of the active-edge masks, colour - background,
  // tile plan of the region, "has a corner join"
  int *s_nv = reinterpret_cast<int *>(smem + L.off_meta);
  int *s_pymax = s_nv + S, *s_r0 = s_pymax + S, *s_rcnt = s_r0 + S, *s_nh = s_rcnt + S;
  int *s_bsh = s_nh + S, *s_dr = s_bsh + S;
  int *s_pny = s_dr + S, *s_pnx = s_pny + S;
  int *s_pinvh = s_pnx + S, *s_hasov = s_pinvh + S;
  unsigned *s_emask = reinterpret_cast<unsigned *>(s_hasov + S);  // [S][8] active edges per row bucket
  // edge table: start vertex, slope, corner-join overrides on the first / last row
  int *e_x0 = reinterpret_cast<int *>(smem + L.off_edge_i);
  int *e_y0 = e_x0 + S * EV;
  float *e_dx = reinterpret_cast<float *>(smem + L.off_edge_f);
  float *e_ovs = e_dx + S * EV, *e_ove = e_ovs + S * EV;
  uint32_t *e_yr = reinterpret_cast<uint32_t *>(smem + L.off_edge_yr);
  short *s_hl = reinterpret_cast<short *>(smem + L.off_hl);  // [S][EV][3]
  short *s_region = reinterpret_cast<short *>(smem + L.off_region);  // [S][4] yo0,yo1,xo0,xo1
  uint8_t *s_nseg = smem + L.off_nseg;
  uint32_t *s_segs = reinterpret_cast<uint32_t *>(smem + L.off_segs);  // xs | xe<<12 | sprite<<24
  const int SEGCAP = L.segcap;
  const int M = rd.max_spans;
  int32_t *s_prefix = reinterpret_cast<int32_t *>(smem + L.off_prefix);
  uint32_t *s_xwin = reinterpret_cast<uint32_t *>(smem + L.off_xwin);
  uint32_t *s_ywin = reinterpret_cast<uint32_t *>(smem + L.off_ywin);
  uint32_t *s_ht = reinterpret_cast<uint32_t *>(smem + L.off_scratch);
  uint8_t *s_frame = smem + L.off_frame;
  // phase-B view of the scratch area: per-row crossing lists
  const int CAP = L.cap;
  __shared__ int s_overflow;
  __shared__ uint8_t *s_dst[SWB_MAX_PEERS];  // kPeers: the targets, indexable at run time

  const int yo_b0 = band * rd.band_rows;
  const int yo_b1 = min(yo_b0 + rd.band_rows, rd.H);
  const int n_yo = yo_b1 - yo_b0;
  // canvas rows this band's vertical windows can touch
  const int row_b0 = rd.ay.win_min[yo_b0];
  const int row_b1 = rd.ay.win_min[yo_b1 - 1] + rd.ay.win_len[yo_b1 - 1];  // exclusive
  const int n_rows = row_b1 - row_b0;

#ifdef SWB_PHASE_CLOCKS
  long long mark_ = clock64();
#endif
  // ---- phase A: a warp per sprite, a lane per vertex / edge ------------------------------
  // Everything Pillow derives from the vertex list before it scans rows: integer vertices,
  // the edge table (add_edge), horizontal edges, extents, the corner joins -- plus this
  // kernel's own per-sprite plan (rows, output region, tiles, active-edge masks).  Neighbour
  // vertices come by shuffle, extents by warp reductions, "edges that share a start row" by
  // match_any; nothing leaves the warp until the barrier that ends the phase.
  constexpr unsigned FULL = 0xFFFFFFFFu;
  constexpr int NWARP = R_THREADS / 32;
  const int lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_overflow = 0;
  if (kPeers && tid < SWB_MAX_PEERS) s_dst[tid] = targets.dst[tid];
  const int cur = st.cursor[e];
  // tables first: their loads overlap the sprite records' dependent loads below
#pragma unroll 1
  for (int i = tid; i < rd.ncls_x * 33; i += R_THREADS) s_prefix[i] = rd.ax.prefix[i];
#pragma unroll 1
  for (int i = tid; i < (n_rows + 3) / 4; i += R_THREADS) reinterpret_cast<uint32_t *>(s_nseg)[i] = 0u;
#pragma unroll 1
  for (int i = tid; i < rd.W; i += R_THREADS)
    s_xwin[i] = (uint32_t)(uint16_t)rd.ax.win_min[i] | ((uint32_t)rd.ax.win_len[i] << 16) |
                ((uint32_t)rd.ax.win_cls[i] << 24);
#pragma unroll 1
  for (int i = tid; i < n_yo; i += R_THREADS)
    s_ywin[i] = (uint32_t)(uint16_t)rd.ay.win_min[yo_b0 + i] | ((uint32_t)rd.ay.win_len[yo_b0 + i] << 16) |
                ((uint32_t)rd.ay.win_cls[yo_b0 + i] << 24);
  for (int s = warp; s < S; s += NWARP) {
    // sprite record of this env's current scene (warp-uniform loads)
    const int scene = (e * st.K + cur) * S + s;
    const int shape = st.p_shape[scene];
    const uint32_t col = st.p_rgb[scene];
    const double px = st.pos_x[e * S + s], py = st.pos_y[e * S + s];
    const double m00 = st.p_m00[scene], m01 = st.p_m01[scene];
    const double m10 = st.p_m10[scene], m11 = st.p_m11[scene];
    const int nv = shape ? st.shape_n[shape] : 0;
    const int t = s * EV + lane;
    // integer canvas vertex of this lane
    int ivx = 0, ivy = 0;
    if (lane < nv) {
      const double2 v = reinterpret_cast<const double2 *>(st.shape_verts)[(size_t)shape * EV + lane];
      // centred path (sprite.py:96-101): (a*x + c*y) + 0 ; world (sprite.py:128-133): + pos
      const double cx = __dadd_rn(__dadd_rn(__dmul_rn(m00, v.x), __dmul_rn(m01, v.y)), 0.0);
      const double cy = __dadd_rn(__dadd_rn(__dmul_rn(m10, v.x), __dmul_rn(m11, v.y)), 0.0);
      const double wx = __dadd_rn(cx, px), wy = __dadd_rn(cy, py);
      // canvas_size * vertices (pil_renderer.py:81), then Pillow's (int) cast
      ivx = __double2int_rz(__dmul_rn((double)rd.CW, wx));
      ivy = __double2int_rz(__dmul_rn((double)rd.CH, wy));
    }
    // edge lane -> lane+1 (Pillow add_edge); the closing edge exists only if the last vertex
    // differs from the first
    const int jn = (lane + 1 >= nv) ? 0 : lane + 1;
    const int x0 = ivx, y0 = ivy;
    const int x1 = __shfl_sync(FULL, ivx, jn), y1 = __shfl_sync(FULL, ivy, jn);
    const bool exists = lane < nv && ((lane + 1 < nv) || (x0 != x1 || y0 != y1));
    const bool scanned = exists && y0 != y1;
    const int ymin = min(y0, y1), ymax = max(y0, y1);
    const float dx = scanned ? __fdiv_rn((float)(x1 - x0), (float)(y1 - y0)) : 0.0f;
    // horizontal edges are drawn directly as hline(xmin, y, xmax)
    const int hxs = max(min(x0, x1), 0), hxe = min(max(x0, x1), rd.CW - 1);
    const bool hline = exists && y0 == y1 && y0 >= 0 && y0 < rd.CH && hxs <= hxe;
    const unsigned hmask = __ballot_sync(FULL, hline);
    if (hline) {
      short *h = s_hl + ((size_t)s * EV + __popc(hmask & ((1u << lane) - 1u))) * 3;
      h[0] = (short)y0; h[1] = (short)hxs; h[2] = (short)hxe;
    }
    // rows the edge crosses, clamped to int16 (canvas rows are < 4096); empty if not scanned
    const uint32_t yr = scanned ? ((uint32_t)(uint16_t)(short)max(ymin, -32768) |
                                   ((uint32_t)(uint16_t)(short)min(ymax, 32767) << 16))
                                : 0x80007FFFu;
    e_x0[t] = x0;
    e_y0[t] = y0;
    e_dx[t] = dx;
    e_yr[t] = yr;
    e_ovs[t] = nanf("");
    e_ove[t] = nanf("");
    // extents over the vertices
    const bool isv = lane < nv;
    const int xmn = __reduce_min_sync(FULL, isv ? ivx : INT_MAX), xmx = __reduce_max_sync(FULL, isv ? ivx : INT_MIN);
    const int ymn = __reduce_min_sync(FULL, isv ? ivy : INT_MAX), ymx = __reduce_max_sync(FULL, isv ? ivy : INT_MIN);
    // Pillow: ymin = min(ysize-1, edges), ymax = max(0, edges); then clip to [0, ysize]
    const int pymin = max(min(rd.CH - 1, ymn), 0), p_ymax = min(max(0, ymx), rd.CH);
    const int r0 = max(pymin, row_b0), r1 = min(min(p_ymax, rd.CH - 1), row_b1 - 1);
    const int rcnt = (nv > 0 && r1 >= r0) ? (r1 - r0 + 1) : 0;
    // active-edge masks of eight equal row buckets: B1 only looks at the edges of its bucket
    const int bsh = rcnt > 8 ? (32 - __clz(rcnt - 1) - 3) : 0;
    unsigned my_bucket = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int lo = r0 + (b << bsh), hi = lo + (1 << bsh) - 1;
      const unsigned m = __ballot_sync(FULL, scanned && ymin <= hi && ymax >= lo);
      if (lane == b) my_bucket = m;
    }
    if (lane < 8) s_emask[s * 8 + lane] = my_bucket;
    __syncwarp();  // the edge table of this sprite is visible to the join search below

    // corner joins ("connect discontiguous corners"): edge i and the FIRST earlier edge k
    // that leaves the same corner in the same x direction -- both starting on row y, or both
    // ending on the polygon's last row -- move k's crossing on that row towards the adjacent
    // row's span (oracle/sw_raster_oracle.c).  Applied in B1 through e_ovs / e_ove.
    const unsigned lt = (1u << lane) - 1u;
    const unsigned pos = __ballot_sync(FULL, scanned && dx > 0.0f);
    const unsigned neg = __ballot_sync(FULL, scanned && dx < 0.0f);
    const unsigned same_dir = dx > 0.0f ? pos : (dx < 0.0f ? neg : 0u);
    const unsigned start_grp = __match_any_sync(FULL, scanned ? (int)(short)(yr & 0xFFFFu) : (0x40000000 | lane));
    const unsigned end_grp = __ballot_sync(FULL, scanned && (int)(short)(yr >> 16) == p_ymax);
    bool any_join = false;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int y = pass == 0 ? ymin : p_ymax;
      unsigned cand = (pass == 0 ? start_grp : end_grp) & lt & same_dir;
      if (!scanned || (pass == 1 && ymax != p_ymax) || y < 0 || y >= rd.CH) cand = 0;
      int jk = -1;
      float jv = 0.0f;
      if (cand) {
        const float x = edge_x_at(y, y0, dx, x0);
        while (cand) {
          const int k = __ffs(cand) - 1;
          cand &= cand - 1;
          const int u = s * EV + k;
          const float odx = e_dx[u];
          const int oy0 = e_y0[u], ox0 = e_x0[u];
          const float ox = edge_x_at(y, oy0, odx, ox0);
          if (roundf(x) != roundf(ox)) continue;
          const int off = (y == p_ymax) ? -1 : 1;
          const float adj = edge_x_at(y + off, y0, dx, x0);
          const float adjo = edge_x_at(y + off, oy0, odx, ox0);
          const bool right = (y == ymax) ? (dx < 0.0f) : (dx > 0.0f);
          float nvx = right ? __fsub_rn(fminf(adj, adjo), 1.0f) : __fadd_rn(fmaxf(adj, adjo), 1.0f);
          nvx = floorf(__fadd_rn(nvx, 0.5f));
          jv = right ? fmaxf(nvx, x) : fminf(nvx, x);
          jk = k;
          break;
        }
      }
      // several later edges may pick the same k: the last one in table order wins
      const unsigned grp = __match_any_sync(FULL, jk >= 0 ? jk : (64 + lane));
      if (jk >= 0 && (grp >> lane) == 1u) {
        (pass == 0 ? e_ovs : e_ove)[s * EV + jk] = jv;
      }
      any_join |= __any_sync(FULL, jk >= 0);
    }

    if (lane == 0) {
      s_nv[s] = nv;
      {  // colour - background per channel, three signed 10-bit fields
        const int dr = (int)(col & 255u) - (int)(rd.bg & 255u);
        const int dg = (int)((col >> 8) & 255u) - (int)((rd.bg >> 8) & 255u);
        const int db = (int)((col >> 16) & 255u) - (int)((rd.bg >> 16) & 255u);
        s_dr[s] = (dr & 1023) | ((dg & 1023) << 10) | ((db & 1023) << 20);
      }
      s_nh[s] = __popc(hmask);
      s_hasov[s] = any_join ? 1 : 0;
      s_pymax[s] = p_ymax;
      s_r0[s] = r0;
      s_rcnt[s] = rcnt;
      s_bsh[s] = bsh;
      // output region whose 2-D tap window can see this sprite's bounding box
      short yo0 = 0, yo1 = -1, xo0 = 0, xo1 = -1;
      if (nv > 0 && xmx >= 0 && xmn < rd.CW && ymx >= 0 && ymn < rd.CH) {
        const int cx0 = max(xmn, 0), cx1 = min(xmx, rd.CW - 1);
        const int cy0 = max(ymn, 0), cy1 = min(ymx, rd.CH - 1);
        xo0 = rd.ax.first_out[cx0]; xo1 = rd.ax.last_out[cx1];
        yo0 = max((int)rd.ay.first_out[cy0], yo_b0);
        yo1 = min((int)rd.ay.last_out[cy1], yo_b1 - 1);
      }
      s_region[s * 4 + 0] = yo0; s_region[s * 4 + 1] = yo1;
      s_region[s * 4 + 2] = xo0; s_region[s * 4 + 3] = xo1;
      // tile plan: equal row blocks whose canvas rows (<= ny*aa + 32) fit the H buffer, equal
      // column blocks of at most TILE_X_MAX (columns need no halo, rows do)
      int pny = 1, pnx = 1;
      if (yo1 >= yo0 && xo1 >= xo0) {
        // regions are at most band_rows x W outputs; the reciprocal table covers divisors <= 64
        const int rh = yo1 - yo0 + 1, rw = xo1 - xo0 + 1;
        const int ny_cap = rd.ny_cap;  // (ny-1)*aa + len <= HT_ROWS
        const int nty = ny_cap <= 64 ? div20(rh + ny_cap - 1, ny_cap) : 1;
        pny = nty <= 64 ? div20(rh + nty - 1, nty) : (rh + nty - 1) / nty;
        const int ntx = div20(rw + TILE_X_MAX - 1, TILE_X_MAX);
        pnx = ntx <= 64 ? div20(rw + ntx - 1, ntx) : (rw + ntx - 1) / ntx;
      }
      s_pny[s] = pny; s_pnx[s] = pnx;
      s_pinvh[s] = (int)c_inv20[(pnx + H_NC - 1) / H_NC];
    }
  }
  __syncthreads();
  SWB_MARK(3);

  // ---- phase B: visible segments per canvas row -----------------------------------------
  // B1 (sprite, row)-parallel: Pillow's scan conversion of one canvas row of one sprite
  //    (crossings of the edges that span the row, with the corner-join overrides; sort;
  //    ROUND_UP/ROUND_DOWN pairing; horizontal edges) -> <= M merged spans.
  // B2 row-parallel: fold the sprites' spans front to back into the row's visible segments.
  // Sprites are taken front to back in chunks whose scratch (crossing lists + spans) fits.
  for (int s_hi = S; s_hi > 0;) {
    int s_lo = s_hi, rows_used = 0;
    if (s_hi == S && S <= 32) {  // common case: every sprite fits in one chunk
      const int rows_all = __reduce_add_sync(FULL, lane < S ? s_rcnt[lane] : 0);
      if (rows_all * CAP * 4 + S * n_rows * M * 4 <= L.scratch_bytes) { s_lo = 0; rows_used = rows_all; }
    }
    while (s_lo > 0) {
      const int rows_next = rows_used + s_rcnt[s_lo - 1];
      const int need = rows_next * CAP * 4 + (s_hi - s_lo + 1) * n_rows * M * 4;
      if (s_lo != s_hi && need > L.scratch_bytes) break;
      rows_used = rows_next;
      --s_lo;
    }
    uint32_t *s_spans = reinterpret_cast<uint32_t *>(smem + L.off_scratch) + rows_used * CAP;
    const bool fits = rows_used * CAP * 4 + (s_hi - s_lo) * n_rows * M * 4 <= L.scratch_bytes;
    if (!fits) { s_overflow = 1; }
    SWB_MARK(4);
    // B1
    for (int t = tid; fits && t < rows_used; t += R_THREADS) {
      int s = s_lo, rem = t;
      while (rem >= s_rcnt[s]) { rem -= s_rcnt[s]; ++s; }
      const int y = s_r0[s] + rem;
      const int base = s * EV, p_ymax = s_pymax[s];
      const bool hasov = s_hasov[s] != 0;
      float *xx = reinterpret_cast<float *>(smem + L.off_scratch) + (size_t)t * CAP;
      int j = 0;
      bool ovf = false;
      // edges whose rows overlap this row's bucket, in table order
      unsigned active = s_emask[s * 8 + (rem >> s_bsh[s])];
      while (active) {
        const int u = base + __ffs(active) - 1;
        active &= active - 1;
        const uint32_t yr = e_yr[u];
        const int ymin = (int)(short)(yr & 0xFFFFu), ymax = (int)(short)(yr >> 16);
        if (y < ymin || y > ymax) continue;
        float x = edge_x_at(y, e_y0[u], e_dx[u], e_x0[u]);
        if (hasov) {  // corner-join overrides (rare: acute same-direction corners only)
          if (y == ymin) { const float o = e_ovs[u]; if (!isnan(o)) x = o; }
          if (y == p_ymax && y == ymax) { const float o = e_ove[u]; if (!isnan(o)) x = o; }
        }
        const int twice = (y == ymax && y < p_ymax) ? 2 : 1;  // edge ending on an interior row
        if (j + twice <= CAP) {
          xx[j] = x;
          if (twice == 2) xx[j + 1] = x;
          j += twice;
        } else {
          ovf = true;
        }
      }
      for (int a = 1; a < j; ++a) {  // insertion sort, ascending
        const float v = xx[a];
        int b = a - 1;
        while (b >= 0 && xx[b] > v) { xx[b + 1] = xx[b]; --b; }
        xx[b + 1] = v;
      }
      uint32_t *dst = s_spans + ((size_t)(s - s_lo) * n_rows + (y - row_b0)) * M;
      const int nh = s_nh[s];
      if (M == 1) {
        // convex shapes: every span of the row merges into one [lo, hi] (kept in registers)
        int lo = 1, hi = 0;
        auto merge1 = [&](int xs, int xe) {
          if (lo > hi) { lo = xs; hi = xe; }
          else if (xs <= hi + 1 && lo <= xe + 1) { lo = min(lo, xs); hi = max(hi, xe); }
          else ovf = true;  // a second, disjoint span: the engine was sized for convex shapes
        };
        int x_pos = 0;
        for (int i = 1; i < j; i += 2) {
          const int x_end = round_down_f(xx[i]);
          if (x_end < x_pos) continue;
          if (xx[i - 1] > (float)x_pos) {
            x_pos = round_up_f(xx[i - 1]);
            if (x_end < x_pos) continue;
          }
          const int xs = max(x_pos, 0), xe = min(x_end, rd.CW - 1);
          if (xs <= xe) merge1(xs, xe);
          x_pos = x_end + 1;
        }
        for (int h = 0; h < nh; ++h) {
          const short *hl = s_hl + ((size_t)s * EV + h) * 3;
          if (hl[0] == y) merge1(hl[1], hl[2]);
        }
        dst[0] = lo <= hi ? ((uint32_t)lo | ((uint32_t)hi << 16)) : 0x0000FFFFu;
      } else {
        int lxs[MAX_ROW_SPANS], lxe[MAX_ROW_SPANS];
        int n = 0;
        int x_pos = 0;
        for (int i = 1; i < j; i += 2) {
          const int x_end = round_down_f(xx[i]);
          if (x_end < x_pos) continue;
          if (xx[i - 1] > (float)x_pos) {
            x_pos = round_up_f(xx[i - 1]);
            if (x_end < x_pos) continue;
          }
          const int xs = max(x_pos, 0), xe = min(x_end, rd.CW - 1);
          if (xs <= xe) add_span(lxs, lxe, n, xs, xe, ovf);
          x_pos = x_end + 1;
        }
        for (int h = 0; h < nh; ++h) {
          const short *hl = s_hl + ((size_t)s * EV + h) * 3;
          if (hl[0] == y) add_span(lxs, lxe, n, hl[1], hl[2], ovf);
        }
        if (n > M) ovf = true;
        for (int k = 0; k < M; ++k)
          dst[k] = k < n ? ((uint32_t)lxs[k] | ((uint32_t)lxe[k] << 16)) : 0x0000FFFFu;
      }
      if (ovf) s_overflow = 1;
    }
    __syncthreads();
    SWB_MARK(5);
    // B2
    for (int ry = tid; fits && ry < n_rows; ry += R_THREADS) {
      const int y = row_b0 + ry;
      uint32_t *seg = s_segs + (size_t)ry * SEGCAP;
      int nseg = s_nseg[ry];
      bool ovf = false;
      for (int s = s_hi - 1; s >= s_lo; --s) {
        const int rem = y - s_r0[s];
        if (rem < 0 || rem >= s_rcnt[s]) continue;
        const uint32_t *spn = s_spans + ((size_t)(s - s_lo) * n_rows + ry) * M;
        for (int k = 0; k < M; ++k) {
          const uint32_t w = spn[k];
          const int b = (int)(w >> 16);
          int cursor = (int)(w & 0xFFFFu);
          // the span minus what nearer sprites already cover, kept sorted by x
          for (int i = 0; i <= nseg && cursor <= b; ++i) {
            int gap_end = b;
            int next_cursor = b + 1;
            if (i < nseg) {
              const int xs = (int)(seg[i] & 0xFFFu), xe = (int)((seg[i] >> 12) & 0xFFFu);
              if (xe < cursor) continue;
              gap_end = min(b, xs - 1);
              next_cursor = xe + 1;
            }
            if (cursor <= gap_end) {
              if (nseg < SEGCAP) {
                for (int m = nseg; m > i; --m) seg[m] = seg[m - 1];
                seg[i] = (uint32_t)cursor | ((uint32_t)gap_end << 12) | ((uint32_t)s << 24);
                ++nseg;
                ++i;  // the segment we compared against moved one slot up
              } else {
                ovf = true;
              }
            }
            cursor = max(cursor, next_cursor);
          }
        }
      }
      s_nseg[ry] = (uint8_t)nseg;
      if (ovf) s_overflow = 1;
    }
    __syncthreads();
    SWB_MARK(6);
    s_hi = s_lo;
  }

  {  // background fill of the staged frame (the scratch area is free again)
    const uint32_t r = rd.bg & 255u, g = (rd.bg >> 8) & 255u, b = (rd.bg >> 16) & 255u;
    const int n_bytes = n_yo * rd.W * 3;
    if (r == g && g == b) {
      const uint32_t w = r * 0x01010101u;
      if ((n_bytes & 15) == 0) {
        uint4 *f128 = reinterpret_cast<uint4 *>(s_frame);
#pragma unroll 1
        for (int i = tid; i < (n_bytes >> 4); i += R_THREADS) f128[i] = make_uint4(w, w, w, w);
      } else {
        uint32_t *f32 = reinterpret_cast<uint32_t *>(s_frame);
        for (int i = tid; i < (n_bytes + 3) / 4; i += R_THREADS) f32[i] = w;
      }
    } else {
      for (int i = tid; i < n_yo * rd.W; i += R_THREADS) {
        s_frame[3 * i] = (uint8_t)r; s_frame[3 * i + 1] = (uint8_t)g; s_frame[3 * i + 2] = (uint8_t)b;
      }
    }
  }
  __syncthreads();
  SWB_MARK(7);

  // ---- phase C: per sprite region, in tiles sized to the H buffer -----------------------
  // A region of h x w outputs is cut into ceil(h/32) row blocks and, per row block, into as
  // few equal column blocks as fit HT_ITEMS H values (columns need no halo, rows do).
  const int bg_r = rd.bg & 255u, bg_g = (rd.bg >> 8) & 255u, bg_b = (rd.bg >> 16) & 255u;
  const uint32_t bg_h = (uint32_t)bg_r | ((uint32_t)bg_g << 10) | ((uint32_t)bg_b << 20);
  for (int s = 0; s < S; ++s) {
    const int ryo0 = s_region[s * 4 + 0], ryo1 = s_region[s * 4 + 1];
    const int rxo0 = s_region[s * 4 + 2], rxo1 = s_region[s * 4 + 3];
    if (ryo1 < ryo0 || rxo1 < rxo0) continue;
    const int ny_blk = s_pny[s], nx_blk = s_pnx[s];
    for (int ty0 = ryo0; ty0 <= ryo1; ty0 += ny_blk) {
      const int ny = min(ny_blk, ryo1 - ty0 + 1);
      const uint32_t yw0 = s_ywin[ty0 - yo_b0], yw1 = s_ywin[ty0 + ny - 1 - yo_b0];
      const int tr0 = (int)(int16_t)(yw0 & 0xFFFFu);
      const int tr1 = (int)(int16_t)(yw1 & 0xFFFFu) + (int)((yw1 >> 16) & 0xFFu);  // exclusive
      const int nr = tr1 - tr0;
      for (int tx0 = rxo0; tx0 <= rxo1; tx0 += nx_blk) {
        const int nx = min(nx_blk, rxo1 - tx0 + 1);
        // ---- H pass: a thread owns NC columns (c0, c0 + cs, ...) and strides over the canvas
        // rows, so the window start/length and tap prefix table are loop invariants and the
        // row's segment records are decoded once for all its columns ----
        {
          const int cs = (nx + H_NC - 1) / H_NC;  // column stride = threads per canvas row
          const uint32_t inv_cs = nx == nx_blk ? (uint32_t)s_pinvh[s] : c_inv20[cs];
          const int rgroup = (int)(((uint32_t)tid * inv_cs) >> 20);  // tid / cs
          const int c0 = tid - rgroup * cs;
          const int rstride = (int)(((uint32_t)R_THREADS * inv_cs) >> 20);  // row groups per pass
          if (rgroup < rstride) {
            const uint32_t prefix0 = (uint32_t)__cvta_generic_to_shared(s_prefix);
            int xmin[H_NC], len[H_NC], kk[H_NC];
            uint32_t pp[H_NC];
            bool on[H_NC];
#pragma unroll
            for (int q = 0; q < H_NC; ++q) {
              const int c = c0 + q * cs;
              on[q] = c < nx;
              const uint32_t xw = s_xwin[tx0 + (on[q] ? c : c0)];
              xmin[q] = (int)(int16_t)(xw & 0xFFFFu);
              len[q] = (int)((xw >> 16) & 0xFFu);
              pp[q] = prefix0 + (xw >> 24) * 33u * 4u;
              kk[q] = lds_s32(pp[q] + ((uint32_t)len[q] << 2));  // sum of the window's taps
            }
            const uint32_t dcol_addr = (uint32_t)__cvta_generic_to_shared(s_dr);
            // raw shared-window addresses, advanced by one row group per iteration
            uint32_t nseg_addr = (uint32_t)__cvta_generic_to_shared(s_nseg) + (uint32_t)(tr0 + rgroup - row_b0);
            uint32_t seg_row = (uint32_t)__cvta_generic_to_shared(s_segs) +
                               (uint32_t)((tr0 + rgroup - row_b0) * SEGCAP) * 4u;
            uint32_t ht_addr = (uint32_t)__cvta_generic_to_shared(s_ht) + (uint32_t)(rgroup * TILE_X_MAX + c0) * 4u;
            const uint32_t seg_step = (uint32_t)(rstride * SEGCAP) * 4u;
            const uint32_t ht_step = (uint32_t)(rstride * TILE_X_MAX) * 4u;
            const uint32_t ht_col = (uint32_t)cs * 4u;
            for (int r = rgroup; r < nr; r += rstride, nseg_addr += rstride, seg_row += seg_step, ht_addr += ht_step) {
              const int nseg = (int)lds_u8(nseg_addr);
              uint32_t hv[H_NC];
#pragma unroll
              for (int q = 0; q < H_NC; ++q) hv[q] = bg_h;
              if (nseg) {
                uint32_t seg_addr = seg_row;
                int ar[H_NC], ag[H_NC], ab[H_NC];
#pragma unroll
                for (int q = 0; q < H_NC; ++q) {
                  ar[q] = bg_r * kk[q] + (1 << 21);
                  ag[q] = bg_g * kk[q] + (1 << 21);
                  ab[q] = bg_b * kk[q] + (1 << 21);
                }
                int j = nseg;
#pragma unroll 1
                do {
                  const uint32_t w = (uint32_t)lds_s32(seg_addr);
                  seg_addr += 4u;
                  const int xs = (int)(w & 0xFFFu), xe1 = (int)((w >> 12) & 0xFFFu) + 1;
                  const int d = lds_s32(dcol_addr + ((w >> 24) << 2));
                  const int dr = (d << 22) >> 22, dg = (d << 12) >> 22, db = (d << 2) >> 22;
#pragma unroll
                  for (int q = 0; q < H_NC; ++q) {
                    const int a = min(max(xs - xmin[q], 0), len[q]), b = min(max(xe1 - xmin[q], 0), len[q]);
                    const int wt = lds_s32(pp[q] + ((uint32_t)b << 2)) - lds_s32(pp[q] + ((uint32_t)a << 2));
                    ar[q] += dr * wt; ag[q] += dg * wt; ab[q] += db * wt;
                  }
                } while (--j);
#pragma unroll
                for (int q = 0; q < H_NC; ++q)
                  hv[q] = clip8_q22(ar[q]) | (clip8_q22(ag[q]) << 10) | (clip8_q22(ab[q]) << 20);
              }
#pragma unroll
              for (int q = 0; q < H_NC; ++q)
                if (on[q]) sts_u32(ht_addr + (uint32_t)q * ht_col, hv[q]);
            }
          }
        }
        __syncthreads();
        SWB_MARK(8);
        // ---- V pass: item = (output row ly, column c).  Consecutive interior rows start 5
        // canvas rows = 100 words = 4 banks apart in the H tile, so a warp takes 16 columns of
        // two output rows four apart (16 banks apart): its 32 loads of one tap hit 32 banks.
        // Rows go in blocks of eight, (j, j + 4); a last block of <= 4 rows pairs (j, j + h);
        // columns past 16 (at most four) go as warps of 8 rows x 4 columns ----
        const int nb8 = ny >> 3, m8 = ny & 7;
        const int hstep = m8 > 4 ? 4 : ((m8 + 1) >> 1);
        const int n_pair = 4 * nb8 + hstep;
        const int n_chunk = n_pair + (nx > 16 ? ((ny + 7) >> 3) : 0);
        for (int chunk = warp; chunk < n_chunk; chunk += NWARP) {
          int ly, c;
          if (chunk < n_pair) {
            const int blk = chunk >> 2;
            ly = (blk << 3) + (chunk & 3) + ((lane >> 4) ? (blk < nb8 ? 4 : hstep) : 0);
            c = lane & 15;
          } else {
            ly = ((chunk - n_pair) << 3) + (lane >> 2);
            c = 16 + (lane & 3);
          }
          if (ly >= ny || c >= nx) continue;
          const int yo = ty0 + ly;
          const uint32_t yw = s_ywin[yo - yo_b0];
          const int rbase = (int)(int16_t)(yw & 0xFFFFu) - tr0;
          const int cls = (int)(yw >> 24);
          int ar = 1 << 21, ag = 1 << 21, ab = 1 << 21;
          const uint32_t *col_ht = s_ht + rbase * TILE_X_MAX + c;
          if (cls == rd.a5_cls) {
            // interior rows of a 5x reduction: taps (k, 28-k) pair up, 4/9/19/24/29 are zero,
            // 14 is the centre -> immediate load offsets, coefficients from the constant bank
            constexpr int A5[12] = {0, 1, 2, 3, 5, 6, 7, 8, 10, 11, 12, 13};
#pragma unroll
            for (int k = 0; k < 12; ++k) {
              const int kk = rd.a5_coef[k];
              // two taps of equal coefficient: the 10-bit fields hold sums up to 510
              const uint32_t t = col_ht[A5[k] * TILE_X_MAX] + col_ht[(28 - A5[k]) * TILE_X_MAX];
              ar += (int)(t & 1023u) * kk;
              ag += (int)((t >> 10) & 1023u) * kk;
              ab += (int)(t >> 20) * kk;
            }
            const int kk = rd.a5_coef[12];
            const uint32_t t = col_ht[14 * TILE_X_MAX];
            ar += (int)(t & 1023u) * kk;
            ag += (int)((t >> 10) & 1023u) * kk;
            ab += (int)(t >> 20) * kk;
          } else {
            const int32_t *prog = rd.ay.program + cls * PROG_STRIDE;  // global, read-only
            const int np = __ldg(prog), ns = __ldg(prog + 1) & 0xFFFF;
            const int2 *pp = reinterpret_cast<const int2 *>(prog + 2);
            const int2 *ps = reinterpret_cast<const int2 *>(prog + 2 + 2 * 16);
            for (int k = 0; k < np; ++k) {
              const int2 pk = __ldg(pp + k);  // (row a | row b << 8, coefficient)
              const uint32_t t = col_ht[(pk.x & 255) * TILE_X_MAX] + col_ht[(pk.x >> 8) * TILE_X_MAX];
              ar += (int)(t & 1023u) * pk.y;
              ag += (int)((t >> 10) & 1023u) * pk.y;
              ab += (int)(t >> 20) * pk.y;
            }
            for (int k = 0; k < ns; ++k) {
              const int2 pk = __ldg(ps + k);
              const uint32_t t = col_ht[pk.x * TILE_X_MAX];
              ar += (int)(t & 1023u) * pk.y;
              ag += (int)((t >> 10) & 1023u) * pk.y;
              ab += (int)(t >> 20) * pk.y;
            }
          }
          // staged in destination order: output row yo is row H-1-yo of the frame (np.flipud)
          uint8_t *px = s_frame + ((size_t)(yo_b1 - 1 - yo) * rd.W + tx0 + c) * 3;
          px[0] = (uint8_t)clip8_q22(ar);
          px[1] = (uint8_t)clip8_q22(ag);
          px[2] = (uint8_t)clip8_q22(ab);
        }
        __syncthreads();
        SWB_MARK(9);
      }
    }
  }

  // ---- phase D: staged frame -> HBM.  The frame is staged in destination order (rows already
  // flipped, np.flipud), so the band is one contiguous block: one thread hands it to the
  // bulk-copy engine (TMA, cp.async.bulk shared -> global), once per target.  With several
  // targets the same block also goes to the other ranks' buffers over NVLink peer memory:
  // the frame gather of the multi-GPU path is issued by the kernel that produced the frame
  // and costs it one instruction per rank -------------------------------------------------
  const int row_bytes = rd.W * 3;
  const int n_bytes = n_yo * row_bytes;
  const size_t band_off = (size_t)(targets.env_offset + e) * rd.H * row_bytes + (size_t)(rd.H - yo_b1) * row_bytes;
  if ((n_bytes & 15) == 0 && (band_off & 15) == 0) {
    // the tiles were written through the generic proxy; make them visible to the async proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(s_frame);
      // peers in an order rotated by rank and CTA, so that at any moment the ranks' copies are
      // spread over all receivers instead of all hitting rank 0 first, then rank 1, ...
      int t = 0;
      if (kPeers) {  // (self + 1 + blockIdx.x) mod n without a division (n <= 8)
        t = targets.self + 1 + (int)(blockIdx.x & 7u);
        while (t >= targets.n) t -= targets.n;
      }
      for (int i = 0; i < (kPeers ? targets.n : 1); ++i) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     : : "l"((kPeers ? s_dst[t] : targets.dst[0]) + band_off), "r"(src), "r"(n_bytes) : "memory");
        if (kPeers && ++t == targets.n) t = 0;
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      // the CTA's shared memory must stay until the engine has read it
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  } else {
    for (int i = tid; i < n_bytes; i += R_THREADS) {
      const uint8_t val = s_frame[i];
      targets.dst[0][band_off + i] = val;
      if (kPeers)
        for (int t = 1; t < targets.n; ++t) targets.dst[t][band_off + i] = val;
    }
  }
  // several bands of one env may race here, but they all OR in the same bit
  if (tid == 0 && s_overflow) st.render_status[e] |= (uint8_t)SWB_ENV_SPAN_OVERFLOW;
  SWB_MARK(10);
}

}  // namespace swb
