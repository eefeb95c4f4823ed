// Closed-form index maps of the Earth-specific window partition (bit-exact contracts,
// SURVEY.md Appendix A; reference models/layers.py:188-247, 441-451, 480-489).
#pragma once
#include <stdint.h>

namespace pg {

struct Geo {
  int Z, H, W;   // un-padded token grid
  int Hp;        // H + 5 (padding_back, models/layers.py:145)
  int nH;        // Hp / 6 latitude windows
  int nLon;      // W / 12 longitude windows
  int types;     // (Z/2) * nH window types
};

__host__ __device__ inline Geo make_geo(int Z, int H, int W) {
  Geo g;
  g.Z = Z; g.H = H; g.W = W;
  g.Hp = H + 5;
  g.nH = g.Hp / 6;
  g.nLon = W / 12;
  g.types = (Z / 2) * g.nH;
  return g;
}

// window-ordered row  [lw][t = zw*nH + hw][k = zl*72 + hl*12 + wl]  ->  natural token
// (z*H + h)*W + w, or -1 for a zero pad token.  shift = (1,3,6) when rolled.
__host__ __device__ inline int win_row_to_token(const Geo& g, int row, int roll) {
  const int k = row % 144;
  const int wt = row / 144;
  const int t = wt % g.types, lw = wt / g.types;
  const int zw = t / g.nH, hw = t % g.nH;
  const int zl = k / 72, hl = (k / 12) % 6, wl = k % 12;
  int zp = 2 * zw + zl, hp = 6 * hw + hl, wp = 12 * lw + wl;
  if (roll) {
    zp = (zp + 1) % g.Z;
    hp = (hp + 3) % g.Hp;
    wp = (wp + 6) % g.W;
  }
  if (hp >= g.H) return -1;
  return (zp * g.H + hp) * g.W + wp;
}

// natural token -> window-ordered row (inverse of the above on real tokens)
__host__ __device__ inline int token_to_win_row(const Geo& g, int tok, int roll) {
  int w = tok % g.W;
  int h = (tok / g.W) % g.H;
  int z = tok / (g.W * g.H);
  if (roll) {
    z = (z + g.Z - 1) % g.Z;
    h = (h + g.Hp - 3) % g.Hp;
    w = (w + g.W - 6) % g.W;
  }
  const int zw = z >> 1, zl = z & 1;
  const int hw = h / 6, hl = h % 6;
  const int lw = w / 12, wl = w % 12;
  return ((lw * g.types + zw * g.nH + hw) * 144) + zl * 72 + hl * 12 + wl;
}

// UpSample pixel shuffle (models/layers.py:480-489): row = low-res token (z, h2, w2) on the
// grid (Z, H2=(H+1)/2, W2=W/2) with `g` describing the HIGH-res grid; group = dh*2 + dw.
// Returns the high-res token or -1 when cropped (lat index == H).
__host__ __device__ inline int upsample_row_to_token(const Geo& g, int row, int group) {
  const int W2 = g.W / 2, H2 = (g.H + 1) / 2;
  const int w2 = row % W2, h2 = (row / W2) % H2, z = row / (W2 * H2);
  const int h = 2 * h2 + (group >> 1), w = 2 * w2 + (group & 1);
  if (h >= g.H) return -1;
  return (z * g.H + h) * g.W + w;
}

}  // namespace pg
