// Host-side driver for the closed-form index maps in csrc/geometry.cuh (they are
// __host__ __device__): dumps the maps so tests/test_host.py can compare them bit-exactly
// with the oracle / reference fixtures without a GPU.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../pangu_pytorch_b200/csrc/geometry.cuh"

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  const int Z = atoi(argv[1]), H = atoi(argv[2]), W = atoi(argv[3]), roll = atoi(argv[4]);
  FILE* f = fopen(argv[5], "wb");
  if (!f) return 3;
  const pg::Geo g = pg::make_geo(Z, H, W);
  const int Tp = g.nLon * g.types * 144, T = Z * H * W;
  std::vector<int> w2t(Tp), t2w(T), up(4 * Z * ((H + 1) / 2) * (W / 2));
  for (int r = 0; r < Tp; ++r) w2t[r] = pg::win_row_to_token(g, r, roll);
  for (int t = 0; t < T; ++t) t2w[t] = pg::token_to_win_row(g, t, roll);
  const int T2 = Z * ((H + 1) / 2) * (W / 2);
  for (int grp = 0; grp < 4; ++grp)
    for (int r = 0; r < T2; ++r) up[grp * T2 + r] = pg::upsample_row_to_token(g, r, grp);
  fwrite(w2t.data(), 4, w2t.size(), f);
  fwrite(t2w.data(), 4, t2w.size(), f);
  fwrite(up.data(), 4, up.size(), f);
  fclose(f);
  return 0;
}
