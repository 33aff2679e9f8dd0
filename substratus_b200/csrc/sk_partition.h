// sk_partition.h — how the stream-K projections (tc_gemm_sk_kernel) cut a weight matrix into per-CTA ranges, and how the
// owner of a tile finds the CTAs that hold the rest of it.  Plain integer arithmetic shared by the kernel, its launcher
// and a host-side exhaustive check (tests/test_sk_partition_cpu.py), so the two sides cannot drift.
//
// Units: the matrix is n_ntiles tiles of 128 weight rows, each tile nkb k-blocks of 64; unit u = tile * nkb + k-block,
// U = n_ntiles * nkb.  CTA b of a grid of G streams units [sk_begin(b), sk_begin(b + 1)).  A range is cut at tile
// boundaries into segments; a segment that does not start at k-block 0 is a CONTRIBUTOR part (always the first segment of
// its CTA, dumped to workspace slot 2 b), the segment of a tile that starts at k-block 0 belongs to the tile's OWNER, which
// adds the parts of the following CTAs (slots 2 (b + 1), 2 (b + 2), ...) and runs the fused epilogue.
#pragma once
#if defined(__CUDACC__)
#define SK_HD __host__ __device__ __forceinline__
#else
#define SK_HD inline
#endif

constexpr int SK_MAX_CONTRIB = 12;  // the launcher sizes the grid so that a tile never spans more than 10 ranges

// first unit of CTA b's range; 32-bit on purpose (64-bit division is a software routine on the GPU): sk_fits() guards it
SK_HD int sk_begin(int b, int U, int G) { return (int)((unsigned)b * (unsigned)U / (unsigned)G); }
SK_HD bool sk_fits(long long U, int G) { return U * (G + 1) < (1ll << 31); }

// grid of the launch: as many CTAs as SMs, but never so many that a tile (nkb units) spans more than 8 full + 2 partial ranges
SK_HD int sk_grid(long long U, int nkb, int n_sm) {
  const long long min_units = (nkb + 7) / 8;
  const long long gmax = U / min_units;
  return (int)(gmax < 1 ? 1 : (gmax < n_sm ? gmax : n_sm));
}

// owner side: CTA b finished its own segment of the tile at unit seg_end; the tile ends at t_end.  Number of following
// CTAs (b + 1, b + 2, ...) whose first segment lies inside this tile.  Needs non-empty ranges (G <= U).
SK_HD int sk_contributors(int b, int U, int G, int seg_end, int t_end) {
  int n = 0;
  for (int cu = seg_end; cu < t_end && n < SK_MAX_CONTRIB; ++n) cu = sk_begin(b + n + 2, U, G);
  return n;
}
