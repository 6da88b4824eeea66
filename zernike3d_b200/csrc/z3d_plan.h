// zernike3d_b200 - plan structures shared by the host planner and the sm_100a kernels.
//
// Geometry of the problem (reference: zernike_master.py, "zm"):
//   voxel [i][j][k] of an N^3 box sits at y=side[i], x=side[j], z=side[k], side = linspace(-1/sqrt3,
//   1/sqrt3, N) (zm:2240-2242).  Along the contiguous axis k only z changes, so phi = atan2(y, x) and
//   rho^2 = x^2 + y^2 are per-row constants; the grid is mirror symmetric in all three axes.
//   A "row group" is the 4 rows (+-y, +-x) of one folded (i, j); folding z as well gives H = ceil(N/2)
//   "folded voxels" per row group, each standing for 8 mirror images that share R_nl and, up to
//   signs, P_lm, cos(m phi), sin(m phi).
//
// Work decomposition:
//   pass   = a subset of the l values (same parity) whose output tiles fit one CTA's registers;
//   split  = a subset of the row groups; grid = npass * nsplit CTAs, 2 resident per SM;
//   tile   = 8 consecutive n  x  8 consecutive (m, re/im) columns of one l  (64 accumulators): with
//            16 operand words loaded per 64 FMAs the contraction is balanced between the shared-memory
//            read path (128 B/clk/SM) and the FP32 pipes (128 FMA/clk/SM); smaller tiles are LDS bound.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace z3d {

constexpr int NT = 320;          // threads per CTA: one 8x8 tile each, 2 CTAs (20 warps) per SM, <= 96 registers
constexpr int NWARP = NT / 32;
constexpr int TS = 1;            // tile slots per thread
constexpr int TILE = 8;          // tile edge
constexpr int TW = TILE * TILE;  // accumulator words per tile
constexpr int KC = 8;            // folded voxels per chunk
constexpr int SUBS = 32 / KC;    // lanes per voxel
constexpr int CPL = 1;           // recurrence chains per lane
constexpr int GRP = SUBS * CPL;  // chains a warp advances together
constexpr int RS = 452;          // smem words per voxel for the R panel   (== 4 mod 32: conflict-free k stride)
constexpr int PS = 1476;         // smem words per voxel for the a*P panel (== 4 mod 32)
constexpr int GD_MAX = 1536;     // float2 entries of the per-pass Legendre coefficient table in smem
constexpr int RK_MAX = 640;      // float4 entries of the per-pass radial coefficient table in smem
constexpr int MAXTASK = 8;       // generation tasks per warp
constexpr int FLUSH_ITEMS = 8;    // row-group items between flushes of the register tiles (two-level summation)
constexpr int CTAS_PER_SM = 2;
constexpr int SLOT_WORDS = NT * TS * TW;  // accumulator words one CTA owns

struct TileDesc {  // smem word offsets of the two 16-byte halves of the R rows and P columns of one tile slot.
    int r0, r1, p0, p1;  // The half loaded first is chosen per lane so that a quarter-warp's LDS.128 hits 8
};                     // distinct bank groups; the planner's output maps absorb the resulting permutation.

struct DevPlan {
    int order, dim, H;
    int npass, nsplit;
    int dim_vec4;               // dim % 4 == 0 -> float4 row loads
    const double *side;         // [dim]   grid coordinate, float64 like the reference
    const float2 *gd;           // [npass][GD_MAX] (gamma, -delta) of the parity-decoupled Legendre recurrence
    const int *goff;            // [npass][L+1] offset of step 0 of chain m in the pass's gd table (pads in front)
    const float4 *rk;           // [npass][RK_MAX] (K1, K2, K3, 0) of zm:1992-1998 for the pass's own l, l-major
    const int *rbase;           // [L+1]   offset of (l, n=l) in its pass's rk table
    const int *roffw;           // [L+1]   word offset of l's R rows inside its pass's R panel
    const int *poffw;           // [L+1]   word offset of l's columns inside its pass's P panel
    const int *ownl;            // [npass][L+1] own l values per pass (ascending), -1 padded
    const int *nown;            // [npass]
    const int *ntiles;          // [npass]
    const TileDesc *tiles;      // [npass][TS][NT]
    const int *gen_tasks;       // [npass][NWARP][MAXTASK]  (type << 16 | base), type 0 = end
};

}  // namespace z3d
