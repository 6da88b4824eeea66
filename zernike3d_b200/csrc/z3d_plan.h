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
//   tile   = 4 consecutive n  x  4 consecutive (m, re/im) columns of one l  (16 accumulators).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace z3d {

constexpr int NT = 256;          // threads per CTA
constexpr int NWARP = NT / 32;
constexpr int TS = 4;            // tile slots per thread (TS * 16 accumulators)
constexpr int KC = 8;            // folded voxels per chunk
constexpr int SUBS = 32 / KC;    // recurrence chains a warp advances together (one per KC lanes)
constexpr int RS = 420;          // smem words per voxel for the R table  (== 4 mod 32: conflict-free k stride)
constexpr int PS = 1412;         // smem words per voxel for the a*P table (== 4 mod 32)
constexpr int MAXTASK = 12;      // generation tasks per warp
constexpr int CTAS_PER_SM = 2;
constexpr int SLOT_WORDS = NT * TS * 16;  // accumulator words one CTA owns

struct DevPlan {
    int order, dim, H;
    int npass, nsplit;
    int dim_vec4;               // dim % 4 == 0 -> float4 row loads
    const double *side;         // [dim]   grid coordinate, float64 like the reference
    const float *beta;          // [(L+1)^2] monic Legendre recurrence coefficient, beta[m*(L+1)+l]
    const float4 *rk;           // [num_radial] (K1, K2, K3, 0) of zm:1992-1998, l-major
    const int *rbase;           // [L+1]   offset of (l, n=l) in rk
    const int *roffw;           // [L+1]   word offset of l's R rows inside its pass's R table
    const int *poffw;           // [L+1]   word offset of l's columns inside its pass's P table
    const int *lpass;           // [L+1]   pass that owns l
    const int *ownl;            // [npass][L+1] own l values per pass (ascending), -1 padded
    const int *nown;            // [npass]
    const int *ntiles;          // [npass]
    const int2 *tiles;          // [npass][NT*TS] (R word offset, P word offset) per tile slot
    const int *gen_tasks;       // [npass][NWARP][MAXTASK]  (type << 16 | base), type 0 = end
};

}  // namespace z3d
