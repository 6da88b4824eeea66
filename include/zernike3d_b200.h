/*
 * zernike3d_b200 - C ABI of the B200-native spharm 3D-Zernike decomposition / reconstruction path.
 *
 * The reference (charbj/zernike3d) has no FFI / plugin seam: its hot path is three Python methods on
 * ZernikeSpherical (zernike_master.py, "zm" below).  Each entry point here names the reference
 * code it replaces; INTEGRATION.md shows the ctypes binding a maintainer would add to zm.
 *
 * Conventions
 *   - every function returns 0 on success and a Z3D_ERR_* code otherwise (no exceptions or aborts
 *     cross the boundary); z3d_last_error() returns a thread-local description of the last failure;
 *   - the caller owns all buffers; "d_" pointers are device pointers on the plan's device, "h_" are
 *     host pointers; device entry points are asynchronous on `stream` (a cudaStream_t passed as
 *     void*, NULL = legacy default stream);
 *   - volumes are float32 [batch][N][N][N], C-contiguous, indexed exactly like `mrc.data` in
 *     zm:2851-2852 / the 'xy' meshgrid of zm:2241 (array index [i][j][k] -> y=side[i], x=side[j],
 *     z=side[k]);
 *   - moments are complex64 (float re, float im) [batch][z3d_num_moments(order)] in the reference's
 *     flat order: sorted (n, l, m), m >= 0 only (zm:378-382, indexing.py:191-228), already scaled by
 *     coeff(n) = 2(2n+3) / (3*sqrt(3)*pi*N^3) (zm:2198);
 *   - there is no CPU fallback: every entry point that computes needs a CUDA device.
 */
#ifndef ZERNIKE3D_B200_H
#define ZERNIKE3D_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define Z3D_OK 0
#define Z3D_ERR_INVALID 1   /* bad argument */
#define Z3D_ERR_CUDA 2      /* CUDA runtime error (see z3d_last_error) */
#define Z3D_ERR_NODEVICE 3  /* no CUDA device / wrong architecture */
#define Z3D_ERR_WORKSPACE 4 /* workspace too small */
#define Z3D_ERR_UNSUPPORTED 5

typedef struct z3d_plan z3d_plan;

/* Library version (major * 10000 + minor * 100 + patch). */
int z3d_version(void);

/* Thread-local text of the last error raised by any z3d_* call on this thread. */
const char *z3d_last_error(void);

/* Index arithmetic (host only, no device needed).
 * z3d_num_moments  = mu_(order) + 1        indexing.py:140-162   (12051 at order 50)
 * z3d_num_radial   = number of (n, l) pairs = mu_R(order) + 1    indexing.py:13-28 (676 at order 50)
 * z3d_num_harmonics= number of (l, m>=0)    = mu_Y(order)        indexing.py:87-99 (1326 at order 50)
 * z3d_nlm_table fills nlm[3 * num_moments] with (n, l, m) per flat index (indexing.py:191-228). */
int z3d_num_moments(int order);
int z3d_num_radial(int order);
int z3d_num_harmonics(int order);
int z3d_nlm_table(int order, int *nlm);

/* Plan = everything ZernikeSpherical.CalculatePolynomials (zm:2237-2323) prepares, minus the
 * materialised basis: recurrence coefficient tables (zm:1992-1998, 2013-2024), per-l tile tables and
 * the launch geometry, uploaded once to `device`.  No R_nl / Y_lm voxel grid is ever built.  Orders 0..74 and box sizes
 * 2..4096 are supported (Z3D_ERR_UNSUPPORTED beyond: the per-pass coefficient tables are sized for order 74; the
 * reference's documented use is order <= 60).  For batches of >= 8 volumes (12 without the streamed-T table) the first z3d_decompose call additionally
 * builds the plan's folded R_nl / Q_lm factor tables on the device (see z3d_plan_info entry 12). */
int z3d_plan_create(int order, int dim, int device, z3d_plan **out);
int z3d_plan_destroy(z3d_plan *plan);

/* Plan queries (host only). info[]: 0 order, 1 dim, 2 num_moments, 3 passes, 4 CTAs per launch,
 * 5 threads per CTA, 6 tiles per thread, 7 dynamic smem bytes, 8 voxels per chunk, 9 SM count,
 * 10 useful accumulators, 11 padded accumulators, 12 smallest batch routed to the TABLE-FED batched analysis kernels
 * (0 = unavailable for this order / box), 13 row groups of the 32-volume tile kernel, 14 of the 64-volume one,
 * 15 smallest remaining batch that takes a 64-volume tile, 16 / 17 rounds (contraction launches per tile) of the two;
 * streamed-T variant of the batched path (preferred when its table fits in device memory): 18 state (0 unavailable,
 * 1 planned, 2 table resident), 19 column sets, 20 splits of the chunk list (sets x splits CTAs), 21 moment rows padded
 * to 16, 22 size of the T table in MiB, 23 chunks of 8 folded voxels per volume;
 * table-free tensor-core variant (the default for batches): 24 available, 25 column sets, 26 splits, 27 accumulator columns over all
 * sets, 28 depth of the shared-memory T ring, 29 smallest batch routed to it, 30 basis-table level in use (z3d_plan_prepare_tables),
 * 31 columns of the widest set; its two-tile layout (while more than 64 volumes are left, 128 volumes share every generated
 * T block): 32 in use by the analysis (opt-in), 33 column sets, 34 splits, 35 depth of the T ring;
 * tensor-core batched synthesis (z3d_reconstruct_batched): 36 available, 37 smallest batch of maps worth routing to it,
 * 38 stored class-partial slots (column sets, minus one per cluster pair of sets of the same kind). */
int z3d_plan_info(const z3d_plan *plan, int *info, int n_info);

/* Host-only planner query (no device needed): the column-set layout of the streamed-T batched analysis kernel for `order` on a
 * device with `num_sms` SMs.  out[]: 0 available, 1 column sets, 2 splits of the chunk list, 3 moment rows padded to 16,
 * 4 columns of the widest set, 5 violated invariants (0 = the layout passed its self-check), 6 / 7 most panels / segments of
 * a set. */
int z3d_stream_layout(int order, int num_sms, int *out, int n_out);

/* Host-only planner query (no device needed): the column-set layout of the table-free tensor-core batched analysis kernel
 * (the default for batches) for `order` on a device with `num_sms` SMs.  out[]: 0 available, 1 column sets, 2 splits of the chunk
 * list, 3 accumulator columns over all sets, 4 columns of the widest set, 5 violated invariants (0 = self-check passed),
 * 6 / 7 / 8 most a panels / MMA segments / recurrence-chain tasks of a set, 9 most seed rows, 10 estimated clocks per chunk. */
int z3d_gen_layout(int order, int num_sms, int *out, int n_out);
/* ... of the layout for `ntile` = 1 or 2 tiles of 64 volumes per generated T block (2: the sets hold half the columns - two
 * accumulator banks and two sets of a-panel slots share the 512 tensor-memory columns - and the chunk list has half the splits). */
int z3d_gen_layout_tiles(int order, int num_sms, int ntile, int *out, int n_out);

/* Opt-in basis tables.  The default path never materialises a basis function: batches run through a tensor-core kernel that
 * generates R_nl * Q_lm in registers.  A caller with memory to spare may instead ask for the table-fed kernels of round 1:
 * level 1 = folded R_nl / Q_lm factor tables (1.1 GB at 128^3 / order 50), level 2 = additionally the folded product table
 * T = R Q (13.7 GB) that the TMA engine streams into the tensor core's operand layout.  Synchronous (the tables are complete on
 * return, usable from any stream afterwards); max_bytes caps the T table (0 = 45 % of the free device memory); *level_out
 * receives the level reached (a table that does not fit is not an error); level 0 frees the tables again.  Not thread-safe
 * against concurrent calls on the same plan. */
int z3d_plan_prepare_tables(z3d_plan *plan, int level, size_t max_bytes, int *level_out);

/* Bytes of device scratch z3d_decompose / z3d_reconstruct need for `batch` volumes. */
size_t z3d_workspace_bytes(const z3d_plan *plan, int batch);

/* Analysis  (replaces ZernikeSpherical.Decompose -> _calculate_Z_GPU, zm:2492-2504, 2194-2217):
 *   moments[b][mu(n,l,m)] = coeff(n) * sum_v vol[b][v] * R_nl(r_v) * conj(Y_lm(v)).
 * `half_lo`, `half_hi` select the voxel slab this call sums over, in units of mirrored plane pairs of
 * array axis 0: pair p in [half_lo, half_hi) is planes i = p and i = N-1-p (0 <= p < ceil(N/2)).
 * Pass 0, ceil(N/2) for the whole volume; a sub-range yields the slab's partial (already scaled)
 * moments, which add up across slabs (the multi-GPU path all-reduces them).  Only the planes of the
 * selected pairs are read from d_vol. */
int z3d_decompose(const z3d_plan *plan, const float *d_vol, int batch, int half_lo, int half_hi,
                  float *d_moments /* [batch][num_moments][2] */, void *d_workspace, size_t workspace_bytes,
                  void *stream);

/* Same sum, sharded by WORK instead of by planes: the full-volume work list (row groups paired with their
 * transposes, which is what lets the kernel fold 16 mirror images per evaluated basis value) is cut into `nparts`
 * equal ranges and this call sums range `part`.  d_vol must hold the whole volume(s); the partial moments of all
 * parts add up to z3d_decompose over the whole volume.  This is the balanced way to spread ONE volume over
 * several GPUs (replicate the 4*N^3 bytes, all-reduce the moments). */
int z3d_decompose_part(const z3d_plan *plan, const float *d_vol, int batch, int part, int nparts,
                       float *d_moments, void *d_workspace, size_t workspace_bytes, void *stream);

/* Fused analysis + all-reduce for ONE volume (batch) spread over `world` GPUs, one process per GPU.  No reference
 * counterpart (the reference has no multi-GPU path).  z3d_comm_create allocates this rank's peer-visible buffer and
 * returns its 64-byte CUDA IPC handle; the ranks exchange the handles out of band (torch.distributed) and call
 * z3d_comm_connect with all of them.  z3d_decompose_allreduce then runs the rank's share of the work list
 * (as z3d_decompose_part(rank, world)) followed by ONE kernel that reduces the rank's partials, pushes them into every
 * rank's peer-mapped buffer with posted NVLink stores (block by block, one flag per block and source rank) and sums the
 * rows that arrived in its own buffer in rank order: a one-shot all-gather + ordered sum, every rank ends with the full,
 * bit-identical moments, without an NCCL call on the data path.  d_vol must hold the whole volume(s).
 * z3d_comm_error returns 1 if an exchange timed out (~2 s; the affected results are NaN) - it never hangs. */
typedef struct z3d_comm z3d_comm;
int z3d_comm_create(const z3d_plan *plan, int batch_max, z3d_comm **out, unsigned char *handle64);
int z3d_comm_connect(z3d_comm *comm, int rank, int world, const unsigned char *handles /* [world][64] */);
int z3d_comm_destroy(z3d_comm *comm);
int z3d_comm_error(z3d_comm *comm);
int z3d_decompose_allreduce(const z3d_plan *plan, z3d_comm *comm, const float *d_vol, int batch, float *d_moments,
                            void *d_workspace, size_t workspace_bytes, void *stream);

/* Synthesis (replaces ZernikeSpherical.Reconstruct -> _calculate_cum_GPU, zm:2642-2715, 2596-2614):
 *   vol[b][v] = Re sum_{n,l,m>=0} [ O R_nl Y_lm + (m != 0) conj(O) R_nl conj(Y_lm) ].
 * Only the planes of pairs [half_lo, half_hi) are written (the rest of d_vol is left untouched), so
 * slab-sharded reconstruction needs no collective. */
int z3d_reconstruct(const z3d_plan *plan, const float *d_moments /* [batch][num_moments][2] */, int batch,
                    int half_lo, int half_hi, float *d_vol /* [batch][N][N][N] */, void *d_workspace,
                    size_t workspace_bytes, void *stream);

/* Batched synthesis on the tensor cores, for ensembles / latent trajectories of maps (the loop over volumes of Reconstruct,
 * zm:2642-2715, and of the 3DVA consumer, moment_analysis_3dva.py:197-246): the same sum as z3d_reconstruct over the WHOLE box,
 * evaluated for tiles of 64 moment sets as a GEMM over the moment rows (tcgen05.mma kind::tf32, FP32-accurate 3xTF32 split; the
 * basis operand R_nl Q_lm is generated inside the kernel, nothing is materialised).  A tile costs the same whatever it holds, so
 * callers route batches below z3d_plan_info entry 37 to z3d_reconstruct.  Needs its own (larger) workspace. */
size_t z3d_reconstruct_batched_workspace_bytes(const z3d_plan *plan, int batch);
int z3d_reconstruct_batched(const z3d_plan *plan, const float *d_moments /* [batch][num_moments][2] */, int batch,
                            float *d_vol /* [batch][N][N][N] */, void *d_workspace, size_t workspace_bytes, void *stream);

/* Rotational invariants F_nl = || Omega_nl. ||_2 over m = -l..l  (Moments.calculate_invariants,
 * zm:62-76): d_invariants [batch][num_radial] float32 in sorted (n, l) order. */
int z3d_invariants(const z3d_plan *plan, const float *d_moments, int batch, float *d_invariants, void *stream);

/* 3DVA latent scaling on the device (replaces Moments.apply_scaling, zm:78-95, for many latent coordinates at once):
 *   out[p][mu] = dims[0][mu] + sum_k latents[p][k] * dims[k + 1][mu],   p < npoints, k < ncomp
 * d_dims [ncomp + 1][num_moments] complex64 (base map, then components: the .npz's arr_0..arr_K), d_latents [npoints][ncomp]
 * float32 (rows of the .npz's `latents`), d_out [npoints][num_moments] complex64 - ready for z3d_reconstruct. */
int z3d_scale_moments(const z3d_plan *plan, const float *d_dims, int ncomp, const float *d_latents, int npoints,
                      float *d_out, void *stream);

/* Per-order correlation of two moment sets ("Zernike FSC"; replaces Z_3DVA.calc_ZCC, moment_analysis_3dva.py:402-422):
 *   out[b][n] = sum |A conj(B)| / sqrt(sum |A|^2 sum |B|^2) over the stored moments (m >= 0) of order n, n = 0..order.
 * d_moments_a / d_moments_b [batch][num_moments] complex64, d_out [batch][order + 1] float32; 0 / 0 gives NaN like numpy. */
int z3d_zcc(const z3d_plan *plan, const float *d_moments_a, const float *d_moments_b, int batch, float *d_out, void *stream);

/* Pre-processing of an ensemble on the device (replaces ZernikeSpherical.bin_crop_array_center zm:2781-2815 and the --bin branch
 * of run() zm:2867-2876, which loop over host arrays with scipy): crop = 0 keeps the box, else the centred crop
 * v[c - half : c + half], half = (crop - crop % 2) / 2, around (cz, cy, cx); then nearest-neighbour resampling to n_out^3 with
 * the index arithmetic of scipy.ndimage.zoom(order = 0) (n_out = round(box / bin) is the caller's; n_out = box: crop only).
 * d_in [batch][n_in]^3, d_out [batch][n_out]^3 float32.  No plan needed. */
int z3d_crop_bin(const float *d_in, int batch, int n_in, int cz, int cy, int cx, int crop, int n_out, float *d_out, void *stream);

/* Host-buffer convenience calls: the same operations for callers that hold numpy / C arrays, the way
 * zm's Decompose(volume, ...) receives `mrc.data` and Reconstruct() returns a host map.  They stage
 * through pinned memory owned by the plan, overlap H2D copies with compute in chunks of volumes, and
 * return after the results are in the host buffers.  Not thread-safe per plan.  z3d_reconstruct_host sends staged groups of at
 * least z3d_plan_info entry 37 maps through the tensor-core synthesis kernel (z3d_reconstruct_batched; its workspace is allocated
 * by the plan on first use). */
int z3d_decompose_host(z3d_plan *plan, const float *h_vol, int batch, float *h_moments);
int z3d_reconstruct_host(z3d_plan *plan, const float *h_moments, int batch, float *h_vol);

/* Streaming form of z3d_decompose_host for ensembles that arrive batch after batch: _submit enqueues the copies and kernels of one
 * batch and returns; the next batch's host-to-device copies then overlap this batch's last kernels and its device-to-host copy.
 * Both host buffers must stay valid (and h_moments unread) until z3d_host_wait returns.  Pinned host memory is needed for the
 * copies to be asynchronous. */
int z3d_decompose_host_submit(z3d_plan *plan, const float *h_vol, int batch, float *h_moments);
int z3d_host_wait(z3d_plan *plan);

/* Number of kernels launched by this library since load (all plans, all threads); bench.py reports
 * the difference across its timed region as "gpu_launches". */
unsigned long long z3d_launch_count(void);

/* Kernel timing for the roofline report: when enabled, z3d_decompose / z3d_reconstruct record CUDA events on
 * the caller's stream around their dominant kernel (k_decompose / k_reconstruct); z3d_profile_last_ms waits
 * for the last such kernel and returns its duration. */
int z3d_profile_enable(z3d_plan *plan, int on);
int z3d_profile_last_ms(z3d_plan *plan, float *ms);

/* FP32 FMA micro-benchmark used to pin the roofline denominator: runs `iters`
 * dependent-chain FFMA rounds on every SM and returns achieved TFLOP/s in *tflops.
 * variant 0 = scalar FFMA, 1 = packed fma.rn.f32x2. */
int z3d_fma_peak(int device, int variant, int iters, double *tflops);

/* Measured dense TF32 tensor-pipe throughput of `device` in TFLOP/s: tcgen05.mma kind::tf32 (M 128, N 256, K 8) issued back to
 * back on every SM, iters x 8 instructions per SM.  The roofline denominator of the batched analysis kernel in bench.py. */
int z3d_tf32_peak(int device, int iters, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* ZERNIKE3D_B200_H */
