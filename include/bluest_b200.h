/*
 * bluest_b200.h -- C ABI of libbluest_b200.so: the B200 (sm_100a, FP64) replacement for the
 * sample-allocation hot path of BLUEST (reference: croci/bluest).
 *
 * Two levels, mirroring SURVEY.md section 8(b):
 *
 *   Level 1  stateless drop-ins for the five routines of the reference's pybind11 module
 *            `_cmisc_bluest` (bluest/cmisc.cpp:99-110).  Same argument order, same
 *            "caller-allocated, pre-zeroed, accumulated in place" convention, HOST pointers.
 *   Level 2  a device-resident problem context that replaces `SAP.__init__` +
 *            `SAP.get_variance_functions` (bluest/sap.py:53-143) and the numpy orchestration in
 *            bluest/misc.py:453-516.  This is the boundary that matters for speed: the packed
 *            per-group inverses stay in HBM, one call evaluates Phi / variance / gradient / Hessian.
 *
 * All entry points are extern "C", take plain pointers and sizes, return an int status
 * (BLU_OK == 0) and never throw.  blu_last_error() returns a thread-local message for the last
 * non-zero status.  Pointers named h_* / without prefix are host memory; d_* are device
 * memory of the context's device.  Floating point is FP64, group indices are int64 (the reference's
 * `long int`, cmisc.cpp:10).  A context is used from one host thread at a time.
 *
 * Naming (reference convention): N = number of models, K = largest group
 * size, "size class k" = the Lk groups with exactly k models, L = sum_k Lk = number of groups.
 */
#ifndef BLUEST_B200_H
#define BLUEST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLU_OK            0
#define BLU_ERR_ARG       1   /* bad argument (NULL, size out of range, unsorted group, ...) */
#define BLU_ERR_CUDA      2   /* a CUDA runtime call or kernel failed */
#define BLU_ERR_STATE     3   /* call order violated (e.g. evaluate before inverses are set) */
#define BLU_ERR_NOMEM     4
#define BLU_ERR_NODEVICE  5   /* no CUDA device: the product path has no CPU fallback */

#define BLU_MAX_MODELS    32  /* group membership is a 32-bit mask */

/* status bits reported by the evaluation calls (out-parameter `flags`) */
#define BLU_FLAG_TINY      1u /* max|m| < 0.05: reference returns inf (misc.py:464,484,510) */
#define BLU_FLAG_NO_MODEL0 2u /* model 0 not in the support of m (misc.py:470 asserts) */
#define BLU_FLAG_PARTIAL   4u /* support is a strict subset of the models (singular Phi) */
#define BLU_FLAG_PEER_TIMEOUT 8u /* fused multi-GPU path: a peer rank never published its partial Phi (variance = NaN) */

/* `which` for blu_ctx_device_ptr */
#define BLU_BUF_M      0   /* (L)        sample vector of the last evaluation */
#define BLU_BUF_PHI    1   /* (N*N + 40) Phi(m) = delta I + sum m_i Psi_i, then 33 SUM-reducible indicators (sharded mode) */
#define BLU_BUF_PINV   2   /* (N,N)      pinv(Phi), symmetric */
#define BLU_BUF_GRAD   3   /* (L)        gradient of the variance */
#define BLU_BUF_U      4   /* (Lpad,NP)  row i = u_i = R_i^T Cinv_i R_i x, NP = 4*ceil(N/4) */
#define BLU_BUF_V      5   /* (Lpad,NP)  row i = 2 pinv(Phi) u_i; written only by evaluations that need it
                            * (dense Hessian, want_hess 1 or 2 / want_uv 1), not by the operator's (want_hess 3) */
#define BLU_BUF_HESS   6   /* (L,ldH)    dense Hessian, row pitch ldH = 16*ceil(L/16) doubles */
#define BLU_BUF_CINV   7   /* packed upper-triangular per-group inverses */
#define BLU_BUF_SCAL   8   /* (8) doubles: [0]=variance, [1]=max|m|, [2]=sweeps, [3]=lambda_max */

typedef struct blu_ctx blu_ctx;

const char *blu_last_error(void);
int  blu_device_count(void);               /* 0 when no usable CUDA device */
const char *blu_version(void);

/* ------------------------------------------------------------------------------------------
 * Level 2: device-resident context
 * ------------------------------------------------------------------------------------------ */

/* Replaces the bookkeeping half of SAP.__init__ (sap.py:57-87): uploads the group tables.
 *   sizes[k-1]   = Lk for k = 1..K (empty classes allowed, sap.py:78-79)
 *   groups_flat  = concatenation over k of the (Lk,k) row-major int64 tables (sap.py:77);
 *                  every group strictly increasing, entries in [0,N).
 * No covariance yet: call blu_ctx_set_covariance or blu_ctx_set_invcovs next. */
int blu_ctx_create(int device, int N, int K, const int64_t *sizes, const int64_t *groups_flat,
                   blu_ctx **out);
int blu_ctx_destroy(blu_ctx *ctx);
/* A second evaluation lane on the same problem: shares the parent's read-only HBM data (group tables, packed
 * inverses, work lists of the parent's current slice), owns its stream and the small per-evaluation buffers.
 * Evaluations on parent and clone run concurrently (the one-CTA serial tail of one overlaps the streaming kernels
 * of the other).  The parent must outlive the clone and keep its slice; set_covariance / set_invcovs / set_slice
 * on a clone fail with BLU_ERR_STATE. */
int blu_ctx_clone(blu_ctx *parent, blu_ctx **out);

/* Kernel (1): batched inversion of C[g_i,g_i] for every group (replaces the L calls of
 * np.linalg.pinv at sap.py:72-74).  One warp-slice per group, shuffle-based Gauss-Jordan; groups
 * whose pivots drop below `pivot_rtol` x diagonal are redone with a Jacobi-eigenvalue pseudo-inverse
 * using numpy's cutoff (1e-15 x sigma_max).  C is (N,N) row-major and may hold NaN where models are
 * never coupled.  n_fallback (optional) receives the number of groups that took the Jacobi route. */
int blu_ctx_set_covariance(blu_ctx *ctx, const double *C, double pivot_rtol, int64_t *n_fallback);

/* Ingest / read back inverses in the reference's layout: invcovs_k is the flat (Lk,k,k) array
 * `SAP.invcovs[k-1]` (sap.py:78).  Ingest symmetrises ((A+A^T)/2) and packs the upper triangle. */
int blu_ctx_set_invcovs(blu_ctx *ctx, int k, const double *invcovs_k);
int blu_ctx_get_invcovs(blu_ctx *ctx, int k, double *invcovs_k);

/* Dense psi (N*N, L) row-major, the matrix the host SDP builders read (sap.py:129,250,328). */
int blu_ctx_assemble_psi(blu_ctx *ctx, double *psi);

/* get_phi (misc.py:459-461): phi is (N,N) row-major. */
int blu_get_phi(blu_ctx *ctx, const double *m, double delta, double *phi);

/* variance (misc.py:463-477).  *var = inf with BLU_FLAG_TINY when max|m| < 0.05.
 * BLU_FLAG_NO_MODEL0 is where the reference asserts. */
int blu_variance(blu_ctx *ctx, const double *m, double delta, double *var, unsigned *flags);

/* variance_GH (misc.py:479-505).  grad: (L).  hess: (L,L) row-major host array or NULL
 * (nohess).  With BLU_FLAG_TINY set, *var = inf and grad is filled with inf (the reference's
 * 2-tuple early-out); hess is left untouched. */
int blu_variance_GH(blu_ctx *ctx, const double *m, double delta, double *var, double *grad,
                    double *hess, unsigned *flags);

/* Split form of blu_variance / blu_variance_GH: _begin queues the upload, the kernels and the
 * downloads on the context's stream and returns immediately; _end waits and delivers.  Contexts own
 * their streams, so the outputs of a MOSAP (mosap.py:91-100 loops over them) overlap on the device.
 * m is copied into pinned staging inside _begin (the caller's buffer may be reused at once); hess
 * (may be NULL) must stay valid until _end. */
int blu_variance_GH_begin(blu_ctx *ctx, const double *m, double delta, int want_grad, double *hess);
int blu_variance_GH_end(blu_ctx *ctx, double *var, double *grad, unsigned *flags);

/* The Hessian as an operator (replaces the dense result of hessKQ_c cmisc.cpp:74-97 + `hess += hess.T`
 * misc.py:497-503 wherever the caller only multiplies by it: scipy trust-constr's projected CG,
 * sap.py:410).  blu_variance_GH_factored = variance_GH with the Hessian kept FACTORED in HBM
 * (H = V U^T, U and V (L,NP)); blu_hess_matvec then returns out[v] = H p[v] for nvec host vectors
 * of length L stored back to back.  BLU_ERR_STATE when no factors are resident (never evaluated,
 * or the last factored evaluation took the BLU_FLAG_TINY early-out).  The _device form takes HBM
 * pointers and is asynchronous on the context's stream. */
int blu_variance_GH_factored(blu_ctx *ctx, const double *m, double delta, double *var, double *grad,
                             unsigned *flags);
int blu_hess_matvec(blu_ctx *ctx, const double *p, int nvec, double *out);
int blu_hess_matvec_device(blu_ctx *ctx, const double *d_p, double *d_out);

/* get_cleanup_matrix (misc.py:507-516): X is (N,L) row-major.  mode 0 reproduces the reference's
 * assignment semantics (cmisc.cpp:51, only l = k-1 survives); mode 1 returns the intended
 * X[:,i] = u_i.  Returns BLU_ERR_ARG-free status with BLU_FLAG_TINY when the reference would
 * raise ValueError. */
int blu_cleanup_matrix(blu_ctx *ctx, const double *m, double delta, int mode, double *X,
                       unsigned *flags);

/* BLUE estimator (compute_BLUE_estimator sap.py:99-119 + PHIinvY0 misc.py:518-544) -- the step after
 * sampling, "next" row of the scope table.  samples: (L) sample counts; sums_flat: the per-group
 * sample sums, k doubles per group in flat group order (sum_k Lk*k doubles).  Outputs: *mu, *var
 * (inf with BLU_FLAG_TINY), y (N, optional) = sum_i R_i^T Cinv_i sums_i.  BLU_FLAG_NO_MODEL0 where
 * the reference asserts. */
int blu_blue_estimator(blu_ctx *ctx, const double *samples, const double *sums_flat, double *mu,
                       double *var, double *y, unsigned *flags);

/* Integer projection, batched candidate evaluation ("next" row; misc.py:368-369 inside
 * best_closest_integer_solution_BLUE): for every column c of ms (LL, ncand; int64, row-major -- the
 * floor/ceil combinations of the LL <= 24 groups idx[]) form Phi_c = basephi + sum_t ms[t,c] Psi_idx[t]
 * and return Vs[c] = pinv(Phi_c, hermitian, rcond)[0,0].  basephi is (N,N) = psi @ baseval. */
int blu_candidate_variances(blu_ctx *ctx, const double *basephi, int LL, const int64_t *idx,
                            const int64_t *ms, int64_t ncand, double rcond, double *Vs);

/* Batched evaluation of SMALL problems: P problems x B sample vectors in ONE kernel launch, one CTA per pair --
 * the device form of the loop over outputs of mosap.py:86-100 (`SAPS[n].variance_GH(m[mappings[n]])`) and of the
 * instances of a budget / tolerance sweep, which are pure launch latency one at a time.
 *   ctxs   P contexts on one device, inverses set.  maps == NULL: one input vector is the concatenation of the
 *          problems' own sample vectors.  maps[p] (L_p indices into a shared vector of length Lm): mosap's mappings.
 *   blu_batch_eval: m (B, Lm) host row-major -> var, flags (P, B); grad (optional): for problem p a (B, L_p) block at
 *          offset B * (L_0 + ... + L_{p-1}), filled with inf where BLU_FLAG_TINY (misc.py:484).  No Hessian.
 *          Batches of up to 128 KB each way are evaluated directly on mapped pinned host buffers (no copy operations
 *          around the kernel); environment variable BLU_BATCH_NO_ZEROCOPY forces the explicit copies. */
typedef struct blu_batch blu_batch;
int blu_batch_create(blu_ctx **ctxs, int P, const int64_t *const *maps, int64_t Lm, blu_batch **out);
int blu_batch_eval(blu_batch *batch, const double *m, int B, double delta, double *var, unsigned *flags, double *grad);
int blu_batch_destroy(blu_batch *batch);

/* Structure-exploiting KKT solve ("next" row f1) for the semidefinite programme SAP.cvxopt_solve builds
 * (sap.py:242-307) -- what a `kktsolver` callback handed to cvxopt.solvers.sdp (sap.py:289 passes none, so cvxopt
 * factorises the dense KKT matrix) would call once per interior-point iteration and right-hand side.
 *   variables   x in R^n, n = L + has_t (has_t = 1: budget mode, x = [t, m/budget], sap.py:259-275; 0: tolerance mode)
 *   G0 = [-I_n; Gx]  Gx (nlin, n) row-major: the dense rows of the linear cone (cost, coverage, sample caps)
 *   G1          column of group i = -scales * vec(pad(Psi_i)) ((N+1) x (N+1), last row/column zero), t column = -E_NN
 *   d (n+nlin)  Nesterov-Todd scaling of the linear cone (W['d']), r ((N+1)^2) of the semidefinite block (W['r'][0])
 *   bx (n), bz (n + nlin + (N+1)^2)   right-hand side;  ux, uz: same shapes, the solution of
 *        [0 G^T; G -W^T W] [ux; uz] = [bx; bz]     (uz UNSCALED: a cvxopt kktsolver returns W uz)
 * Diagonal + rank (N+1)(N+2)/2 + nlin: one weighted Gram contraction over the packed inverses (FP64 tensor cores)
 * and a Cholesky of that order, instead of the dense (L+1)^3/3.  device_ms (optional): CUDA-event time of the
 * device part.  Limits: N <= 21 and (N+1)(N+2)/2 + nlin <= 255.  The device workspace (n x (Q+1) doubles and the
 * partial Gram tiles) stays with the context between calls and is released by blu_ctx_destroy. */
int blu_kkt_solve(blu_ctx *ctx, int has_t, double scales, int nlin, const double *Gx, const double *d, const double *r,
                  const double *bx, const double *bz, double *ux, double *uz, float *device_ms);

/* Device-resident evaluation: d_m lives on the context's device (or NULL to reuse BLU_BUF_M).
 * want_grad / want_hess select the work (want_hess == 2: the U,V factors only, no Hessian -- the
 * caller then asks for row panels with blu_shard_hess; want_hess == 3: the U factor only, all the
 * Hessian operator needs: H p = U (S (U^T p)), S = 2 pinv(Phi)); results stay in the context's buffers
 * (blu_ctx_device_ptr).  The call is asynchronous on the context's stream. */
int blu_eval_device(blu_ctx *ctx, const double *d_m, double delta, int want_grad, int want_hess);
int blu_ctx_sync(blu_ctx *ctx);
int blu_ctx_device_ptr(blu_ctx *ctx, int which, void **ptr, int64_t *nbytes);
int blu_ctx_stream(blu_ctx *ctx, void **cuda_stream);
/* Read back the status flags / variance of the last evaluation (synchronises). */
int blu_ctx_last_result(blu_ctx *ctx, double *var, unsigned *flags);
/* Time the device part of the last blu_eval_device call (CUDA events on the context's stream):
 * ms[0]=Phi+pinv, ms[1]=grad/U, ms[2]=Hessian, ms[3]=total. */
int blu_ctx_last_timing(blu_ctx *ctx, float *ms);
/* Per-evaluation event log: the next `capacity` blu_eval_device calls record their phase events into
 * a log (no host sync inside a timed loop); blu_ctx_timing_read synchronises and returns (n,4)
 * floats [phi+pinv, grad/U, Hessian, total] in ms. */
int blu_ctx_timing_log(blu_ctx *ctx, int capacity);
int blu_ctx_timing_read(blu_ctx *ctx, float *ms, int *n);
/* Options.  "soa" (default 1): gradient / U,V kernels read a second, group-interleaved copy of the
 * inverses, one group per lane (0: the entry-per-lane kernels on the group-major copy).
 * "sym_download" (default 1, used when L >= 4096): blu_variance_GH moves only the upper
 * block-triangle of the (exactly symmetric) dense Hessian over PCIe and mirrors it with host threads
 * (0: one plain copy of all 8 L^2 bytes).  "mirror_threads" (default 0 = automatic): host threads of
 * that mirroring.  "sym_full_rows_pct" (default -1 = adaptive, starting at 10): share (%) of the lower triangle that still
 * travels by DMA (the bottom rows, whole); adaptive = raised when the host threads lag the DMA, lowered when they idle.
 * "hess_onebuf": single staging buffer in the dense-Hessian kernel.  "phi_stages" (2, the default, or 4): ring
 * depth of the Phi kernel (4 = 2 KB chunks; rebuilds the kernel's work list; parent contexts only). */
int blu_ctx_set_option(blu_ctx *ctx, const char *name, int value);
/* Current value of an option (for "sym_full_rows_pct": the share the next download will use). */
int blu_ctx_get_option(blu_ctx *ctx, const char *name, int *value);
/* Number of kernels the last evaluation launched. */
int blu_ctx_last_launches(blu_ctx *ctx);

/* Profiling aid: 16 %globaltimer stamps (ns) left by the last CTA of the fused Phi kernel at its milestones
 * ([1] own stream done, [2] last of its group, [3] final fold starts, [4] rank sums complete, [5] peer exchange
 * complete, [6] Phi ready, [7] pinv written, [9] finish done).  Synchronises. */
int blu_ctx_last_stamps(blu_ctx *ctx, unsigned long long *stamps16);

/* CUDA graphs: everything the device-resident calls (blu_eval_device, blu_shard_eval_fused,
 * blu_hess_matvec_device, blu_ctx_save_result) enqueue between _begin and _end is recorded instead of run;
 * blu_ctx_graph_launch replays it `times` times on the context's stream (asynchronous).  Run the same calls
 * once eagerly first: lazily allocated buffers cannot be created while recording.  The reference has no
 * analogue; this is the device-side form of the solver loops of sap.py:410-416 / mosap.py:595-607, where
 * small problems are pure launch latency. */
int blu_ctx_graph_begin(blu_ctx *ctx);
int blu_ctx_graph_end(blu_ctx *ctx, int *graph_id);
int blu_ctx_graph_launch(blu_ctx *ctx, int graph_id, int times);
/* Stream-ordered copy of the last enqueued evaluation's variance / status flags to device memory of the
 * caller (either may be NULL): keeps the scalars of every evaluation of a batch or a graph. */
int blu_ctx_save_result(blu_ctx *ctx, double *d_var, unsigned *d_flags);
/* Gradient destination of the following evaluations (L doubles in HBM; NULL = BLU_BUF_GRAD). */
int blu_ctx_set_grad_output(blu_ctx *ctx, double *d_grad);

/* Group-sharded evaluation: a context may own a contiguous slice
 * [lo,hi) of the flat group enumeration.  The three phases are exposed separately so the host
 * can put the NCCL exchange between them:
 *   blu_shard_phi      partial Phi of the slice -> BLU_BUF_PHI (N*N doubles, no delta, no pinv)
 *   (all-reduce BLU_BUF_PHI across ranks)
 *   blu_shard_finish   delta*I, pinv, variance on the reduced Phi; grad and U,V rows of the slice
 *                      (want_uv: 0 none, 1 U and V, 2 U only -- enough for blu_shard_hv_*)
 *   (all-gather U,V rows)
 *   blu_shard_hess     rows [row_lo,row_hi) of the Hessian against all L columns.  Once U,V are
 *                      gathered any rank can produce any rows, so the row panels are balanced by
 *                      row count (equal bytes written) independently of the group slices, which
 *                      are balanced by k^2 work.
 *
 * Fused variant (one process per GPU on one NVLink box): the all-reduce is done INSIDE the finish
 * kernel over peer memory -- every rank publishes its partial Phi in a CUDA-IPC shared slot and reads
 * the other ranks' slots directly over NVLink, in rank order (bit-identical Phi on all ranks):
 *   blu_ctx_peer_handle    64-byte cudaIpcMemHandle_t of this rank's exchange buffer
 *   (exchange the handles between the processes with any host-side all-gather)
 *   blu_ctx_peer_connect   map the peers' buffers (handles = world x 64 bytes, rank order)
 *   blu_shard_eval_fused   partial Phi -> [local reduce + NVLink all-reduce + pinv + variance] ->
 *                          gradient (and U,V) of the slice; every rank must call it the same number
 *                          of times (the epoch counter is the call count). */
int blu_ctx_set_slice(blu_ctx *ctx, int64_t lo, int64_t hi);
int blu_ctx_peer_handle(blu_ctx *ctx, void *handle64);
int blu_ctx_peer_connect(blu_ctx *ctx, int world, int rank, const void *handles);
int blu_shard_eval_fused(blu_ctx *ctx, const double *d_m, double delta, int want_grad, int want_uv);
int blu_shard_phi(blu_ctx *ctx, const double *d_m);
int blu_shard_finish(blu_ctx *ctx, double delta, int want_grad, int want_uv);
int blu_shard_hess(blu_ctx *ctx, int64_t row_lo, int64_t row_hi);
/* Sharded Hessian operator (N = 20: the dense matrix is 8.8 TB and cannot exist).  p, out are (L,)
 * in HBM, only the owned rows [lo,hi) are read / written:
 *   blu_shard_hv_partial   d_t[0..31] = sum over owned rows of p_i u_i (zero beyond N)
 *   (all-reduce d_t across ranks: N doubles, the only exchange)
 *   blu_shard_hv_apply     d_out[i] = v_i . t for the owned rows */
int blu_shard_hv_partial(blu_ctx *ctx, const double *d_p, double *d_t);
int blu_shard_hv_apply(blu_ctx *ctx, const double *d_t, double *d_out);

/* Page-locked host memory for large results (the dense Hessian): pageable destinations make
 * the D2H copy several times slower. */
int blu_host_alloc(size_t bytes, void **out);
int blu_host_free(void *p);

/* ------------------------------------------------------------------------------------------
 * Kernel (4): pilot-sample covariance (blue_fn.py:159-167, blue_models.py:333).
 * Y is (n,N) row-major.  Outputs (host): s1 (N) column sums, S2 (N,N) Gram matrix Y^T Y,
 * C_hat (N,N) = S2/n - s1 s1^T/n^2.  FP64 DMMA contraction, deterministic two-stage reduce.
 * y_on_device != 0: Y is a device pointer. */
int blu_pilot_covariance(int device, const double *Y, int64_t n, int N, int y_on_device,
                         double *s1, double *S2, double *C_hat, float *kernel_ms);

/* The sums behind it, for n_out outputs at once and for a row split across devices (blue_fn.py:147-167 incl. the
 * MLMC difference sums, blue_fn.py:177-187 is the reduction between ranks):
 *   Y            n_out sample matrices (n, N) row-major, `ystride` doubles apart (0 = n*N); host or device
 *   telescoped   0: sums of Y.  1: sums of Z, Z_0 = Y_0, Z_j = Y_j - Y_{j-1} -- the difference sums of the MLMC
 *                pairs then come from Gram entries of differences and do not cancel against the model variances
 *   producer_stream  (cudaStream_t) stream that wrote a device-resident Y, or NULL: the kernel is ordered behind it
 *   sums         n_out x (N*N + N) doubles: per output the (symmetric) Gram matrix, then the N column sums; host
 *                memory, or device memory (sums_on_device) so that ranks holding different sample rows can add
 *                their sums with one all-reduce of (N*N + N) n_out doubles before blu_pilot_finalize.
 * (one extra tile column carries the ones that produce the column sums: N <= 32 means at most 5 tiles of 8). */
int blu_pilot_sums(int device, const double *Y, int64_t n, int N, int n_out, int64_t ystride, int y_on_device,
                   int telescoped, void *producer_stream, double *sums, int sums_on_device, float *kernel_ms);
/* Host arithmetic on the (all-reduced) sums of ONE output (n_total = samples over all ranks).  Outputs, each may
 * be NULL: s1 (N) = sumse, S2 (N,N) = sumsc, C_hat (N,N) = S2/n - s1 s1^T/n^2 (blue_models.py:333), d1 / d2 (N,N)
 * = sumsd1[i][j] / sumsd2[i][j] for i < j, 0 elsewhere (blue_fn.py:147-157), dV (N,N) = d2/n - (d1/n)^2 for i < j,
 * NaN elsewhere (blue_models.py:339). */
int blu_pilot_finalize(const double *sums, int64_t n_total, int N, int telescoped, double *s1, double *S2, double *C_hat,
                       double *d1, double *d2, double *dV);

/* ------------------------------------------------------------------------------------------
 * Level 1: drop-ins for `_cmisc_bluest` (bluest/cmisc.cpp).  Host pointers, in-place "+=".
 * ------------------------------------------------------------------------------------------ */
/* assemble_psi_c, cmisc.cpp:10-23 */
int blu_assemble_psi_c(double *psi, int N, int k, int Lk, const int64_t *groupsk,
                       const double *invcovsk);
/* objectiveK_c<double>, cmisc.cpp:25-40,104 ; objectiveK_c<long int>, cmisc.cpp:105 */
int blu_objectiveK_c(double *PHI, int N, int k, int Lk, const double *mk, const int64_t *groupsk,
                     const double *invcovsk);
int blu_objectiveK_c_i64(double *PHI, int N, int k, int Lk, const int64_t *mk,
                         const int64_t *groupsk, const double *invcovsk);
/* cleanupK_c, cmisc.cpp:42-56 (assignment semantics); N = leading dimension of X = number of models */
int blu_cleanupK_c(double *X, int N, int k, int Lk, const int64_t *groupsk, const double *invcovsk,
                   const double *invPHI_0);
/* gradK_c, cmisc.cpp:58-72; N = length of invPHI_0 */
int blu_gradK_c(double *grad, int N, int k, int Lk, const int64_t *groupsk, const double *invcovsk,
                const double *invPHI_0);
/* hessKQ_c, cmisc.cpp:74-97 */
int blu_hessKQ_c(double *hess, int N, int k, int q, int Lk, int Lq, const int64_t *groupsk,
                 const int64_t *groupsq, const double *invcovsk, const double *invcovsq,
                 const double *invPHI);

#ifdef __cplusplus
}
#endif
#endif /* BLUEST_B200_H */
