/*
 * torchsde_b200 — C ABI of the B200 (sm_100a) SDE-integration hot path.
 *
 * This header is the drop-in boundary of the package.  Every entry point is a
 * plain `extern "C"` function taking raw device pointers, sizes, scalars and a
 * CUDA stream handle; no torch type crosses it.  The Python host
 * (torchsde_b200/_cabi.py, ctypes) is the only caller inside this repository;
 * INTEGRATION.md shows the binding a torchsde maintainer would add.
 *
 * Each function replaces one "kernel-equivalent call site" of the reference
 * (google-research/torchsde v0.2.6).  The reference file:line a function
 * follows is cited above its declaration; paths are relative to the
 * reference tree (`torchsde/...`).
 *
 * Conventions
 *   - All tensors are dense, row-major, contiguous device buffers.
 *       state-like    : (rows, d)
 *       diagonal g    : (rows, d)          (d == m)
 *       general g     : (rows, d, m)       (scalar noise: m == 1; additive: same layout)
 *       noise W, U    : (rows, m)
 *   - `dtype` is TSDE_F32 or TSDE_F64 and applies to every tensor of a call.
 *   - Scalars (dt, coefficients) are passed as double and rounded ONCE to the
 *     tensor dtype on the host side of the kernel launch, which is what the
 *     reference's `tensor * python_scalar` / `tensor * 0-d tensor` does.
 *   - Arithmetic inside a tableau follows the reference's left-to-right
 *     evaluation order with separate IEEE roundings (no FMA contraction), so
 *     a diagonal-noise step fed identical inputs is bit-identical to the
 *     reference's sequence of ATen elementwise ops.
 *   - Every function enqueues work on `stream` and returns immediately; it
 *     never allocates and never synchronises, so it may be captured in a
 *     CUDA graph.  Return value: 0 on success, otherwise the cudaError_t of
 *     the launch (or TSDE_EINVAL for a contract violation).
 */
#ifndef TORCHSDE_B200_H_
#define TORCHSDE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSDE_ABI_VERSION 1
#define TSDE_EINVAL (-22)

enum { TSDE_F32 = 0, TSDE_F64 = 1 };

/* noise layouts (torchsde/settings.py:41-45 NOISE_TYPES) */
enum {
  TSDE_NOISE_DIAGONAL = 0, /* g:(rows,d)      W:(rows,d)                   base_sde.py:98-99   */
  TSDE_NOISE_GENERAL  = 1  /* g:(rows,d,m)    W:(rows,m)   also additive   base_sde.py:101-102 */
                           /* scalar noise is GENERAL with m == 1                               */
};

/* where a kernel takes the Brownian increment from */
enum {
  TSDE_SRC_MEMORY  = 0, /* read W (and U) from device buffers (any BaseBrownian)         */
  TSDE_SRC_COUNTER = 1, /* regenerate in registers from the Philox counter (fast path)   */
  TSDE_SRC_UNIT    = 2  /* W == 1: `g` already holds the user's g_prod (base_sde.py:51-56) */
};

/* Launch flags (tsde_noise.flags).
 * TSDE_FLAG_G_BROADCAST: every (rows,d,m) diffusion operand of this launch is ONE dense (d,m) block shared by
 * all rows (row stride 0) — what `sigma.expand(B, d, m)` is for additive noise whose diffusion does not depend on
 * y (the reference's own additive test problem materialises it with `repeat`, tests/problems.py:113-116; the
 * product base_sde.py:101-102 / misc.py:62-63 only ever reads it).  Honoured by the general-noise entry points
 * (noise_type GENERAL, m > 1); the pointer then addresses d*m elements. */
#define TSDE_FLAG_G_BROADCAST 1

/* Shape / stream descriptor shared by all launches. */
typedef struct tsde_launch {
  int32_t dtype;      /* TSDE_F32 | TSDE_F64                       */
  int32_t noise_type; /* TSDE_NOISE_*                              */
  int64_t rows;       /* trajectories held by this rank; 0 is a valid launch that does nothing (operand pointers
                         of an empty batch may be NULL)            */
  int64_t d;          /* state channels                            */
  int64_t m;          /* Brownian channels                         */
  void*   stream;     /* cudaStream_t                              */
} tsde_launch;

/*
 * Brownian increment over one solver step.
 *
 * Counter mode restates what BrownianInterval hands the solver
 * (torchsde/_brownian/brownian_interval.py:589-687) for an interval that is
 * the merge of `n_cells` consecutive primary cells of a grid node:
 *     W_c = sqrt(h_c) * N_W(key; cell_id + c; row; channel)          (:553-554)
 *     H_c = sqrt(h_c / 12) * N_H(key; cell_id + c; row; channel)     (:555-558)
 * merged left to right with the reference's aggregation rule (:643-672) and
 * returned as W and U = h (W/2 + H) (:102-103).  N_* are Philox4x32-10
 * Box-Muller normals; see torchsde_b200/csrc/philox.cuh for the bit-level
 * definition and oracle/philox.py for its CPU restatement.
 */
typedef struct tsde_noise {
  int32_t     source;     /* TSDE_SRC_*                                                    */
  int32_t     want_u;     /* also produce U (space-time Levy area, srk.py:61)              */
  const void* w;          /* MEMORY: (rows, m) increments                                  */
  const void* u;          /* MEMORY: (rows, m) U = h(W/2 + H), or NULL                     */
  const void* key;        /* COUNTER: device pointer to the 64-bit Philox key              */
  uint64_t    cell_id;    /* COUNTER: counter words 2,3 of the first cell                  */
  int64_t     row_offset; /* COUNTER: global index of local row 0 (batch sharding)         */
  int32_t     n_cells;    /* COUNTER: number of consecutive cells merged (>= 1)            */
  int32_t     flags;      /* TSDE_FLAG_* of this launch (0 = none)                         */
  double      h;          /* COUNTER: length of each cell when cell_h == NULL              */
  const double* cell_h;   /* COUNTER: DEVICE pointer to n_cells lengths, or NULL (uniform) */
  double      h_total;    /* COUNTER: tb - ta of the whole step (U = h_total (W/2 + H))    */
} tsde_noise;

int tsde_abi_version(void);
/* Last CUDA error string of this library's runtime instance (for diagnostics). */
const char* tsde_error_string(int code);
/* Diagnostics: how many launches of a kernel family this process has issued so far (tests use it to
   check which general-noise tile kernel a shape was routed to). */
#define TSDE_KERNEL_GEN_CTA 0  /* per-thread-load tile kernel                       */
#define TSDE_KERNEL_GEN_TMA 1  /* TMA-staged persistent tile kernel (bulk copies)   */
int64_t tsde_kernel_launches(int32_t family);

/* ------------------------------------------------------------------------ */
/* Brownian source  (replaces torchsde/_brownian/brownian_interval.py)       */
/* ------------------------------------------------------------------------ */

/*
 * Materialise the increment described by `nz` (COUNTER source) into device
 * buffers: out_w (rows,m) always; out_u (rows,m) = U if non-NULL; out_h
 * (rows,m) = H if non-NULL.  Replaces BrownianInterval.__call__ for an
 * interval made of whole cells: brownian_interval.py:589-687 (_randn :30-32,
 * top-level draw :551-558, merge :643-672, _H_to_U :102-103).
 */
int tsde_brownian_cells(const tsde_launch* L, const tsde_noise* nz,
                        void* out_w, void* out_u, void* out_h);

/*
 * Brownian-bridge descent: given the (W,H) of an ancestor interval in
 * in_w/in_h (in_h may be NULL when no Levy area is tracked), walk `depth`
 * binary splits down to a descendant and write its (W,H).
 * Level l uses node id ids[l] for its two normals X1,X2, `is_left[l]`, and the
 * split geometry (parent start/mid/end) times[3*l..3*l+2] (host doubles).
 * Replaces _Interval._increment_and_space_time_levy_area,
 * brownian_interval.py:188-241 (with H :199-225, W only :226-237).
 */
int tsde_brownian_bridge(const tsde_launch* L, const void* key, int64_t row_offset,
                         int32_t depth, const uint64_t* ids, const int32_t* is_left,
                         const double* times, const void* in_w, const void* in_h,
                         void* out_w, void* out_h);

/*
 * Merge the increment of [s,u] (w0,h0) with that of the adjacent [u,t]
 * (w1,h1): W = W0 + W1, H per brownian_interval.py:649-658, in place into
 * (w0,h0).  len0 = u - s, len1 = t - u, tot = t - s (each as the host computed it).
 * h pointers may be NULL.
 */
int tsde_brownian_merge(const tsde_launch* L, void* w0, void* h0, const void* w1,
                        const void* h1, double len0, double len1, double tot);

/*
 * One launch for the query pattern of the Levy-area methods on a solver grid: W (rows,m), U = h (W/2 + H) (rows,m)
 * and the Davie / Foster area A (rows,m,m) of ONE primary cell (nz: COUNTER source, n_cells == 1, cell nz->cell_id of
 * length nz->h), with H drawn from the counter as well and never materialised.  Replaces BrownianInterval.__call__
 * (ta, tb, return_U=True, return_A=True) for a whole cell: brownian_interval.py:589-687 with _randn :30-32, the
 * top-level law :551-558, _davie_foster_approximation :78-99 and _H_to_U :102-103.  Requires m % 4 == 0, 4 <= m <= 64.
 */
int tsde_brownian_cell_levy(const tsde_launch* L, const tsde_noise* nz, uint64_t a_id, int32_t foster, void* out_w,
                            void* out_u, void* out_a);

/*
 * `logqp=True` (diagonal noise): the KL-integrand augmentation of SDELogqp.f_and_g_diagonal
 * (base_sde.py:266-283, misc.py:66-68):  u = (f - h) / stable(g),  f_aug = [f, 0.5 sum_d u^2],  g_aug = [g, 0].
 * f, g, h are (rows,d); f_aug, g_aug are (rows,d+1).  L: noise_type DIAGONAL, d = m = state channels WITHOUT the
 * log-ratio channel.  (General noise solves a least-squares problem per row, `pinverse`: it stays in the host's
 * linear algebra.)
 */
int tsde_logqp_augment(const tsde_launch* L, const void* f, const void* g, const void* h, double eps,
                       void* f_aug, void* g_aug);

/*
 * GA = g A for the log-ODE correction (base_sde.py:170,191: `ga = torch.bmm(g, a)` inside
 * dg_ga_jvp_column_sum_v1/_v2; used by methods/log_ode.py:39-56).  g is (rows,d,m), a is (rows,m,m) (the Levy
 * area of the step), out_t is (m, rows, d): column l of g A as a contiguous (rows,d) slab, which is what the
 * column-wise jvp's through the user's g consume (base_sde.py:173-184).  noise_type must be GENERAL.
 */
int tsde_bmm_ga(const tsde_launch* L, const void* g, const void* a, void* out_t);

/* U = h (W/2 + H)   brownian_interval.py:102-103 */
int tsde_brownian_h_to_u(const tsde_launch* L, const void* w, const void* hh, double h, void* out_u);

/*
 * Davie / Foster Levy-area approximation of one interval
 * (brownian_interval.py:78-99): A = H (x) W - W (x) H + std * (N - N^T); foster != 0 selects Foster's std.
 * The antisymmetric noise N - N^T is drawn as one Philox normal per pair i < j of counter id `a_id`
 * (N_ij = z_ij / sqrt 2 = -N_ji: the law of the reference's antisymmetrised iid matrix, half the draws).
 * out_a is (rows, m, m).
 */
int tsde_brownian_levy_area(const tsde_launch* L, const void* key, int64_t row_offset,
                            uint64_t a_id, const void* w, const void* hh, double h,
                            int32_t foster, void* out_a);

/* A-merge of two adjacent intervals, brownian_interval.py:659-671, in place into a0. */
int tsde_brownian_merge_area(const tsde_launch* L, void* a0, const void* a1,
                             const void* w0, const void* w1);

/* ------------------------------------------------------------------------ */
/* Step tableaus  (replace torchsde/_core/methods/<name>.py  .step bodies)        */
/* `g*` arguments are (rows,d) for DIAGONAL and (rows,d,m) for GENERAL.      */
/* ------------------------------------------------------------------------ */

/* y1 = y0 + f*dt + g.dW                    methods/euler.py:36
 * also Heun predictor heun.py:42, midpoint corrector midpoint.py:43,
 * additive-noise Milstein milstein.py:72 with gdg == 0 (base_sde.py:157-158) */
int tsde_step_euler(const tsde_launch* L, const tsde_noise* nz, const void* y0,
                    const void* f, const void* g, double dt, void* y1);

/* grad_outputs of Milstein's vjp: go = g * (0.5 * v), v = dW^2 - dt (Ito) or dW^2
 * (Stratonovich).  methods/milstein.py:56,69,80-81,90-91; base_sde.py:127-155.
 * DIAGONAL: go (rows,d).  GENERAL (scalar noise): go (rows,d,m) = g * v2[:,None,:]. */
int tsde_milstein_vjp_seed(const tsde_launch* L, const tsde_noise* nz, const void* g,
                           double dt, int32_t ito, void* go);

/* y1 = y0 + f*dt + g.dW + gdg              methods/milstein.py:72 */
int tsde_step_milstein(const tsde_launch* L, const tsde_noise* nz, const void* y0,
                       const void* f, const void* g, const void* gdg, double dt, void* y1);

/* derivative-free Milstein, predictor: y' = y0 + (Ito ? dt*f : 0) + g*sqrt_dt
 * methods/milstein.py:58-63,83-84,93-94.  g is (rows,d) also for scalar noise (squeezed). */
int tsde_milstein_gf_predict(const tsde_launch* L, const void* y0, const void* f,
                             const void* g, double dt, double sqrt_dt, int32_t ito, void* yp);

/* derivative-free Milstein, corrector:
 * y1 = y0 + f*dt + g.dW + ((g'-g).v) / (2*sqrt_dt)      methods/milstein.py:64-72 */
int tsde_step_milstein_gf(const tsde_launch* L, const tsde_noise* nz, const void* y0,
                          const void* f, const void* g, const void* gp, double dt,
                          double two_sqrt_dt, int32_t ito, void* y1);

/* y1 = y0 + (dt*(f+f') + g.dW + g'.dW) * 0.5             methods/heun.py:46 */
int tsde_step_heun(const tsde_launch* L, const tsde_noise* nz, const void* y0, const void* f,
                   const void* fp, const void* g, const void* gp, double dt, void* y1);

/* y' = y0 + half_dt*f + 0.5*(g.dW)                        methods/midpoint.py:36-38 */
int tsde_midpoint_predict(const tsde_launch* L, const tsde_noise* nz, const void* y0,
                          const void* f, const void* g, double half_dt, void* yp);

/* y' = y0 + g.dW                                          methods/euler_heun.py:36 */
int tsde_euler_heun_predict(const tsde_launch* L, const tsde_noise* nz, const void* y0,
                            const void* g, void* yp);

/* y1 = y0 + dt*f + (g.dW + g'.dW)*0.5                     methods/euler_heun.py:40 */
int tsde_step_euler_heun(const tsde_launch* L, const tsde_noise* nz, const void* y0,
                         const void* f, const void* g, const void* gp, double dt, void* y1);

/* z1 = 2*y0 - z0 + f0*dt + g0.dW                          methods/reversible_heun.py:69 */
int tsde_reversible_heun_z(const tsde_launch* L, const tsde_noise* nz, const void* y0,
                           const void* z0, const void* f0, const void* g0, double dt, void* z1);

/* y1 = y0 + (f0+f1)*(0.5*dt) + (g0+g1).(0.5*dW)           methods/reversible_heun.py:71 */
int tsde_step_reversible_heun(const tsde_launch* L, const tsde_noise* nz, const void* y0,
                              const void* f0, const void* f1, const void* g0, const void* g1,
                              double half_dt, void* y1);

/*
 * Roessler SRI2 (srid2), diagonal or scalar noise: methods/srk.py:57-88 with
 * methods/tableaus/srid2.py:19-54.  The reference re-evaluates earlier stages;
 * only 3 distinct f and 4 distinct g evaluations exist (rows with zero
 * coefficients add exact zeros).  g arguments are (rows,d) (squeezed for scalar).
 *   stage1: H0_1 = y0 + f0*dt ; H1_1 = y0 + 1/4 f0 dt - 1/2 g0 sqrt_dt          (srk.py:74-75, s=1)
 *   stage2: H0_2 = y0 + 1/4 f0 dt + g0 U/dt + 1/4 f1 dt + 1/2 g1 U/dt ;
 *           H1_2 = y0 + f0 dt + g0 sqrt_dt                                      (s=2)
 *   stage3: H1_3 = y0 + 2 g0 sqrt_dt - g1 sqrt_dt + 1/4 f2 dt + 1/2 g2 sqrt_dt  (s=3)
 *   final : y1 = y0 + sum_s alpha_s f_s dt + g_s * gw_s                         (srk.py:80-87)
 */
int tsde_srk_diag_stage1(const tsde_launch* L, const void* y0, const void* f0, const void* g0,
                         double dt, double sqrt_dt, void* h0_1, void* h1_1);
int tsde_srk_diag_stage2(const tsde_launch* L, const tsde_noise* nz, const void* y0,
                         const void* f0, const void* g0, const void* f1, const void* g1,
                         double dt, double rdt, double sqrt_dt, void* h0_2, void* h1_2);
int tsde_srk_diag_stage3(const tsde_launch* L, const void* y0, const void* g0, const void* g1,
                         const void* f2, const void* g2, double dt, double sqrt_dt, void* h1_3);
int tsde_step_srk_diag(const tsde_launch* L, const tsde_noise* nz, const void* y0,
                       const void* f0, const void* f1, const void* f2, const void* g0,
                       const void* g1, const void* g2, const void* g3, double dt, double rdt,
                       double sqrt_dt, double three_dt, void* y1);

/*
 * Roessler SRA1, additive noise: methods/srk.py:90-111 with tableaus/sra1.py:19-36.
 *   stage : H0_1 = y0 + 3/4 f0 dt + gA.(3/2 U/dt)            gA = g(t1, y0)
 *   final : y1 = y0 + 1/3 f0 dt + gA.(W - U/dt) + 2/3 f1 dt + gB.(U/dt)   gB = g(t0, y0)
 */
int tsde_srk_additive_stage(const tsde_launch* L, const tsde_noise* nz, const void* y0,
                            const void* f0, const void* ga, double dt, double rdt, void* h0_1);
int tsde_step_srk_additive(const tsde_launch* L, const tsde_noise* nz, const void* y0,
                           const void* f0, const void* f1, const void* ga, const void* gb,
                           double dt, double rdt, void* y1);

/* ys[i] = (t1-t)/(t1-t0)*y0 + (t-t0)/(t1-t0)*y1     _core/interp.py:15-18; w0,w1 host-computed */
int tsde_linear_interp(const tsde_launch* L, const void* y0, const void* y1, double w0,
                       double w1, void* out);

/* Adaptive stepping: out[0] (double) = sum over all (rows*d) elements of ((y11 - y12)/tol)^2 with
 * tol = clamp_min(rtol*max(|y11|,|y12|) + atol, eps)   _core/adaptive_stepping.py:42-76 (compute_error, _rms).
 * scratch: device buffer of >= 592 doubles.  The host finishes with sqrt(out/numel).clamp_min(eps). */
int tsde_adaptive_error_sumsq(const tsde_launch* L, const void* y11, const void* y12, double rtol,
                              double atol, double eps, void* scratch, void* out);

/* ------------------------------------------------------------------------ */
/* Reversible-Heun adjoint  (methods/reversible_heun.py:98-144)              */
/* ------------------------------------------------------------------------ */

/* First half (:103-115): z1 = 2*y0 - z0 - f0*dt - g0.dW ;
 *   adj_f0' = adj_f0 + adj_y0*half_dt ; adj_g0' = adj_g0 + adj_of_prod(adj_y0, half_dW).
 * adj_g* are (rows,d) for DIAGONAL and (rows,d,m) otherwise (outer product :95-96). */
int tsde_adjoint_reversible_heun_a(const tsde_launch* L, const tsde_noise* nz, const void* y0,
                                   const void* z0, const void* f0, const void* g0,
                                   const void* adj_y0, const void* adj_f0, const void* adj_g0,
                                   double dt, double half_dt, void* z1, void* adj_f0_out,
                                   void* adj_g0_out);

/* Second half (:130-140): adj_z0' = adj_z0 + vjp_z ;
 *   y1 = y0 - (f0+f1)*half_dt - (g0+g1).half_dW ; adj_y1 = adj_y0 + 2*adj_z0' ;
 *   adj_z1 = -adj_z0' ; adj_f1 = adj_y0*half_dt + adj_z0'*dt ;
 *   adj_g1 = adj_of_prod(adj_y0, half_dW) + adj_of_prod(adj_z0', dW). */
int tsde_adjoint_reversible_heun_b(const tsde_launch* L, const tsde_noise* nz, const void* y0,
                                   const void* f0, const void* f1, const void* g0,
                                   const void* g1, const void* adj_y0, const void* adj_z0,
                                   const void* vjp_z, double dt, double half_dt, void* y1,
                                   void* adj_y1, void* adj_z1, void* adj_f1, void* adj_g1);

#ifdef __cplusplus
}
#endif
#endif /* TORCHSDE_B200_H_ */
