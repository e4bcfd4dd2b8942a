/* ffb200.h -- C ABI of libffb200.so: B200 (sm_100a) kernels for the flowfusion hot path.
 *
 * The reference (Cosmo-Pop/flowfusion) has no FFI: its narrowest seam is the call
 *   torchdiffeq.odeint(func, y0, t, rtol=, atol=, method=, options=)
 * made from diffusion.py:621,631,734,744, flow.py:288,299,358,371,781,792,855,869 and
 * symplectic.py:237, plus the two hand-rolled loops diffusion.py:543-562 (Euler-Maruyama)
 * and symplectic.py:192-197 (forward Euler).  The entry points below are what a binding for
 * that seam needs: they take plain device pointers, sizes and a cudaStream_t; no torch types.
 *
 * Conventions
 *   - all tensors are FP32, row-major, contiguous, resident on the current CUDA device;
 *   - every call only enqueues work on `stream` (a cudaStream_t passed as void*);
 *   - return value 0 = OK, negative = ffb_status; ffb_last_error() gives the message;
 *   - the caller owns every buffer, including `scratch` (ffb_scratch_bytes) and `partials`;
 *   - per-evaluation scalars (time features, SDE coefficients) are computed by the HOST in
 *     the reference's exact FP32 op order and passed in ffb_eval_scalars, so the kernels
 *     never re-derive beta(t), sigma(t), sin/cos(2 pi t W) with different rounding.
 */
#ifndef FFB200_H
#define FFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFB_ABI_VERSION 5
#define FFB_MAX_LAYERS 16   /* Linear layers per network (more than 8: wide engine, FP32 pipe)     */
#define FFB_MAX_WIDTH 512   /* widest hidden layer; above 128 the wide engine (FP32 pipe) runs the field */
#define FFB_MAX_TFEAT 32    /* time-feature columns (embedding_dimensions or 1)  */
#define FFB_MAX_STATE 128   /* ODE state columns                                 */
#define FFB_TILE_ROWS 128   /* rows (trajectories x (1 + tangents)) per CTA tile */
#define FFB_NPART 16        /* doubles per tile in the `partials` buffer         */

typedef enum {
  FFB_OK = 0,
  FFB_ERR_ARG = -1,        /* bad argument / unsupported shape       */
  FFB_ERR_CUDA = -2,       /* CUDA runtime error (see ffb_last_error) */
  FFB_ERR_NOGPU = -3       /* no sm_100 device                        */
} ffb_status;

/* device-side status word bits (written with atomicOr by the kernels) */
#define FFB_ST_NONFINITE_STATE 1   /* torchdiffeq: "non-finite values in state `y`"  */
#define FFB_ST_NAN_SAMPLE 2        /* diffusion.py:560-562 "Diffusion is not stable" */

/* field kinds */
#define FFB_FIELD_NET 0    /* x_dot = sign * net(x)                 flow.py:118, symplectic.py:120-123 */
#define FFB_FIELD_SCORE 1  /* x_dot = sign * (a*x - c*score)        diffusion.py:276-279 / :553         */

/* divergence modes */
#define FFB_DIV_NONE 0
#define FFB_DIV_EXACT 1      /* D forward-mode tangents  (flow.py:157-161, diffusion.py:483-503) */
#define FFB_DIV_HUTCH 2      /* one tangent along a fixed Rademacher probe (diffusion.py:327-334) */

/* hidden-layer activations (ffb_net_desc.activation).  SiLU is the reference's default (diffusion.py:38,
 * flow.py:41, symplectic.py:25); the others are what a user may pass as `activation=`.  Non-SiLU activations run
 * on the chunk-pipelined tensor-core engines and on the wide engine; the debug engines (ffb_set_engine(0) / (2)) and
 * samples wider than a tile return FFB_ERR_ARG for them. */
#define FFB_ACT_SILU 0
#define FFB_ACT_TANH 1
#define FFB_ACT_RELU 2
#define FFB_ACT_SOFTPLUS 3   /* torch.nn.Softplus(beta=1, threshold=20) */
#define FFB_ACT_GELU 4       /* torch.nn.GELU(approximate='none')       */

/* fixed-grid methods */
#define FFB_M_EULER 0        /* torchdiffeq 'euler'; also symplectic.py:192-197 */
#define FFB_M_MIDPOINT 1     /* torchdiffeq 'midpoint'                          */
#define FFB_M_RK4 2          /* torchdiffeq 'rk4' (3/8 rule)                    */
#define FFB_M_EM 3           /* Euler-Maruyama, diffusion.py:543-562            */
#define FFB_M_LEAPFROG 4     /* kick-drift-kick (extension; no reference oracle) */

typedef struct ffb_net ffb_net; /* opaque: packed weights of one MLP */

/* One MLP exactly as torch.nn.Linear stores it (diffusion.py:67-72, flow.py:63-74,
 * symplectic.py:71-78) plus the column map of its layer-0 input:
 *   diffusion.MLP      [ sin|cos (t_dim) | x (x_dim) | cond (c_dim) ]     diffusion.py:101-113
 *   flow.*ODEFlow      [ x | t (t_dim = 1) | cond ]                       flow.py:112-115, 583-586
 *   symplectic         [ p or q | cond | sin|cos ]                        symplectic.py:109-114   */
typedef struct {
  int32_t n_layers;
  int32_t in_features;
  int32_t widths[FFB_MAX_LAYERS];       /* out_features of each Linear                    */
  const float* weight[FFB_MAX_LAYERS];  /* device, [out][in] row-major                    */
  const float* bias[FFB_MAX_LAYERS];    /* device, [out]                                  */
  int32_t x_col, x_dim;
  int32_t c_col, c_dim;
  int32_t t_col, t_dim;
  int32_t activation;                   /* FFB_ACT_*                                      */
} ffb_net_desc;

/* A vector field built from one or two networks acting on column blocks of the state. */
typedef struct {
  int32_t n_calls;              /* 1, or 2 for symplectic (mlp_q then mlp_p)           */
  const ffb_net* net[2];
  int32_t in_off[2];            /* first state column fed to net[i]                    */
  int32_t out_off[2];           /* first state column whose derivative net[i] produces */
  float out_sign[2];            /* +1 / -1 (symplectic.py:121)                         */
  int32_t state_dim;            /* columns of the state tensor                         */
  int32_t cond_dim;
  int32_t kind;                 /* FFB_FIELD_*                                         */
  int32_t use_sigma;            /* score = net / sigma(t)  (diffusion.py:236)          */
  int32_t has_drift;            /* 0 for VE (diffusion.py:905)                         */
  int32_t div_mode;             /* FFB_DIV_*                                           */
} ffb_field;

/* Host-computed scalars of ONE function evaluation. */
typedef struct {
  float tfeat[FFB_MAX_TFEAT];   /* sin|cos(((t*W)*2)*pi) or raw t                              */
  float a;                      /* drift coefficient, e.g. fl(-0.5*beta(t))                    */
  float c;                      /* score coefficient: fl(0.5*g^2) (PF-ODE) or fl(g^2) (SDE)    */
  float sigma;                  /* sigma(t) when use_sigma                                      */
  float sign;                   /* -1 when torchdiffeq integrates reversed time, else +1       */
} ffb_eval_scalars;

/* ---- single evaluation (+ the norms of torchdiffeq's initial-step heuristic) ---------- */
typedef struct {
  int64_t batch;
  const float* y;               /* (B, state_dim)                                              */
  const float* fbase;           /* optional: evaluate at y + h*fbase                           */
  const float* dlpbase;         /* optional: d(logp)/dt that goes with fbase                   */
  float h;
  const float* cond;            /* (B, cond_dim) network input (already normalised) or NULL    */
  const float* cond_state;      /* (B, cond_dim) raw conditional as carried in the ODE state   */
  const float* probes;          /* (B, state_dim) Rademacher, FFB_DIV_HUTCH only               */
  float* f;                     /* out (B, state_dim) or NULL                                  */
  float* dlp;                   /* out (B,) divergence, or NULL                                */
  ffb_eval_scalars ev;
  float atol, rtol;
  int32_t norms;                /* 0 none; 1: sum (y/sc)^2, (f/sc)^2; 2: sum ((f-fbase)/sc)^2  */
  int32_t cond_in_state;        /* ConditionalODEFlow carries cond in the ODE state (flow.py:857-861) */
  double* partials;             /* (n_tiles, FFB_NPART)                                        */
  int32_t* status;
  void* scratch;
  float* jac;                   /* optional out (B, x_dim, x_dim), FFB_DIV_EXACT on the tensor-core engine only:
                                   jac[b][j][n] = d net_n / d x_j of the NETWORK output (before a*x - c*score);
                                   row j is what a reverse sweep seeded with e_n collects for x_j, i.e. the
                                   matrix a VJP multiplies by (diffusion.py:361-374), consumed by
                                   ffb_trace_estimate                                              */
} ffb_eval_args;

/* ---- one attempted dopri5 step (6 fused evaluations + error norm + dense output) ------- */
typedef struct {
  int64_t batch;
  const float* y0;   const float* f0;       /* state and FSAL derivative at t0                 */
  const float* lp0;  const float* dlp0;     /* log-det column and its derivative (or NULL)     */
  const float* cond; const float* probes;
  float* y1;  float* f1;  float* lp1;  float* dlp1;
  float* y_out;  float* lp_out;             /* interpolant at t_end, written when final != 0   */
  ffb_eval_scalars ev[6];                   /* stages 2..7                                     */
  float cb[6][6];                           /* fl32(beta_ij * dt)                              */
  float ce[7];                              /* fl32(dt * c_err_j)                              */
  float cm[7];                              /* fl32(dt * c_mid_j)                              */
  float dt, atol, rtol, x_interp;
  int32_t final;
  double* partials;
  int32_t* status;
  void* scratch;
  const struct ffb_dopri5_ctl* ctl;         /* NULL: the step is described by the fields above (host-driven loop);
                                               else: device-resident step + controller state, see below            */
} ffb_dopri5_args;

/* ---- device-side dopri5 controller ------------------------------------------------------------
 * torchdiffeq's adaptive loop (accept test, step-size controller, step_t clipping, stage times) and the
 * per-evaluation scalars of the next attempt, evaluated by a one-block kernel between two attempts, so that
 * the host never sits between two launches: it enqueues attempt -> reduce [-> all-reduce] -> control a few
 * times ahead and only polls `done`.  Attempt kernels launched after the solve has finished return at once.
 * The same code compiles for the host (ffb_dopri5_control_host) so that the controller is testable without a
 * GPU.  Semantics: SURVEY.md section 8c T3-T12, identical to flowfusion_b200/solver.py::dopri5 (host loop). */
#define FFB_PROG_RAW_T 0      /* tfeat[0] = t                                  flow.py:112-115             */
#define FFB_PROG_FOURIER 1    /* tfeat = sin|cos(((t*W)*2)*pi)                 diffusion.py:109-110,
                                                                               symplectic.py:103           */
#define FFB_SDE_NONE 0        /* a = c = 0, sigma = 1 (flows, symplectic)                                  */
#define FFB_SDE_VP 1          /* diffusion.py:1047-1156                                                    */
#define FFB_SDE_VE 2          /* diffusion.py:852-905                                                      */
#define FFB_SDE_SUBVP 3       /* diffusion.py:1223-1342                                                    */
#define FFB_MAX_FREQ (FFB_MAX_TFEAT / 2)

/* How ffb_eval_scalars derive from the (user) time of an evaluation.  Every constant is the FP32 value the
 * reference's eager ops would use (Python scalars rounded to FP32 where they meet an FP32 tensor). */
typedef struct {
  int32_t time_features;        /* FFB_PROG_*                                                  */
  int32_t n_freq;               /* FOURIER: embedding_dimensions / 2                           */
  float W[FFB_MAX_FREQ];        /* FOURIER: frequencies                                         */
  float pi;                     /* FOURIER: fl32(pi)                                            */
  int32_t sde;                  /* FFB_SDE_*                                                    */
  int32_t use_sigma;            /* fill ev.sigma with sigma(t) (else 1)                         */
  int32_t sde_mode;             /* c = g^2 (reverse SDE) instead of 0.5*g^2 (PF-ODE)            */
  float T;
  float beta_min, beta_diff;    /* VP / subVP: fl32(beta_min), fl32(beta_max - beta_min)        */
  float half_beta_diff;         /* fl32(0.5*(beta_max - beta_min))                              */
  float m2_beta_min;            /* fl32(-2*beta_min)                                            */
  float sigma_min, sigma_ratio; /* VE: sigma_min, fl32(sigma_max / sigma_min)                   */
  float ve_gfac;                /* VE: fl32(sqrt(2*(log sigma_max - log sigma_min)/T))          */
} ffb_time_program;

#define FFB_CTL_MAX_GRID 16   /* options['step_t'] points                                       */
#define FFB_CTL_HIST 256      /* attempts whose (dt, ratio, accept) are kept for SolveStats      */
#define FFB_CTL_NOTIFY_SLOTS 1024
/* ffb_dopri5_ctl.done */
#define FFB_CTL_RUNNING 0
#define FFB_CTL_FINISHED 1
#define FFB_CTL_NONFINITE (-1)      /* torchdiffeq: "non-finite values in state `y`"   */
#define FFB_CTL_DT_UNDERFLOW (-2)   /* torchdiffeq: "underflow in dt"                   */
#define FFB_CTL_MAX_STEPS (-3)      /* torchdiffeq: "max_num_steps exceeded"            */

/* constants of one solve (passed by value to the control kernel) */
typedef struct {
  double t_end;                 /* solver time (negated for descending spans, T2)               */
  double min_step, max_step, safety, ifactor, dfactor;
  int64_t n_x, n_lp, n_cond;    /* GLOBAL element counts of the RMS norms (0 = component absent) */
  int32_t reverse;              /* user time = -solver time, ev.sign = -1                        */
  int32_t max_num_steps;
  int32_t n_grid;
  int32_t _pad;
  double grid[FFB_CTL_MAX_GRID];/* ascending solver-time step_t points >= t0                     */
  float alpha[6];               /* Dormand-Prince tableau rounded to FP32 as torchdiffeq does    */
  float beta[6][6];
  float c_err[7];
  float c_mid[7];
  ffb_time_program prog;
} ffb_dopri5_ctl_params;

/* device-resident block: the next attempt's step + the controller's state */
typedef struct ffb_dopri5_ctl {
  /* read by ffb_dopri5_attempt (same meaning as the same-named fields of ffb_dopri5_args) */
  ffb_eval_scalars ev[6];
  float cb[6][6];
  float ce[7];
  float cm[7];
  float dt, x_interp;
  int32_t final;
  int32_t cur;                  /* 0: (y0,f0,lp0,dlp0) hold the current state and (y1,..) receive the
                                   candidate; 1: the roles are swapped (flips on every accepted step)   */
  int32_t done;                 /* FFB_CTL_*                                                             */
  int32_t grid_idx;
  /* controller state */
  double t, dt_next;            /* initialise: t = t0, dt_next = first step                              */
  double cur_t1, cur_dt;        /* the attempt in flight                                                 */
  int32_t cur_on_grid;
  int32_t n_attempts, n_accepted, n_rejected;
  int32_t n_turns;              /* controller turns after an attempt, including those after the solve ended  */
  int32_t _pad;
  int32_t* notify;              /* optional: FFB_CTL_NOTIFY_SLOTS int32 of PINNED HOST memory (device-accessible);
                                   turn k stores (k << 8) | (done & 0xff) into slot (k - 1) % slots, so the host
                                   can follow the solve without any stream operation                          */
  double hist_dt[FFB_CTL_HIST];
  float hist_ratio[FFB_CTL_HIST];
  uint8_t hist_accept[FFB_CTL_HIST];
} ffb_dopri5_ctl;

/* ---- fixed-grid integrators, whole trajectory on-chip ---------------------------------- */
#define FFB_STEP_STRIDE 8   /* floats per step in step_table: dt, g, sqrt(-dt), then method constants */
typedef struct {
  int64_t batch;
  int32_t method;
  int32_t nsteps;
  const float* x0;               /* (B, state_dim)                                             */
  const float* lp0;              /* (B,) or NULL                                               */
  const float* cond;
  const float* probes;
  const float* noise;            /* EM comparison mode: (nsteps, B, state_dim) unit normals     */
  uint64_t philox_seed;          /* EM throughput mode (noise == NULL)                          */
  uint64_t philox_offset;
  int64_t row_offset;            /* global index of row 0 (multi-GPU shards share one stream)   */
  float* x_out;                  /* (B, state_dim); EM: x_mean of the last step (diffusion.py:563) */
  float* lp_out;
  const float* step_table;       /* device (nsteps, FFB_STEP_STRIDE)                            */
  const ffb_eval_scalars* ev_table; /* device (nsteps, evals_per_step)                          */
  int32_t* status;               /* device, TWO words: [0] FFB_ST_* bits (atomicOr); [1] index of the first
                                    Euler-Maruyama step after which any element of x was NaN (atomicMin; the
                                    caller initialises it to INT32_MAX) -- diffusion.py:560-563 stops there      */
  void* scratch;
} ffb_fixed_args;

int ffb_abi_version(void);
const char* ffb_last_error(void);
/* sm count, opt-in shared memory per block, compute capability, SM clock (kHz) */
int ffb_device_info(int32_t* sm_count, int32_t* smem_optin, int32_t* cc_major, int32_t* cc_minor,
                    int32_t* clock_khz);

int ffb_net_create(const ffb_net_desc* desc, void* stream, ffb_net** out);
void ffb_net_destroy(ffb_net* net);
/* algorithmic FLOPs (2*MAC over the Linear layers) of one forward pass of one row */
int64_t ffb_net_flops(const ffb_net* net);

int64_t ffb_num_tiles(const ffb_field* field, int64_t batch);
size_t ffb_scratch_bytes(const ffb_field* field);

int ffb_field_eval(const ffb_field* field, const ffb_eval_args* args, void* stream);
int ffb_dopri5_attempt(const ffb_field* field, const ffb_dopri5_args* args, void* stream);
int ffb_integrate_fixed(const ffb_field* field, const ffb_fixed_args* args, void* stream);

/* 1 when ffb_dopri5_attempt accepts args->ctl for this field (the chunk-pipelined tensor-core engines) */
int ffb_dopri5_ctl_supported(const ffb_field* field);
/* One controller turn on the device (one block, enqueued on `stream`).  after_attempt = 0: prepare the first
 * attempt from ctl->t / ctl->dt_next; 1: judge the attempt whose FFB_NPART sums are in `sums` (device, FP64,
 * already reduced over tiles and ranks), update the state, prepare the next attempt or set ctl->done.
 * Single-GPU shortcut: with partials != NULL the same launch first reduces partials (n_tiles, FFB_NPART) into
 * sums, like ffb_reduce_partials. */
int ffb_dopri5_control(const ffb_dopri5_ctl_params* params, double* sums, const double* partials, int64_t n_tiles,
                       ffb_dopri5_ctl* ctl, int32_t after_attempt, void* stream);
/* the same turn on the CPU (host pointers): test twin of the kernel above, no CUDA call */
int ffb_dopri5_control_host(const ffb_dopri5_ctl_params* params, const double* sums, ffb_dopri5_ctl* ctl,
                            int32_t after_attempt);
/* rows of ffb_eval_scalars for n user times (host arrays); on_device != 0 evaluates them in a kernel on the
 * current device (synchronous, for tests), else on the CPU twin */
int ffb_time_program_rows(const ffb_time_program* prog, const float* times, int32_t n, float sign,
                          ffb_eval_scalars* out, int32_t on_device);

/* ---- staged solves: Hutch++ / XTrace divergence estimators (diffusion.py:336-481) ---------------
 * Both estimators need a per-sample thin QR BETWEEN two rounds of Jacobian products, so one evaluation is
 * staged over two launches instead of living inside the fused attempt kernel:
 *   ffb_field_eval(div_mode = FFB_DIV_EXACT, jac = J)   field + the full network Jacobian (D tangent rows)
 *   ffb_trace_estimate(J, probes)                        the reference's estimator algebra, one thread per sample
 * and a dopri5 attempt is 6 x (ffb_rk_combine, ffb_field_eval, ffb_trace_estimate) + ffb_rk_finish, which leaves
 * the same FFB_NPART partial sums the fused attempt kernel does (the host controller does not change).      */
#define FFB_TRACE_HUTCHPP 1   /* diffusion.py:336-400 */
#define FFB_TRACE_XTRACE 2    /* diffusion.py:402-481 */
#define FFB_TRACE_MAX_DIM 124 /* state columns D (above 32: thread-per-sample kernel)  */
#define FFB_TRACE_MAX_RANK 8  /* Hutch++ r = min(hpp_rank, D); XTrace m = xt_vecs  */
#define FFB_STAGED_BLOCKS 1024 /* rows of the `partials` buffer the staged kernels write (caller zero-fills once) */
typedef struct {
  int64_t batch;
  int32_t dim;                  /* D                                                             */
  int32_t kind;                 /* FFB_TRACE_*                                                   */
  int32_t rank;                 /* Hutch++: rows of S (r);  XTrace: rows of O (m <= D)            */
  int32_t nvec;                 /* Hutch++: rows of G (m >= 1); XTrace: unused                    */
  const float* jac;             /* (B, D, D) from ffb_field_eval                                  */
  const float* S;               /* (rank, B, D): Hutch++ S / XTrace O, reference layout (diffusion.py:710, 721) */
  const float* G;               /* (nvec, B, D): Hutch++ G (diffusion.py:711)                     */
  int32_t score;                /* field kind FFB_FIELD_SCORE: J_f = a I - c J_net [/ sigma]      */
  int32_t use_sigma;
  int32_t has_drift;
  float a, c, sigma, sign;      /* the evaluation's ffb_eval_scalars                              */
  float* dlp;                   /* out (B,): sign * estimated divergence                          */
  /* optional norms of torchdiffeq's initial-step heuristic on the log-det column (as ffb_eval_args.norms) */
  int32_t norms;                /* 0 none; 1: sum (dlp/atol)^2 -> P_LP_F; 2: sum ((dlp - dlpbase)/atol)^2 -> P_LP_DF */
  float atol;
  const float* dlpbase;
  double* partials;             /* (FFB_STAGED_BLOCKS, FFB_NPART), required when norms != 0        */
} ffb_trace_args;
int ffb_trace_estimate(const ffb_trace_args* args, void* stream);
/* the same algebra on the CPU (host pointers, no CUDA call; norms ignored): test twin of the kernel */
int ffb_trace_estimate_host(const ffb_trace_args* args);

/* out = y0 + sum_{j < n_terms} coef[j] * k[j] over n flat elements: the stage input of torchdiffeq's
 * `y0 + k[..., :i+1] @ (beta_i * dt)` */
typedef struct {
  int64_t n;
  int32_t n_terms;
  const float* y0;
  const float* k[7];
  float coef[7];
  float* out;
} ffb_rk_combine_args;
int ffb_rk_combine(const ffb_rk_combine_args* args, void* stream);

/* end of one staged attempt of an adaptive Runge-Kutta method (dopri5: 7 stage derivatives; bosh3: 4; adaptive_heun: 2;
 * fehlberg2: 3): log-det column of y1, error partial sums (P_X_ERR, P_LP_ERR, P_NONFINITE), and the dense output at
 * t_end when `final` (same statements as the fused attempt kernel).  f1 is the LAST stage derivative, k[n_k - 1]
 * (torchdiffeq rk_common._runge_kutta_step, also for tableaus that are not FSAL). */
typedef struct {
  int64_t batch;
  int32_t dim;
  int32_t final;
  int32_t n_k;                            /* stage derivatives in k / dlp: 2..7 (0 = 7)     */
  int32_t _pad;
  const float* y0; const float* y1;       /* (B, D): y1 = y0 + k @ (dt c_sol)               */
  const float* k[7];                      /* k1..k_{n_k}, (B, D) each; the rest is ignored  */
  const float* lp0; const float* dlp[7];  /* log-det column and its 7 stage derivatives, or NULL */
  float* lp1;
  float cl[6];                            /* fl32(c_sol_j * dt), j < min(n_k, 6)            */
  float ce[7], cm[7];
  float dt, atol, rtol, x_interp;
  float* y_out; float* lp_out;
  double* partials;                       /* (FFB_STAGED_BLOCKS, FFB_NPART)                 */
} ffb_rk_finish_args;
int ffb_rk_finish(const ffb_rk_finish_args* args, void* stream);

/* ---- fused training step (diffusion.py:1369-1463, flow.py:191-256, :679-747) ----------------
 * Every loss of the reference is  loss = scale * sum_{b,d} ( alpha_b * net(X)_{b,d} + beta_{b,d} )^2  for per-row scalars
 * alpha and per-element offsets beta that do not depend on the weights (denoising score matching: alpha = 1 or sigma_b,
 * beta = z; likelihood weighting: alpha = g_b / sigma_b or g_b, beta = (g_b / sigma_b) z; flow matching: alpha = 1,
 * beta = x0 - xT).  ffb_train_step evaluates the network on X (B, in_features: the layer-0 input rows in the
 * network's own column order, time features included), the loss and its gradient with respect to every weight and
 * bias (torch.nn.Linear layout, written -- not accumulated) and optionally X, in four launches; deterministic.
 * `net` is read as raw torch weights (weight / bias / widths / in_features / activation; the column map is unused):
 * the weights change every optimiser step, so nothing is cached between calls. */
typedef struct {
  int64_t batch;
  const float* x_in;                  /* (B, in_features)                                   */
  const float* alpha;                 /* (B,) or NULL (= 1)                                 */
  const float* beta;                  /* (B, out_features)                                  */
  float scale;
  int32_t _pad;
  float* grad_w[FFB_MAX_LAYERS];      /* out: (out, in) row-major                           */
  float* grad_b[FFB_MAX_LAYERS];      /* out: (out,)                                        */
  float* grad_x;                      /* out: (B, in_features) or NULL                      */
  double* loss;                       /* out: one double (device)                           */
  float* work;                        /* ffb_train_work_bytes(net, batch, grad_x != NULL)   */
  /* vector-Jacobian mode (the adjoint ODE of SURVEY 8f rank 3, diffusion.py:620-629, flow.py:286-295): with cot != NULL
   * the "loss" is scale * sum(cot * net(X)), i.e. grad_w / grad_b / grad_x are scale * cot^T d net / d (W, b, X);
   * alpha and beta are ignored. */
  const float* cot;                   /* (B, out_features) or NULL                          */
  float* out;                         /* out: (B, out_features) network output, or NULL     */
} ffb_train_args;
/* Forward only: with `out` set and loss, grad_x, grad_w[0], grad_b[0] all NULL the call just evaluates out = net(X) at
 * per-row inputs (two launches; e.g. the reference's score(t, x) with one time per sample, diffusion.py:82-121); any width
 * up to FFB_MAX_WIDTH fits.  Size `work` with want_grad_x = 2 for that mode. */
size_t ffb_train_work_bytes(const ffb_net_desc* net, int64_t batch, int32_t want_grad_x);
int ffb_train_step(const ffb_net_desc* net, const ffb_train_args* args, void* stream);

/* ---- Hamiltonian leapfrog (BASELINE.json north_star: "the leapfrog integrator's dH/dq, dH/dp is a fused forward+backward
 * MLP kernel"; configs[4]).  H(q, p [, cond]) is a scalar-output MLP on the input rows [q | p | cond] (raw torch weights as
 * in ffb_train_step); one launch carries every trajectory through n_steps kick-drift-kick steps
 *     p -= dt/2 dH/dq(q, p);  q += dt dH/dp(q, p);  p -= dt/2 dH/dq(q, p)
 * the gradient being a forward pass that keeps the pre-activations in shared memory and the backward sweep to the inputs,
 * entirely on-chip.  Extension: the reference's own symplectic networks output dq/dt and dp/dt directly
 * (symplectic.py:80-123) and run on the two-network field of ffb_integrate_fixed (FFB_M_EULER / FFB_M_LEAPFROG). */
typedef struct {
  int64_t batch;
  int32_t dim;                        /* D: q and p have D columns each                       */
  int32_t cond_dim;
  const float* z0;                    /* (B, 2D) [q | p]                                      */
  const float* cond;                  /* (B, cond_dim) or NULL                                */
  float* z_out;                       /* out: (B, 2D)                                         */
  float* h_out;                       /* out: (B, 2) H at the start and at the end, or NULL   */
  int32_t n_steps;
  float dt;
  float* work;                        /* ffb_train_work_bytes(net, 0, 1)                      */
} ffb_hamiltonian_args;
int ffb_hamiltonian_leapfrog(const ffb_net_desc* net, const ffb_hamiltonian_args* args, void* stream);

/* sums[FFB_NPART] = sum over tiles of partials, in tile order (deterministic) */
int ffb_reduce_partials(const double* partials, int64_t n_tiles, double* sums, void* stream);
/* out[b] = sum_d ( -0.5*x^2 - 0.5*log(2*pi*var) ) + (add ? add[b] : 0)   (flow.py:434, diffusion.py:814) */
int ffb_gaussian_logprob(const float* x, const float* add, float* out, int64_t batch, int32_t dim,
                         float sigma, void* stream);
/* the normals the EM kernel would draw for (seed, offset, step): for statistical tests */
int ffb_philox_normal(float* out, int64_t batch, int32_t dim, uint64_t seed, uint64_t offset,
                      int32_t step, int64_t row_offset, void* stream);
/* FP32 FFMA2 peak probe: returns achieved TFLOP/s through *tflops (roofline denominator) */
int ffb_ffma_peak(int32_t iters, float* tflops, void* stream);
/* contraction engine: 1 = tcgen05 tensor cores, 3xTF32 (default: the dual-tile engine takes the dopri5 attempts of
 * fields without a divergence when the batch holds at least two tiles per SM, the single-tile engine everything else);
 * 3 = single-tile engine only; 4 = dual-tile engine for every dopri5 attempt it can hold (tests); 2 = the older
 * whole-layer hand-off tile engine; 0 = FP32 FFMA2 (debug / A-B checks); 5 = the wide engine (FP32 pipe, 32-row passes)
 * for every field -- by default it only takes networks with a hidden width above 128 or more than 8 Linear layers, whatever
 * this setting.  Engines 1, 3 and 4 give the same bits.
 * The environment variable FFB_ENGINE = ffma | tc_tile | rr | rd | wide selects 0 | 2 | 3 | 4 | 5 at load time. */
int ffb_set_engine(int engine);
int ffb_get_engine(void);
/* debug: hand-off timelines of CTA 0 (libraries built with -DFFB_TRACE only; scripts/trace_rr.py, scripts/trace_rd.py).
 * buf: device buffer of 3 (single-tile engines) / 5 (dual-tile engine) roles x 2048 x 2 int64, or NULL to stop. */
int ffb_debug_trace(long long* buf);
int ffb_debug_trace_rd(long long* buf);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t ffb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FFB200_H */
