/*
 * flowstate_b200 - C ABI of the B200-native NF-MCMC sampling hot path.
 *
 * The reference (Inesalmansa/flow-state) is pure Python and has no FFI of its
 * own: its boundary is the Python class API that the experiment
 * drivers call.  Each entry point below is what a ctypes/cffi binding placed
 * under that class API binds; the comment on each names the reference
 * function(s) it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the comment
 *     says "host"; no allocation happens inside the hot calls;
 *   - `stream` is the caller's cudaStream_t passed as void* (NULL = default);
 *   - return value 0 = ok, non-zero = error (fs_last_error() has the text);
 *   - positions are float32 [B, N, 2] in MC-box coordinates [0, L];
 *   - running per-chain energies/virials, max displacement and counters keep
 *     the reference's float64 / int64 types.
 */
#ifndef FLOWSTATE_B200_H
#define FLOWSTATE_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FS_OK 0
#define FS_ERR_INVALID 1
#define FS_ERR_CUDA 2
#define FS_ERR_UNSUPPORTED 3

/* External potential + pair cut-offs.  Reference: MCMC/energy_calculator.py:10-21
 * (num_wells, V0_list, r0, k), MCMC/potential.py:3 (cutoff 2.5, shifted),
 * MCMC/energy_calculator.py:73,150 (hard core r < 0.5 -> inf). */
typedef struct fs_pot {
    int    num_wells;   /* 0, 1 or 2 */
    double V0[2];       /* the reference keeps these as Python floats (float64): r0 = 1.2 is not a float32 value, */
    double r0;          /* and the steep wall (k = 15) turns its float32 rounding into 4e-6 of energy            */
    double k;
    double r_cut;       /* 2.5 */
    double r_core;      /* 0.5 */
} fs_pot;

/* Random source of the Metropolis kernels.
 *   FS_RNG_PCG64  : bit-exact emulation of numpy.random.Generator(PCG64) as the
 *                   reference uses it (MCMC/monte_carlo.py:92-95,153,161,215,287):
 *                   integers(N) = buffered 32-bit Lemire, random() = 53-bit double.
 *                   state: uint64 [B, 6] = {state_hi, state_lo, inc_hi, inc_lo,
 *                   has_uint32, uinteger}, updated in place.
 *   FS_RNG_PHILOX : Philox4x32-10 keyed by (seed, global chain id, per-chain step
 *                   counter = attempts), so results do not depend on how chains
 *                   are sharded over GPUs or split over launches.
 *   FS_RNG_REPLAY : recorded draws; idx [B, idx_stride] consumed one per local
 *                   step, u [B, u_stride] consumed through cursor [B, 2] =
 *                   {idx cursor, u cursor} (third uniform only on finite uphill
 *                   moves, like the reference). */
#define FS_RNG_PCG64 0
#define FS_RNG_PHILOX 1
#define FS_RNG_REPLAY 2
#define FS_RNG_PHILOX_REF 3   /* the FS_RNG_PHILOX draws through the reference-order parity kernel (tests) */
typedef struct fs_rng {
    int                 kind;
    unsigned long long* pcg_state;
    unsigned long long  philox_seed;
    long long           chain_id0;       /* global id of chain 0 of this call */
    const int*          replay_idx;
    const double*       replay_u;
    int                 idx_stride;
    int                 u_stride;
    int*                replay_cursor;
} fs_rng;

const char* fs_last_error(void);
int fs_version(void);
/* Binds the library's CUDA runtime to `device` (one process per GPU: the device that owns the caller's buffers). */
int fs_set_device(int device);
/* kernels launched by this library in this process so far (bench bookkeeping) */
unsigned long long fs_launch_count(void);

/* Element-wise helpers behind the reference's small public functions:
 *   SimulationBox.apply_pbc (MCMC/simulation_box.py:19-29): pos [n,2] wrapped in place (numpy floor-mod);
 *   SimulationBox.compute_distances (:31-65): r[i] = |min_image(p1 - p2[i])|, p1 one point or [n,2];
 *   lennard_jones_energy_virial (MCMC/potential.py:3-29): shifted, cut pair energy / virial of r[i];
 *   double_well_potential (MCMC/potential.py:55-116): V_ext of each position. */
int fs_apply_pbc(float* pos, long long n, float Lx, float Ly, void* stream);
int fs_distances(const float* p1, int p1_is_single, const float* p2, long long n, float Lx, float Ly,
                 float* r, void* stream);
int fs_lj_pair(const float* r, long long n, const fs_pot* pot /*host*/, float* e, float* w, void* stream);
int fs_double_well(const float* pos, long long n, float Lx, float Ly, const fs_pot* pot /*host*/, float* v,
                   void* stream);

/* EnergyCalculator.calculate_total_energy_virial  (MCMC/energy_calculator.py:121-203)
 * for B configurations at once: E[b] = sum_{i<j} LJ + sum_i V_ext, W[b] = sum virial,
 * overlap[b] = 1 and E = W = +inf when any pair is closer than r_core. */
int fs_energy_total(const float* pos, int B, int N, float Lx, float Ly, const fs_pot* pot /*host*/,
                    float* E, float* W, unsigned char* overlap, void* stream);

/* EnergyCalculator.calculate_particle_energy_virial  (MCMC/energy_calculator.py:48-108)
 * for particle idx[b] of each configuration, at its stored position (new_xy == NULL)
 * or moved to new_xy[b] = (x, y). */
int fs_energy_particle(const float* pos, const int* idx, const float* new_xy, int B, int N,
                       float Lx, float Ly, const fs_pot* pot /*host*/,
                       float* e, float* w, unsigned char* overlap, void* stream);

/* `steps` calls of MonteCarlo.particle_displacement  (MCMC/monte_carlo.py:146-223)
 * on each of B independent chains; chain state stays in shared memory for the
 * whole launch.  Optional traces (NULL to skip): trace_accept/trace_idx [B, steps],
 * trace_e [B, steps, 2] = (e_old, e_new) of the moved particle. */
int fs_local_sweep(float* pos, double* E, double* W, const double* max_disp,
                   long long* attempts, long long* accepted,
                   int B, int N, int steps, float Lx, float Ly, double beta,
                   const fs_pot* pot /*host*/, const fs_rng* rng /*host*/,
                   unsigned char* trace_accept, int* trace_idx, float* trace_e, void* stream);

/* MonteCarlo.adjust_displacement  (MCMC/monte_carlo.py:375-403), per chain. */
int fs_adjust_displacement(double* max_disp, const long long* attempts, const long long* accepted,
                           long long* prev_attempts, long long* prev_accepted,
                           double target, int B, void* stream);

/* Steps 3-5 of MonteCarlo.nf_big_move  (MCMC/monte_carlo.py:264-303):
 * ratio = exp(-beta (E_new - E) - (nll_new - nll_old)), accept if >= 1 or u < ratio
 * (uniform drawn only when ratio < 1), masked in-place copy prop -> pos, E/W and
 * counter updates.  u == NULL draws from rng. */
int fs_accept_global(float* pos, const float* prop, double* E, double* W,
                     const float* E_new, const float* W_new,
                     const float* logq_old, const float* logq_new,
                     const double* u, const fs_rng* rng /*host, may be NULL if u*/,
                     double beta, long long* attempts, long long* accepted,
                     unsigned char* accept_mask, int B, int N, void* stream);

/* The whole tail of MonteCarlo.nf_big_move in ONE kernel (MCMC/monte_carlo.py:243-303): the energy and virial of the
 * proposal (calculate_total_energy_virial, MCMC/energy_calculator.py:121-203, what fs_energy_total computes), the
 * acceptance rule of fs_accept_global on it and the masked in-place update - the proposal is staged in shared memory
 * once for its O(N^2) pair walk and an accepted one is written to `pos` from there.  E_new / W_new [B] (required)
 * receive the proposals' energies and virials (inf on hard-core overlap -> the move is rejected).  `prop` must not
 * alias `pos`.  Tiles beyond the packed kernel's shared-memory budget run as fs_energy_total + fs_accept_global. */
int fs_accept_global_fused(float* pos, const float* prop, double* E, double* W,
                           float* E_new, float* W_new,
                           const float* logq_old, const float* logq_new,
                           const double* u, const fs_rng* rng /*host, may be NULL if u*/,
                           double beta, long long* attempts, long long* accepted,
                           unsigned char* accept_mask, int B, int N,
                           float Lx, float Ly, const fs_pot* pot /*host*/, void* stream);

/* ---- flow (NF/normflows) -------------------------------------------------- */

/* Parameters of one CircularCoupledRationalQuadraticSpline layer in the
 * reference's state_dict layout, HOST pointers, float32:
 * prefix flows.<i>.prqct.  */
typedef struct fs_layer_params {
    const float* init_w;    /* transform_net.initial_layer.weight  [H, 2N] */
    const float* init_b;    /* transform_net.initial_layer.bias    [H] */
    const float* bn_w;      /* blocks.<b>.batch_norm_layers.<j>.weight        [n_blocks, 2, H] */
    const float* bn_b;      /* ... .bias */
    const float* bn_mean;   /* ... .running_mean */
    const float* bn_var;    /* ... .running_var */
    const float* lin_w;     /* blocks.<b>.linear_layers.<j>.weight  [n_blocks, 2, H, H] */
    const float* lin_b;     /* blocks.<b>.linear_layers.<j>.bias    [n_blocks, 2, H] */
    const float* final_w;   /* transform_net.final_layer.weight  [N (3 nb + 1), H] */
    const float* final_b;   /* transform_net.final_layer.bias    [N (3 nb + 1)] */
    const float* un_w;      /* unconditional_transform.unnormalized_widths      [N, nb] */
    const float* un_h;      /* unconditional_transform.unnormalized_heights     [N, nb] */
    const float* un_d;      /* unconditional_transform.unnormalized_derivatives [N, nb + 1] */
} fs_layer_params;

#define FS_PREC_FP32 0   /* CUDA-core FP32 conditioner (reference arithmetic) */
#define FS_PREC_TF32 1   /* tcgen05 kind::tf32 conditioner, FP32 accumulate in TMEM */

typedef struct fs_flow_desc {
    int K;            /* number of coupling layers */
    int N;            /* particles; D = 2 N features, N identity + N transformed */
    int H;            /* hidden width of the residual conditioner */
    int n_blocks;     /* residual blocks */
    int nb;           /* spline bins (<= 32) */
    double bound;     /* tail_bound = half box (the reference keeps it as a Python float) */
    float bn_eps;     /* 1e-3 (nets/resnet.py:25) */
    const int* identity_features;    /* host [N]  (prqct.identity_features) */
    const int* transform_features;   /* host [N]  (prqct.transform_features) */
    const fs_layer_params* layers;   /* host [K] */
} fs_flow_desc;

typedef struct fs_flow fs_flow;

/* Packs a flow for inference: folds eval-mode BatchNorm, precomputes the
 * unconditional spline knots, uploads weights (and, for the tensor path, lays
 * them out as TMA-ready swizzled tiles).  Replaces nothing in the reference -
 * it is the analogue of model.eval().to(device)
 * (hybrid_NF_MCMC/main_algorithm_1.py:284,331). */
int fs_flow_create(const fs_flow_desc* desc /*host*/, fs_flow** out);
void fs_flow_destroy(fs_flow* flow);
size_t fs_flow_workspace_bytes(const fs_flow* flow, int B, int precision);

/* Re-packs `flow` in place from parameters that live on the DEVICE: `desc` has the shapes of the descriptor the flow
 * was created from, but the pointers inside desc->layers[] are device pointers (identity_features / transform_features
 * are ignored).  Every packed buffer (BatchNorm / bias folding, FP32 matrices, swizzled TF32 / FP16 tile streams, knot
 * tables) is recomputed by a few kernels on `stream`: no device-to-host copy, no allocation after the first call.
 * This is what `model.eval()` after an optimizer step amounts to in Algorithm 2, once per training cycle
 * (hybrid_NF_MCMC/main_algorithm_2.py:450-451 -> 476). */
int fs_flow_update(fs_flow* flow, const fs_flow_desc* desc /*host struct, device pointers in layers*/, void* stream);

/* ResidualNet.forward of the conditioner of layer `layer`  (NF/normflows/nets/resnet.py:92-104, eval mode)
 * on ready-made periodic features [rows, 2N] (NF/normflows/utils/nn.py:120-137) -> theta [rows, 3 nb + 1, N]:
 * the library keeps the spline parameters parameter-major (the reference's column j (3nb+1) + k is k N + j here). */
int fs_flow_conditioner(fs_flow* flow, int layer, const float* features, int rows, float* theta,
                        void* workspace, size_t workspace_bytes, int precision, void* stream);

/* 1 when the packed flow has the tensor-core conditioner (sm_100, H = 128 / 256, at least one residual block), so
 * FS_PREC_TF32 can be passed to the calls below; 0 otherwise (FS_PREC_FP32 only). */
int fs_flow_has_tensor_path(const fs_flow* flow);

/* Conditioner + conditional spline of the transformed half of one coupling layer in ONE kernel (tensor-core path
 * with the fused spline epilogue; FS_ERR_UNSUPPORTED for flow shapes without it: H not 128 / 256 or nb > 32):
 * PiecewiseRationalQuadraticCoupling forward / inverse on the transformed features
 * (NF/normflows/flows/neural_spline/coupling.py:86-102 / 126-135).  direction 1 = density direction, 2 = sampling.
 * features [rows, 2N] as for fs_flow_conditioner; xin [rows, D] is the layer input, the transformed half of
 * xout [rows, D] is written (rolled by D/2 in the density direction) and logdet [rows] (may be NULL) accumulated. */
int fs_flow_coupling(fs_flow* flow, int layer, int direction, const float* features, const float* xin, float* xout,
                     float* logdet, int rows, int* nan_flag, void* stream);

/* All K coupling layers of a pass in ONE launch of K x ceil(rows / 128) thread blocks (the schedule fs_flow_inverse /
 * fs_flow_forward use on the fused tensor path when it pays - fs_flow_uses_layer_parallel - and the identity features
 * stay identity features under the roll by D/2 - every even-N configuration of the reference's drivers, SURVEY.md
 * A.4-Q2: the conditioner inputs of all layers then follow from the unconditional splines alone and only the spline
 * chain of the transformed features is sequential: a step's thread block starts its spline chunks when the previous
 * step of its row tile has written all of its coordinates).
 * features: K row-tiled feature matrices (FS_FEATURES_TILED layout) in step order - density (direction 1): layers
 * K-1 .. 0, sampling (2): 0 .. K-1 - fs_flow_tiled_features_bytes(rows, 2N) bytes apart; buf0 [rows, D]: input (only its
 * transformed columns are read); the steps alternate between buf0 and buf1, the last one writes buf[K & 1];
 * logdet_parts [K, rows]: one log-det partial per step; scratch: K * ceil(rows / 128) * 8 ints.
 * FS_ERR_UNSUPPORTED when the flow has no such path. */
int fs_flow_coupling_all(fs_flow* flow, int direction, const float* features, float* buf0, float* buf1,
                         float* logdet_parts, int* scratch, int rows, int* nan_flag, void* stream);

/* 1 when fs_flow_inverse / fs_flow_forward run a pass of `rows` rows at `precision` as one layer-parallel launch, 0 when
 * they launch layer by layer (no such path for the flow, or one launch per layer already fills the GPU). */
int fs_flow_uses_layer_parallel(const fs_flow* flow, int rows, int precision);

/* Scheduling hint for the passes of this flow: 0 = automatic (default: layer-parallel only when measured to pay next to
 * concurrent streams), 1 = prefer - the caller's fs_flow_inverse / fs_flow_forward calls have the GPU to themselves
 * (Algorithm 2's per-cycle sample / log_prob, hybrid_NF_MCMC/main_algorithm_2.py:476-500), so every pass with at
 * least two whole layer steps resident runs layer-parallel, 2 = never.  Results do not depend on the mode beyond
 * the summation order of the per-layer log-det partials. */
int fs_flow_set_layer_parallel(fs_flow* flow, int mode);

/* Feature layout of the tensor path.  fs_flow_inverse / fs_flow_forward keep the periodic features of a chunk in
 * 128-row tiles of quads, element (row b, feature k) at [b / 128][k / 4][b % 128][k % 4], so that the kernel's
 * row-per-thread loads are coalesced.  A caller of fs_flow_coupling that already holds the features in that layout
 * (fs_flow_tile_features converts a row-major [rows, K0] matrix; fs_flow_tiled_features_bytes sizes the buffer) passes
 * direction | FS_FEATURES_TILED; without the flag the row-major matrix is read as it is (slower feature stage). */
#define FS_FEATURES_TILED 0x10
size_t fs_flow_tiled_features_bytes(int rows, int K0);
int fs_flow_tile_features(const float* features, int rows, int K0, float* tiled, void* stream);

/* NormalizingFlow.inverse_and_log_det / log_prob  (NF/normflows/core.py:71-86,198-214):
 * x [B, D] -> z [B, D], logdet [B]; x_in = x - in_shift (MC-box -> centred coords,
 * MCMC/monte_carlo.py:251-258).  logq (nullable) = logdet + UniformParticle.log_prob(z)
 * (NF/normflows/Energy/Uniform.py:50-74).  nan_flag (nullable) is set non-zero if a
 * NaN was produced (the reference raises ValueError, utils/splines.py:176-183). */
int fs_flow_inverse(fs_flow* flow, const float* x, int B, double in_shift,
                    float* z, float* logdet, float* logq, int* nan_flag,
                    void* workspace, size_t workspace_bytes, int precision, void* stream);

/* NormalizingFlow.forward_and_log_det / sample  (NF/normflows/core.py:28-56,178-196):
 * z [B, D] -> x [B, D] (+ out_shift: centred -> MC-box coords), logdet [B] (nullable). */
int fs_flow_forward(fs_flow* flow, const float* z, int B, double out_shift,
                    float* x, float* logdet, int* nan_flag,
                    void* workspace, size_t workspace_bytes, int precision, void* stream);

/* Development aid (FS_TC_DEBUG=1): wait-cycle counters of the last tensor-core conditioner launch,
 * 16 int64 per CTA; returns the number of CTAs copied to `host`. */
int fs_tc_debug_read(long long* host, int max_ctas);

/* ---- training target of Algorithm 2 ---------------------------------------------------- */

/* SimpleLJ._energy / DoubleWellLJ._energy (NF/normflows/Energy/SimpleLJ.py:15-39, 114-128), the target energy of
 * NormalizingFlow.reverse_kld (NF/normflows/core.py:139-141): x [B, n_particles, 2] centred coordinates ->
 * E [B] = (soft-core LJ over the n + 1 particles incl. one at the origin, coordinates wrapped, no minimum image)
 * / temperature + double well (pot->num_wells = 0: SimpleLJ).  dEdx [B, 2 n_particles] (nullable) = gradient of E
 * with respect to x, for autograd.  pot: V0, r0, k of the wells (r_cut / r_core unused), host pointer, may be NULL. */
int fs_target_energy(const float* x, int B, int n_particles, double bound, double temperature,
                     const fs_pot* pot /*host*/, float* E, float* dEdx, void* stream);

/* Rational-quadratic spline of the density direction for TRAINING, and its reverse-mode derivative
 * (unconstrained_rational_quadratic_spline / rational_quadratic_spline, NF/normflows/utils/splines.py:16-161, 203-222,
 * as autograd differentiates them under NormalizingFlow.forward_kld, NF/normflows/core.py:88-108).
 * x [rows, N]; the parameters of coordinate j of row r are the 3 nb + 1 floats at theta + r * theta_row_stride + j (3 nb + 1)
 * = [nb widths | nb heights | nb + 1 derivatives] (theta_row_stride = 0: shared by all rows, the unconditional spline);
 * width / height logits are multiplied by `scale` (1 / sqrt(hidden) for the conditional spline, coupling.py:340-342).
 * fwd: y, logdet [rows, N].  bwd: grad_x [rows, N] and grad_theta [rows, N, 3 nb + 1] from grad_y, grad_logdet [rows, N]
 * (the forward is recomputed; sum grad_theta over rows when the parameters are shared). */
int fs_spline_train_fwd(const float* x, const float* theta, long long theta_row_stride, int rows, int N, int nb,
                        double bound, double scale, float* y, float* logdet, void* stream);
int fs_spline_train_bwd(const float* x, const float* theta, long long theta_row_stride, int rows, int N, int nb,
                        double bound, double scale, const float* grad_y, const float* grad_logdet,
                        float* grad_x, float* grad_theta, void* stream);

/* ---- one forward-KL training step of the whole flow ------------------------------------------- */

/* NormalizingFlow.forward_kld(x) + loss.backward() (NF/normflows/core.py:88-108; the minibatch step of
 * hybrid_NF_MCMC/main_algorithm_1.py:297-320 and main_algorithm_2.py:437-452) for a flow of K
 * CircularCoupledRationalQuadraticSpline layers (flows/neural_spline/wrapper.py:16-93, coupling.py:86-102) whose
 * conditioner is the ResidualNet of nets/resnet.py:7-104 with BatchNorm1d in training mode and dropout 0:
 *     loss = -mean_rows( sum_layers log|det J| ),  gradients of every trainable tensor WRITTEN (not accumulated).
 * The descriptor holds HOST arrays of DEVICE pointers.  Per layer (flow index 0 .. K-1) PER = 2 + 8 n_blocks + 5 tensors:
 *     initial_layer.weight [H, 2N], .bias [H],
 *     per block: batch_norm_layers.0.weight, .bias [H], linear_layers.0.weight [H, H], .bias [H],
 *                batch_norm_layers.1.weight, .bias,     linear_layers.1.weight,       .bias,
 *     final_layer.weight [N (3 nb + 1), H], .bias,
 *     unconditional_transform.unnormalized_widths [N, nb], .unnormalized_heights [N, nb], .unnormalized_derivatives [N, nb + 1];
 * grads: the same order (e.g. views of one flat bucket); bn_running: per layer 4 n_blocks tensors
 * (block b: batch_norm_layers.0.running_mean, .running_var, batch_norm_layers.1.running_mean, .running_var),
 * updated in place with `bn_momentum` and the unbiased batch variance when update_running != 0 (num_batches_tracked is
 * the caller's).  feature_scale: PeriodicFeaturesElementwise.scale (utils/nn.py:65-137).  The pointers must stay valid
 * (optimizers update in place).  FS_ERR_UNSUPPORTED when the identity features are not closed under the roll by D/2
 * (odd N) or N (3 nb + 1) is not a multiple of 4: the caller keeps the autograd path. */
typedef struct fs_train fs_train;
typedef struct fs_train_desc {
    int K, N, H, n_blocks, nb;
    double bound, feature_scale, bn_eps, bn_momentum;
    const int* transform_features; /* host [N] */
    const int* identity_features;  /* host [N] */
    float* const* params;          /* host [K * PER] */
    float* const* grads;           /* host [K * PER] */
    float* const* bn_running;      /* host [K * 4 n_blocks] */
} fs_train_desc;
int fs_train_create(const fs_train_desc* desc, fs_train** out);
void fs_train_destroy(fs_train* t);
/* x [B, 2N] (centred coordinates), B >= 2; loss: one float on the device.  Repeatable bit for bit. */
int fs_train_forward_kld(fs_train* t, const float* x, int B, float* loss, int update_running, void* stream);

/* torch.optim.Adam(params, lr, weight_decay).step() on ONE flat float32 buffer (the optimizer of both drivers:
 * hybrid_NF_MCMC/main_algorithm_1.py:297-320, main_algorithm_2.py:440-451): params / grads / exp_avg / exp_avg_sq [n]
 * (16-byte aligned), L2 weight decay, bias-corrected moments.  state: 4 device floats - [0] steps applied so far (zero
 * it, with the moments, for a new optimizer), [1] 1 if this call applied its step, [2] lr / bc1, [3] 1 / sqrt(bc2).
 * The step is skipped ON THE DEVICE when *skip != 0 (nullable; e.g. a flag all-reduced over the ranks) or *loss is
 * NaN / Inf (nullable; the reference's `if ~(isnan(loss) | isinf(loss))`, main_algorithm_2.py:449). */
int fs_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                 float* state, const float* skip, const float* loss, float lr, float beta1, float beta2,
                 float eps, float weight_decay, void* stream);

/* ---- affine (RealNVP) coupling and periodic shifts ----------------------------------------- */

/* MaskedAffineFlow.forward / inverse (NF/normflows/flows/affine/coupling.py:163-229) and the element-wise half of
 * AffineCoupling.forward / inverse (:99-160).  z, out [rows, n] with row strides; mask [n] (1 = feature unchanged) or
 * NULL (all n features transformed); scale / shift read at row * row_stride + j * elem_stride (elem_stride 2 for
 * AffineCoupling's interleaved parameters), NULL = none.  scale_map 0 = exp, 1 = sigmoid, 2 = sigmoid_inv;
 * inverse != 0 applies the inverse map; nan_rule != 0 turns non-finite parameters into NaN (MaskedAffineFlow).
 * logdet [rows] (nullable) is WRITTEN (sum over the n features). */
int fs_affine_coupling(const float* z, long long z_row_stride, const float* mask, const float* scale,
                       long long scale_row_stride, int scale_elem_stride, const float* shift,
                       long long shift_row_stride, int shift_elem_stride, int rows, int n, int scale_map, int inverse,
                       int nan_rule, float* out, long long out_row_stride, float* logdet, void* stream);

/* PeriodicShift.forward / inverse, PeriodicWrap.inverse (NF/normflows/flows/periodic.py:6-73): out = z with
 * remainder(z + shift + bound, 2 bound) - bound on the columns with col_slot[col] = s >= 0 (bound[s], shift[s];
 * col_slot = -1 leaves a column unchanged). */
int fs_periodic_shift(const float* z, int rows, int D, const int* col_slot, const float* bound, const float* shift,
                      float* out, void* stream);

/* ---- observables of sampled configurations ------------------------------------------- */

/* classify_particles + the per-configuration part of calculate_well_statistics
 * (hybrid_NF_MCMC/utils.py:61-141): pos [B, N, 2] MC-box coordinates ->
 * cls [B, N] (0 outside, 1 well A around (Lx/4, Ly/2), 2 well B around (3Lx/4, Ly/2); radius 1.1 r0, minimum image),
 * state [B] (1 all particles in A, 2 all in B, 0 otherwise), avg_x [B] mean x.  Any output may be NULL. */
int fs_classify_wells(const float* pos, int B, int N, double Lx, double Ly, double r0, unsigned char* cls,
                      unsigned char* state, double* avg_x, void* stream);

/* Per-configuration pair-distance histogram of calculate_pair_correlation (hybrid_NF_MCMC/utils.py:530-556):
 * cfg [B, N, 2] centred coordinates in [-bound, bound], float32 minimum-image distances as numpy computes them,
 * bins [k dr, (k+1) dr), k < nbins, last bin closed, zero distances dropped, every unordered pair counted twice
 * -> counts [B, nbins]. */
int fs_pair_histogram(const float* cfg, int B, int N, double bound, double dr, int nbins, unsigned int* counts,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOWSTATE_B200_H */
