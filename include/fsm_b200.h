/* fsm_b200.h — C ABI of the B200-native pseudo-spectral stepper (libfsm_b200.so).
 *
 * Drop-in boundary for ONE path of qiauil/torchfsm (v0.0.4): the per-step update behind
 * Operator.integrate(u_0, mesh, dt, step). The reference has no FFI; its seams are Python
 * protocols. Each entry point below names the reference code it replaces
 * (paths relative to the reference root). INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - plain pointers and sizes only; every device buffer is allocated and owned by the caller
 *    (PyTorch); the library never allocates or frees device memory and keeps no pointer beyond
 *    those stored in the plan descriptor (tables), which the caller must keep alive.
 *  - all work is enqueued on the caller's stream, asynchronously; no device synchronisation.
 *  - return 0 on success, negative errno-style codes otherwise: -EINVAL bad argument,
 *    -ENOSYS unsupported configuration, -ENOMEM workspace too small, -EIO CUDA error;
 *    fsm_last_error() returns a thread-local message. No C++ exception crosses the ABI.
 *
 * Spectral layout ("rot-half"): real fields are stored as half spectra (Hermitian redundancy
 * removed along the last physical axis) with the x wavenumber fastest:
 *     1-D  [kx < n0/2+1]
 *     2-D  [ky < n1/2+1][kx < n0]
 *     3-D  [ky < n1][kz < n2/2+1][kx < n0]
 * per (batch, channel), complex interleaved (re, im) in the plan dtype. fsm_half_to_full /
 * fsm_full_to_half convert from/to the reference's full C2C layout (B, C, n0, n1, n2).
 */
#ifndef FSM_B200_H
#define FSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSM_ABI_VERSION 1

typedef struct fsm_plan fsm_plan;

enum fsm_dtype { FSM_F32 = 0, FSM_F64 = 1 };

/* Nonlinear program = the fused form of the reference's nonlinear core(s). */
enum fsm_program {
    FSM_PROG_LINEAR = 0,      /* no nonlinear term (operator/_base.py:451-453 -> ETDRK0)            */
    FSM_PROG_CONVECTION = 1,  /* u.grad(u), channels == ndim   (operator/generic/_convection.py:18-48) */
    FSM_PROG_KS = 2,          /* 1/2|grad phi|^2 - mean        (operator/dedicated/_ks_convection.py:18-38) */
    FSM_PROG_NS2D_VORT = 3,   /* vorticity convection, 2-D     (operator/dedicated/_navier_stokes.py:27-46) */
    FSM_PROG_NS3D = 4         /* convection + pressure projection, 2-D or 3-D, channels == ndim
                                 (operator/dedicated/_navier_stokes.py:231-254) */
};

/* Time integrators (integrator/_etdrk.py:10-93, integrator/_stable_etdrk/_uncached.py:18-228,
 * integrator/_stable_etdrk/_setdrk_step.py:5-82, integrator/_rk.py:43-58,142-155). */
enum fsm_integrator {
    FSM_INT_ETDRK0 = 0, FSM_INT_ETDRK1 = 1, FSM_INT_ETDRK2 = 2,
    FSM_INT_SETDRK1 = 3, FSM_INT_SETDRK2 = 4, FSM_INT_SETDRK3 = 5, FSM_INT_SETDRK4 = 6,
    FSM_INT_RK4 = 7
};

/* Plan descriptor. All pointers are DEVICE pointers in the plan dtype unless stated otherwise.
 * Tables are real arrays in the rot-half layout, shape [tab_channels][modes], built by the
 * caller with the reference's own expressions (the plain-ETDRK tables cancel
 * catastrophically in fp32, so they are inputs, not something the library may "improve"). */
typedef struct fsm_desc {
    int32_t struct_size;      /* sizeof(fsm_desc), ABI guard                                        */
    int32_t dtype;            /* enum fsm_dtype                                                     */
    int32_t ndim;             /* 1..3                                                               */
    int32_t n[3];             /* grid points per axis (powers of two, 8..1024); unused = 1           */
    int32_t batch;            /* B                                                                  */
    int32_t channels;         /* C                                                                  */
    int32_t program;          /* enum fsm_program                                                   */
    int32_t integrator;       /* enum fsm_integrator                                                */
    int32_t kmax[3];          /* dealiasing: modes with |m_i| <= kmax[i] are kept (mesh.py:443-461)  */
    int32_t ks_remove_mean;   /* KS: subtract the batch+space mean (_ks_convection.py:34-36)         */
    int32_t tab_channels;     /* 1 = one table for all channels, or == channels                      */
    int32_t chunk;            /* samples processed per pass launch; 0 = choose automatically         */
    double dt;                /* time step (RK4 only uses it; ETD tables already contain it)         */
    double nl_coef;           /* scalar coefficient of the convective nonlinear term                 */
    double ks_ext_sum;        /* reserved (0)                                                         */
    int32_t dynamic_force;    /* FSM_PROG_NS3D: 1 = the stages will be driven one by one through fsm_stage_run with a
                                 state-dependent force spectrum (keeps the reference's grouping of the two-stage step)  */
    int32_t tab_complex;      /* 1: every coefficient table holds complex entries (2 reals per mode): odd-order
                               * linear terms such as KdV's dispersion (_spatial_derivative.py:7-20). 1-D grids only */
    const void* dk[3];        /* per-axis 2*pi*f(m), Nyquist entry zeroed (length n[i])  mesh.py:399-404 */
    const void* dkraw[3];     /* per-axis 2*pi*f(m) as the reference has it (length n[i])           */
    const void* tab_exp;      /* exp(dt L)            _etdrk.py:21 / _uncached.py:30                 */
    const void* tab_half_exp; /* exp(dt L / 2)        _uncached.py:135,191                           */
    const void* tab_coef[6];  /* coef_1..coef_6       _etdrk.py:43-45,66-70 / _uncached.py:72-211. ETDRK2 / SETDRK2 use
                                 coef_1 and coef_2 only; a caller may put coef_1 - coef_2 into the coef_3 slot, the step then
                                 runs as d = E u + (coef_1 - coef_2) N0, u' = d + coef_2 N(a): one array read less */
    const void* tab_lin;      /* L itself (RK4 and fsm_rhs)   operator/_base.py:339-357              */
    const void* source_hat;   /* optional constant source spectrum, complex [C][modes], coefficient
                                 folded in (operator/_base.py:994-1015)                              */
    int32_t slab_rank;        /* slab decomposition of ONE 3-D grid over slab_nranks GPUs (0 / 1 = off):  */
    int32_t slab_nranks;      /* spectral state and tables hold the local ky lines [n1/P][n2/2+1][n0] with CYCLIC
                                 ownership (local line t of rank r is ky = r + P t: every rank owns the same share of
                                 the dealiased band, and the inverse-side exchange ships kept lines only); physical
                                 fields the local x slab [n0/P][n1][n2] (blocked); see fsm_slab_phase      */
    int32_t lanes;            /* ensembles: independent sample ranges stepped concurrently on internal streams
                                 (fork/join on the caller's stream); 0 = choose, 1 = everything on the caller's stream */
    int32_t tab_batched;      /* 1: every table holds one copy PER SAMPLE, [batch][tab_channels][modes] (tensor-valued
                                 coefficients on linear terms: operator/_base.py:339-357 builds L of shape (B, C, N...)) */
    const void* force_hat;    /* FSM_PROG_NS3D only: constant spectrum added to coef * conv_hat BEFORE the pressure
                                 projection, complex [C][modes] = -coef * f_hat of NSPressureConvection(external_force)
                                 for a force that does not depend on u (_navier_stokes.py:237-254) */
    const void* nl_coef_b;    /* optional per-sample coefficient of the convective term, real [batch] in the plan dtype
                                 (tensor-valued coefficient, operator/_base.py:375-403); replaces nl_coef when set */
} fsm_desc;

/* replaces: OperatorLike._build_integrator (operator/_base.py:441-526) */
int fsm_plan_create(fsm_plan** out, const fsm_desc* desc);
void fsm_plan_destroy(fsm_plan* plan);
/* bytes of caller-provided device workspace every call below needs */
size_t fsm_workspace_bytes(const fsm_plan* plan);

/* replaces: the hot loop `for i in range(step): u_hat = integrator.forward(u_hat, dt)`
 * (operator/_base.py:732-735). u_hat is updated in place (rot-half layout, [B][C][modes]). */
int fsm_step(fsm_plan* plan, void* u_hat, void* workspace, size_t ws_bytes, int n_steps, void* stream);

/* replaces: the operator closure L*u_hat + N(u_hat) (operator/_base.py:408-439), used by
 * Operator.__call__ (operator/_base.py:753-790). out_hat must not alias u_hat. */
int fsm_rhs(fsm_plan* plan, const void* u_hat, void* out_hat, void* workspace, size_t ws_bytes, void* stream);

/* replaces: FourierMesh.fft on a real field (mesh.py:481-485): u (B,C,n0,n1,n2) real -> rot-half */
int fsm_r2c(fsm_plan* plan, const void* u, void* u_hat, void* workspace, size_t ws_bytes, void* stream);
/* replaces: FourierMesh.ifft(...).real (mesh.py:487-491): rot-half -> (B,C,n0,n1,n2) real */
int fsm_c2r(fsm_plan* plan, const void* u_hat, void* u, void* workspace, size_t ws_bytes, void* stream);

/* rot-half <-> the reference's full complex spectrum (B,C,n0,n1,n2): u_0_fft input,
 * return_in_fourier and recorder frames (operator/_base.py:727-751, traj_recorder.py:46-55) */
int fsm_half_to_full(fsm_plan* plan, const void* u_hat, void* full_hat, void* stream);
int fsm_full_to_half(fsm_plan* plan, const void* full_hat, void* u_hat, void* stream);

/* Point-wise spectral map in the rot-half layout (no transform inside):
 *     out_hat[b][co][k] = sum_t coef_t * prod_a (i k_a)^power_t[a] * (1/lap(k))^inv_laplacian_t * in_hat[b][in_channel_t][k]
 * summed over the terms with out_channel_t == co; 1/lap is the reference's invert_laplacian (0 -> 1, mesh.py:412-418).
 * replaces the symbol products of the channel-changing cores, which the reference evaluates as full-size tensor
 * multiplications: _GradCore (operator/generic/_grad.py:6-15), _DivCore (_div.py:9-23), _Curl2DCore/_Curl3DCore
 * (_curl.py:9-55), _Vorticity2VelocityCore (operator/dedicated/_navier_stokes.py:75-92) and the pressure solve of
 * _Velocity2PressureCore/_Vorticity2PressureCore (:158-163, :213-217); linear cores in the same operator sum map to
 * terms too (Operator.__call__, operator/_base.py:753-790). Symbols are Hermitian-projected (a term vanishes where
 * the powers on the axes sitting on their Nyquist index add up to an odd number), which equals `.real` of the
 * reference's inverse transform. dealias != 0 zeroes input modes outside the plan's kmax box first
 * (operator/_base.py:381-393). c_in, c_out <= 6, n_terms <= 32; in_hat [B][c_in][modes], out_hat [B][c_out][modes]
 * must not alias. Slab-decomposed plans act on their local ky lines. */
typedef struct fsm_map_term {
    int32_t out_channel, in_channel;
    int32_t power[3];
    int32_t inv_laplacian;
    double coef;
} fsm_map_term;
int fsm_spectral_map(fsm_plan* plan, const void* in_hat, int32_t c_in, void* out_hat, int32_t c_out,
                     const fsm_map_term* terms, int32_t n_terms, int32_t dealias, void* stream);

/* Nonlinear terms evaluated OUTSIDE the fused programs (plans created with FSM_PROG_LINEAR and any integrator): the
 * caller composes the evaluation from this library's passes (fsm_c2r, a point-wise physical-space function,
 * fsm_r2c, fsm_spectral_map) and hands the spectrum of the nonlinear term to the integrator stage. This is how the
 * reference's open NonlinearFunc protocol (operator/_base.py:76-131) is served for cores without a fused program:
 * _ImplicitFuncSourceCore (generic/_source.py:20-42: fft(f(ifft(u).real)), f a user callable) and
 * _ConservativeConvectionCore (generic/_conservative_convection.py:8-27).
 *   fsm_stage_input:   byte offset inside the workspace of the array stage `stage` evaluates its nonlinear term on, or
 *                      -1 when that array is the caller's u_hat; stage -1 = the right-hand side. Returns the stage
 *                      count of the plan (>= 0) or a negative error.
 *   fsm_stage_combine: run the combine of stage `stage` (integrator/_etdrk.py:47-82, _setdrk_step.py:5-82,
 *                      _rk.py:43-58) with fresh_hat [B][C][modes] as N(stage input) (+ the plan's constant source);
 *                      stage -1 writes L u + N into rhs_out instead (operator/_base.py:408-439).
 *   fsm_mask_state:    zero the modes outside the plan's dealiasing box in place (operator/_base.py:381-385).
 *   fsm_sym_outer:     physical space, out[b][pair(i,j)] = u[b][i] * u[b][j] for i <= j (row-major pairs), channels <= 6. */
int fsm_stage_input(const fsm_plan* plan, int32_t stage, int64_t* ws_offset);
int fsm_stage_combine(fsm_plan* plan, int32_t stage, void* u_hat, const void* fresh_hat, void* rhs_out, void* workspace,
                      size_t ws_bytes, void* stream);
int fsm_mask_state(fsm_plan* plan, void* state_hat, int32_t channels, void* stream);
/* NSPressureConvection(external_force) with a force operator that DEPENDS on the state (_navier_stokes.py:237-254):
 * the caller evaluates f_hat = external_force(stage input) [B][C][modes] with this library's passes BEFORE the stage
 * runs (the reference evaluates the force on the un-dealiased state, then dealiases that state in place), and
 * fsm_stage_run executes the fused evaluation + combine of stage `stage` with it: coef * (P(conv - f) + f).
 * Plans must be created with dynamic_force = 1; stage -1 = the right-hand side into rhs_out. */
int fsm_stage_run(fsm_plan* plan, int32_t stage, void* u_hat, const void* force_hat, void* rhs_out, void* workspace,
                  size_t ws_bytes, void* stream);
int fsm_sym_outer(fsm_plan* plan, const void* u, void* out, int32_t channels, void* stream);
/* out = base + sum_j coefs[j] * terms[j] over `count` complex elements of the plan dtype (n_terms <= 8, out may alias
 * base): the stage states and the update of the explicit Runge-Kutta family other than RK4, whose right-hand sides come
 * from fsm_rhs (replaces `x_t + dt * sum([a_i * k for a_i, k in zip(ca_i[1:], ks)])`, integrator/_rk.py:43-58). */
int fsm_lincomb(fsm_plan* plan, void* out, const void* base, const void* const* terms, const double* coefs, int32_t n_terms,
                int64_t count, void* stream);

/* introspection for benchmarks: kernels launched per step, algorithmic bytes per step
 * (transform-pass model), modes per field, chunk size */
int fsm_plan_info(const fsm_plan* plan, int64_t* launches_per_step, int64_t* algo_bytes_per_step,
                  int64_t* modes_per_field, int32_t* chunk);

/* bytes the passes of each class {IX, MID, PHYS, FX + combine} really read + write per step: kept (dealiased) modes
 * only on the inverse side, coefficient tables excluded. The section 8d model (fsm_plan_info / fsm_profile_read) counts
 * whole fields per pass; this is what a pass's GB/s must be quoted on to stay below the HBM peak. */
int fsm_plan_traffic(const fsm_plan* plan, int64_t* touched_bytes_per_step4);

/* introspection for tests: number of integrator stages; kinds[i] = index of the compile-time combine
 * structure stage i runs with in the forward-x epilogue (-1 = generic data-driven path) */
int fsm_stage_kinds(const fsm_plan* plan, int32_t* kinds, int32_t capacity);

/* Slab-decomposed evaluation (single large 3-D grids): the library runs the local passes of
 * one phase; the caller performs the all-to-all between phases (torch.distributed all_to_all_single over
 * NCCL/NVLink, equal splits) on the exchange buffers, which the kernels read and write directly in
 * rank-blocked layouts (no pack/unpack pass).
 *   op FSM_SLAB_STEP / FSM_SLAB_RHS, stage s:  phase 0: state -> x-transform -> send          [all-to-all]
 *                                              phase 1: recv -> y, z, product, z, y -> send    [all-to-all]
 *                                              phase 2: recv -> x-transform + integrator combine
 * The local x range can be split into nsub sub-slabs (power of two): the exchange buffers are then laid out
 * [sub-slab][rank][...], phase 0 and 2 cover all sub-slabs, phase 1 runs sub-slab `sub`, so the all-to-all of
 * one sub-slab overlaps the local chain of another (caller-side streams/events).
 *   op FSM_SLAB_R2C: phase 1 (aux = local physical slab -> send), phase 2 (recv -> u_hat)
 *   op FSM_SLAB_C2R: phase 0 (u_hat -> send), phase 1 (recv -> aux = local physical slab)
 * fsm_slab_info returns the complex-element counts of the two exchanges for an op and the stage count. */
enum fsm_slab_op { FSM_SLAB_STEP = 0, FSM_SLAB_RHS = 1, FSM_SLAB_R2C = 2, FSM_SLAB_C2R = 3 };
int fsm_slab_phase(fsm_plan* plan, int op, int stage, int phase, int sub, int nsub, void* u_hat, void* aux,
                   void* workspace, size_t ws_bytes, void* send, void* recv, void* stream);
int fsm_slab_info(const fsm_plan* plan, int op, int64_t* exch1_elems, int64_t* exch2_elems, int32_t* n_stages);

/* Direct exchange: register, for exchange 1 (spectral ky-slabs -> physical x-slabs, written by phase 0) or
 * exchange 2 (x-slabs -> ky-slabs, written by phase 1), the receive buffer of EVERY rank as addressable from
 * this process (CUDA: NVLink peer mappings, e.g. torch symmetric memory; emulator: shared host memory). The
 * transform kernels then store straight into the destination ranks' buffers -- the transfer rides inside the
 * kernel, no send buffer, no separate all-to-all. The caller separates writers and readers with a barrier
 * across ranks on the stream (after phase 0 / phase 1, before the next phase reads). n = 0 or ptrs = NULL
 * returns to the send-buffer path. Replaces the all_to_all a torch.distributed slab FFT would issue. */
int fsm_slab_peers(fsm_plan* plan, int exchange, const void* const* ptrs, int32_t n);

/* KS ensembles sharded over ranks: the batch mean of _ks_convection.py:34-36 spans every rank, but it only
 * touches the k=0 bin, which never feeds back. With a log attached, every nonlinear evaluation appends the
 * LOCAL sum of the per-sample zero modes (plan dtype, `capacity` entries, device memory owned by the caller;
 * a new call resets the position). One all-reduce of the log after the run gives the exact zero-mode correction
 * (torchfsm_b200.FusedStepper._step_half_ks_sharded, Operator.set_ensemble_group) - no collective inside the step. */
int fsm_ks_log(fsm_plan* plan, void* log, int64_t capacity);

/* per-pass device timing for benchmarks: when enabled every pass launch is bracketed by CUDA
 * events on the caller's stream. fsm_profile_read waits for the recorded events, returns summed
 * milliseconds and launch counts per pass class {0: IX, 1: MID, 2: PHYS, 3: FX+combine} since the
 * last read, and the algorithmic bytes each class moves per step (SURVEY.md §8d model). */
int fsm_profile_enable(fsm_plan* plan, int on);
int fsm_profile_read(fsm_plan* plan, double* ms4, int64_t* launches4, int64_t* algo_bytes_per_step4);

const char* fsm_last_error(void);
int fsm_abi_version(void);
/* 0 = CUDA sm_100a build (the product); 1 = host emulator build used only by the CPU test-suite */
int fsm_backend(void);

#ifdef __cplusplus
}
#endif
#endif /* FSM_B200_H */
