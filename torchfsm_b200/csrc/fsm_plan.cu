// Plan, stage programs and the C ABI (include/fsm_b200.h) of libfsm_b200.so.
#include "fsm_b200.h"
#include "fsm_launch.h"
#include "fsm_pointwise.cuh"

#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <utility>

#ifndef FSM_LANES_DEFAULT
#define FSM_LANES_DEFAULT 2
#endif
#ifndef FSM_PF_DEFAULT
#define FSM_PF_DEFAULT 5   // PF_IX | PF_FX_OPS (measured on C3: 1.908 -> 1.861 ms/step; PF_PHYS and PF_FX_WIN cost time)
#endif

namespace fsm {

#define FSM_DECL_TABLE(N)                      \
    const LaunchTable<float>* table_f32_##N(); \
    const LaunchTable<double>* table_f64_##N();
FSM_DECL_TABLE(8)
FSM_DECL_TABLE(16)
FSM_DECL_TABLE(32)
FSM_DECL_TABLE(64)
FSM_DECL_TABLE(128)
FSM_DECL_TABLE(256)
FSM_DECL_TABLE(512)
FSM_DECL_TABLE(1024)

template <> const LaunchTable<float>* launch_table<float>(int N) {
    switch (N) {
        case 8: return table_f32_8();
        case 16: return table_f32_16();
        case 32: return table_f32_32();
        case 64: return table_f32_64();
        case 128: return table_f32_128();
        case 256: return table_f32_256();
        case 512: return table_f32_512();
        case 1024: return table_f32_1024();
        default: return nullptr;
    }
}
template <> const LaunchTable<double>* launch_table<double>(int N) {
    switch (N) {
        case 8: return table_f64_8();
        case 16: return table_f64_16();
        case 32: return table_f64_32();
        case 64: return table_f64_64();
        case 128: return table_f64_128();
        case 256: return table_f64_256();
        case 512: return table_f64_512();
        case 1024: return table_f64_1024();
        default: return nullptr;
    }
}

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// --------------------------------------------------------------------------------------------
// Stage programs. Arrays: 0 = U (caller state, updated in place), 1..4 = scratch state arrays,
// 5 = external output (fsm_rhs). Tables: see enum below. X_0 is the fresh nonlinear term.
// --------------------------------------------------------------------------------------------
enum { ARR_U = 0, ARR_S1 = 1, ARR_S2 = 2, ARR_S3 = 3, ARR_S4 = 4, ARR_EXT = 5, ARR_COUNT = 6 };
enum { TAB_EXP = 0, TAB_HALF = 1, TAB_C1 = 2, TAB_C2, TAB_C3, TAB_C4, TAB_C5, TAB_C6, TAB_LIN = 8, TAB_COUNT = 9 };
constexpr int NO_TERM = -2, SCALAR = -1;

struct Coef {
    int tab = NO_TERM;
    double a = 0, b = 0;  // coefficient = a + b * table + b2 * table2
    int tab2 = -1;
    double b2 = 0;
};
struct Stage {
    int input = ARR_U;          // array the nonlinear term is evaluated on
    int n_in = 0, in[FSM_MAX_IN] = {0, 0, 0};
    int n_out = 0, out[FSM_MAX_OUT] = {0, 0, 0};
    bool has_next = false;      // row FSM_MAX_OUT = next stage state handed over in registers (fused FX+IX)
    int model_io = -1;          // arrays the reference's stage formula reads + writes (SURVEY.md section 8d byte model) when
                                // that differs from n_in + n_out of this formulation
    Coef c[FSM_MAX_OUT + 1][FSM_MAX_IN + 1];
};
static Coef tabc(int tab, double b = 1.0) { Coef c; c.tab = tab; c.a = 0; c.b = b; return c; }
static Coef scal(double a) { Coef c; c.tab = SCALAR; c.a = a; c.b = 0; return c; }
static Coef affine(double a, int tab, double b) { Coef c; c.tab = tab; c.a = a; c.b = b; return c; }
static Coef tab2c(int tab, double b, int tab2, double b2) { Coef c; c.tab = tab; c.a = 0; c.b = b; c.tab2 = tab2; c.b2 = b2; return c; }

static std::vector<Stage> build_stages(int integ, double dt, bool has_lin, bool has_c12) {
    std::vector<Stage> st;
    auto mk_stage = [](int input, std::initializer_list<int> ins, std::initializer_list<int> outs) {
        Stage s;
        s.input = input;
        for (int a : ins) s.in[s.n_in++] = a;
        for (int a : outs) s.out[s.n_out++] = a;
        return s;
    };
    switch (integ) {
        case FSM_INT_ETDRK0: {  // u' = E u                      (_etdrk.py:23-27)
            Stage s = mk_stage(ARR_U, {ARR_U}, {ARR_U});
            s.c[0][1] = tabc(TAB_EXP);
            st.push_back(s);
            break;
        }
        case FSM_INT_ETDRK1:
        case FSM_INT_SETDRK1: {  // u' = E u + c1 N(u)            (_etdrk.py:47-51, _setdrk_step.py:5-11)
            Stage s = mk_stage(ARR_U, {ARR_U}, {ARR_U});
            s.c[0][0] = tabc(TAB_C1);
            s.c[0][1] = tabc(TAB_EXP);
            st.push_back(s);
            break;
        }
        case FSM_INT_ETDRK2:
        case FSM_INT_SETDRK2: {  // a = E u + c1 N0; u' = a + c2 (N(a) - N0)   (_etdrk.py:72-82)
            if (has_c12) {
                // the same step with one array read less: stage 1 also writes d = a - c2 N0 = E u + (c1 - c2) N0 (table
                // coef_1 - coef_2 supplied in the coef_3 slot) in place of N0, stage 2 is u' = d + c2 N(a) and never
                // reads a again. Rounding differs from the reference's grouping by ~1 ulp of c2 N0.
                Stage s1 = mk_stage(ARR_U, {ARR_U}, {ARR_S1, ARR_S2});
                s1.c[0][0] = tabc(TAB_C1); s1.c[0][1] = tabc(TAB_EXP);
                s1.c[1][0] = tabc(TAB_C3); s1.c[1][1] = tabc(TAB_EXP);
                Stage s2 = mk_stage(ARR_S1, {ARR_S2}, {ARR_U});
                s2.c[0][0] = tabc(TAB_C2); s2.c[0][1] = scal(1.0);
                s2.model_io = 3;
                st.push_back(s1);
                st.push_back(s2);
                break;
            }
            Stage s1 = mk_stage(ARR_U, {ARR_U}, {ARR_S1, ARR_S2});
            s1.c[0][0] = tabc(TAB_C1);
            s1.c[0][1] = tabc(TAB_EXP);
            s1.c[1][0] = scal(1.0);
            Stage s2 = mk_stage(ARR_S1, {ARR_S1, ARR_S2}, {ARR_U});
            s2.c[0][0] = tabc(TAB_C2);
            s2.c[0][1] = scal(1.0);
            s2.c[0][2] = tabc(TAB_C2, -1.0);
            st.push_back(s1);
            st.push_back(s2);
            break;
        }
        case FSM_INT_SETDRK3: {  // _setdrk_step.py:28-52 ; S1 = stage state, S2 = N0, S3 = running sum
            Stage s1 = mk_stage(ARR_U, {ARR_U}, {ARR_S1, ARR_S2, ARR_S3});
            s1.c[0][0] = tabc(TAB_C1); s1.c[0][1] = tabc(TAB_HALF);
            s1.c[1][0] = scal(1.0);
            s1.c[2][0] = tabc(TAB_C3); s1.c[2][1] = tabc(TAB_EXP);
            Stage s2 = mk_stage(ARR_S1, {ARR_U, ARR_S2, ARR_S3}, {ARR_S1, ARR_S3});
            s2.c[0][0] = tabc(TAB_C2, 2.0); s2.c[0][1] = tabc(TAB_EXP); s2.c[0][2] = tabc(TAB_C2, -1.0);
            s2.c[1][0] = tabc(TAB_C4); s2.c[1][3] = scal(1.0);
            Stage s3 = mk_stage(ARR_S1, {ARR_S3}, {ARR_U});
            s3.c[0][0] = tabc(TAB_C5); s3.c[0][1] = scal(1.0);
            st.push_back(s1); st.push_back(s2); st.push_back(s3);
            break;
        }
        case FSM_INT_SETDRK4: {  // _setdrk_step.py:55-82 ; S1 = a, S2 = N0, S3 = running sum, S4 = b then c
            Stage s1 = mk_stage(ARR_U, {ARR_U}, {ARR_S1, ARR_S2, ARR_S3});
            s1.c[0][0] = tabc(TAB_C1); s1.c[0][1] = tabc(TAB_HALF);
            s1.c[1][0] = scal(1.0);
            s1.c[2][0] = tabc(TAB_C4); s1.c[2][1] = tabc(TAB_EXP);
            Stage s2 = mk_stage(ARR_S1, {ARR_U, ARR_S3}, {ARR_S4, ARR_S3});
            s2.c[0][0] = tabc(TAB_C2); s2.c[0][1] = tabc(TAB_HALF);
            s2.c[1][0] = tabc(TAB_C5, 2.0); s2.c[1][2] = scal(1.0);
            Stage s3 = mk_stage(ARR_S4, {ARR_S1, ARR_S2, ARR_S3}, {ARR_S4, ARR_S3});
            s3.c[0][0] = tabc(TAB_C3, 2.0); s3.c[0][1] = tabc(TAB_HALF); s3.c[0][2] = tabc(TAB_C3, -1.0);
            s3.c[1][0] = tabc(TAB_C5, 2.0); s3.c[1][3] = scal(1.0);
            Stage s4 = mk_stage(ARR_S4, {ARR_S3}, {ARR_U});
            s4.c[0][0] = tabc(TAB_C6); s4.c[0][1] = scal(1.0);
            st.push_back(s1); st.push_back(s2); st.push_back(s3); st.push_back(s4);
            break;
        }
        case FSM_INT_RK4: {  // k = L s + N(s)  (_rk.py:43-58,142-155); S1 = stage state, S2 = running sum
            auto lin = [&](double a, double b) { return has_lin ? affine(a, TAB_LIN, b) : scal(a); };
            Stage s1 = mk_stage(ARR_U, {ARR_U}, {ARR_S1, ARR_S2});
            s1.c[0][0] = scal(dt / 2); s1.c[0][1] = lin(1.0, dt / 2);
            s1.c[1][0] = scal(dt / 6); s1.c[1][1] = lin(1.0, dt / 6);
            Stage s2 = mk_stage(ARR_S1, {ARR_U, ARR_S1, ARR_S2}, {ARR_S1, ARR_S2});
            s2.c[0][0] = scal(dt / 2); s2.c[0][1] = scal(1.0); s2.c[0][2] = lin(0.0, dt / 2);
            s2.c[1][0] = scal(dt / 3); s2.c[1][2] = lin(0.0, dt / 3); s2.c[1][3] = scal(1.0);
            Stage s3 = s2;
            s3.c[0][0] = scal(dt); s3.c[0][2] = lin(0.0, dt);
            Stage s4 = mk_stage(ARR_S1, {ARR_S1, ARR_S2}, {ARR_U});
            s4.c[0][0] = scal(dt / 6); s4.c[0][1] = lin(0.0, dt / 6); s4.c[0][2] = scal(1.0);
            st.push_back(s1); st.push_back(s2); st.push_back(s3); st.push_back(s4);
            break;
        }
        default: break;
    }
    return st;
}

}  // namespace fsm

using namespace fsm;

#define FSM_MAX_LANES 4
struct fsm_plan {
    fsm_desc d;
    int ndim, n[3], nh, ph, B, C, prog, kprog;  // kprog = kernel-side Prog enum
    bool f64;
    long nmodes, ntot;
    int chunk;
    int nlanes = 1;              // independent sample ranges stepped concurrently on internal streams (see do_step)
    size_t lane_w_bytes[4] = {0, 0, 0, 0};   // per-lane size of the W1, W2, W3, W2b regions
    int nf_ix, nfi, nout;  // fields: IX per channel, PHYS inputs per sample, PHYS outputs per sample
    void* peers[2][FSM_MAX_PEERS] = {};   // direct exchange: receive buffers of every rank (exchange 1, 2)
    int n_peers[2] = {0, 0};
    int pf = 0;                  // L2 prefetch switches (PF_* in fsm_passes.cuh)
    void* ks_log = nullptr;      // optional device log of the per-evaluation local KS zero-mode sums
    long ks_log_cap = 0;
    mutable long ks_log_pos = 0;
    int P = 1, rank = 0, kyl = 0, nxl = 0, nkz1 = 0;  // slab decomposition (P > 1): local ky / x extents, kept kz planes
    int gap_at = 0, gap = 0, nky1 = 0;                // cyclic ky ownership: dropped run of local lines, kept local lines
    std::vector<Stage> stages;
    Stage rhs_stage;
    // workspace (offsets in bytes)
    size_t off_arr[ARR_COUNT], off_w1, off_w2, off_w3, off_w2b, off_dc, ws_bytes;
    size_t cap_fields;  // how many independent fields the W buffers can hold at once (r2c/c2r)
    int64_t launches_per_step, algo_bytes_per_step;
    int64_t pass_units[4];  // algorithmic field-units moved per step by each pass class (IX, MID, PHYS, FX)
    int64_t pass_touched[4];  // bytes the passes of each class really read + write per step (kept modes only, tables excluded)
    // optional per-pass device timing
    bool profile = false;
#ifndef FSM_EMU
    std::vector<cudaEvent_t> ev_pool;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev_used;
    cudaStream_t lane_stream[FSM_MAX_LANES] = {};
    cudaEvent_t lane_fork = nullptr, lane_join[FSM_MAX_LANES] = {};
    int lane_device = -1;
#endif
};

namespace {

enum { PASS_IX = 0, PASS_MID = 1, PASS_PHYS = 2, PASS_FX = 3 };

#ifndef FSM_EMU
struct ProfScope {
    fsm_plan* p;
    int cls;
    cudaStream_t st;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    static cudaEvent_t get(fsm_plan* p) {
        if (!p->ev_pool.empty()) { cudaEvent_t e = p->ev_pool.back(); p->ev_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    ProfScope(const fsm_plan* cp, int c, cudaStream_t s) : p(const_cast<fsm_plan*>(cp)), cls(c), st(s) {
        if (p->profile) { e0 = get(p); e1 = get(p); cudaEventRecord(e0, st); }
    }
    ~ProfScope() {
        if (e0) { cudaEventRecord(e1, st); p->ev_used.push_back({cls, {e0, e1}}); }
    }
};
#else
struct ProfScope { ProfScope(const fsm_plan*, int, cudaStream_t) {} };
#endif

// error of the launch just issued (the pass launches check theirs in the launch layer)
int launch_status(const char* what) {
#ifndef FSM_EMU
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-EIO, "%s launch failed: %s", what, cudaGetErrorString(e));
#endif
    (void)what;
    return 0;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

template <typename T>
Geom<T> make_geom(const fsm_plan* p, bool nomask) {
    Geom<T> g;
    g.ndim = p->ndim;
    for (int i = 0; i < 3; ++i) {
        g.n[i] = p->n[i];
        g.kmax[i] = nomask ? p->n[i] / 2 : p->d.kmax[i];
        g.dk[i] = static_cast<const T*>(p->d.dk[i]);
        g.dkraw[i] = static_cast<const T*>(p->d.dkraw[i]);
    }
    g.ky0 = (p->P > 1) ? p->rank : 0;
    g.kys = (p->P > 1) ? p->P : 1;
    g.gap_at = (p->P > 1 && !nomask) ? p->gap_at : (1 << 30);
    g.gap = (p->P > 1 && !nomask) ? p->gap : 0;
    g.pf = p->pf;
    g.pf_wave = 0;
    g.nh = p->nh;
    g.ph = p->ph;
    g.nmodes = p->nmodes;
    g.inv_ntot = T(1.0 / (double)p->ntot);
    return g;
}

template <typename T>
const void* table_ptr(const fsm_plan* p, int tab) {
    switch (tab) {
        case TAB_EXP: return p->d.tab_exp;
        case TAB_HALF: return p->d.tab_half_exp;
        case TAB_LIN: return p->d.tab_lin;
        default: return (tab >= TAB_C1 && tab <= TAB_C6) ? p->d.tab_coef[tab - TAB_C1] : nullptr;
    }
}

// Translate a stage into the device-side Combine descriptor.
template <typename T>
int make_combine(const fsm_plan* p, const Stage& s, cplx<T>* const* arr, bool fresh, Combine<T>* out) {
    Combine<T> cb;
    memset(&cb, 0, sizeof(cb));
    cb.n_in = s.n_in;
    cb.n_out = s.n_out;
    cb.use_fresh = fresh ? 1 : 0;
    cb.tab_cstride = (p->d.tab_channels > 1) ? p->nmodes : 0;
    cb.tab_bstride = p->d.tab_batched ? (long)p->d.tab_channels * p->nmodes : 0;
    cb.tab_cplx = p->d.tab_complex ? 1 : 0;
    int slot_of[TAB_COUNT];
    for (int i = 0; i < TAB_COUNT; ++i) slot_of[i] = -1;
    for (int i = 0; i < s.n_in; ++i) cb.in[i] = arr[s.in[i]];
    for (int r = 0; r < s.n_out; ++r) cb.out[r] = arr[s.out[r]];
    cb.has_next = s.has_next ? 1 : 0;
    auto slot = [&](int tab) -> int {
        if (slot_of[tab] < 0) {
            const void* tp = table_ptr<T>(p, tab);
            if (!tp) return -100;
            if (cb.n_tab >= FSM_MAX_TAB) return -101;
            slot_of[tab] = cb.n_tab;
            cb.tab[cb.n_tab++] = static_cast<const T*>(tp);
        }
        return slot_of[tab];
    };
    for (int r = 0; r < FSM_MAX_OUT + 1; ++r)
        for (int m = 0; m < FSM_MAX_IN + 1; ++m) {
            const Coef& c = s.c[r][m];
            cb.ct[r][m] = NO_TERM;
            cb.ct2[r][m] = -1;
            const bool row_on = (r < FSM_MAX_OUT) ? (r < s.n_out) : s.has_next;
            if (!row_on || c.tab == NO_TERM) continue;
            if (m == 0 && !fresh) continue;
            if (m > s.n_in) return fail(-EINVAL, "stage coefficient refers to a missing input");
            cb.ca[r][m] = (T)c.a;
            cb.cb[r][m] = (T)c.b;
            cb.cb2[r][m] = (T)c.b2;
            if (c.tab == SCALAR) {
                cb.ct[r][m] = -1;
            } else {
                const int sl = slot(c.tab);
                if (sl == -100) return fail(-EINVAL, "integrator needs table %d but the descriptor has none", c.tab);
                if (sl == -101) return fail(-EINVAL, "too many tables in one stage");
                cb.ct[r][m] = sl;
            }
            if (c.tab2 >= 0) {
                cb.any_ct2 = 1;
                const int sl = slot(c.tab2);
                if (sl < 0) return fail(-EINVAL, "integrator needs table %d but the descriptor has none", c.tab2);
                cb.ct2[r][m] = sl;
            }
        }
    cb.kind = match_combine_shape(cb);
#ifdef FSM_EMU   // test seam of the CPU suite only (generic vs specialised combine); the CUDA build reads no environment
    if (getenv("FSM_GENERIC_COMBINE") != nullptr) cb.kind = -1;
#endif
    *out = cb;
    return 0;
}

template <typename T>
struct Buffers {
    cplx<T>* arr[ARR_COUNT];
    cplx<T>*w1, *w2, *w3, *w2b;
    T* dc;
};

template <typename T>
Buffers<T> carve(const fsm_plan* p, void* u_hat, void* ws, void* ext, int lane = 0) {
    Buffers<T> b;
    char* base = static_cast<char*>(ws);
    b.arr[ARR_U] = static_cast<cplx<T>*>(u_hat);
    for (int i = ARR_S1; i <= ARR_S4; ++i) b.arr[i] = reinterpret_cast<cplx<T>*>(base + p->off_arr[i]);
    b.arr[ARR_EXT] = static_cast<cplx<T>*>(ext);
    b.w1 = reinterpret_cast<cplx<T>*>(base + p->off_w1 + lane * p->lane_w_bytes[0]);
    b.w2 = reinterpret_cast<cplx<T>*>(base + p->off_w2 + lane * p->lane_w_bytes[1]);
    b.w3 = reinterpret_cast<cplx<T>*>(base + p->off_w3 + lane * p->lane_w_bytes[2]);
    b.w2b = reinterpret_cast<cplx<T>*>(base + p->off_w2b + lane * p->lane_w_bytes[3]);
    b.dc = reinterpret_cast<T*>(base + p->off_dc);
    return b;
}

MidSpec mid_spec_inverse(const fsm_plan* p) {
    MidSpec s;
    memset(&s, 0, sizeof(s));
    if (p->kprog == PROG_KS) {          // in: phi, dx phi -> phi, dx phi, dy phi
        s.nfo = 3;
        s.src[0] = 0; s.src[1] = 1; s.src[2] = 0; s.deriv[2] = 1;
    } else {                            // CONV / NS3D: in [c][u, dx u] -> u_c, dx u_c, dy u_c
        s.nfo = 9;
        for (int c = 0; c < 3; ++c) {
            s.src[c] = 2 * c;
            s.src[3 + c] = 2 * c + 1;
            s.src[6 + c] = 2 * c; s.deriv[6 + c] = 1;
        }
    }
    return s;
}
MidSpec mid_spec_identity(int nf) {
    MidSpec s;
    memset(&s, 0, sizeof(s));
    s.nfo = nf;
    for (int i = 0; i < nf; ++i) s.src[i] = i;
    return s;
}

// forward half of an evaluation: [MID forward] -> FX + combine, for `nb` samples starting at b0
template <typename T>
int run_forward_tail(const fsm_plan* p, const Buffers<T>& bf, const Geom<T>& g, int nfields_per_sample, int C,
                     const Combine<T>& cb, const FxEpilogue<T>& ep, int b0, int nb, cudaStream_t st) {
    const cplx<T>* fx_in = bf.w2;
    if (p->ndim == 3) {
        const LaunchTable<T>* ty = launch_table<T>(p->n[1]);
        MidArgs<T> m;
        m.g = g; m.in = bf.w2; m.out = bf.w2b;
        m.in_fstride = (long)p->nh * p->n[0] * p->n[1]; m.out_fstride = p->nmodes;
        m.in_t_stride = p->n[1]; m.in_o_stride = (long)p->n[0] * p->n[1];
        m.out_o_stride = p->n[0]; m.out_e_stride = (long)p->nh * p->n[0];
        m.nfi = nfields_per_sample; m.n_t = p->n[0]; m.n_outer = p->nh; m.nb = nb;
        m.spec = mid_spec_identity(nfields_per_sample);
        { ProfScope ps(p, PASS_MID, st); if (int e = ty->mid(-1, m, st)) return fail(e, "MID forward launch failed"); }
        fx_in = bf.w2b;
    }
    const LaunchTable<T>* tx = launch_table<T>(p->n[0]);
    FxArgs<T> f;
    f.g = g; f.win = fx_in; f.win_fstride = p->nmodes; f.cb = cb; f.ep = ep;
    f.nlines = (int)(p->nmodes / p->n[0]); f.b0 = b0; f.nb = nb;
    // independent channels (no projection, no per-channel table or source): run them as separate
    // single-channel fields -> no three-channel register tile, C times the parallelism
    if (C > 1 && !ep.project && !ep.source && !ep.dc_out && !ep.nl_coef_b && cb.tab_cstride == 0 && cb.tab_bstride == 0) {
        f.b0 = b0 * C; f.nb = nb * C;
        C = 1;
    }
    { ProfScope ps(p, PASS_FX, st); if (int e = tx->fx(C, f, st)) return fail(e, "FX launch failed (channels=%d, n0=%d)", C, p->n[0]); }
    return 0;
}

// NSPressureConvection(external_force) dealiases the integrator's own stage state in place before it evaluates
// (dedicated/_navier_stokes.py:241); every later use of that state sees the masked values, except in
// the one-stage schemes, whose `exp * u` is formed before the evaluation runs (_etdrk.py:47-51, _setdrk_step.py:5-11), and
// in the right-hand side (operator/_base.py:425-433: `linear_coef * u_fft` first).
template <typename T>
int mask_stage_input(const fsm_plan* p, const Geom<T>& g, const Stage& s, cplx<T>* arr, int b_lo, int b_hi, cudaStream_t st,
                     bool dyn_force = false) {
    if ((!p->d.force_hat && !dyn_force) || &s == &p->rhs_stage || p->stages.size() < 2) return 0;
    const long total = (long)(b_hi - b_lo) * p->C * p->nmodes;
    auto kern = k_mask_state<T>;
    FSM_LAUNCH(kern, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, g, arr + (long)b_lo * p->C * p->nmodes, total);
    return launch_status("stage-state dealiasing");
}

// One nonlinear evaluation of `stage_in` followed by the stage combine.
template <typename T>
int run_stage(const fsm_plan* p, const Buffers<T>& bf, const Stage& s, cudaStream_t st, int b_lo = 0, int b_hi = -1,
              const cplx<T>* ext_fresh = nullptr) {
    if (b_hi < 0) b_hi = p->B;
    Combine<T> cb;
    const bool fused = p->prog != FSM_PROG_LINEAR;
    // no convective term: the "nonlinear" term is the constant source (if any) or one evaluated outside (ext_fresh)
    const bool fresh = fused || ext_fresh != nullptr || p->d.source_hat != nullptr;
    if (int e = make_combine<T>(p, s, bf.arr, fresh, &cb)) return e;
    if (!fused) {
        const long total = (long)p->B * p->C * p->nmodes;
        auto kern = k_combine_only<T>;
        FSM_LAUNCH(kern, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, cb, p->nmodes, p->C, total, ext_fresh,
                   static_cast<const cplx<T>*>(p->d.source_hat));
        return launch_status("linear combine");
    }
    const Geom<T> g = make_geom<T>(p, false);
    const cplx<T>* stage_in = bf.arr[s.input];
    if (int e = mask_stage_input<T>(p, g, s, bf.arr[s.input], b_lo, b_hi, st, ext_fresh != nullptr)) return e;
    const LaunchTable<T>* tx = launch_table<T>(p->n[0]);
    const LaunchTable<T>* tl = launch_table<T>(p->n[p->ndim - 1]);
    const LaunchTable<T>* ty = (p->ndim == 3) ? launch_table<T>(p->n[1]) : nullptr;
    FxEpilogue<T> ep;
    ep.nl_coef = (T)p->d.nl_coef;
    ep.nl_coef_b = static_cast<const T*>(p->d.nl_coef_b);
    ep.source = static_cast<const cplx<T>*>(p->d.source_hat);
    ep.dc_out = (p->prog == FSM_PROG_KS && p->d.ks_remove_mean) ? bf.dc : nullptr;
    ep.project = (p->prog == FSM_PROG_NS3D) ? 1 : 0;
    ep.force = static_cast<const cplx<T>*>(p->d.force_hat);
    ep.force_dyn = ext_fresh;     // fused programs: the externally evaluated array is the state-dependent force (fsm_stage_run)
    for (int b0 = b_lo; b0 < b_hi; b0 += p->chunk) {
        const int nb = (b_hi - b0 < p->chunk) ? (b_hi - b0) : p->chunk;
        IxArgs<T> a;
        a.g = g; a.state = stage_in + (long)b0 * p->C * p->nmodes; a.w1 = bf.w1;
        a.state_bstride = p->nmodes; a.nbc = nb * p->C;
        PhysArgs<T> ph;
        ph.g = g; ph.wout = bf.w2; ph.phys_in = nullptr; ph.phys_out = nullptr; ph.nb = nb;
        if (p->ndim == 2) {
            const bool zlines = (p->kprog == PROG_NS2D || p->kprog == PROG_KS2D);
            const int w1_pitch = zlines ? p->n[1] : p->ph;   // Z-line programs: full complex rows
            a.w1_fstride = (long)p->n[0] * w1_pitch;
            a.in_t_stride = p->n[0]; a.in_o_stride = 0; a.out_o_stride = 0; a.out_e_stride = w1_pitch;
            a.n_t = (p->d.kmax[1] + 1 < p->nh) ? p->d.kmax[1] + 1 : p->nh; a.n_outer = 1;
            { ProfScope ps(p, PASS_IX, st); if (int e = tx->ix(p->kprog, a, st)) return fail(e, "IX launch failed"); }
            ph.win = bf.w1; ph.win_fstride = a.w1_fstride; ph.wout_fstride = p->nmodes;
            ph.in_t_stride = w1_pitch; ph.in_o_stride = 0; ph.out_o_stride = 0; ph.out_e_stride = p->n[0];
            ph.n_t = p->n[0]; ph.n_outer = 1;
        } else {
            const long plane = (long)p->n[0] * p->n[1];
            a.w1_fstride = (long)p->nh * plane;
            a.in_t_stride = (long)p->nh * p->n[0]; a.in_o_stride = p->n[0];
            a.out_o_stride = plane; a.out_e_stride = p->n[1];
            a.n_t = p->n[1]; a.n_outer = (p->d.kmax[2] + 1 < p->nh) ? p->d.kmax[2] + 1 : p->nh;
            { ProfScope ps(p, PASS_IX, st); if (int e = tx->ix(p->kprog, a, st)) return fail(e, "IX launch failed"); }
            MidArgs<T> m;
            m.g = g; m.in = bf.w1; m.out = bf.w3;
            m.in_fstride = a.w1_fstride; m.out_fstride = plane * p->ph;
            m.in_t_stride = plane; m.in_o_stride = p->n[1];
            m.out_o_stride = (long)p->n[1] * p->ph; m.out_e_stride = p->ph;
            m.nfi = p->C * p->nf_ix; m.n_t = a.n_outer; m.n_outer = p->n[0]; m.nb = nb;
            m.spec = mid_spec_inverse(p);
            { ProfScope ps(p, PASS_MID, st); if (int e = ty->mid(+1, m, st)) return fail(e, "MID inverse launch failed"); }
            ph.win = bf.w3; ph.win_fstride = m.out_fstride; ph.wout_fstride = (long)p->nh * plane;
            ph.in_t_stride = p->ph; ph.in_o_stride = (long)p->n[1] * p->ph;
            ph.out_o_stride = p->n[1]; ph.out_e_stride = plane;
            ph.n_t = p->n[1]; ph.n_outer = p->n[0];
        }
        { ProfScope ps(p, PASS_PHYS, st); if (int e = tl->phys(p->kprog, p->ndim, ph, st)) return fail(e, "PHYS launch failed"); }
        if (int e = run_forward_tail<T>(p, bf, g, p->nout, p->C, cb, ep, b0, nb, st)) return e;
    }
    if (ep.dc_out) {
        auto kern = k_ks_dc_fix<T>;
        T* slot = nullptr;
        if (p->ks_log && p->ks_log_pos < p->ks_log_cap) slot = static_cast<T*>(p->ks_log) + p->ks_log_pos++;
        FSM_LAUNCH(kern, dim3(1), dim3(128), sizeof(T) * 2, st, cb, (const T*)bf.dc, p->B, p->nmodes, slot);
        if (int e = launch_status("KS zero-mode fix")) return e;
    }
    return 0;
}

// 1-D: every stage of every step inside one launch (one CTA per sample)
template <typename T>
int run_1d(const fsm_plan* p, const Buffers<T>& bf, const Stage* stages, int n_stages, int n_steps, cudaStream_t st) {
    Step1dArgs<T> a;
    memset(&a.sl, 0, sizeof(a.sl));
    a.g = make_geom<T>(p, false);
    a.sl.n_stages = n_stages;
    for (int i = 0; i < n_stages; ++i) {
        a.sl.input[i] = bf.arr[stages[i].input];
        if (int e = make_combine<T>(p, stages[i], bf.arr, true, &a.sl.cb[i])) return e;
    }
    a.ep.nl_coef = (T)p->d.nl_coef;
    a.ep.nl_coef_b = static_cast<const T*>(p->d.nl_coef_b);
    a.ep.source = static_cast<const cplx<T>*>(p->d.source_hat);
    a.ep.dc_out = nullptr;
    a.ep.project = 0;
    a.ep.force = nullptr;
    a.ep.force_dyn = nullptr;
    a.n_steps = n_steps;
    a.nb = p->B;
    const LaunchTable<T>* tx = launch_table<T>(p->n[0]);
    if (int e = tx->step1d(a, st)) return fail(e, "1-D step launch failed");
    return 0;
}

template <typename T>
int do_step(fsm_plan* p, void* u_hat, void* ws, int n_steps, cudaStream_t st) {
    Buffers<T> bf = carve<T>(p, u_hat, ws, nullptr);
    if (p->ndim == 1 && p->prog != FSM_PROG_LINEAR) {
        if (p->stages.size() > FSM_MAX_STAGES) return fail(-ENOSYS, "too many stages for the 1-D kernel");
        return run_1d<T>(p, bf, p->stages.data(), (int)p->stages.size(), n_steps, st);
    }
#ifndef FSM_EMU
    // Samples are independent (every program but KS with its batch mean): the batch is split into `nlanes` ranges,
    // each advanced through ALL steps on its own internal stream with its own intermediates. The kernels of one lane
    // fill the wave tails of the other, and passes bound by different resources (FX: HBM, last-axis pass: shared
    // memory, inverse-x: latency) share the SMs instead of following each other. Fork/join with events on the
    // caller's stream; per-pass profiling keeps the single-stream order so that its timings do not overlap.
    if (p->nlanes > 1 && !p->profile && n_steps > 0 && p->prog != FSM_PROG_LINEAR) {
        int dev = -1;
        cudaGetDevice(&dev);
        if (p->lane_device != dev) {       // first use (or another device): create the streams and events here
            for (int l = 0; l < p->nlanes; ++l) {
                if (cudaStreamCreateWithFlags(&p->lane_stream[l], cudaStreamNonBlocking) != cudaSuccess ||
                    cudaEventCreateWithFlags(&p->lane_join[l], cudaEventDisableTiming) != cudaSuccess)
                    return fail(-EIO, "could not create the lane streams");
            }
            if (cudaEventCreateWithFlags(&p->lane_fork, cudaEventDisableTiming) != cudaSuccess) return fail(-EIO, "event");
            p->lane_device = dev;
        }
        const int per = (p->B + p->nlanes - 1) / p->nlanes;
        cudaEventRecord(p->lane_fork, st);
        for (int l = 0; l < p->nlanes; ++l) {
            const int lo = l * per, hi = (lo + per < p->B) ? lo + per : p->B;
            if (lo >= hi) continue;
            cudaStream_t ls = p->lane_stream[l];
            cudaStreamWaitEvent(ls, p->lane_fork, 0);
            Buffers<T> lb = carve<T>(p, u_hat, ws, nullptr, l);
            for (int i = 0; i < n_steps; ++i)
                for (const Stage& s : p->stages)
                    if (int e = run_stage<T>(p, lb, s, ls, lo, hi)) return e;
            cudaEventRecord(p->lane_join[l], ls);
            cudaStreamWaitEvent(st, p->lane_join[l], 0);
        }
        return 0;
    }
#endif
    for (int i = 0; i < n_steps; ++i)
        for (const Stage& s : p->stages)
            if (int e = run_stage<T>(p, bf, s, st)) return e;
    return 0;
}

template <typename T>
int do_rhs(fsm_plan* p, const void* u_hat, void* out, void* ws, cudaStream_t st) {
    Buffers<T> bf = carve<T>(p, const_cast<void*>(u_hat), ws, out);
    if (p->ndim == 1 && p->prog != FSM_PROG_LINEAR) return run_1d<T>(p, bf, &p->rhs_stage, 1, 1, st);
    return run_stage<T>(p, bf, p->rhs_stage, st);
}

template <typename T>
int do_r2c(fsm_plan* p, const void* u, void* u_hat, void* ws, cudaStream_t st) {
    if (p->ndim == 1) {
        if (int e = launch_table<T>(p->n[0])->line1d(MODE1D_R2C, u, u_hat, (long)p->B * p->C, st))
            return fail(e, "1-D R2C launch failed");
        return 0;
    }
    Buffers<T> bf = carve<T>(p, u_hat, ws, nullptr);
    const Geom<T> g = make_geom<T>(p, true);
    const LaunchTable<T>* tl = launch_table<T>(p->n[p->ndim - 1]);
    const long nfields = (long)p->B * p->C;
    const long nreal = p->ntot;
    Stage s;  // out = fresh
    s.input = ARR_U; s.n_in = 0; s.n_out = 1; s.out[0] = ARR_U;
    s.c[0][0] = scal(1.0);
    for (long f0 = 0; f0 < nfields; f0 += (long)p->cap_fields) {
        const int nf = (int)((nfields - f0 < (long)p->cap_fields) ? nfields - f0 : (long)p->cap_fields);
        PhysArgs<T> ph;
        ph.g = g; ph.win = nullptr; ph.wout = bf.w2; ph.phys_in = static_cast<const T*>(u) + f0 * nreal;
        ph.phys_out = nullptr; ph.win_fstride = 0; ph.nb = nf;
        if (p->ndim == 2) {
            ph.wout_fstride = p->nmodes; ph.in_t_stride = 0; ph.in_o_stride = 0; ph.out_o_stride = 0;
            ph.out_e_stride = p->n[0]; ph.n_t = p->n[0]; ph.n_outer = 1;
        } else {
            const long plane = (long)p->n[0] * p->n[1];
            ph.wout_fstride = (long)p->nh * plane; ph.in_t_stride = 0; ph.in_o_stride = 0;
            ph.out_o_stride = p->n[1]; ph.out_e_stride = plane; ph.n_t = p->n[1]; ph.n_outer = p->n[0];
        }
        if (int e = tl->phys(PROG_R2C, p->ndim, ph, st)) return fail(e, "R2C PHYS launch failed");
        // every field is an independent single-channel "sample" for the tail
        cplx<T>* arr[ARR_COUNT] = {bf.arr[ARR_U] + f0 * p->nmodes, nullptr, nullptr, nullptr, nullptr, nullptr};
        Combine<T> cb;
        fsm_plan tmp = *p;  // tables unused; channel stride irrelevant
        tmp.d.tab_channels = 1;
        if (int e = make_combine<T>(&tmp, s, arr, true, &cb)) return e;
        FxEpilogue<T> ep;
        ep.nl_coef = T(1); ep.nl_coef_b = nullptr; ep.source = nullptr; ep.dc_out = nullptr; ep.project = 0; ep.force = nullptr; ep.force_dyn = nullptr;
        if (int e = run_forward_tail<T>(p, bf, g, 1, 1, cb, ep, 0, nf, st)) return e;
    }
    return 0;
}

template <typename T>
int do_c2r(fsm_plan* p, const void* u_hat, void* u, void* ws, cudaStream_t st) {
    if (p->ndim == 1) {
        if (int e = launch_table<T>(p->n[0])->line1d(MODE1D_C2R, u_hat, u, (long)p->B * p->C, st))
            return fail(e, "1-D C2R launch failed");
        return 0;
    }
    Buffers<T> bf = carve<T>(p, const_cast<void*>(u_hat), ws, nullptr);
    const Geom<T> g = make_geom<T>(p, true);
    const LaunchTable<T>* tx = launch_table<T>(p->n[0]);
    const LaunchTable<T>* tl = launch_table<T>(p->n[p->ndim - 1]);
    const long nfields = (long)p->B * p->C;
    for (long f0 = 0; f0 < nfields; f0 += (long)p->cap_fields) {
        const int nf = (int)((nfields - f0 < (long)p->cap_fields) ? nfields - f0 : (long)p->cap_fields);
        IxArgs<T> a;
        a.g = g; a.state = static_cast<const cplx<T>*>(u_hat) + f0 * p->nmodes; a.w1 = bf.w1;
        a.state_bstride = p->nmodes; a.nbc = nf;
        PhysArgs<T> ph;
        ph.g = g; ph.wout = nullptr; ph.phys_in = nullptr; ph.phys_out = static_cast<T*>(u) + f0 * p->ntot;
        ph.wout_fstride = 0; ph.out_o_stride = 0; ph.out_e_stride = 0; ph.nb = nf;
        if (p->ndim == 2) {
            a.w1_fstride = (long)p->n[0] * p->ph;
            a.in_t_stride = p->n[0]; a.in_o_stride = 0; a.out_o_stride = 0; a.out_e_stride = p->ph;
            a.n_t = p->nh; a.n_outer = 1;
            if (int e = tx->ix(PROG_C2R, a, st)) return fail(e, "C2R IX launch failed");
            ph.win = bf.w1; ph.win_fstride = a.w1_fstride; ph.in_t_stride = p->ph; ph.in_o_stride = 0;
            ph.n_t = p->n[0]; ph.n_outer = 1;
        } else {
            const long plane = (long)p->n[0] * p->n[1];
            const LaunchTable<T>* ty = launch_table<T>(p->n[1]);
            a.w1_fstride = (long)p->nh * plane;
            a.in_t_stride = (long)p->nh * p->n[0]; a.in_o_stride = p->n[0];
            a.out_o_stride = plane; a.out_e_stride = p->n[1];
            a.n_t = p->n[1]; a.n_outer = p->nh;
            if (int e = tx->ix(PROG_C2R, a, st)) return fail(e, "C2R IX launch failed");
            MidArgs<T> m;
            m.g = g; m.in = bf.w1; m.out = bf.w3;
            m.in_fstride = a.w1_fstride; m.out_fstride = plane * p->ph;
            m.in_t_stride = plane; m.in_o_stride = p->n[1];
            m.out_o_stride = (long)p->n[1] * p->ph; m.out_e_stride = p->ph;
            m.nfi = 1; m.n_t = p->nh; m.n_outer = p->n[0]; m.nb = nf;
            m.spec = mid_spec_identity(1);
            if (int e = ty->mid(+1, m, st)) return fail(e, "C2R MID launch failed");
            ph.win = bf.w3; ph.win_fstride = m.out_fstride;
            ph.in_t_stride = p->ph; ph.in_o_stride = (long)p->n[1] * p->ph;
            ph.n_t = p->n[1]; ph.n_outer = p->n[0];
        }
        if (int e = tl->phys(PROG_C2R, p->ndim, ph, st)) return fail(e, "C2R PHYS launch failed");
    }
    return 0;
}

int ilog2(int x) { int s = 0; while ((1 << s) < x) ++s; return s; }

// ---------------------------------------------------------------------------------------------
// Slab decomposition: local passes of one phase (see include/fsm_b200.h, fsm_slab_phase)
//   exchange 1 (inverse side)  [dst rank q][field][kz < nkz][x_local][ky_local]
//   exchange 2 (forward side)  [dst rank q][field][ky_local][kz < nh][x_local]
// ---------------------------------------------------------------------------------------------
// `nsub` sub-slabs split the local x range so the exchange of one sub-slab can overlap the local chain of
// another: exchange buffers are laid out [sub-slab h][rank q][field][...], each h a contiguous all-to-all.
template <typename T>
int slab_ix(const fsm_plan* p, const Geom<T>& g, int kprog, const cplx<T>* state, cplx<T>* send, int nfields_in, int nf,
            int nkz, int nsub, cudaStream_t st) {
    const LaunchTable<T>* tx = launch_table<T>(p->n[0]);
    const int nxh = p->nxl / nsub;
    const int nky = p->kyl - g.gap;     // local lines shipped: the kept ones (everything for the plain transforms)
    IxArgs<T> a;
    a.g = g; a.state = state; a.w1 = send; a.state_bstride = p->nmodes; a.nbc = nfields_in;
    a.w1_fstride = (long)nkz * nxh * nky;
    a.in_t_stride = (long)p->nh * p->n[0]; a.in_o_stride = p->n[0];
    a.out_o_stride = (long)nxh * nky; a.out_e_stride = nky;
    a.n_t = nky; a.n_outer = nkz;
    a.eb.shift = ilog2(p->nxl); a.eb.stride = (long)nfields_in * nf * a.w1_fstride;
    a.eb.shift2 = ilog2(nxh); a.eb.stride2 = (long)p->P * a.eb.stride;
    if (p->n_peers[0] > 0) {
        if (nsub != 1) return fail(-EINVAL, "the direct exchange runs with one sub-slab");
        a.pe.n = p->n_peers[0];
        for (int r = 0; r < a.pe.n; ++r) a.pe.base[r] = p->peers[0][r];
        a.pe.self_off = (long)p->rank * a.eb.stride;
    }
    ProfScope ps(p, PASS_IX, st);
    if (int e = tx->ix(kprog, a, st)) return fail(e, "slab IX launch failed");
    return 0;
}

template <typename T>
int slab_mid_inverse(const fsm_plan* p, const Geom<T>& g, const cplx<T>* recv, cplx<T>* w3, int nfi_in, const MidSpec& spec,
                     int nkz, int nb, int sub, int nsub, cudaStream_t st) {
    const LaunchTable<T>* ty = launch_table<T>(p->n[1]);
    const int nxh = p->nxl / nsub;
    const int nky = p->kyl - g.gap;
    MidArgs<T> m;
    m.g = g;
    m.in_fstride = (long)nkz * nxh * nky; m.out_fstride = (long)p->nxl * p->n[1] * p->ph;
    m.in_t_stride = (long)nxh * nky; m.in_o_stride = nky;
    m.out_o_stride = (long)p->n[1] * p->ph; m.out_e_stride = p->ph;
    m.nfi = nfi_in; m.n_t = nkz; m.n_outer = nxh; m.nb = nb; m.spec = spec;
    m.ib.shift = 0; m.ib.cyc = ilog2(p->P); m.ib.gap_at = g.gap_at; m.ib.gap = g.gap;   // ky: cyclic owners, compact lines
    m.ib.stride = (long)nb * nfi_in * m.in_fstride;
    m.in = recv + (long)sub * p->P * m.ib.stride;
    m.out = w3 + (long)sub * nxh * m.out_o_stride;
    ProfScope ps(p, PASS_MID, st);
    if (int e = ty->mid(+1, m, st)) return fail(e, "slab MID inverse launch failed");
    return 0;
}

template <typename T>
int slab_mid_forward(const fsm_plan* p, const Geom<T>& g, const cplx<T>* w2a, cplx<T>* send, int nf, int nb, int sub, int nsub,
                     cudaStream_t st) {
    const LaunchTable<T>* ty = launch_table<T>(p->n[1]);
    const int nxh = p->nxl / nsub;
    MidArgs<T> m;
    m.g = g;
    m.in_fstride = (long)p->nh * p->nxl * p->n[1]; m.out_fstride = (long)p->kyl * p->nh * nxh;
    m.in_t_stride = p->n[1]; m.in_o_stride = (long)p->nxl * p->n[1];
    m.out_o_stride = nxh; m.out_e_stride = (long)p->nh * nxh;
    m.nfi = nf; m.n_t = nxh; m.n_outer = p->nh; m.nb = nb; m.spec = mid_spec_identity(nf);
    m.eb.shift = 0; m.eb.cyc = ilog2(p->P); m.eb.gap_at = 1 << 30; m.eb.gap = 0;    // ky: cyclic owners, every line
    m.eb.stride = (long)nb * nf * m.out_fstride;
    m.in = w2a + (long)sub * nxh * m.in_t_stride;
    m.out = send + (long)sub * p->P * m.eb.stride;
    if (p->n_peers[1] > 0) {
        if (nsub != 1) return fail(-EINVAL, "the direct exchange runs with one sub-slab");
        m.pe.n = p->n_peers[1];
        for (int r = 0; r < m.pe.n; ++r) m.pe.base[r] = p->peers[1][r];
        m.pe.self_off = (long)p->rank * m.eb.stride;
    }
    ProfScope ps(p, PASS_MID, st);
    if (int e = ty->mid(-1, m, st)) return fail(e, "slab MID forward launch failed");
    return 0;
}

template <typename T>
int slab_phys(const fsm_plan* p, const Geom<T>& g, int kprog, const cplx<T>* w3, cplx<T>* w2a, const T* phys_in, T* phys_out,
              int nb, int sub, int nsub, cudaStream_t st) {
    const LaunchTable<T>* tl = launch_table<T>(p->n[2]);
    const int nxh = p->nxl / nsub;
    PhysArgs<T> ph;
    ph.g = g; ph.phys_in = phys_in; ph.phys_out = phys_out; ph.nb = nb;
    ph.win_fstride = (long)p->nxl * p->n[1] * p->ph; ph.wout_fstride = (long)p->nh * p->nxl * p->n[1];
    ph.in_t_stride = p->ph; ph.in_o_stride = (long)p->n[1] * p->ph;
    ph.out_o_stride = p->n[1]; ph.out_e_stride = (long)p->nxl * p->n[1];
    ph.n_t = p->n[1]; ph.n_outer = nxh;
    ph.win = w3 ? w3 + (long)sub * nxh * ph.in_o_stride : nullptr;
    ph.wout = w2a ? w2a + (long)sub * nxh * ph.out_o_stride : nullptr;
    ProfScope ps(p, PASS_PHYS, st);
    if (int e = tl->phys(kprog, 3, ph, st)) return fail(e, "slab PHYS launch failed");
    return 0;
}

template <typename T>
int slab_fx(const fsm_plan* p, const Geom<T>& g, const cplx<T>* recv, int C, int nb, const Combine<T>& cb, const FxEpilogue<T>& ep,
            int nsub, cudaStream_t st) {
    const LaunchTable<T>* tx = launch_table<T>(p->n[0]);
    const int nxh = p->nxl / nsub;
    FxArgs<T> f;
    f.g = g; f.win = recv; f.win_fstride = (long)p->kyl * p->nh * nxh; f.cb = cb; f.ep = ep;
    f.nlines = p->kyl * p->nh; f.b0 = 0; f.nb = nb;
    f.ib.shift = ilog2(p->nxl); f.ib.stride = (long)nb * C * f.win_fstride;
    f.ib.shift2 = ilog2(nxh); f.ib.stride2 = (long)p->P * f.ib.stride;
    f.line_stride = nxh;
    ProfScope ps(p, PASS_FX, st);
    if (int e = tx->fx(C, f, st)) return fail(e, "slab FX launch failed");
    return 0;
}

template <typename T>
int do_slab_phase(fsm_plan* p, int op, int stage, int phase, int sub, int nsub, void* u_hat, void* aux, void* ws, void* send,
                  void* recv, cudaStream_t st) {
    Buffers<T> bf = carve<T>(p, u_hat, ws, (op == FSM_SLAB_RHS) ? aux : nullptr);
    cplx<T>* snd = static_cast<cplx<T>*>(send);
    const cplx<T>* rcv = static_cast<const cplx<T>*>(recv);
    if (op == FSM_SLAB_STEP || op == FSM_SLAB_RHS) {
        if (p->prog == FSM_PROG_LINEAR) return fail(-EINVAL, "linear operators have no slab phases; call fsm_step");
        if (op == FSM_SLAB_STEP && p->d.force_hat && p->d.integrator == FSM_INT_RK4)
            return fail(-ENOSYS, "NS pressure convection with an external force is not supported with the RK integrators");
        const Stage& s = (op == FSM_SLAB_RHS) ? p->rhs_stage : p->stages[stage];
        const Geom<T> g = make_geom<T>(p, false);
        if (phase == 0) {
            if (int e = mask_stage_input<T>(p, g, s, bf.arr[s.input], 0, p->B, st)) return e;
            return slab_ix<T>(p, g, p->kprog, bf.arr[s.input], snd, p->B * p->C, p->nf_ix, p->nkz1, nsub, st);
        }
        if (phase == 1) {
            if (int e = slab_mid_inverse<T>(p, g, rcv, bf.w3, p->C * p->nf_ix, mid_spec_inverse(p), p->nkz1, p->B, sub, nsub, st)) return e;
            if (int e = slab_phys<T>(p, g, p->kprog, bf.w3, bf.w2, nullptr, nullptr, p->B, sub, nsub, st)) return e;
            return slab_mid_forward<T>(p, g, bf.w2, snd, p->nout, p->B, sub, nsub, st);
        }
        Combine<T> cb;
        if (int e = make_combine<T>(p, s, bf.arr, true, &cb)) return e;
        FxEpilogue<T> ep;
        ep.nl_coef = (T)p->d.nl_coef;
        ep.nl_coef_b = static_cast<const T*>(p->d.nl_coef_b);
        ep.source = static_cast<const cplx<T>*>(p->d.source_hat);
        ep.dc_out = nullptr;
        ep.project = (p->prog == FSM_PROG_NS3D) ? 1 : 0;
        ep.force = static_cast<const cplx<T>*>(p->d.force_hat);
        ep.force_dyn = nullptr;
        return slab_fx<T>(p, g, rcv, p->C, p->B, cb, ep, nsub, st);
    }
    if (nsub != 1) return fail(-EINVAL, "the plain transforms run with one sub-slab");
    const Geom<T> g = make_geom<T>(p, true);
    const int nf = p->B * p->C;
    if (op == FSM_SLAB_R2C) {
        if (phase == 1) {
            if (int e = slab_phys<T>(p, g, PROG_R2C, nullptr, bf.w2, static_cast<const T*>(aux), nullptr, nf, 0, 1, st)) return e;
            return slab_mid_forward<T>(p, g, bf.w2, snd, 1, nf, 0, 1, st);
        }
        Stage s;
        s.input = ARR_U; s.n_in = 0; s.n_out = 1; s.out[0] = ARR_U;
        s.c[0][0] = scal(1.0);
        Combine<T> cb;
        fsm_plan tmp = *p;
        tmp.d.tab_channels = 1;
        if (int e = make_combine<T>(&tmp, s, bf.arr, true, &cb)) return e;
        FxEpilogue<T> ep;
        ep.nl_coef = T(1); ep.nl_coef_b = nullptr; ep.source = nullptr; ep.dc_out = nullptr; ep.project = 0; ep.force = nullptr; ep.force_dyn = nullptr;
        return slab_fx<T>(p, g, rcv, 1, nf, cb, ep, 1, st);
    }
    if (op == FSM_SLAB_C2R) {
        if (phase == 0) return slab_ix<T>(p, g, PROG_C2R, static_cast<const cplx<T>*>(u_hat), snd, nf, 1, p->nh, 1, st);
        if (int e = slab_mid_inverse<T>(p, g, rcv, bf.w3, 1, mid_spec_identity(1), p->nh, nf, 0, 1, st)) return e;
        return slab_phys<T>(p, g, PROG_C2R, bf.w3, nullptr, nullptr, static_cast<T*>(aux), nf, 0, 1, st);
    }
    return fail(-EINVAL, "unknown slab op %d", op);
}

template <typename T>
MapTerms<T> make_map_terms(const fsm_map_term* terms, int n_terms, int c_in, int c_out, int dealias) {
    MapTerms<T> mt;
    memset(&mt, 0, sizeof(mt));
    mt.n_terms = n_terms; mt.c_in = c_in; mt.c_out = c_out; mt.dealias = dealias ? 1 : 0;
    for (int t = 0; t < n_terms; ++t) {
        mt.out_ch[t] = (signed char)terms[t].out_channel;
        mt.in_ch[t] = (signed char)terms[t].in_channel;
        for (int a = 0; a < 3; ++a) mt.pw[t][a] = (signed char)terms[t].power[a];
        mt.inv_lap[t] = (signed char)terms[t].inv_laplacian;
        mt.coef[t] = (T)terms[t].coef;
    }
    return mt;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int fsm_abi_version(void) { return FSM_ABI_VERSION; }
int fsm_backend(void) {
#ifdef FSM_EMU
    return 1;
#else
    return 0;
#endif
}
const char* fsm_last_error(void) { return g_err; }

int fsm_plan_create(fsm_plan** out, const fsm_desc* d) {
    if (!out || !d) return fail(-EINVAL, "null argument");
    if (d->struct_size != (int32_t)sizeof(fsm_desc))
        return fail(-EINVAL, "fsm_desc size mismatch: caller %d, library %d", d->struct_size, (int)sizeof(fsm_desc));
    if (d->ndim < 1 || d->ndim > 3) return fail(-EINVAL, "ndim=%d: grids must be 1-D, 2-D or 3-D", d->ndim);
    if (d->dtype != FSM_F32 && d->dtype != FSM_F64) return fail(-EINVAL, "bad dtype %d", d->dtype);
    if (d->batch < 1 || d->channels < 1) return fail(-EINVAL, "batch and channels must be positive");
    fsm_plan* p = new fsm_plan();
    p->d = *d;
    p->ndim = d->ndim;
    p->f64 = d->dtype == FSM_F64;
    p->B = d->batch;
    p->C = d->channels;
    p->prog = d->program;
    p->ntot = 1;
    for (int i = 0; i < 3; ++i) {
        p->n[i] = (i < d->ndim) ? d->n[i] : 1;
        p->ntot *= p->n[i];
        if (i < d->ndim) {
            const bool ok = p->f64 ? (launch_table<double>(p->n[i]) != nullptr) : (launch_table<float>(p->n[i]) != nullptr);
            if (!is_pow2(p->n[i]) || !ok) {
                delete p;
                return fail(-ENOSYS, "axis %d has %d points: only powers of two in [8, 1024] are supported", i, d->n[i]);
            }
            if (!d->dk[i] || !d->dkraw[i]) { delete p; return fail(-EINVAL, "missing wavenumber table for axis %d", i); }
            if (d->kmax[i] < 0 || d->kmax[i] > p->n[i] / 2) { delete p; return fail(-EINVAL, "bad kmax[%d]=%d", i, d->kmax[i]); }
        }
    }
    const int nlast = p->n[p->ndim - 1];
    p->nh = nlast / 2 + 1;
    p->ph = (p->nh + 7) / 8 * 8;
    p->nmodes = (long)p->nh * (p->ntot / nlast);
    if (d->slab_nranks > 1) {
        const int P = d->slab_nranks;
        if (p->ndim != 3) { delete p; return fail(-EINVAL, "slab decomposition needs a 3-D grid"); }
        if (!is_pow2(P) || p->n[0] % P || p->n[1] % P || d->slab_rank < 0 || d->slab_rank >= P) {
            delete p;
            return fail(-EINVAL, "slab decomposition: %d ranks must be a power of two dividing n0=%d and n1=%d", P, p->n[0], p->n[1]);
        }
        p->P = P; p->rank = d->slab_rank; p->kyl = p->n[1] / P; p->nxl = p->n[0] / P;
        p->nkz1 = (d->kmax[2] + 1 < p->nh) ? d->kmax[2] + 1 : p->nh;
        // cyclic ky ownership (rank r holds ky = r + P t): the same compact set of local lines covers the kept band
        // |ky| <= kmax on every rank: t < Llo (largest count, rank 0) and t >= kyl - Lhi (largest count, rank P-1)
        {
            const int km = d->kmax[1], n1 = p->n[1];
            int llo = km / P + 1;
            int first_hi = (n1 - km - (P - 1) + P - 1) / P;      // ceil((n1 - km - (P-1)) / P)
            if (first_hi < 0) first_hi = 0;
            int lhi = p->kyl - first_hi;
            if (llo + lhi >= p->kyl || p->prog == FSM_PROG_LINEAR) { llo = p->kyl; lhi = 0; }
            p->gap_at = llo;
            p->gap = p->kyl - llo - lhi;
            p->nky1 = p->kyl - p->gap;
        }
        p->nmodes = (long)p->kyl * p->nh * p->n[0];   // local spectral slab
    }
    // program -> kernel program and field counts
    switch (p->prog) {
        case FSM_PROG_LINEAR: p->kprog = PROG_NONE; p->nf_ix = 1; p->nfi = 1; p->nout = 1; break;
        case FSM_PROG_CONVECTION:
            if (p->C != p->ndim) { delete p; return fail(-EINVAL, "convection needs channels == ndim"); }
            p->kprog = PROG_CONV; p->nf_ix = 2; p->nfi = (p->ndim == 2) ? 4 : 9; p->nout = p->C; break;
        case FSM_PROG_KS:
            if (p->C != 1) { delete p; return fail(-EINVAL, "KS convection needs one channel"); }
            if (p->ndim == 1) { delete p; return fail(-ENOSYS, "KS convection on 1-D grids is not supported by the fused CUDA path"); }
            p->kprog = (p->ndim == 2) ? PROG_KS2D : PROG_KS; p->nf_ix = 2; p->nfi = (p->ndim == 2) ? 2 : 3; p->nout = 1; break;
        case FSM_PROG_NS2D_VORT:
            if (p->C != 1 || p->ndim != 2) { delete p; return fail(-EINVAL, "vorticity convection needs a 2-D scalar field"); }
            p->kprog = PROG_NS2D; p->nf_ix = 4; p->nfi = 4; p->nout = 1; break;
        case FSM_PROG_NS3D:   // velocity form with the pressure projected out: the convection program + projection in FX
            if (p->C != p->ndim || p->ndim < 2) { delete p; return fail(-EINVAL, "NS pressure convection needs a 2-D or 3-D field with channels == ndim"); }
            p->kprog = (p->ndim == 3) ? PROG_NS3D : PROG_CONV; p->nf_ix = 2; p->nfi = (p->ndim == 2) ? 4 : 9; p->nout = p->C; break;
        default: delete p; return fail(-ENOSYS, "unknown program %d", p->prog);
    }
    if (p->prog == FSM_PROG_LINEAR && d->integrator != FSM_INT_ETDRK0 && d->integrator != FSM_INT_RK4) {
        // a purely linear operator under an ETD scheme: the nonlinear term is identically zero
    }
    if (p->P > 1 && p->prog != FSM_PROG_CONVECTION && p->prog != FSM_PROG_NS3D && p->prog != FSM_PROG_LINEAR) {
        delete p;
        return fail(-ENOSYS, "slab decomposition supports the convection and Navier-Stokes programs only");
    }
    if (p->prog != FSM_PROG_LINEAR && d->integrator == FSM_INT_ETDRK0) {
        delete p;
        return fail(-EINVAL, "The ETDRK0 integrator only supports linear term");
    }
    if (d->nl_coef_b && (d->force_hat || d->dynamic_force)) {
        delete p;
        return fail(-ENOSYS, "a per-sample coefficient on NS pressure convection cannot be combined with an external force");
    }
    if (d->force_hat && p->prog != FSM_PROG_NS3D) {
        delete p;
        return fail(-EINVAL, "force_hat belongs to the NS pressure-convection program");
    }
    if (d->tab_complex && p->ndim != 1) {
        delete p;
        return fail(-ENOSYS, "complex coefficient tables (odd-order linear terms) are supported on 1-D grids only");
    }
    for (int i = 0; i < p->ndim; ++i) {
        const int e = p->f64 ? launch_table<double>(p->n[i])->prepare() : launch_table<float>(p->n[i])->prepare();
        if (e) { delete p; return fail(e, "could not initialise the twiddle tables"); }
    }
    p->stages = build_stages(d->integrator, d->dt, d->tab_lin != nullptr,
                             (d->integrator == FSM_INT_ETDRK2 || d->integrator == FSM_INT_SETDRK2) && d->tab_coef[2] != nullptr &&
                                 !d->force_hat && !d->dynamic_force);   // with a force the reference masks a in place but not N0
    if (p->stages.empty()) { delete p; return fail(-ENOSYS, "unknown integrator %d", d->integrator); }
    p->pf = FSM_PF_DEFAULT;
    // right-hand side L u + N(u)
    {
        Stage s;
        s.input = ARR_U; s.n_in = 1; s.in[0] = ARR_U; s.n_out = 1; s.out[0] = ARR_EXT;
        s.c[0][0] = scal(1.0);
        s.c[0][1] = d->tab_lin ? tabc(TAB_LIN) : scal(0.0);
        p->rhs_stage = s;
    }
    // chunking: keep the per-chunk intermediates of 2-D runs inside the 126 MB L2
    const size_t esz = p->f64 ? 16 : 8;
    const long plane = (long)p->n[0] * p->n[1];
    size_t w1_per = 0, w2_per = 0, w3_per = 0, w2b_per = 0;
    if (p->ndim == 1) {
        // the 1-D kernels keep everything on chip: no intermediates
    } else if (p->ndim == 2) {
        w1_per = (p->kprog == PROG_NS2D || p->kprog == PROG_KS2D) ? (size_t)2 * p->n[0] * p->n[1] * esz
                                         : (size_t)p->C * p->nf_ix * p->n[0] * p->ph * esz;
        w2_per = (size_t)p->nout * p->nmodes * esz;
    } else {
        if (p->P > 1) {   // the exchange buffers are the caller's; only the x-slab intermediates live here
            const int nf3 = (p->nfi > p->C) ? p->nfi : p->C;
            w1_per = 0;
            w3_per = (size_t)nf3 * p->nxl * p->n[1] * p->ph * esz;
            w2_per = (size_t)((p->nout > p->C) ? p->nout : p->C) * p->nh * p->nxl * p->n[1] * esz;
            w2b_per = 0;
        } else {
            w1_per = (size_t)p->C * p->nf_ix * p->nh * plane * esz;
            w3_per = (size_t)p->nfi * plane * p->ph * esz;
            w2_per = (size_t)p->nout * p->nh * plane * esz;
            w2b_per = (size_t)p->nout * p->nmodes * esz;
        }
    }
    // lanes: independent sample ranges stepped concurrently (do_step). KS couples the samples through its batch mean.
    int nlanes = d->lanes;
    const bool independent = !(p->prog == FSM_PROG_KS && d->ks_remove_mean) && p->prog != FSM_PROG_LINEAR;
    if (nlanes <= 0) nlanes = (independent && p->ndim >= 2 && p->P == 1 && p->B >= 4) ? FSM_LANES_DEFAULT : 1;
    if (!independent || p->ndim == 1 || p->P > 1) nlanes = 1;
    if (nlanes > FSM_MAX_LANES) nlanes = FSM_MAX_LANES;
    if (nlanes > p->B) nlanes = p->B;
#ifdef FSM_EMU
    nlanes = 1;
#endif
    p->nlanes = nlanes;
    const int lane_B = (p->B + nlanes - 1) / nlanes;   // samples per lane
    int chunk = d->chunk;
    if (chunk <= 0) {
        // as many samples per launch as an 8 GiB budget for the per-chunk intermediates allows (measured on
        // C3: 64 samples per launch 2.19 ms/step, 16 -> 2.42, 4 -> 3.4: longer grids hide the wave tails
        // and launch gaps better than L2 residency of the intermediates pays back), then balance the chunks.
        const size_t per = w1_per + w2_per + w3_per + w2b_per;
        const size_t budget = (size_t)8 << 30;
        chunk = lane_B;
        if (per > 0 && (size_t)chunk * per > budget && p->P == 1) chunk = (int)(budget / per);
        if (chunk < 1) chunk = 1;
        if (chunk > lane_B) chunk = lane_B;
        const int nch = (lane_B + chunk - 1) / chunk;
        chunk = (lane_B + nch - 1) / nch;
    }
    if (p->P > 1) chunk = p->B;
    if (chunk > lane_B) chunk = lane_B;
    p->chunk = chunk;
    // workspace layout
    size_t off = 0;
    const size_t state_bytes = (size_t)p->B * p->C * p->nmodes * esz;
    int n_scratch = 0;
    for (const Stage& s : p->stages) {
        for (int i = 0; i < s.n_out; ++i) if (s.out[i] >= ARR_S1 && s.out[i] <= ARR_S4 && s.out[i] > n_scratch) n_scratch = s.out[i];
    }
    for (int i = 0; i < ARR_COUNT; ++i) p->off_arr[i] = 0;
    for (int i = ARR_S1; i <= ARR_S4; ++i) {
        p->off_arr[i] = off;
        if (i <= n_scratch) off = align_up(off + state_bytes, 256);
    }
    p->lane_w_bytes[0] = align_up(w1_per * chunk, 256);
    p->lane_w_bytes[1] = align_up(w2_per * chunk, 256);
    p->lane_w_bytes[2] = align_up(w3_per * chunk, 256);
    p->lane_w_bytes[3] = align_up(w2b_per * chunk, 256);
    p->off_w1 = off; off += p->lane_w_bytes[0] * nlanes;
    p->off_w2 = off; off += p->lane_w_bytes[1] * nlanes;
    p->off_w3 = off; off += p->lane_w_bytes[2] * nlanes;
    p->off_w2b = off; off += p->lane_w_bytes[3] * nlanes;
    p->off_dc = off; off = align_up(off + (size_t)p->B * (p->f64 ? 8 : 4), 256);
    p->ws_bytes = off;
    // generic transforms (r2c / c2r) process this many independent fields per launch: bounded by what
    // each intermediate buffer can hold when every field takes one plain slot
    {
        size_t cap;
        if (p->ndim == 1) {
            cap = (size_t)p->B * p->C;
        } else if (p->ndim == 2) {
            cap = (p->lane_w_bytes[0] * nlanes) / ((size_t)p->n[0] * p->ph * esz);    // the lane regions are contiguous
            const size_t c2 = (p->lane_w_bytes[1] * nlanes) / ((size_t)p->nmodes * esz);
            if (c2 < cap) cap = c2;
        } else if (p->P > 1) {
            cap = (size_t)p->B * p->C;
        } else {
            cap = (p->lane_w_bytes[0] * nlanes) / ((size_t)p->nh * plane * esz);
            const size_t c3 = (p->lane_w_bytes[2] * nlanes) / ((size_t)plane * p->ph * esz);
            const size_t c2 = (p->lane_w_bytes[1] * nlanes) / ((size_t)p->nh * plane * esz);
            const size_t c2b = (p->lane_w_bytes[3] * nlanes) / ((size_t)p->nmodes * esz);
            if (c3 < cap) cap = c3;
            if (c2 < cap) cap = c2;
            if (c2b < cap) cap = c2b;
        }
        p->cap_fields = cap < 1 ? 1 : cap;
    }
    // bookkeeping for benchmarks (SURVEY.md §8d transform-pass model)
    {
        const int nchunks = nlanes * ((lane_B + chunk - 1) / chunk);
        int64_t launches = 0, units = 0;  // units of one real field
        for (int i = 0; i < 4; ++i) p->pass_units[i] = 0;
        for (const Stage& s : p->stages) {
            if (p->prog == FSM_PROG_LINEAR) {
                launches += 1;
            } else if (p->ndim == 1) {
                launches = 1;   // the whole time loop is one launch
            } else {
                launches += (int64_t)nchunks * (p->ndim == 2 ? 3 : 5) + ((p->prog == FSM_PROG_KS && d->ks_remove_mean) ? 1 : 0);
                const int cnf = p->C * p->nf_ix;
                p->pass_units[PASS_IX] += p->C + cnf;
                p->pass_units[PASS_PHYS] += p->nfi + p->nout;
                p->pass_units[PASS_FX] += p->nout;
                if (p->ndim == 3) p->pass_units[PASS_MID] += (cnf + p->nfi) + 2 * p->nout;
            }
            p->pass_units[PASS_FX] += (int64_t)(s.model_io >= 0 ? s.model_io : s.n_in + s.n_out) * p->C;
        }
        // bytes really touched: the inverse side moves kept (dealiased) modes only
        for (int i = 0; i < 4; ++i) p->pass_touched[i] = 0;
        if (p->prog != FSM_PROG_LINEAR && p->ndim >= 2) {
            const int64_t esz8 = p->f64 ? 16 : 8;
            const int64_t n0 = p->n[0], n1 = p->n[1], nh = p->nh;
            const int64_t kx = (2 * d->kmax[0] + 1 < n0) ? 2 * d->kmax[0] + 1 : n0;
            const int64_t cnf = (int64_t)p->C * p->nf_ix;
            int64_t t_ix, t_mid = 0, t_phys;
            if (p->ndim == 2) {
                const int64_t ly = (d->kmax[1] + 1 < nh) ? d->kmax[1] + 1 : nh;
                const bool zl = (p->kprog == PROG_NS2D || p->kprog == PROG_KS2D);
                const int64_t w1 = zl ? (int64_t)(p->nfi / 2) * n0 * (2 * ly - 1) : cnf * n0 * ly;
                t_ix = p->C * ly * kx + w1;
                t_phys = w1 + (int64_t)p->nout * nh * n0;
            } else {
                const int64_t ky = (2 * d->kmax[1] + 1 < n1) ? 2 * d->kmax[1] + 1 : n1;
                const int64_t kz = (d->kmax[2] + 1 < nh) ? d->kmax[2] + 1 : nh;
                t_ix = p->C * ky * kz * kx + cnf * kz * n0 * ky;
                t_mid = (int64_t)p->nfi * kz * n0 * ky + (int64_t)p->nfi * n0 * n1 * kz + 2 * (int64_t)p->nout * nh * n0 * n1;
                t_phys = (int64_t)p->nfi * n0 * n1 * kz + (int64_t)p->nout * nh * n0 * n1;
            }
            const int64_t full = (p->P > 1) ? p->nmodes * p->P : p->nmodes;   // modes of one field of the whole grid
            for (const Stage& s : p->stages) {
                p->pass_touched[PASS_IX] += t_ix * esz8 * p->B;
                p->pass_touched[PASS_MID] += t_mid * esz8 * p->B;
                p->pass_touched[PASS_PHYS] += t_phys * esz8 * p->B;
                p->pass_touched[PASS_FX] += ((int64_t)p->nout + (int64_t)(s.n_in + s.n_out) * p->C) * full * esz8 * p->B;
            }
        }
        for (int i = 0; i < 4; ++i) units += p->pass_units[i];
        p->launches_per_step = launches;
        p->algo_bytes_per_step = units * (int64_t)p->ntot * (p->f64 ? 8 : 4) * p->B;
    }
    *out = p;
    return 0;
}

void fsm_plan_destroy(fsm_plan* plan) {
    if (!plan) return;
#ifndef FSM_EMU
    for (auto& u : plan->ev_used) { cudaEventDestroy(u.second.first); cudaEventDestroy(u.second.second); }
    for (auto e : plan->ev_pool) cudaEventDestroy(e);
    if (plan->lane_device >= 0) {
        for (int l = 0; l < plan->nlanes; ++l) {
            if (plan->lane_stream[l]) cudaStreamDestroy(plan->lane_stream[l]);
            if (plan->lane_join[l]) cudaEventDestroy(plan->lane_join[l]);
        }
        if (plan->lane_fork) cudaEventDestroy(plan->lane_fork);
    }
#endif
    delete plan;
}

size_t fsm_workspace_bytes(const fsm_plan* plan) { return plan ? plan->ws_bytes : 0; }

int fsm_plan_info(const fsm_plan* plan, int64_t* launches_per_step, int64_t* algo_bytes_per_step,
                  int64_t* modes_per_field, int32_t* chunk) {
    if (!plan) return fail(-EINVAL, "null plan");
    if (launches_per_step) *launches_per_step = plan->launches_per_step;
    if (algo_bytes_per_step) *algo_bytes_per_step = plan->algo_bytes_per_step;
    if (modes_per_field) *modes_per_field = plan->nmodes;
    if (chunk) *chunk = plan->chunk;
    return 0;
}

int fsm_slab_peers(fsm_plan* plan, int exchange, const void* const* ptrs, int32_t n) {
    if (!plan || plan->P <= 1) return fail(-EINVAL, "plan has no slab decomposition");
    if (exchange != 1 && exchange != 2) return fail(-EINVAL, "exchange must be 1 (inverse side) or 2 (forward side)");
    if (!ptrs || n == 0) { plan->n_peers[exchange - 1] = 0; return 0; }
    if (n != plan->P || n > FSM_MAX_PEERS) return fail(-EINVAL, "need one receive buffer per rank (%d), at most %d", plan->P, FSM_MAX_PEERS);
    for (int r = 0; r < n; ++r) {
        if (!ptrs[r]) return fail(-EINVAL, "null receive buffer for rank %d", r);
        plan->peers[exchange - 1][r] = const_cast<void*>(ptrs[r]);
    }
    plan->n_peers[exchange - 1] = n;
    return 0;
}

int fsm_stage_kinds(const fsm_plan* plan, int32_t* kinds, int32_t capacity) {
    if (!plan) return fail(-EINVAL, "null plan");
    const int n = (int)plan->stages.size();
    for (int i = 0; i < n && i < capacity; ++i) {
        int kind = -1;
        if (plan->f64) {
            cplx<double>* arr[ARR_COUNT] = {};
            Combine<double> cb;
            if (make_combine<double>(plan, plan->stages[i], arr, plan->prog != FSM_PROG_LINEAR, &cb) == 0) kind = cb.kind;
        } else {
            cplx<float>* arr[ARR_COUNT] = {};
            Combine<float> cb;
            if (make_combine<float>(plan, plan->stages[i], arr, plan->prog != FSM_PROG_LINEAR, &cb) == 0) kind = cb.kind;
        }
        if (kinds) kinds[i] = kind;
    }
    return n;
}

int fsm_slab_info(const fsm_plan* plan, int op, int64_t* exch1_elems, int64_t* exch2_elems, int32_t* n_stages) {
    if (!plan || plan->P <= 1) return fail(-EINVAL, "plan has no slab decomposition");
    const int64_t blk1 = (int64_t)plan->nxl * plan->kyl, blk2 = (int64_t)plan->kyl * plan->nh * plan->nxl;
    int64_t e1, e2;
    if (op == FSM_SLAB_STEP || op == FSM_SLAB_RHS) {
        e1 = (int64_t)plan->P * plan->B * plan->C * plan->nf_ix * plan->nkz1 * plan->nxl * plan->nky1;
        e2 = (int64_t)plan->P * plan->B * plan->nout * blk2;
    } else {
        e1 = (int64_t)plan->P * plan->B * plan->C * plan->nh * blk1;
        e2 = (int64_t)plan->P * plan->B * plan->C * blk2;
    }
    if (exch1_elems) *exch1_elems = e1;
    if (exch2_elems) *exch2_elems = e2;
    if (n_stages) *n_stages = (int32_t)plan->stages.size();
    return 0;
}

int fsm_slab_phase(fsm_plan* plan, int op, int stage, int phase, int sub, int nsub, void* u_hat, void* aux, void* workspace,
                   size_t ws_bytes, void* send, void* recv, void* stream) {
    if (!plan || plan->P <= 1) return fail(-EINVAL, "plan has no slab decomposition");
    if (!workspace || ws_bytes < plan->ws_bytes) return fail(-ENOMEM, "workspace too small");
    if (op == FSM_SLAB_STEP && (stage < 0 || stage >= (int)plan->stages.size())) return fail(-EINVAL, "bad stage %d", stage);
    if (phase < 0 || phase > 2) return fail(-EINVAL, "bad phase %d", phase);
    if (nsub < 1 || (nsub & (nsub - 1)) || plan->nxl % nsub || sub < 0 || sub >= nsub)
        return fail(-EINVAL, "bad sub-slab %d of %d (local x extent %d)", sub, nsub, plan->nxl);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return plan->f64 ? do_slab_phase<double>(plan, op, stage, phase, sub, nsub, u_hat, aux, workspace, send, recv, st)
                     : do_slab_phase<float>(plan, op, stage, phase, sub, nsub, u_hat, aux, workspace, send, recv, st);
}

int fsm_ks_log(fsm_plan* plan, void* log, int64_t capacity) {
    if (!plan) return fail(-EINVAL, "null plan");
    plan->ks_log = log;
    plan->ks_log_cap = log ? capacity : 0;
    plan->ks_log_pos = 0;
    return 0;
}

int fsm_profile_enable(fsm_plan* plan, int on) {
    if (!plan) return fail(-EINVAL, "null plan");
    plan->profile = on != 0;
    return 0;
}

int fsm_profile_read(fsm_plan* plan, double* ms, int64_t* launches, int64_t* algo_bytes_per_step) {
    if (!plan || !ms || !launches) return fail(-EINVAL, "bad argument");
    for (int i = 0; i < 4; ++i) { ms[i] = 0; launches[i] = 0; }
    if (algo_bytes_per_step)
        for (int i = 0; i < 4; ++i)
            algo_bytes_per_step[i] = plan->pass_units[i] * (int64_t)plan->ntot * (plan->f64 ? 8 : 4) * plan->B;
#ifndef FSM_EMU
    for (auto& u : plan->ev_used) {
        cudaEventSynchronize(u.second.second);
        float t = 0;
        cudaEventElapsedTime(&t, u.second.first, u.second.second);
        ms[u.first] += t;
        launches[u.first] += 1;
        plan->ev_pool.push_back(u.second.first);
        plan->ev_pool.push_back(u.second.second);
    }
    plan->ev_used.clear();
#endif
    return 0;
}

#define FSM_CHECK_WS(plan, ws, bytes)                                                        \
    if (!(plan)) return fail(-EINVAL, "null plan");                                          \
    if (!(ws) || (bytes) < (plan)->ws_bytes)                                                 \
        return fail(-ENOMEM, "workspace too small: %zu bytes given, %zu needed", (size_t)(bytes), (plan)->ws_bytes);

int fsm_step(fsm_plan* plan, void* u_hat, void* workspace, size_t ws_bytes, int n_steps, void* stream) {
    FSM_CHECK_WS(plan, workspace, ws_bytes);
    if (!u_hat || n_steps < 0) return fail(-EINVAL, "bad argument");
    if (plan->P > 1 && plan->prog != FSM_PROG_LINEAR) return fail(-EINVAL, "slab-decomposed plans are driven through fsm_slab_phase");
    // the reference dealiases the stage state in place when a force is attached (quirk Q5); what that does to the
    // temporaries of _rk.py:43-58 is an accident of evaluation order this path does not restate
    if (plan->d.force_hat && plan->d.integrator == FSM_INT_RK4)
        return fail(-ENOSYS, "NS pressure convection with an external force is not supported with the RK integrators");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return plan->f64 ? do_step<double>(plan, u_hat, workspace, n_steps, st) : do_step<float>(plan, u_hat, workspace, n_steps, st);
}

int fsm_rhs(fsm_plan* plan, const void* u_hat, void* out_hat, void* workspace, size_t ws_bytes, void* stream) {
    FSM_CHECK_WS(plan, workspace, ws_bytes);
    if (!u_hat || !out_hat || u_hat == out_hat) return fail(-EINVAL, "bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return plan->f64 ? do_rhs<double>(plan, u_hat, out_hat, workspace, st) : do_rhs<float>(plan, u_hat, out_hat, workspace, st);
}

int fsm_r2c(fsm_plan* plan, const void* u, void* u_hat, void* workspace, size_t ws_bytes, void* stream) {
    FSM_CHECK_WS(plan, workspace, ws_bytes);
    if (!u || !u_hat) return fail(-EINVAL, "bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return plan->f64 ? do_r2c<double>(plan, u, u_hat, workspace, st) : do_r2c<float>(plan, u, u_hat, workspace, st);
}

int fsm_c2r(fsm_plan* plan, const void* u_hat, void* u, void* workspace, size_t ws_bytes, void* stream) {
    FSM_CHECK_WS(plan, workspace, ws_bytes);
    if (!u || !u_hat) return fail(-EINVAL, "bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return plan->f64 ? do_c2r<double>(plan, u_hat, u, workspace, st) : do_c2r<float>(plan, u_hat, u, workspace, st);
}

int fsm_half_to_full(fsm_plan* plan, const void* u_hat, void* full_hat, void* stream) {
    if (!plan || !u_hat || !full_hat) return fail(-EINVAL, "bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long nf = (long)plan->B * plan->C;
    const long total = nf * plan->ntot;
    dim3 grid((unsigned)((total + 255) / 256)), block(256);
    if (plan->f64) {
        auto kern = k_half_to_full<double>;
        FSM_LAUNCH(kern, grid, block, 0, st, make_geom<double>(plan, true), (const cplx<double>*)u_hat, (cplx<double>*)full_hat, nf);
    } else {
        auto kern = k_half_to_full<float>;
        FSM_LAUNCH(kern, grid, block, 0, st, make_geom<float>(plan, true), (const cplx<float>*)u_hat, (cplx<float>*)full_hat, nf);
    }
    return launch_status("half_to_full");
}

int fsm_full_to_half(fsm_plan* plan, const void* full_hat, void* u_hat, void* stream) {
    if (!plan || !u_hat || !full_hat) return fail(-EINVAL, "bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long nf = (long)plan->B * plan->C;
    const long total = nf * plan->nmodes;
    dim3 grid((unsigned)((total + 255) / 256)), block(256);
    if (plan->f64) {
        auto kern = k_full_to_half<double>;
        FSM_LAUNCH(kern, grid, block, 0, st, make_geom<double>(plan, true), (const cplx<double>*)full_hat, (cplx<double>*)u_hat, nf);
    } else {
        auto kern = k_full_to_half<float>;
        FSM_LAUNCH(kern, grid, block, 0, st, make_geom<float>(plan, true), (const cplx<float>*)full_hat, (cplx<float>*)u_hat, nf);
    }
    return launch_status("full_to_half");
}

int fsm_spectral_map(fsm_plan* plan, const void* in_hat, int32_t c_in, void* out_hat, int32_t c_out,
                     const fsm_map_term* terms, int32_t n_terms, int32_t dealias, void* stream) {
    if (!plan || !in_hat || !out_hat || in_hat == out_hat || !terms) return fail(-EINVAL, "bad argument");
    if (c_in < 1 || c_in > FSM_MAP_MAX_CH || c_out < 1 || c_out > FSM_MAP_MAX_CH)
        return fail(-EINVAL, "spectral map: channel counts must lie in [1, %d]", FSM_MAP_MAX_CH);
    if (n_terms < 0 || n_terms > FSM_MAP_MAX_TERMS) return fail(-EINVAL, "spectral map: at most %d terms", FSM_MAP_MAX_TERMS);
    for (int t = 0; t < n_terms; ++t) {
        const fsm_map_term& m = terms[t];
        if (m.out_channel < 0 || m.out_channel >= c_out || m.in_channel < 0 || m.in_channel >= c_in)
            return fail(-EINVAL, "spectral map: term %d refers to a channel outside the fields", t);
        if (m.inv_laplacian < 0 || m.inv_laplacian > 8) return fail(-EINVAL, "spectral map: bad inverse-Laplacian power in term %d", t);
        for (int a = 0; a < 3; ++a) {
            if (m.power[a] < 0 || m.power[a] > 16) return fail(-EINVAL, "spectral map: bad derivative order in term %d", t);
            if (a >= plan->ndim && m.power[a] != 0) return fail(-EINVAL, "spectral map: term %d differentiates along a missing axis", t);
        }
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long total = (long)plan->B * plan->nmodes;
    dim3 grid((unsigned)((total + 255) / 256)), block(256);
    if (plan->f64) {
        const MapTerms<double> mt = make_map_terms<double>(terms, n_terms, c_in, c_out, dealias);
        auto kern = k_spectral_map<double>;
        FSM_LAUNCH(kern, grid, block, 0, st, make_geom<double>(plan, dealias == 0), mt, (const cplx<double>*)in_hat, (cplx<double>*)out_hat, total);
    } else {
        const MapTerms<float> mt = make_map_terms<float>(terms, n_terms, c_in, c_out, dealias);
        auto kern = k_spectral_map<float>;
        FSM_LAUNCH(kern, grid, block, 0, st, make_geom<float>(plan, dealias == 0), mt, (const cplx<float>*)in_hat, (cplx<float>*)out_hat, total);
    }
    return launch_status("spectral map");
}

int fsm_stage_input(const fsm_plan* plan, int32_t stage, int64_t* ws_offset) {
    if (!plan || !ws_offset) return fail(-EINVAL, "bad argument");
    if (stage < -1 || stage >= (int)plan->stages.size()) return fail(-EINVAL, "bad stage %d", stage);
    const int arr = (stage < 0) ? ARR_U : plan->stages[stage].input;
    *ws_offset = (arr == ARR_U) ? -1 : (int64_t)plan->off_arr[arr];
    return (int)plan->stages.size();
}

int fsm_stage_combine(fsm_plan* plan, int32_t stage, void* u_hat, const void* fresh_hat, void* rhs_out, void* workspace,
                      size_t ws_bytes, void* stream) {
    FSM_CHECK_WS(plan, workspace, ws_bytes);
    if (!u_hat || !fresh_hat) return fail(-EINVAL, "bad argument");
    if (plan->prog != FSM_PROG_LINEAR) return fail(-EINVAL, "fsm_stage_combine belongs to plans without a fused nonlinear program");
    if (stage < -1 || stage >= (int)plan->stages.size()) return fail(-EINVAL, "bad stage %d", stage);
    if (stage < 0 && (!rhs_out || rhs_out == u_hat)) return fail(-EINVAL, "the right-hand side needs its own output array");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Stage& s = (stage < 0) ? plan->rhs_stage : plan->stages[stage];
    if (plan->f64) {
        Buffers<double> bf = carve<double>(plan, u_hat, workspace, rhs_out);
        return run_stage<double>(plan, bf, s, st, 0, -1, static_cast<const cplx<double>*>(fresh_hat));
    }
    Buffers<float> bf = carve<float>(plan, u_hat, workspace, rhs_out);
    return run_stage<float>(plan, bf, s, st, 0, -1, static_cast<const cplx<float>*>(fresh_hat));
}

int fsm_mask_state(fsm_plan* plan, void* state_hat, int32_t channels, void* stream) {
    if (!plan || !state_hat || channels < 1) return fail(-EINVAL, "bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long total = (long)plan->B * channels * plan->nmodes;
    dim3 grid((unsigned)((total + 255) / 256)), block(256);
    if (plan->f64) {
        auto kern = k_mask_state<double>;
        FSM_LAUNCH(kern, grid, block, 0, st, make_geom<double>(plan, false), (cplx<double>*)state_hat, total);
    } else {
        auto kern = k_mask_state<float>;
        FSM_LAUNCH(kern, grid, block, 0, st, make_geom<float>(plan, false), (cplx<float>*)state_hat, total);
    }
    return launch_status("mask state");
}

int fsm_sym_outer(fsm_plan* plan, const void* u, void* out, int32_t channels, void* stream) {
    if (!plan || !u || !out || u == out) return fail(-EINVAL, "bad argument");
    if (channels < 1 || channels > FSM_MAP_MAX_CH) return fail(-EINVAL, "channel count must lie in [1, %d]", FSM_MAP_MAX_CH);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long npts = plan->ntot / plan->P;      // local physical slab of a decomposed grid
    const long total = (long)plan->B * npts;
    dim3 grid((unsigned)((total + 255) / 256)), block(256);
    if (plan->f64) {
        auto kern = k_sym_outer<double>;
        FSM_LAUNCH(kern, grid, block, 0, st, (const double*)u, (double*)out, (int)channels, npts, total);
    } else {
        auto kern = k_sym_outer<float>;
        FSM_LAUNCH(kern, grid, block, 0, st, (const float*)u, (float*)out, (int)channels, npts, total);
    }
    return launch_status("symmetric products");
}

int fsm_plan_traffic(const fsm_plan* plan, int64_t* touched_bytes_per_step4) {
    if (!plan || !touched_bytes_per_step4) return fail(-EINVAL, "bad argument");
    for (int i = 0; i < 4; ++i) touched_bytes_per_step4[i] = plan->pass_touched[i];
    return 0;
}

int fsm_stage_run(fsm_plan* plan, int32_t stage, void* u_hat, const void* force_hat, void* rhs_out, void* workspace,
                  size_t ws_bytes, void* stream) {
    FSM_CHECK_WS(plan, workspace, ws_bytes);
    if (!u_hat || !force_hat) return fail(-EINVAL, "bad argument");
    if (plan->prog != FSM_PROG_NS3D || !plan->d.dynamic_force)
        return fail(-EINVAL, "fsm_stage_run belongs to NS pressure-convection plans created with dynamic_force = 1");
    if (plan->P > 1) return fail(-ENOSYS, "state-dependent forces are not available on slab-decomposed grids");
    if (plan->d.integrator == FSM_INT_RK4 && stage >= 0)
        return fail(-ENOSYS, "NS pressure convection with an external force is not supported with the RK integrators");
    if (stage < -1 || stage >= (int)plan->stages.size()) return fail(-EINVAL, "bad stage %d", stage);
    if (stage < 0 && (!rhs_out || rhs_out == u_hat)) return fail(-EINVAL, "the right-hand side needs its own output array");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Stage& s = (stage < 0) ? plan->rhs_stage : plan->stages[stage];
    if (plan->f64) {
        Buffers<double> bf = carve<double>(plan, u_hat, workspace, rhs_out);
        return run_stage<double>(plan, bf, s, st, 0, -1, static_cast<const cplx<double>*>(force_hat));
    }
    Buffers<float> bf = carve<float>(plan, u_hat, workspace, rhs_out);
    return run_stage<float>(plan, bf, s, st, 0, -1, static_cast<const cplx<float>*>(force_hat));
}

int fsm_lincomb(fsm_plan* plan, void* out, const void* base, const void* const* terms, const double* coefs, int32_t n_terms,
                int64_t count, void* stream) {
    if (!plan || !out || !base || (n_terms > 0 && (!terms || !coefs)) || count < 0) return fail(-EINVAL, "bad argument");
    if (n_terms < 0 || n_terms > FSM_LINCOMB_MAX) return fail(-EINVAL, "at most %d terms", FSM_LINCOMB_MAX);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((unsigned)((count + 255) / 256)), block(256);
    if (count == 0) return 0;
    if (plan->f64) {
        LinComb<double> lc;
        lc.n = n_terms;
        for (int j = 0; j < n_terms; ++j) { lc.term[j] = static_cast<const cplx<double>*>(terms[j]); lc.coef[j] = coefs[j]; }
        auto kern = k_lincomb<double>;
        FSM_LAUNCH(kern, grid, block, 0, st, (cplx<double>*)out, (const cplx<double>*)base, lc, (long)count);
    } else {
        LinComb<float> lc;
        lc.n = n_terms;
        for (int j = 0; j < n_terms; ++j) { lc.term[j] = static_cast<const cplx<float>*>(terms[j]); lc.coef[j] = (float)coefs[j]; }
        auto kern = k_lincomb<float>;
        FSM_LAUNCH(kern, grid, block, 0, st, (cplx<float>*)out, (const cplx<float>*)base, lc, (long)count);
    }
    return launch_status("linear combination");
}

}  // extern "C"
