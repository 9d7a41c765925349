// One warp = one FJSP environment.  This file is the whole per-environment algorithm:
// dispatching-rule selection, dispatch, discrete-event clock, order arrival with the
// fluid LP, state features, reward, auto-reset.  It is written against a tiny warp
// abstraction (fj_lane / fj_sync / fj_bcast / reductions) so that the SAME source
//   * compiles with nvcc for sm_100a with 32 cooperating lanes (the product), and
//   * compiles with g++ as a 1-lane program (tests/hostsim) so the CPU test-suite can
//     check the compressed state machine against the oracle without a GPU.
//
// Reference behaviour being reproduced (see DESIGN.md for the mapping):
//   environments/SO_DFJSP.py, MO_DFJSP.py, MO_DFJSP_breakdown.py (step / state_extract /
//   update_parameter / task_select / machine_select / compute_reward) on
//   environments/class_FJSP.py, class_MODFJSP.py (reset_parameter, reset_object_add,
//   fluid_model, update_fluid_parameter).
//
// Exactness rules: every float that feeds a dispatching decision (urgency, estimated
// delay, gap, machine gap_ave, fluid rates) is computed with the reference's operation
// order, CPython's compensated sum() included, using explicitly rounded add/mul (no FMA
// contraction).  Floats that only feed the observation vector use warp tree sums.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "fjsp_layout.h"

#if defined(__CUDACC__) && defined(__CUDA_ARCH__)
#define FJ_DEVICE_CODE 1
#endif

#ifdef __CUDACC__
#define FJ_FN __device__ __forceinline__
#define FJ_MFN __device__ __forceinline__
#define FJ_FN_NOINLINE __device__ __noinline__
#define FJ_OUTLINE __device__ __noinline__
#define FJ_NOUNROLL _Pragma("unroll 1")
#define FJ_UNROLL4 _Pragma("unroll 4")
#define FJ_NL 32
FJ_FN int fj_lane() { return threadIdx.x & 31; }
FJ_FN void fj_sync() { __syncwarp(); }
FJ_FN int fj_bcast_i(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
FJ_FN double fj_bcast_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
FJ_FN long long fj_bcast_ll(long long v, int src) { return __shfl_sync(0xffffffffu, v, src); }
FJ_FN int fj_any(int p) { return __any_sync(0xffffffffu, p); }
FJ_FN int fj_xor_i(int v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
FJ_FN double fj_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
FJ_FN long long fj_xor_ll(long long v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
// explicitly rounded double arithmetic: never contracted into FMA
FJ_FN double fj_add(double a, double b) { return __dadd_rn(a, b); }
FJ_FN double fj_sub(double a, double b) { return __dsub_rn(a, b); }
FJ_FN double fj_mul(double a, double b) { return __dmul_rn(a, b); }
FJ_FN_NOINLINE double fj_div(double a, double b) { return __ddiv_rn(a, b); }
#else
#define FJ_FN static inline
#define FJ_MFN inline
#define FJ_FN_NOINLINE static
#define FJ_OUTLINE static
#define FJ_NOUNROLL
#define FJ_UNROLL4
#define FJ_NL 1
FJ_FN int fj_lane() { return 0; }
FJ_FN void fj_sync() {}
FJ_FN int fj_bcast_i(int v, int) { return v; }
FJ_FN double fj_bcast_d(double v, int) { return v; }
FJ_FN long long fj_bcast_ll(long long v, int) { return v; }
FJ_FN int fj_any(int p) { return p != 0; }
FJ_FN int fj_xor_i(int v, int) { return v; }
FJ_FN double fj_xor_d(double v, int) { return v; }
FJ_FN long long fj_xor_ll(long long v, int) { return v; }
// the host simulation is built with -ffp-contract=off
FJ_FN double fj_add(double a, double b) { return a + b; }
FJ_FN double fj_sub(double a, double b) { return a - b; }
FJ_FN double fj_mul(double a, double b) { return a * b; }
FJ_FN double fj_div(double a, double b) { return a / b; }
#endif

// ---------------------------------------------------------------- warp reductions
// Integer reductions are single redux.sync instructions on the device (the shuffle butterflies
// they replace were ~250 instructions on every step's serial chain).
#ifdef FJ_DEVICE_CODE
FJ_FN int fj_min_i(int v) { return __reduce_min_sync(0xffffffffu, v); }
FJ_FN unsigned fj_or_u(unsigned v) { return __reduce_or_sync(0xffffffffu, v); }
// sum of non-negative 64-bit values: three 22-bit limbs, each summed exactly in 32 bits
FJ_FN long long fj_sum_ll(long long v)
{
    const unsigned long long u = (unsigned long long)v;
    const unsigned long long s0 = __reduce_add_sync(0xffffffffu, (unsigned)(u & 0x3fffffu));
    const unsigned long long s1 = __reduce_add_sync(0xffffffffu, (unsigned)((u >> 22) & 0x3fffffu));
    const unsigned long long s2 = __reduce_add_sync(0xffffffffu, (unsigned)(u >> 44));
    return (long long)(s0 + (s1 << 22) + (s2 << 44));
}
// sum of two packed non-negative 32-bit counters (hi << 32 | lo), each total below 2^32
FJ_FN long long fj_sum_pair(long long v)
{
    const unsigned long long lo = __reduce_add_sync(0xffffffffu, (unsigned)v);
    const unsigned long long hi = __reduce_add_sync(0xffffffffu, (unsigned)((unsigned long long)v >> 32));
    return (long long)((hi << 32) | lo);
}
#else
FJ_FN long long fj_sum_ll(long long v) { return v; }
FJ_FN long long fj_sum_pair(long long v) { return v; }
FJ_FN int fj_min_i(int v) { return v; }
FJ_FN unsigned fj_or_u(unsigned v) { return v; }
#endif
FJ_OUTLINE double fj_sum_d(double v)   // observation-only sums (fixed butterfly order)
{
    for (int m = FJ_NL / 2; m > 0; m >>= 1) v = fj_add(v, fj_xor_d(v, m));
    return v;
}

FJ_FN void fj_sum_d2(double &a, double &b)   // two observation-only sums in one butterfly (same order per value)
{
    for (int m = FJ_NL / 2; m > 0; m >>= 1) { const double oa = fj_xor_d(a, m), ob = fj_xor_d(b, m); a = fj_add(a, oa); b = fj_add(b, ob); }
}

FJ_FN void fj_sum_d4(double &a, double &b, double &c, double &d)
{
    for (int m = FJ_NL / 2; m > 0; m >>= 1) {
        const double oa = fj_xor_d(a, m), ob = fj_xor_d(b, m), oc = fj_xor_d(c, m), od = fj_xor_d(d, m);
        a = fj_add(a, oa); b = fj_add(b, ob); c = fj_add(c, oc); d = fj_add(d, od);
    }
}

// Lexicographic minimum of (key, id) over the warp on an order-preserving integer image of the
// double key (keys are never NaN; -0.0 is canonicalised by the callers): three redux.sync minima.
#define FJ_EMPTY 0x7fffffff
#ifdef __CUDACC__
FJ_FN unsigned long long fj_ord_enc(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
FJ_FN double fj_ord_dec(unsigned hi, unsigned lo)
{
    const unsigned long long u = ((unsigned long long)hi << 32) | lo;
    return __longlong_as_double((long long)((u >> 63) ? (u ^ 0x8000000000000000ull) : ~u));
}
// (hi, lo, id), 0xffffffff everywhere = none.  All lanes return the minimum.
FJ_FN void fj_warp_lexmin(unsigned &hi, unsigned &lo, unsigned &id)
{
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    lo = hi == mh ? lo : 0xffffffffu;
    const unsigned ml = __reduce_min_sync(0xffffffffu, lo);
    id = (hi == mh && lo == ml) ? id : 0xffffffffu;
    id = __reduce_min_sync(0xffffffffu, id);
    hi = mh; lo = ml;
}
#endif

// running "first extremal element" of Python's max()/min() over a list in index order
struct FjBest {
    double key; int idx;
};
FJ_FN void fj_best_init(FjBest &b) { b.key = 0.0; b.idx = 0x7fffffff; }
// local update; candidates arrive in ascending idx on a lane, so strict comparison keeps the first
FJ_FN void fj_best_max(FjBest &b, double key, int idx) { if (b.idx == 0x7fffffff || key > b.key) { b.key = key; b.idx = idx; } }
FJ_FN void fj_best_min(FjBest &b, double key, int idx) { if (b.idx == 0x7fffffff || key < b.key) { b.key = key; b.idx = idx; } }
// warp-wide winner: extremal key, lowest index on ties (only idx is meaningful afterwards)
FJ_FN void fj_best_reduce(FjBest &b, int want_max)
{
#ifdef FJ_DEVICE_CODE
    const bool none = b.idx == 0x7fffffff;
    unsigned long long u = fj_ord_enc(fj_add(b.key, 0.0));   // -0.0 -> +0.0: equal keys, equal images
    if (want_max) u = ~u;
    unsigned hi = none ? 0xffffffffu : (unsigned)(u >> 32), lo = none ? 0xffffffffu : (unsigned)u;
    unsigned id = none ? 0xffffffffu : (unsigned)b.idx;
    fj_warp_lexmin(hi, lo, id);
    b.idx = id == 0xffffffffu ? 0x7fffffff : (int)id;
#else
    (void)b; (void)want_max;
#endif
}

// ---------------------------------------------------------------- CPython sum()
struct FjPySum { double f, c; };
FJ_FN void fj_pysum_init(FjPySum &s) { s.f = 0.0; s.c = 0.0; }
template <int MODE> FJ_FN void fj_pysum_add(FjPySum &s, double x)
{
    if (MODE == 0) { s.f = fj_add(s.f, x); return; }
    double t = fj_add(s.f, x);
    if (fabs(s.f) >= fabs(x)) s.c = fj_add(s.c, fj_add(fj_sub(s.f, t), x));
    else s.c = fj_add(s.c, fj_add(fj_sub(x, t), s.f));
    s.f = t;
}
template <int MODE> FJ_FN double fj_pysum_result(const FjPySum &s)
{
    if (MODE != 0 && s.c != 0.0 && isfinite(s.c)) return fj_add(s.f, s.c);
    return s.f;
}

// ---------------------------------------------------------------- per-env context
struct FjCtx {
    const FjParams *P;
    const int32_t *I;        // instance record in HBM (read-only path)
    int inst;
    int M, K, KT, S, Mx, Kx, Sx;
    unsigned mmask;
    int32_t *scal; double *obs, *obs2, *gapave, *urg, *maxe; uint32_t *avmask, *favmask, *demask, *damask;
    int32_t *mend, *mlast, *mjob, *mD; double *mF, *invnkt; uint16_t *qhead, *qtail, *qlen; int32_t *proc, *fstart; uint32_t *flmask;
    double *rsum, *tsum; uint16_t *cntunp, *cntnow, *pk, *slot; double *fu, *fa, *ff; uint16_t *next;
    uint32_t *unpmask; int32_t *duejob, *mindue;   // SO_FJSSP only (per-job due dates)
    uint32_t *h_elig, *h_mtpack, *h_fo; uint16_t *h_rjinfo; int32_t *h_due, *h_cum, *h_jobbase;   // hot copies of instance statics
    unsigned char *lp;
};

// The launch parameters and every warp's context object live in CTA-shared memory (plain
// statics in the one-lane host build) and are never passed around: each out-of-line function
// picks its warp's context up with FJ_CTX.  (A 450-byte FjCtx handed by reference to
// __noinline__ functions lived in the per-thread stack: 1 KB x 896 threads per SM thrashed L1,
// every `c.field` was an 8-sector local-memory load and half of them missed;
// profiles/README.md r01_v4.  From shared memory it is one broadcast word.)
#define FJ_TRACE_ROWS 33   // trace rows per CTA: one per warp (<= 32) + one for the CTA's LP phases
#ifndef FJ_CTX_WARPS
#define FJ_CTX_WARPS 32   // most warps per CTA of any kernel that uses the contexts
#endif
#ifdef __CUDACC__
static __shared__ FjParams fj_sP;
static __shared__ FjCtx fj_sC[FJ_CTX_WARPS];
#define FJ_WIDX (threadIdx.x >> 5)
// every kernel calls this first (all threads)
FJ_FN void fj_params_to_shared(const FjParams &P)
{
    const int nw = (int)(sizeof(FjParams) / 4);
    const int32_t *src = (const int32_t *)&P;
    int32_t *dst = (int32_t *)&fj_sP;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}
FJ_FN const FjParams &fj_params_bind(const FjParams &) { return fj_sP; }
#else
static FjParams fj_sP;
static FjCtx fj_sC[1];
#define FJ_WIDX 0
FJ_FN const FjParams &fj_params_bind(const FjParams &P) { fj_sP = P; return fj_sP; }
#endif
#define FJ_CTX (fj_sC[FJ_WIDX])

// builds this warp's context for `env` (every lane stores the same words)
FJ_FN FjCtx &fj_ctx_init(const FjParams &, int env, unsigned char *lp, unsigned char *hot = nullptr)
{
    const FjParams &P = fj_sP;
    FjCtx &c = FJ_CTX;
    fj_sync();
    c.P = &fj_sP;
    c.inst = P.env_inst[env];
    c.I = P.inst + (size_t)c.inst * P.io.stride;
    const int32_t *h = c.I + P.io.hdr;
    c.M = h[0]; c.K = h[1]; c.KT = h[2]; c.S = h[3];
    c.Mx = P.d.Mx; c.Kx = P.d.Kx; c.Sx = P.d.Sx;
    c.mmask = h[0] >= 32 ? 0xffffffffu : ((1u << h[0]) - 1u);
    unsigned char *G = P.env + (size_t)env * P.eo.stride;   // record in HBM
    unsigned char *E = hot ? hot : G;                       // hot part, possibly staged in shared memory
    const FjEnvOff &o = P.eo;
    c.scal = (int32_t *)(E + o.scal); c.obs = (double *)(E + o.obs); c.obs2 = (double *)(E + o.obs2);
    c.gapave = (double *)(E + o.gapave); c.urg = (double *)(G + o.urg); c.maxe = (double *)(G + o.maxe);
    c.avmask = (uint32_t *)(E + o.avmask); c.favmask = (uint32_t *)(E + o.favmask);
    c.demask = (uint32_t *)(E + o.demask); c.damask = (uint32_t *)(E + o.damask);
    c.h_elig = (uint32_t *)(E + o.h_elig); c.h_mtpack = (uint32_t *)(E + o.h_mtpack); c.h_fo = (uint32_t *)(E + o.h_fo); c.h_rjinfo = (uint16_t *)(E + o.h_rjinfo); c.h_due = (int32_t *)(E + o.h_due);
    c.h_cum = (int32_t *)(E + o.h_cum); c.h_jobbase = (int32_t *)(E + o.h_jobbase);
    c.mF = (double *)(E + o.mF); c.mD = (int32_t *)(E + o.mD); c.invnkt = (double *)(E + o.invnkt);
    c.mend = (int32_t *)(E + o.mend); c.mlast = (int32_t *)(E + o.mlast); c.mjob = (int32_t *)(E + o.mjob);
    c.qhead = (uint16_t *)(E + o.qhead); c.qtail = (uint16_t *)(E + o.qtail); c.qlen = (uint16_t *)(E + o.qlen);
    c.proc = (int32_t *)(E + o.proc); c.fstart = (int32_t *)(E + o.fstart); c.flmask = (uint32_t *)(E + o.flmask);
    c.rsum = (double *)(E + o.rsum); c.tsum = (double *)(E + o.tsum);
    c.cntunp = (uint16_t *)(E + o.cntunp); c.cntnow = (uint16_t *)(E + o.cntnow);
    c.pk = (uint16_t *)(G + o.pk); c.slot = (uint16_t *)(G + o.slot);
    c.fu = (double *)(G + o.fu); c.fa = (double *)(G + o.fa); c.ff = (double *)(G + o.ff);
    c.next = (uint16_t *)(G + o.next);
    c.unpmask = (uint32_t *)(G + o.unpmask); c.duejob = (int32_t *)(G + o.duejob); c.mindue = (int32_t *)(G + o.mindue);
    c.lp = lp;
    fj_sync();
    return c;
}

// read-only instance table: loaded through the non-coherent path (ld.global.nc: L1-cached, never
// invalidated by the env-state stores).  (Staging its small arrays in shared memory next to the
// env record was measured slower -- it takes L1 away -- and is gone; the per-operation-type
// statics every step reads are copied into the env record's hot prefix at reset() instead.)
struct FjRO {
    const int32_t *p;
    FJ_MFN int operator[](int i) const
    {
#ifdef FJ_DEVICE_CODE
        return __ldg(p + i);
#else
        return p[i];
#endif
    }
    FJ_MFN FjRO operator+(int o) const { FjRO r; r.p = p + o; return r; }
};
FJ_FN FjRO fj_ro(const int32_t *p) { FjRO r; r.p = p; return r; }
FJ_FN FjRO fj_ro_field(const FjCtx &c, int off) { FjRO r; r.p = c.I + off; return r; }
#define FJ_I(c, field) (fj_ro_field((c), (c).P->io.field))
// the per-operation-type statics every step reads (eligible-machine mask, kind / stage / last
// flag) and the orders' due dates are copied into the env record's hot prefix at reset(), so
// the main kernel reads them from shared memory
struct FjEligRO { const uint32_t *p; FJ_MFN int operator[](int q) const { return (int)p[q]; } };
struct FjLastRO { const uint16_t *p; FJ_MFN int operator[](int q) const { return p[q] & 1; } };
struct FjKindRO { const uint16_t *p; FJ_MFN int operator[](int q) const { return p[q] >> 8; } };
struct FjStageRO { const uint16_t *p; FJ_MFN int operator[](int q) const { return (p[q] >> 1) & 0x7f; } };
struct FjDueRO { const int32_t *p; FJ_MFN int operator[](int s) const { return p[s]; } };
FJ_FN FjEligRO fj_elig(const FjCtx &c) { FjEligRO r; r.p = c.h_elig; return r; }
FJ_FN FjLastRO fj_rjlast(const FjCtx &c) { FjLastRO r; r.p = c.h_rjinfo; return r; }
FJ_FN FjKindRO fj_rjkind(const FjCtx &c) { FjKindRO r; r.p = c.h_rjinfo; return r; }
FJ_FN FjStageRO fj_rjstage(const FjCtx &c) { FjStageRO r; r.p = c.h_rjinfo; return r; }
FJ_FN FjDueRO fj_due(const FjCtx &c) { FjDueRO r; r.p = c.h_due; return r; }
FJ_FN FjDueRO fj_cum(const FjCtx &c) { FjDueRO r; r.p = c.h_cum; return r; }          // [(Sx+1)*Kx]
FJ_FN FjDueRO fj_jobbase(const FjCtx &c) { FjDueRO r; r.p = c.h_jobbase; return r; }  // [Kx]
FJ_FN long long fj_get_ll(const int32_t *scal, int i) { return *(const long long *)(scal + i); }
FJ_FN void fj_set_ll(int32_t *scal, int i, long long v) { *(long long *)(scal + i) = v; }
FJ_FN double fj_get_d(const int32_t *scal, int i) { return *(const double *)(scal + i); }
FJ_FN void fj_set_d(int32_t *scal, int i, double v) { *(double *)(scal + i) = v; }
FJ_FN int fj_is_mo(int variant) { return variant == FJSP_MO_DFJSP || variant == FJSP_MO_BREAKDOWN; }
FJ_FN int fj_popc(unsigned v)
{
#ifdef FJ_DEVICE_CODE
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
FJ_FN int fj_ffs0(unsigned v)   // index of lowest set bit, v != 0
{
#ifdef FJ_DEVICE_CODE
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
FJ_OUTLINE int fj_order_of(const int32_t *cum, int S, int Kx, int r, int n)   // which order job n of kind r came with
{
    int s = 0;
    FJ_NOUNROLL
    while (s + 1 < S && n >= cum[(s + 1) * Kx + r]) ++s;
    return s;
}

// ---------------------------------------------------------------- fluid LP
// Same pivoting specification as oracle/fjsp_lp.c (DESIGN.md "fluid LP specification").
// The solver is written once against a "group" policy: FjWarpGroup (the 32 lanes of the
// env's own warp, scratch in global memory: the in-line fallback) or FjCtaGroup (a whole
// CTA, basis inverse in shared memory: the LP kernel).  Both perform the identical
// floating-point operations; only the work split differs.
#define FJ_LP_EPS_D 1e-9
#define FJ_LP_EPS_PIV 1e-9
#define FJ_LP_EPS_ZERO 1e-9

// Group reductions.  argmin() is the lexicographic minimum of (key, idx) over the group's
// candidates (idx 0x7fffffff = none; candidate idx values are distinct), `aux` travels with the
// winner; every member gets the result, and the call orders the group's earlier memory writes
// before its later reads (it contains a group barrier).  On the device the pair is reduced with
// three redux.sync minima over an order-preserving integer image of the double (keys are never
// NaN or -0.0 here): the first version's shuffle butterfly plus a serial scan of the per-warp
// partials by every thread was a third of all LP instructions (profiles/README.md r01_v5).
#ifdef __CUDACC__
FJ_FN void fj_lex_pack(double key, int idx, int aux, unsigned &hi, unsigned &lo, unsigned &id)
{
    const unsigned long long u = fj_ord_enc(key);
    const bool none = idx == FJ_EMPTY;
    hi = none ? 0xffffffffu : (unsigned)(u >> 32);
    lo = none ? 0xffffffffu : (unsigned)u;
    id = none ? 0xffffffffu : (((unsigned)idx << 12) | (unsigned)aux);
}
FJ_FN void fj_lex_unpack(unsigned hi, unsigned lo, unsigned id, double &key, int &idx, int &aux)
{
    if (id == 0xffffffffu) { idx = FJ_EMPTY; return; }
    key = fj_ord_dec(hi, lo); idx = (int)(id >> 12); aux = (int)(id & 0xfffu);
}
#endif

struct FjWarpGroup {
    FJ_MFN int rank() const { return fj_lane(); }
    FJ_MFN int size() const { return FJ_NL; }
    FJ_MFN int lane() const { return fj_lane(); }
    FJ_MFN int warp() const { return 0; }
    FJ_MFN int nwarps() const { return 1; }
    FJ_MFN void sync() const { fj_sync(); }
    FJ_MFN int min_i(int v) { fj_sync(); return fj_min_i(v); }
    FJ_MFN void argmin(double &key, int &idx, int &aux)
    {
        fj_sync();
#ifdef FJ_DEVICE_CODE
        unsigned hi, lo, id;
        fj_lex_pack(key, idx, aux, hi, lo, id);
        fj_warp_lexmin(hi, lo, id);
        fj_lex_unpack(hi, lo, id, key, idx, aux);
#endif
    }
};

#ifdef __CUDACC__
// A group of whole warps of one CTA that synchronises on its own named barrier: the whole CTA
// (base 0, blockDim.x threads, barrier 0 = __syncthreads) in the LP kernel, or one LP group of a server
// CTA of the step kernel (barriers FJ_BAR_SRV0 + group for the group, + 6 for its driver warps).
struct FjCtaGroup {
    int4 *red;    // shared scratch: two buffers of one entry per warp (<= 32 warps)
    int flip;     // which buffer the next reduction uses (same value in every thread)
    int base, nthr, bar;   // first thread, threads (multiple of 32), hardware barrier id
    int dbar;              // a second barrier id of the group (the LP fast path's driver warps)
    FJ_MFN void whole_cta() { base = 0; nthr = (int)blockDim.x; bar = 0; dbar = 1; }
    FJ_MFN int rank() const { return (int)threadIdx.x - base; }
    FJ_MFN int size() const { return nthr; }
    FJ_MFN int lane() const { return threadIdx.x & 31; }
    FJ_MFN int warp() const { return ((int)threadIdx.x - base) >> 5; }
    FJ_MFN int nwarps() const { return nthr >> 5; }
    FJ_MFN void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthr) : "memory"); }
    // one barrier per reduction: the partials of consecutive reductions alternate between two
    // buffers, so a buffer is rewritten only after every thread has passed the next barrier
    FJ_MFN int min_i(int v)
    {
        v = (int)__reduce_min_sync(0xffffffffu, (unsigned)v);   // callers pass non-negative values
        int4 *buf = red + flip * 32; flip ^= 1;
        if (lane() == 0) buf[warp()].x = v;
        sync();
        const int l = lane();
        const unsigned o = l < nwarps() ? (unsigned)buf[l].x : 0xffffffffu;
        return (int)__reduce_min_sync(0xffffffffu, o);
    }
    FJ_MFN void argmin(double &key, int &idx, int &aux)
    {
        unsigned hi, lo, id;
        fj_lex_pack(key, idx, aux, hi, lo, id);
        fj_warp_lexmin(hi, lo, id);
        if (nthr == 32) { fj_lex_unpack(hi, lo, id, key, idx, aux); sync(); return; }
        int4 *buf = red + flip * 32; flip ^= 1;
        if (lane() == 0) buf[warp()] = make_int4((int)hi, (int)lo, (int)id, 0);
        sync();
        const int l = lane();
        hi = lo = id = 0xffffffffu;
        if (l < nwarps()) { const int4 e = buf[l]; hi = (unsigned)e.x; lo = (unsigned)e.y; id = (unsigned)e.z; }
        fj_warp_lexmin(hi, lo, id);
        fj_lex_unpack(hi, lo, id, key, idx, aux);
    }
};
#else
struct int4 { int x, y, z, w; };
struct FjCtaGroup {   // host simulation: one thread
    int4 *red; int flip; int base, nthr, bar, dbar;
    FJ_MFN void whole_cta() { base = 0; nthr = 1; bar = 0; dbar = 1; }
    FJ_MFN int rank() const { return 0; }
    FJ_MFN int size() const { return 1; }
    FJ_MFN int lane() const { return 0; }
    FJ_MFN int warp() const { return 0; }
    FJ_MFN int nwarps() const { return 1; }
    FJ_MFN void sync() const {}
    FJ_MFN int min_i(int v) { return v; }
    FJ_MFN void argmin(double &, int &, int &) {}
};
#endif

struct FjLp {
    double *Binv, *xB, *w, *adem, *rate;
    int *basis, *pos, *colq, *colm, *prec;
    int R, C, NP;
};

// bytes of the small arrays (everything but Binv) and carving of a slab
FJ_FN size_t fj_lp_small_bytes(const FjDims &d)
{
    size_t R = d.Rx, C = d.NPx + 1;
    return (2 * R + 2 * C) * 8 + (R + (C + R) + 2 * C + d.KTx) * 4;
}
FJ_FN void fj_lp_carve(FjLp &L, unsigned char *binv, unsigned char *small_, const FjDims &d)
{
    size_t R = d.Rx, C = d.NPx + 1;
    L.Binv = (double *)binv;
    double *p = (double *)small_;
    L.xB = p; p += R;
    L.w = p; p += R;
    L.adem = p; p += C;
    L.rate = p; p += C;
    int *q = (int *)p;
    L.basis = q; q += R;
    L.pos = q; q += C + R;
    L.colq = q; q += C;
    L.colm = q; q += C;
    L.prec = q; q += d.KTx;
}

// the same carving with the dimensions of ONE LP (R rows at most, C columns): used when the
// scratch lives in a CTA's shared memory.  Returns the bytes of the small arrays.
FJ_FN size_t fj_lp_carve_dims(FjLp &L, unsigned char *binv, unsigned char *small_, size_t R, size_t C, size_t KT)
{
    L.Binv = (double *)binv;
    double *p = (double *)small_;
    L.xB = p; p += R;
    L.w = p; p += R;
    L.adem = p; p += C;
    L.rate = p; p += C;
    int *q = (int *)p;
    L.basis = q; q += R;
    L.pos = q; q += C + R;
    L.colq = q; q += C;
    L.colm = q; q += C;
    L.prec = q; q += KT;
    return (size_t)((unsigned char *)q - small_);
}
FJ_FN size_t fj_lp_dims_bytes(size_t R, size_t C, size_t KT)
{
    return R * R * 8 + (2 * R + 2 * C) * 8 + (R + (C + R) + 2 * C + KT) * 4;
}

// sum_k (negate ? -brow[row_k] : brow[row_k]) * val_k of column j in ascending row order
FJ_FN double fj_lp_colvec_dot(const FjCtx &c, const FjLp &L, const double *brow, int j, int negate)
{
    int M = c.M, KT = c.KT;
    double acc = 0.0;
    if (j == L.NP) {   // the t column: +1 in every demand row
        for (int q = 0; q < KT; ++q) { double b = brow[M + q]; acc = fj_add(acc, fj_mul(negate ? -b : b, 1.0)); }
        return acc;
    }
    int q = L.colq[j], m = L.colm[j];
    double b0 = brow[m];
    acc = fj_add(acc, fj_mul(negate ? -b0 : b0, 1.0));
    double b1 = brow[M + q];
    acc = fj_add(acc, fj_mul(negate ? -b1 : b1, L.adem[j]));
    int stage = fj_rjstage(c)[q];
    if (stage > 0 && L.prec[q - 1] >= 0) { double b = brow[L.prec[q - 1]]; acc = fj_add(acc, fj_mul(negate ? -b : b, L.rate[j])); }
    if (L.prec[q] >= 0) { double b = brow[L.prec[q]]; acc = fj_add(acc, fj_mul(negate ? -b : b, -L.rate[j])); }
    return acc;
}

#if defined(FJ_TRACE) && defined(FJ_DEVICE_CODE)
// per-phase cycles of the CTA-group LP, accumulated by thread 0 in registers and added to trace
// row 15 of its CTA when the LP ends: setup, pricing, entering argmin, w + ratios, leaving
// argmin, xB / pivot row, rank-1 update, iterations
#define FJ_LPT_DECL long long lt_ = clock64(), la_[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define FJ_LPT(k) do { const long long n_ = clock64(); la_[k] += n_ - lt_; lt_ = n_; } while (0)
#define FJ_LPT_ITER() (la_[7] += 1)
#define FJ_LPT_FLUSH() do { if (g.rank() == 0 && g.size() > FJ_NL && fj_sP.trace) { \
    for (int k_ = 0; k_ < 8; ++k_) atomicAdd((unsigned long long *)&fj_sP.trace[((size_t)blockIdx.x * FJ_TRACE_ROWS + FJ_TRACE_ROWS - 1) * 8 + k_], (unsigned long long)la_[k_]); } } while (0)
#else
#define FJ_LPT_DECL
#define FJ_LPT(k)
#define FJ_LPT_ITER()
#define FJ_LPT_FLUSH()
#endif

// Builds the canonical LP from the env record (after fj_arrival_begin) and solves it.
// On return x[0..NP) holds the structural solution (values < 1e-9 flushed to 0).
template <class G>
FJ_FN int fj_lp_solve(const G &g_in, FjCtx &c, FjLp &L, double *x_out, int *iters_out)
{
    G g = g_in;
    const int tid = g.rank(), nt = g.size();
    const int M = c.M, KT = c.KT, Mx = c.Mx;
    const FjRO ptime = FJ_I(c, ptime);
    const FjEligRO elig = fj_elig(c);
    const FjLastRO rjlast = fj_rjlast(c);
    const FjRO colbase = FJ_I(c, colbase);
    FJ_LPT_DECL;
    if (tid == 0) {   // precedence rows: sequential numbering
        int nprec = 0;
        for (int q = 0; q < KT; ++q) {
            L.prec[q] = -1;
            if (!rjlast[q] && c.qlen[q + 1] == 0) L.prec[q] = M + KT + nprec++;
        }
        L.pos[0] = nprec;
    }
    g.sync();
    const int NP = FJ_I(c, hdr)[7], R = M + KT + L.pos[0], C = NP + 1;
    g.sync();
    L.NP = NP; L.R = R; L.C = C;
    for (int q = tid; q < KT; q += nt) {
        unsigned em = (unsigned)elig[q];
        int col = colbase[q];
        double fs = (double)c.fstart[q];
        while (em) {
            int m = fj_ffs0(em); em &= em - 1;
            double rate = fj_div(1.0, (double)ptime[q * Mx + m]);
            L.colq[col] = q; L.colm[col] = m;
            L.rate[col] = rate;
            L.adem[col] = -fj_div(rate, fs);
            ++col;
        }
    }
    for (int j = tid; j < C + R; j += nt) L.pos[j] = j >= C ? j - C : -1;
    for (int i = tid; i < R; i += nt) { L.basis[i] = C + i; L.xB[i] = i < M ? 1.0 : 0.0; }
    for (int e = tid; e < R * R; e += nt) L.Binv[e] = (e / R == e % R) ? 1.0 : 0.0;
    g.sync();
    const int dantzig_iters = 20 * R + 100, hard_iters = 200 * R + 1000;
    const int nvar = C + R, t_col = NP;
    // rank-1 update: warps own contiguous row blocks (four rows in flight), lanes own columns
    const int rpw = (R + g.nwarps() - 1) / g.nwarps();
    const int row_lo = g.warp() * rpw, row_hi = row_lo + rpw < R ? row_lo + rpw : R;
    int it = 0, rc = 0;
    FJ_LPT(0);
    for (;; ++it) {
        FJ_LPT_ITER();
        if (it >= hard_iters) { rc = 2; break; }
        const int pt = L.pos[t_col];
        const double *yrow = L.Binv + (size_t)(pt >= 0 ? pt : 0) * R;
        const int bland = it >= dantzig_iters;
        // pricing: most negative reduced cost (lowest column on ties) / Bland: lowest column
        double ek = 0.0; int ei = FJ_EMPTY, ea = 0;
        for (int j = tid; j < nvar; j += nt) {
            if (L.pos[j] >= 0) continue;
            double d;
            if (j < C) {
                double acc = pt >= 0 ? fj_lp_colvec_dot(c, L, yrow, j, 1) : 0.0;
                d = fj_sub((j == t_col) ? -1.0 : 0.0, acc);
            } else {
                d = pt >= 0 ? -(-yrow[j - C]) : -0.0;
            }
            if (d < -FJ_LP_EPS_D) {
                if (bland) { if (ei == FJ_EMPTY) { ek = d; ei = j; } }
                else if (ei == FJ_EMPTY || d < ek) { ek = d; ei = j; }
            }
        }
        FJ_LPT(1);
        if (bland) ei = g.min_i(ei); else g.argmin(ek, ei, ea);
        FJ_LPT(2);
        const int qin = ei;
        if (qin == FJ_EMPTY) break;   // optimal
        // w = Binv * A_q and, by the same thread, the row's ratio  max(xB,0)/w  (w > eps);
        // ties -> lowest basic variable
        double rk = 0.0; int ri = FJ_EMPTY, rrow = 0;
        for (int i = tid; i < R; i += nt) {
            const double *brow = L.Binv + (size_t)i * R;
            const double wi = qin < C ? fj_lp_colvec_dot(c, L, brow, qin, 0) : brow[qin - C];
            L.w[i] = wi;
            if (wi > FJ_LP_EPS_PIV) {
                const double xb = L.xB[i] > 0.0 ? L.xB[i] : 0.0;
                const double r = fj_div(xb, wi);
                const int bi = L.basis[i];
                if (ri == FJ_EMPTY || r < rk || (r == rk && bi < ri)) { rk = r; ri = bi; rrow = i; }
            }
        }
        FJ_LPT(3);
        g.argmin(rk, ri, rrow);    // also publishes w
        FJ_LPT(4);
        if (ri == FJ_EMPTY) { rc = 3; break; }
        const int p = rrow;
        const double theta = rk, wp = L.w[p];
        for (int i = tid; i < R; i += nt)
            L.xB[i] = (i == p) ? theta : fj_sub(L.xB[i], fj_mul(theta, L.w[i]));
        double *rowp = L.Binv + (size_t)p * R;
        for (int k = tid; k < R; k += nt) rowp[k] = fj_div(rowp[k], wp);
        if (tid == 0) {
            L.pos[L.basis[p]] = -1;
            L.basis[p] = qin;
            L.pos[qin] = p;
        }
        g.sync();
        FJ_LPT(5);
        // every other row:  B[i][k] - w[i] * B[p][k]  with separately rounded multiply and
        // subtract; rows with w[i] == 0 are left alone
        for (int i0 = row_lo; i0 < row_hi; i0 += 4) {
            double wv0 = 0.0, wv1 = 0.0, wv2 = 0.0, wv3 = 0.0;
            if (i0 + 0 < row_hi && i0 + 0 != p) wv0 = L.w[i0 + 0];
            if (i0 + 1 < row_hi && i0 + 1 != p) wv1 = L.w[i0 + 1];
            if (i0 + 2 < row_hi && i0 + 2 != p) wv2 = L.w[i0 + 2];
            if (i0 + 3 < row_hi && i0 + 3 != p) wv3 = L.w[i0 + 3];
            if (wv0 == 0.0 && wv1 == 0.0 && wv2 == 0.0 && wv3 == 0.0) continue;
            double *r0 = L.Binv + (size_t)(i0 + 0) * R, *r1 = L.Binv + (size_t)(i0 + 1) * R;
            double *r2 = L.Binv + (size_t)(i0 + 2) * R, *r3 = L.Binv + (size_t)(i0 + 3) * R;
            for (int k = g.lane(); k < R; k += FJ_NL) {
                const double pk_ = rowp[k];
                double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
                if (wv0 != 0.0) e0 = r0[k];
                if (wv1 != 0.0) e1 = r1[k];
                if (wv2 != 0.0) e2 = r2[k];
                if (wv3 != 0.0) e3 = r3[k];
                if (wv0 != 0.0) r0[k] = fj_sub(e0, fj_mul(wv0, pk_));
                if (wv1 != 0.0) r1[k] = fj_sub(e1, fj_mul(wv1, pk_));
                if (wv2 != 0.0) r2[k] = fj_sub(e2, fj_mul(wv2, pk_));
                if (wv3 != 0.0) r3[k] = fj_sub(e3, fj_mul(wv3, pk_));
            }
        }
        g.sync();
        FJ_LPT(6);
    }
    for (int j = tid; j < NP; j += nt) {
        double x = L.pos[j] >= 0 ? L.xB[L.pos[j]] : 0.0;
        if (x < FJ_LP_EPS_ZERO) x = 0.0;
        x_out[j] = x;
    }
    g.sync();
    FJ_LPT_FLUSH();
    if (iters_out) *iters_out = it;
    return rc;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------- group fast path of the LP
// The same pivoting rules and the same floating-point operations as fj_lp_solve (results are
// bit-identical; tests/test_gpu_parity.py compares both with the oracle), organised around what
// bounds a small dense simplex on an SM: the LENGTH OF THE DEPENDENT CHAIN of one iteration.  A
// warp retires one dependent instruction per ~10 cycles, a CTA-wide argmin costs two warp
// reductions, a trip through shared memory and a barrier, and with "thread i owns row i" most
// phases run one element per thread: the round-1/2 team version spent 11.8 k cycles per iteration
// on ~600 dependent instructions (profiles/README.md).  Here
//   * ONE warp -- the DRIVER -- runs the whole decision chain of an iteration: pricing over all
//     columns (a lane owns columns lane, lane + 32, ...: independent 4-term dot products, so the
//     loads and the multiply-add chains of several columns are in flight together), the entering
//     argmin (three redux.sync, no shared-memory exchange, no barrier), w = B^-1 A_q and the ratio
//     test for rows lane, lane + 32, ..., the leaving argmin, x_B, the scaled pivot row and the new
//     pricing vector;
//   * the group's other warps -- the HELPERS -- apply the rank-1 update of the pivot the driver has
//     just published while the driver already prices the next iteration (pricing needs only the
//     mirrored t row y of B^-1, which the driver updates itself);
//   * two group barriers per iteration tie them together: A "update of pivot n-1 done" (before the
//     driver reads B^-1 for w) and B "pivot n published";
//   * every structural column carries a descriptor: its four row indices (machine row, demand
//     row, the two precedence rows; a missing row points at a padding column of B^-1 that is
//     always 0, which adds +-0 and leaves the accumulator unchanged) and its two coefficients:
//     a reduced cost or an entry of w is 4 loads, 4 multiplies, 4 adds, no branch;
//   * B^-1 is stored COLUMN-major with an odd column stride: "lane i reads B^-1[i][k]" and "lane k
//     reads B^-1[p][k]" are conflict-free shared-memory accesses (coalesced when the scratch is
//     global); the update skips 32-row groups whose w is all zero and column quadruples whose
//     pivot-row entries are all zero (x - w * 0 = x: only the sign of a zero could differ, and a
//     zero's sign never reaches a comparison, a non-zero value or the flushed solution).
// tells the compiler that p points into shared memory (a round trip through the shared window: the generic
// pointer came through a noinline call or a select and lost its address space; with it loads / stores are
// LDS / STS with 32-bit addresses instead of generic LD / ST with 64-bit address arithmetic)
template <typename T> FJ_FN T *fj_as_shared(T *p)
{
    return (T *)__cvta_shared_to_generic(__cvta_generic_to_shared(p));
}

// per-group state of the fast path, carved from the group's shared memory for the batch's largest LP
// (rows <= Rx): w = B^-1 A_q, the scaled pivot row, the pricing vector y (+ the zero padding entry),
// x_B, the basic variable of every row, the precedence row of every operation type, control words
struct FjLpFastSmem {
    double *base; int n, ktx;   // n = Rx + 2 entries per double array
    FJ_MFN double *w() const { return base; }
    FJ_MFN double *pr() const { return base + n; }
    FJ_MFN double *y() const { return base + 2 * n; }
    FJ_MFN double *xb() const { return base + 3 * n; }
    FJ_MFN int *bvar() const { return (int *)(base + 4 * n); }
    FJ_MFN int *ctl() const { return bvar() + n; }                 // [0] nprec [1] pivot row for the helpers (-1: leave) [2] ticket [3] requester [4..16) argmin partials of the driver warps
    FJ_MFN short *prec() const { return (short *)(ctl() + 16); }
};
FJ_FN size_t fj_lpf_state_bytes(const FjDims &d)
{
    const size_t n = (size_t)d.Rx + 2;
    return (4 * n * 8 + n * 4 + 64 + (size_t)d.KTx * 2 + 15) / 16 * 16;
}
FJ_FN FjLpFastSmem fj_lpf_state(unsigned char *p, const FjDims &d)
{
    FjLpFastSmem S; S.base = fj_as_shared((double *)p); S.n = d.Rx + 2; S.ktx = d.KTx;
    return S;
}

// What the LP is built from: the instance's statics and, of the environment's state at the order
// arrival, the unprocessed operations per type and which waiting queues are empty.  Either the env
// record itself (qlen != null: the LP kernel after reset() / for parked envs) or the request an env
// warp posted for the LP servers (empty != null; written by another SM: read through L2).
struct FjLpIn {
    const int32_t *I;
    int M, KT, NP;
    const int32_t *fstart;
    const uint16_t *qlen;
    const uint32_t *empty;
};
FJ_FN int fj_lpin_fstart(const FjLpIn &in, int q) { return in.qlen ? in.fstart[q] : __ldcg(in.fstart + q); }
FJ_FN int fj_lpin_prec(const FjLpIn &in, const FjParams &P, int q)   // a precedence row: not a last stage and nothing waits at the next one
{
    if (__ldg(in.I + P.io.rjlast + q)) return 0;
    if (in.qlen) return in.qlen[q + 1] == 0;
    return (int)(__ldcg(in.empty + ((q + 1) >> 5)) >> ((q + 1) & 31) & 1u);
}
FJ_FN void fj_lpin_from_ctx(FjLpIn &in, const FjCtx &c)
{
    in.I = c.I; in.M = c.M; in.KT = c.KT; in.NP = c.I[c.P->io.hdr + 7];
    in.fstart = c.fstart; in.qlen = c.qlen; in.empty = nullptr;
}

// The 4-term dot product of a structural column with a row-indexed vector v (the pricing vector, or a
// row of B^-1):  ((v0 + v1 * a) + v2 * r) + v3 * (-r),  every multiply and add rounded separately.
// The specification (oracle/fjsp_lp.c) writes it  (((0 + v0 * 1) + v1 * a) + v2 * r) + v3 * (-r)  and, for the
// reduced cost,  0 - sum((-y) * coefficient):  x * 1 = x and 0 + x = x exactly, and rounding to nearest is
// symmetric under negation, so this is the same number (only the sign of a zero can differ, and a zero's
// sign never reaches a comparison, a non-zero value or the flushed solution).
FJ_FN double fj_lpf_dot4(double v0, double v1, double v2, double v3, double a, double r)
{
    double acc = fj_add(v0, fj_mul(v1, a));
    acc = fj_add(acc, fj_mul(v2, r));
    return fj_add(acc, fj_mul(v3, -r));
}
// column descriptor: the four rows of the column as BYTE offsets into a row-indexed double vector
// (row * 8, 16 bits each; a missing precedence row points at the padding entry R, which is always 0);
// bit 0 of the first word: the column is basic
// a / b for b > 0 (exactly __ddiv_rn's result): a zero numerator -- the common case in this degenerate LP: most
// x_B and most entries of a pivot row are 0 -- would send the whole warp through the division's special-operand
// slow path (measured: the ratio test and the pivot-row scaling took 2.6 k and 1.5 k cycles, most of an iteration);
// 0 / b = 0 with the numerator's sign, so those lanes divide 1.0 instead and keep their zero
FJ_FN double fj_lpf_div(double a, double b)
{
    const bool z = a == 0.0;
    const double q = __ddiv_rn(z ? 1.0 : a, b);
    return z ? a : q;
}
FJ_FN double fj_lpf_at(const double *v, unsigned off) { return *(const double *)((const char *)v + (off & 0xfff8u)); }

// rank-1 update of pivot row p by helper h of nh:  B^-1[i][k] -= w[i] * pr[k]  for rows i != p with
// w[i] != 0; lanes own the rows of a 32-row group (w in a register), helpers interleave the columns;
// a quad of columns whose pivot-row entries are all zero is skipped (the pivot row is sparse).
// (Lanes-own-columns with the rows interleaved was measured 20-40 % slower: w is much denser than the
// pivot row, so nothing could be skipped.)
FJ_FN void fj_lpf_rank1(const FjLpFastSmem &S, double *BT, int Rs, int R, int p, int h, int nh, int lane)
{
    const int ngrp = (R + 31) >> 5;
    const double *pr = S.pr();
    for (int grp = 0; grp < ngrp; ++grp) {
        const int i = grp * 32 + lane;
        const double wv = (i < R && i != p) ? S.w()[i] : 0.0;
        const bool on = wv != 0.0;
        if (!__any_sync(0xffffffffu, on)) continue;
        double *b = BT + i;
        int k = h;
        for (; k + 3 * nh < R; k += 4 * nh) {
            const double p0 = pr[k], p1 = pr[k + nh], p2 = pr[k + 2 * nh], p3 = pr[k + 3 * nh];
            if (p0 == 0.0 && p1 == 0.0 && p2 == 0.0 && p3 == 0.0) continue;
            if (on) {
                double *b0 = b + (size_t)k * Rs, *b1 = b0 + (size_t)nh * Rs, *b2 = b1 + (size_t)nh * Rs, *b3 = b2 + (size_t)nh * Rs;
                const double e0 = *b0, e1 = *b1, e2 = *b2, e3 = *b3;
                *b0 = fj_sub(e0, fj_mul(wv, p0));
                *b1 = fj_sub(e1, fj_mul(wv, p1));
                *b2 = fj_sub(e2, fj_mul(wv, p2));
                *b3 = fj_sub(e3, fj_mul(wv, p3));
            }
        }
        for (; k < R; k += nh) {
            const double p0 = pr[k];
            if (p0 != 0.0 && on) { double *b0 = b + (size_t)k * Rs; *b0 = fj_sub(*b0, fj_mul(wv, p0)); }
        }
    }
}

// lexicographic minimum over the group's D driver warps: warp minimum, partials through shared memory,
// one barrier of the drivers, warp minimum of the partials (every driver lane gets the result)
FJ_FN void fj_lpf_lexmin(unsigned &hi, unsigned &lo, unsigned &id, int *part, int dw, int D, int dbar, int lane)
{
    fj_warp_lexmin(hi, lo, id);
    if (D == 1) return;
    if (lane == 0) { part[3 * dw] = (int)hi; part[3 * dw + 1] = (int)lo; part[3 * dw + 2] = (int)id; }
    asm volatile("bar.sync %0, %1;" ::"r"(dbar), "r"(D * 32) : "memory");
    hi = lo = id = 0xffffffffu;
    if (lane < D) { hi = (unsigned)part[3 * lane]; lo = (unsigned)part[3 * lane + 1]; id = (unsigned)part[3 * lane + 2]; }
    fj_warp_lexmin(hi, lo, id);
}

// the simplex iterations (everything is set up: all-slack basis, B^-1 = I).  BINV_SM only tells the
// compiler which address space BT is in.  The group's first D warps are the drivers: driver dw owns the
// columns and rows  lane + 32 (dw + D k);  the other warps are the helpers.
template <bool BINV_SM>
FJ_FN int fj_lpf_iterate(FjCtaGroup g, const FjLpFastSmem &S, double *BT, const double2 *coef, uint2 *cidx, int *pos, int M, int KT, int NP, int R, int *iters_out)
{
    const int tid = g.rank(), lane = tid & 31, wid = tid >> 5, nw = g.size() >> 5;
    const int D = nw >= 12 ? 4 : nw >= 6 ? 2 : 1;
    const int C = NP + 1, Rs = R | 1, t_col = NP;
#ifdef FJ_LP_DEBUG_MAXIT   // timing probe only (tools/lp_probe.py): stop after a few iterations
    const int dantzig_iters = 20 * R + 100, hard_iters = FJ_LP_DEBUG_MAXIT;
#else
    const int dantzig_iters = 20 * R + 100, hard_iters = 200 * R + 1000;
#endif
    int it = 0, rc = 0;
    FJ_LPT_DECL;
    if (wid < D) {
        // ------------------------------------------------ drivers
        const int dw = wid, first = lane + 32 * dw, step = 32 * D;
        int pt = -1;               // row of t in the basis, -1: nonbasic
        unsigned slack_nb = 0;     // bit k: the slack of row first + step k is nonbasic (all slacks start basic)
        double *const y = S.y(), *const w = S.w(), *const xb = S.xb(), *const prow = S.pr();
        int *const bvar = S.bvar(), *const part = S.ctl() + 4;
        for (;; ++it) {
            FJ_LPT_ITER();
            int qin = FJ_EMPTY;
            if (it >= hard_iters) rc = 2;
            else if (pt < 0) qin = t_col;   // y = 0: every structural / slack reduced cost is 0 and t prices at -1
            else {
                // ---- pricing on the mirrored t row: most negative reduced cost, lowest column on ties (Dantzig);
                // after dantzig_iters the lowest column with a negative reduced cost (Bland: every key equal).
                // t itself is basic here.  Two columns per pass: their loads and chains overlap.
                const bool bland = it >= dantzig_iters;
                double ek = 0.0; int ei = FJ_EMPTY;
                int j = first;
                FJ_NOUNROLL
                for (; j + step < NP; j += 2 * step) {
                    const uint2 xa = cidx[j], xc = cidx[j + step];
                    const double2 ca = coef[j], cc = coef[j + step];
                    const double a0 = fj_lpf_at(y, xa.x), a1 = fj_lpf_at(y, xa.x >> 16), a2 = fj_lpf_at(y, xa.y), a3 = fj_lpf_at(y, xa.y >> 16);
                    const double c0 = fj_lpf_at(y, xc.x), c1 = fj_lpf_at(y, xc.x >> 16), c2 = fj_lpf_at(y, xc.y), c3 = fj_lpf_at(y, xc.y >> 16);
                    const double da = fj_lpf_dot4(a0, a1, a2, a3, ca.x, ca.y), dc = fj_lpf_dot4(c0, c1, c2, c3, cc.x, cc.y);
                    const double ka = bland ? -1.0 : da, kc = bland ? -1.0 : dc;
                    if (!(xa.x & 1u) && da < -FJ_LP_EPS_D && (ei == FJ_EMPTY || ka < ek)) { ek = ka; ei = j; }
                    if (!(xc.x & 1u) && dc < -FJ_LP_EPS_D && (ei == FJ_EMPTY || kc < ek)) { ek = kc; ei = j + step; }
                }
                if (j < NP) {
                    const uint2 xa = cidx[j];
                    const double2 ca = coef[j];
                    const double da = fj_lpf_dot4(fj_lpf_at(y, xa.x), fj_lpf_at(y, xa.x >> 16), fj_lpf_at(y, xa.y), fj_lpf_at(y, xa.y >> 16), ca.x, ca.y);
                    const double ka = bland ? -1.0 : da;
                    if (!(xa.x & 1u) && da < -FJ_LP_EPS_D && (ei == FJ_EMPTY || ka < ek)) { ek = ka; ei = j; }
                }
                FJ_NOUNROLL
                for (unsigned nb = slack_nb, k = 0; __any_sync(0xffffffffu, nb != 0u); nb >>= 1, ++k) {
                    if (nb & 1u) {
                        const int i = first + step * (int)k;
                        const double d = y[i], kd = bland ? -1.0 : d;
                        if (d < -FJ_LP_EPS_D && (ei == FJ_EMPTY || kd < ek)) { ek = kd; ei = C + i; }
                    }
                }
                unsigned hi, lo, id; int ea = 0;
                fj_lex_pack(ek, ei, 0, hi, lo, id);
                fj_lpf_lexmin(hi, lo, id, part, dw, D, g.dbar, lane);
                fj_lex_unpack(hi, lo, id, ek, ei, ea);
                qin = ei;
            }
            FJ_LPT(1);
            g.sync();   // A: the helpers have applied the previous pivot
            FJ_LPT(2);
            if (qin == FJ_EMPTY) break;   // optimal (or rc = 2)
            // ---- w = B^-1 A_q and the ratio test  min max(xB, 0) / w  over w > eps (ties: lowest basic variable)
            double rk = 0.0; int ri = FJ_EMPTY, rrow = 0;
            if (qin < NP) {
                const uint2 ix = cidx[qin];
                const double2 cf = coef[qin];
                const double *c0 = BT + (size_t)((ix.x & 0xfff8u) >> 3) * Rs, *c1 = BT + (size_t)((ix.x >> 16 & 0xfff8u) >> 3) * Rs;
                const double *c2 = BT + (size_t)((ix.y & 0xfff8u) >> 3) * Rs, *c3 = BT + (size_t)((ix.y >> 16 & 0xfff8u) >> 3) * Rs;
                FJ_NOUNROLL
                for (int i = first; i < R; i += 2 * step) {
                    // two rows per pass (the second may be past the end: it is computed on row i again and dropped)
                    const int i2 = i + step < R ? i + step : i;
                    const double w1 = fj_lpf_dot4(c0[i], c1[i], c2[i], c3[i], cf.x, cf.y);
                    const double w2 = fj_lpf_dot4(c0[i2], c1[i2], c2[i2], c3[i2], cf.x, cf.y);
                    const double x1 = xb[i], x2 = xb[i2];
                    const int b1 = bvar[i], b2 = bvar[i2];
                    w[i] = w1;
                    if (i2 != i) w[i2] = w2;
                    const bool t1 = w1 > FJ_LP_EPS_PIV, t2 = i2 != i && w2 > FJ_LP_EPS_PIV;
                    if (__any_sync(__activemask(), t1 || t2)) {
                        const double r1 = fj_lpf_div(x1 > 0.0 ? x1 : 0.0, t1 ? w1 : 1.0), r2 = fj_lpf_div(x2 > 0.0 ? x2 : 0.0, t2 ? w2 : 1.0);
                        if (t1 && (ri == FJ_EMPTY || r1 < rk || (r1 == rk && b1 < ri))) { rk = r1; ri = b1; rrow = i; }
                        if (t2 && (ri == FJ_EMPTY || r2 < rk || (r2 == rk && b2 < ri))) { rk = r2; ri = b2; rrow = i2; }
                    }
                }
            } else {
                // t (sum of the demand-row columns, ascending) or a slack (one column of B^-1) enters.  t enters
                // OFTEN: a degenerate ratio test (ratio 0) is won by the lowest basic variable, t ranks below
                // every slack, so t keeps leaving and re-entering.  Its w is a sum in ascending operation-type
                // order per row: the adds are a serial chain, the loads are not -- eight in flight.
                FJ_NOUNROLL
                for (int i = first; i < R; i += step) {
                    const double *brow = BT + i;
                    double wi;
                    if (qin == NP) {
                        double acc = 0.0;
                        const double *b = brow + (size_t)M * Rs;
                        int q = 0;
                        for (; q + 8 <= KT; q += 8, b += (size_t)8 * Rs) {
                            const double v0 = b[0], v1 = b[Rs], v2 = b[2 * (size_t)Rs], v3 = b[3 * (size_t)Rs];
                            const double v4 = b[4 * (size_t)Rs], v5 = b[5 * (size_t)Rs], v6 = b[6 * (size_t)Rs], v7 = b[7 * (size_t)Rs];
                            acc = fj_add(fj_add(fj_add(fj_add(fj_add(fj_add(fj_add(fj_add(acc, v0), v1), v2), v3), v4), v5), v6), v7);
                        }
                        for (; q < KT; ++q, b += Rs) acc = fj_add(acc, b[0]);
                        wi = acc;
                    } else wi = brow[(size_t)(qin - C) * Rs];
                    w[i] = wi;
                    if (wi > FJ_LP_EPS_PIV) {
                        const double xv = xb[i];
                        const double r = fj_lpf_div(xv > 0.0 ? xv : 0.0, wi);
                        const int bi = bvar[i];
                        if (ri == FJ_EMPTY || r < rk || (r == rk && bi < ri)) { rk = r; ri = bi; rrow = i; }
                    }
                }
            }
            FJ_LPT(3);
            {
                unsigned hi, lo, id;
                fj_lex_pack(rk, ri, rrow, hi, lo, id);
                __syncwarp();
                fj_lpf_lexmin(hi, lo, id, part, dw, D, g.dbar, lane);
                fj_lex_unpack(hi, lo, id, rk, ri, rrow);
            }
            __syncwarp();   // every lane's w is visible to the warp (other drivers' rows: the drivers' barrier above)
            FJ_LPT(4);
            if (ri == FJ_EMPTY) { rc = 3; break; }
            const int p = rrow;
            const double theta = rk, wp = w[p];
            const int pt_new = qin == t_col ? p : (ri == t_col ? -1 : pt);
            const double wt = (pt_new >= 0 && pt_new != p) ? w[pt_new] : 0.0;
            const double *bp = BT + p;      // row p: entry k at bp[k * Rs]
            FJ_NOUNROLL
            for (int i = first; i < R; i += 2 * step) {
                // x_B, the scaled pivot row (entry i of it) and the t row of B^-1 for the next pricing; two entries per pass
                const int i2 = i + step < R ? i + step : i;
                const double e1 = bp[(size_t)i * Rs], e2 = bp[(size_t)i2 * Rs];
                const double x1 = xb[i], x2 = xb[i2], w1 = w[i], w2 = w[i2], y1 = y[i], y2 = y[i2];
                double p1 = e1, p2 = e2;       // 0 / wp = 0 (wp > 0): a pass of zeros needs no division
                if (__any_sync(__activemask(), e1 != 0.0 || e2 != 0.0)) { p1 = fj_lpf_div(e1, wp); p2 = fj_lpf_div(e2, wp); }
                const double nx1 = i == p ? theta : fj_sub(x1, fj_mul(theta, w1)), nx2 = i2 == p ? theta : fj_sub(x2, fj_mul(theta, w2));
                double n1 = 0.0, n2 = 0.0;
                if (pt_new == p) { n1 = p1; n2 = p2; }
                else if (pt_new >= 0) { n1 = wt != 0.0 ? fj_sub(y1, fj_mul(wt, p1)) : y1; n2 = wt != 0.0 ? fj_sub(y2, fj_mul(wt, p2)) : y2; }
                xb[i] = nx1; BT[(size_t)i * Rs + p] = p1; prow[i] = p1; y[i] = n1;
                if (i == p) bvar[i] = qin;
                if (i2 != i) {
                    xb[i2] = nx2; BT[(size_t)i2 * Rs + p] = p2; prow[i2] = p2; y[i2] = n2;
                    if (i2 == p) bvar[i2] = qin;
                }
            }
            // bookkeeping: positions (for the final read-out), basic flags of the structural columns, nonbasic slacks
            if (tid == 0) {
                pos[ri] = -1; pos[qin] = p; S.ctl()[1] = p;
                if (qin < NP) cidx[qin].x |= 1u;
                if (ri < NP) cidx[ri].x &= ~1u;
            }
            if (ri >= C) { const int s_ = ri - C, g_ = s_ >> 5; if ((s_ & 31) == lane && g_ % D == dw) slack_nb |= 1u << (g_ / D); }
            if (qin >= C) { const int s_ = qin - C, g_ = s_ >> 5; if ((s_ & 31) == lane && g_ % D == dw) slack_nb &= ~(1u << (g_ / D)); }
            pt = pt_new;
            FJ_LPT(5);
            if (nw == 1) { __syncwarp(); fj_lpf_rank1(S, BT, Rs, R, p, 0, 1, lane); __syncwarp(); }
            g.sync();   // B: pivot published
            FJ_LPT(6);
        }
        if (tid == 0) S.ctl()[1] = -1;
        g.sync();       // B of the last round: the helpers leave
    } else {
        // ------------------------------------------------ helpers
        for (;;) {
            g.sync();   // A
            g.sync();   // B
            const int p = *(volatile int *)&S.ctl()[1];
            if (p < 0) break;
            fj_lpf_rank1(S, BT, Rs, R, p, wid - D, nw - D, lane);
        }
    }
    FJ_LPT_FLUSH();
    if (iters_out) *iters_out = it;
    return rc;
}

// returns -1 when the LP does not fit the fast path (the caller then runs fj_lp_solve).
// `smem` / `smem_bytes`: shared-memory scratch of the calling group (a server group of the step kernel, the
// LP kernel's CTA): the column descriptors and positions live there, and B^-1 too when it still fits (else on
// `slab` in HBM/L2).
FJ_FN int fj_lp_solve_fast(const FjCtaGroup &g_in, const FjParams &P, const FjLpIn &in, const FjLpFastSmem &S, unsigned char *slab, double *x_out,
                           int *iters_out, unsigned char *smem, int smem_bytes)
{
    FjCtaGroup g = g_in;
    const int tid = g.rank(), nt = g.size();
    const int M = in.M, KT = in.KT;
    if (tid < 32) {   // precedence rows, numbered in ascending operation-type order
        int base = 0;
        for (int q0 = 0; q0 < KT; q0 += 32) {
            const int q = q0 + tid;
            const int f = q < KT && fj_lpin_prec(in, P, q);
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (q < KT) S.prec()[q] = (short)(f ? M + KT + base + __popc(bal & ((1u << tid) - 1u)) : -1);
            base += __popc(bal);
        }
        if (tid == 0) S.ctl()[0] = base;
    }
    g.sync();
    const int NP = in.NP, R = M + KT + S.ctl()[0], C = NP + 1, Rs = R | 1;   // odd column stride
    // carve: column descriptors, positions, then B^-1 (R + 1 columns of Rs entries: column R is the all-zero padding)
    const size_t binv_bytes = ((size_t)(R + 1) * Rs * 8 + 15) / 16 * 16;
    const size_t small_bytes = ((size_t)C * 24 + (size_t)((C + R + 1) & ~1) * 4 + 15) / 16 * 16;
    if (R + 2 > S.n || R >= 0xfff || !smem || (size_t)smem_bytes < small_bytes) { g.sync(); return -1; }
    const bool binv_sm = (size_t)smem_bytes >= small_bytes + binv_bytes;
    double2 *coef = fj_as_shared((double2 *)smem);
    uint2 *cidx = (uint2 *)(coef + C);
    int *pos = (int *)(cidx + C);
    double *BT = (double *)(binv_sm ? smem + small_bytes : slab);
    {   // one thread per structural column: descriptor and coefficients (the per-pair rate 1/p is a static of the instance)
        const int32_t *colqm = in.I + P.io.colqm, *rjstage = in.I + P.io.rjstage;
        const double *colrate = (const double *)(in.I + P.io.colrate);
        const short *prec = S.prec();
        for (int j = tid; j < NP; j += nt) {
            const int qm = __ldg(colqm + j), q = qm >> 8, m = qm & 0xff;
            const double rate = __ldg(colrate + j);
            const double fs = (double)fj_lpin_fstart(in, q);
            const int p_prev = (__ldg(rjstage + q) > 0 && prec[q - 1] >= 0) ? prec[q - 1] : R;
            const int p_own = prec[q] >= 0 ? prec[q] : R;
            coef[j] = make_double2(-__ddiv_rn(rate, fs), rate);
            cidx[j] = make_uint2(((unsigned)m << 3) | ((unsigned)(M + q) << 19), ((unsigned)p_prev << 3) | ((unsigned)p_own << 19));
        }
    }
    for (int j = tid; j < C + R; j += nt) pos[j] = j >= C ? j - C : -1;
    for (int e = tid; e < (R + 1) * Rs; e += nt) BT[e] = 0.0;
    for (int k = tid; k < R + 2; k += nt) S.y()[k] = 0.0;
    for (int i = tid; i < R; i += nt) { S.xb()[i] = i < M ? 1.0 : 0.0; S.bvar()[i] = C + i; }
    g.sync();
    for (int i = tid; i < R; i += nt) BT[(size_t)i * Rs + i] = 1.0;
    g.sync();
    int rc;
    if (binv_sm) rc = fj_lpf_iterate<true>(g, S, fj_as_shared(BT), coef, cidx, pos, M, KT, NP, R, iters_out);
    else rc = fj_lpf_iterate<false>(g, S, BT, coef, cidx, pos, M, KT, NP, R, iters_out);
    for (int j = tid; j < NP; j += nt) {
        double x = pos[j] >= 0 ? S.xb()[pos[j]] : 0.0;
        if (x < FJ_LP_EPS_ZERO) x = 0.0;
        x_out[j] = x;
    }
    g.sync();
    return rc;
}
#endif

// class_FJSP.py:218-248 reset_object_add up to the LP: the new order's jobs join the
// counters, fluid start counts are taken, per-pair fluid bookkeeping is cleared.
FJ_FN void fj_arrival_begin(FjCtx &c, int s)
{
    const int lane = fj_lane();
    const int KT = c.KT, Mx = c.Mx, Sx = c.Sx, Kx = c.Kx;
    const FjRO count = FJ_I(c, count);
    const FjKindRO rjkind = fj_rjkind(c);
    const FjStageRO rjstage = fj_rjstage(c);
    FJ_NOUNROLL
    for (int q = lane; q < KT; q += FJ_NL) {
        int cnt = count[s * Kx + rjkind[q]];
        c.cntunp[q * Sx + s] = (uint16_t)cnt;
        if (rjstage[q] == 0) {
            c.cntnow[q * Sx + s] = (uint16_t)cnt;
            int ql = c.qlen[q] + cnt;
            if (ql > 65000) c.scal[FJ_S_ERROR] |= FJ_E_OVERFLOW;
            c.qlen[q] = (uint16_t)ql;
        }
        int fs = 0;
        FJ_NOUNROLL
        for (int k = 0; k < c.S; ++k) fs += c.cntunp[q * Sx + k];
        c.fstart[q] = fs;
        c.flmask[q] = 0;
    }
    FJ_NOUNROLL
    for (int i = lane; i < KT * Mx; i += FJ_NL) { c.pk[i] = 0; c.slot[i] = 0xFFFF; }
    FJ_NOUNROLL
    for (int m = lane; m < c.M; m += FJ_NL) { c.mD[m] = 0; c.mF[m] = 0.0; }
    if (lane == 0) {
        int add = 0;
        FJ_NOUNROLL
        for (int r = 0; r < c.K; ++r) add += count[s * Kx + r];
        c.scal[FJ_S_LEFT] += add;
    }
    if (c.P->variant == FJSP_SO_FJSSP) {
        // class_FJSSP.py:214-218: every job has its own due date
        //   r_due = round(delivery * len(tasks) / count);  due(n) = round(r_due * n / count)
        // (Python round() of a float = round-half-even = rint), and the unprocessed sets are
        // kept per job because the due date now varies inside an order
        const FjRO ntask = FJ_I(c, ntask);
        const FjDueRO cum = fj_cum(c), jobbase = fj_jobbase(c);
        const FjDueRO due = fj_due(c);
        const int NWx = c.P->d.NWx;
        FJ_NOUNROLL
        for (int q = lane; q < KT; q += FJ_NL) {
            const int r = rjkind[q], n0 = cum[s * Kx + r], n1 = cum[(s + 1) * Kx + r];
            FJ_NOUNROLL
            for (int n = n0; n < n1; ++n) c.unpmask[q * NWx + (n >> 5)] |= 1u << (n & 31);
        }
        FJ_NOUNROLL
        for (int r = 0; r < c.K; ++r) {
            const int n0 = cum[s * Kx + r], n1 = cum[(s + 1) * Kx + r], cnt = n1 - n0;
            const long long r_due = (long long)rint(fj_div((double)((long long)due[s] * ntask[r]), (double)cnt));
            FJ_NOUNROLL
            for (int n = n0 + lane; n < n1; n += FJ_NL)
                c.duejob[jobbase[r] + n] = (int)rint(fj_div((double)(r_due * n), (double)cnt));
        }
    }
    fj_sync();
}

FJ_OUTLINE unsigned fj_small_set_order(unsigned seq, int n);

// exclusive prefix sum of one int per lane (and the warp total)
FJ_FN int fj_excl_scan_i(int v, int &total)
{
#ifdef FJ_DEVICE_CODE
    int inc = v;
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if (fj_lane() >= d) inc += o; }
    total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
#else
    total = v;
    return 0;
#endif
}
FJ_FN double fj_ld_x(const double *x, int j)   // an LP solution entry: possibly written by another SM, so never from L1
{
#ifdef FJ_DEVICE_CODE
    return __ldcg(x + j);
#else
    return x[j];
#endif
}
FJ_FN double fj_inst_d(const FjCtx &c, int off, int i)   // double i of an 8-byte aligned array of the instance record
{
    const double *p = (const double *)(c.I + off) + i;
#ifdef FJ_DEVICE_CODE
    return __ldg(p);
#else
    return *p;
#endif
}

// class_FJSP.py:292-316 update_fluid_parameter from the LP solution x (canonical columns).  Lanes own
// operation types; the fluid slots (canonical column order) are numbered with a warp prefix sum and
// filled by their types' lanes (one division per fluid pair, all types at once).
template <int SUM_MODE>
FJ_FN void fj_arrival_finish(FjCtx &c, const double *x, int iters, int rc)
{
    const int lane = fj_lane();
    const int KT = c.KT, Mx = c.Mx, NFx = c.P->d.NFx;
    const FjRO poord = FJ_I(c, poord), nelig = FJ_I(c, nelig);
    const FjEligRO elig = fj_elig(c);
    const FjRO colbase = FJ_I(c, colbase);
    const int rate_off = c.P->io.colrate;   // 1.0 / processing time of every column
    if (lane == 0) {
        c.scal[FJ_S_LPSOLVES] += 1; c.scal[FJ_S_LPITERS] += iters;
        if (rc) c.scal[FJ_S_ERROR] |= FJ_E_LP;
    }
    int base = 0, overflow = 0;
    FJ_NOUNROLL
    for (int q0 = 0; q0 < KT; q0 += FJ_NL) {
        const int q = q0 + lane;
        unsigned em = 0, fm = 0;
        int nfq = 0, cb = 0;
        double rs = 0.0;
        if (q < KT) {
            em = (unsigned)elig[q]; cb = colbase[q];
            unsigned fseq = 0;
            FjPySum ps; fj_pysum_init(ps);
            FJ_NOUNROLL
            for (int k = 0; k < nelig[q]; ++k) {
                const int m = poord[q * Mx + k];
                const int col = cb + fj_popc(em & ((1u << m) - 1u));
                const double xv = fj_ld_x(x, col);
                const double fr = fj_mul(xv, fj_inst_d(c, rate_off, col));
                fj_pysum_add<SUM_MODE>(ps, fr);
                if (xv != 0.0) { fm |= 1u << m; if (nfq < 4) fseq |= (unsigned)m << (8 * nfq); ++nfq; }
            }
            // set(fluid_machine_list) is built from x.items() order: its iteration order (what machine_select
            // walks) is fixed until the next arrival
            c.h_fo[q] = nfq < 5 ? fj_small_set_order(fseq, nfq) : 0u;
            rs = fj_pysum_result<SUM_MODE>(ps);
            c.rsum[q] = rs;
            c.tsum[q] = fj_div(1.0, rs);
            c.flmask[q] = fm;
        }
        int tot;
        int sl = base + fj_excl_scan_i(nfq, tot);
        base += tot;
        if (base > NFx) overflow = 1;
        if (q < KT && !overflow) {
            const double fs = (double)c.fstart[q];
            FJ_NOUNROLL
            while (fm) {
                const int m = fj_ffs0(fm); fm &= fm - 1;
                const int col = cb + fj_popc(em & ((1u << m) - 1u));
                const double fr = fj_mul(fj_ld_x(x, col), fj_inst_d(c, rate_off, col));
                const double arr = fj_div(fj_mul(fs, fr), rs);
                c.slot[q * Mx + m] = (uint16_t)sl;
                c.ff[sl] = fr; c.fa[sl] = arr; c.fu[sl] = arr;
                ++sl;
            }
        }
    }
    fj_sync();
    // a machine's fluid rate sum: its pairs in ascending operation-type order (canonical column order)
    if (!overflow) {
        FJ_NOUNROLL
        for (int m = lane; m < c.M; m += FJ_NL) {
            double acc = c.mF[m];
            FJ_NOUNROLL
            for (int q = 0; q < KT; ++q)
                if (c.flmask[q] >> m & 1u) acc = fj_add(acc, c.ff[c.slot[q * Mx + m]]);
            c.mF[m] = acc;
        }
    }
    if (lane == 0) {
        if (overflow) c.scal[FJ_S_ERROR] |= FJ_E_OVERFLOW;
        c.scal[FJ_S_NFL] = overflow ? 0 : base;
    }
    fj_sync();
}

// in-line arrival: the env's own warp solves the LP on its global scratch slab
template <int SUM_MODE>
FJ_FN_NOINLINE void fj_order_arrives_inline(int s, int do_begin)
{
    FjCtx &c = FJ_CTX;
    if (do_begin) fj_arrival_begin(c, s);
    FjLp L;
    const FjDims &d = c.P->d;
    unsigned char *binv = c.lp;
    unsigned char *small_ = c.lp + (size_t)d.Rx * d.Rx * 8;
    fj_lp_carve(L, binv, small_, d);
    double *x = (double *)(small_ + (fj_lp_small_bytes(d) + 7) / 8 * 8);
    int iters = 0;
    FjWarpGroup g;
    int rc = fj_lp_solve(g, c, L, x, &iters);
    fj_arrival_finish<SUM_MODE>(c, x, iters, rc);
}

// ---------------------------------------------------------------- observation + rule keys
// state_extract + update_parameter (SO_DFJSP.py:81-169 / MO_DFJSP.py:91-187).  Besides v(t)
// (written to c.obs2) it leaves what the NEXT step's task_select needs: availability /
// delay bit masks and, per available operation type, the two keys that come out of the
// walk over its unprocessed operations (delivery urgency, largest estimated delay).  All
// other rule keys are cheap and recomputed lazily by fj_task_select for the one rule used.
template <int SUM_MODE> FJ_FN void fj_neumaier(double &f, double &cc, double x)
{
    // CPython's Neumaier step adds the exact rounding error of f + x to the compensation,
    // picking the Fast2Sum operand order by magnitude.  Knuth's branch-free TwoSum yields
    // the same (exactly representable) error for either order, so the result is identical.
    if (SUM_MODE == 0) { f = fj_add(f, x); return; }
    const double t2 = fj_add(f, x);
    const double bp = fj_sub(t2, f);
    const double err = fj_add(fj_sub(f, fj_sub(t2, bp)), fj_sub(x, bp));
    cc = fj_add(cc, err);
    f = t2;
}

FJ_OUTLINE double fj_gap_mrj(int q, int m, double gt)
{
    FjCtx &c = FJ_CTX;
    const int sl = c.slot[q * c.Mx + m];
    if (sl == 0xFFFF) return fj_sub(0.0, (double)c.pk[q * c.Mx + m]);
    return fj_sub(c.fu[sl], fj_sub(c.fa[sl], fj_mul(gt, c.ff[sl])));
}

// gap_ave of machine m (class_FJSP.py:156-159).  EXACT: CPython's sum order and compensation
// (it is a machine-rule key); otherwise a plain running sum (observation feature only).
template <int SUM_MODE, int EXACT>
FJ_OUTLINE double fj_machine_gap_ave(int m, double gt)
{
    FjCtx &c = FJ_CTX;
    const int KT = c.KT;
    const FjEligRO elig = fj_elig(c);
    double f = 0.0, cc = 0.0;
    int n = 0;
    FJ_NOUNROLL
    for (int q = 0; q < KT; ++q) {
        if (!((unsigned)elig[q] >> m & 1u)) continue;
        const double term = fj_gap_mrj(q, m, gt);
        if (EXACT) fj_neumaier<SUM_MODE>(f, cc, term); else f = fj_add(f, term);
        ++n;
    }
    if (EXACT && SUM_MODE != 0 && cc != 0.0 && isfinite(cc)) f = fj_add(f, cc);
    return fj_div(f, (double)n);
}

template <int VARIANT, int SUM_MODE>
FJ_FN_NOINLINE void fj_observe(int rates_zero)
{
    FjCtx &c = FJ_CTX;
    const int lane = fj_lane();
    const bool MO = (VARIANT == FJSP_MO_DFJSP || VARIANT == FJSP_MO_BREAKDOWN);
    const int M = c.M, KT = c.KT, S = c.S, Sx = c.Sx;
    const FjEligRO elig = fj_elig(c);
    const FjDueRO due = fj_due(c);
    const FjLastRO rjlast = fj_rjlast(c);
    const int t = c.scal[FJ_S_TIME];
    const double td = (double)t;
    const unsigned idle = ~(unsigned)c.scal[FJ_S_BUSY] & c.mmask;
    const double gt = fj_get_d(c.scal, FJ_S_GAPTIME);
    // packed integer partials: (tn, da) (de, jn) (ja, je) and the 64-bit tardiness of waiting jobs
    long long p0 = 0, p1 = 0, p2 = 0, dunp = 0;
    int nav = 0, nfav = 0;
    double s_fr = 0.0, s_gr = 0.0;
    double fr_keep = 0.0, gr_keep = 0.0;   // this lane's two rates when every operation type has its own lane
    const int rounds = (KT + FJ_NL - 1) / FJ_NL;
    FJ_NOUNROLL
    for (int rd = 0; rd < rounds; ++rd) {
        const int q = rd * FJ_NL + lane;
        int av = 0, fav = 0, dle = 0, dla = 0;
        if (q < KT) {
            const int qn = c.qlen[q];
            av = qn > 0 && ((unsigned)elig[q] & idle) != 0;
            fav = qn > 0 && (c.flmask[q] & idle) != 0;
            const double f = c.tsum[q];
            // walk the unprocessed operations in list order: positions are consecutive and the
            // due date is constant inside an order, so per-order counts are all that is needed;
            // inside an order the estimate grows with the position, so its maximum is the last
            int residue = 0, a_cnt = 0, e_cnt = 0;
            double kd = 0.0, max_e = 0.0, sf = 0.0, sc = 0.0;
            bool first = true;
            long long late_sum = 0;
            if (VARIANT == FJSP_SO_FJSSP) {
                // per-job due dates: walk the set of unprocessed jobs in job-number order
                const int NWx = c.P->d.NWx;
                const int r = fj_rjkind(c)[q], jb = fj_jobbase(c)[r];
                const int arrived = fj_cum(c)[c.scal[FJ_S_NEXTORDER] * c.Kx + r];
                int mind = 0x7fffffff;
                FJ_NOUNROLL
                for (int w = 0; w * 32 < arrived; ++w) {
                    unsigned bits = c.unpmask[q * NWx + w];
                    while (bits) {
                        const int n = w * 32 + fj_ffs0(bits); bits &= bits - 1;
                        const int d = c.duejob[jb + n];
                        const double dd = (double)d;
                        ++residue; kd += 1.0;
                        const double est = fj_add(td, fj_mul(f, kd));
                        const double ve = fj_sub(est, dd);
                        if (t > d) { ++a_cnt; late_sum += t - d; }
                        e_cnt += est > dd;
                        if (first || ve > max_e) max_e = ve;
                        first = false;
                        mind = d < mind ? d : mind;
                        fj_neumaier<SUM_MODE>(sf, sc, ve);
                    }
                }
                c.mindue[q] = mind;
            } else {
            // No loop over the operations here.  Inside an order the estimate
            //   g(k) = fl(t + fl(f * k))  is monotone in the position k, so the count of
            // estimated-late operations is a suffix found by bisection on the exactly rounded
            // g, and the order's largest estimated delay is its last term.  The one quantity
            // that needs every term, the compensated SUM behind the delivery urgency, is a key
            // of task rules 1, 2 and 4 only and is computed by fj_task_select when such a
            // rule is actually played (the state is the same then).
            int kpos = 0;
            FJ_NOUNROLL
            for (int s = 0; s < S; ++s) {
                const int cnt = c.cntunp[q * Sx + s];
                if (cnt == 0) continue;
                residue += cnt;
                const int d = due[s];
                const double dd = (double)d;
                if (t > d) { a_cnt += cnt; late_sum += (long long)cnt * (t - d); }
                int lo = kpos + 1, hi = kpos + cnt;
                const double ghi = fj_add(td, fj_mul(f, (double)hi));
                if (ghi > dd) {
                    if (fj_add(td, fj_mul(f, (double)lo)) > dd) e_cnt += cnt;
                    else {   // g(lo) <= dd < g(hi): smallest late position by bisection
                        FJ_NOUNROLL
                        while (hi - lo > 1) {
                            const int mid = (lo + hi) >> 1;
                            if (fj_add(td, fj_mul(f, (double)mid)) > dd) hi = mid; else lo = mid;
                        }
                        e_cnt += kpos + cnt - hi + 1;
                    }
                }
                const double ve = fj_sub(ghi, dd);
                if (first || ve > max_e) max_e = ve;
                first = false;
                kpos += cnt;
            }
            }
            p0 += ((long long)residue << 32) + a_cnt;
            p1 += ((long long)e_cnt << 32);
            if (rjlast[q]) { p1 += residue; p2 += ((long long)a_cnt << 32) + e_cnt; dunp += late_sum; }
            const double fluid_unp = fj_sub((double)c.fstart[q], fj_mul(c.rsum[q], gt));
            const double gap = fj_sub((double)residue, fluid_unp);
            const int pr = c.proc[q];
            fr_keep = fj_div((double)pr, (double)(residue + pr));
            gr_keep = fj_div(gap, (double)c.fstart[q]);
            s_fr = fj_add(s_fr, fr_keep);
            s_gr = fj_add(s_gr, gr_keep);
            // more operation types than lanes: park the two rates for the variance pass (the D-FJSP
            // classes do not use the urg / maxe arrays of the record otherwise)
            if (VARIANT != FJSP_SO_FJSSP && rounds > 1) { c.urg[q] = fr_keep; c.maxe[q] = gr_keep; }
            if (av) {
                if (VARIANT == FJSP_SO_FJSSP) {
                    if (SUM_MODE != 0 && sc != 0.0 && isfinite(sc)) sf = fj_add(sf, sc);
                    c.urg[q] = fj_div(sf, (double)residue);
                }
                if (VARIANT == FJSP_SO_FJSSP) c.maxe[q] = max_e;   // the D-FJSP classes recompute it on demand
                dle = e_cnt > 0; dla = a_cnt > 0;
            }
        }
        // bit masks, one 32-bit word per round
#ifdef FJ_DEVICE_CODE
        const unsigned wa = __ballot_sync(0xffffffffu, av), wf = __ballot_sync(0xffffffffu, fav);
        const unsigned we = __ballot_sync(0xffffffffu, dle), wd = __ballot_sync(0xffffffffu, dla);
        if (lane == 0) { c.avmask[rd] = wa; c.favmask[rd] = wf; c.demask[rd] = we; c.damask[rd] = wd; }
        nav += fj_popc(wa); nfav += fj_popc(wf);
#else
        if ((q & 31) == 0) { c.avmask[q >> 5] = 0; c.favmask[q >> 5] = 0; c.demask[q >> 5] = 0; c.damask[q >> 5] = 0; }
        if (av) c.avmask[q >> 5] |= 1u << (q & 31);
        if (fav) c.favmask[q >> 5] |= 1u << (q & 31);
        if (dle) c.demask[q >> 5] |= 1u << (q & 31);
        if (dla) c.damask[q >> 5] |= 1u << (q & 31);
        nav += av; nfav += fav;
#endif
    }
    p0 = fj_sum_pair(p0); p1 = fj_sum_pair(p1); p2 = fj_sum_pair(p2); dunp = fj_sum_ll(dunp);
    const long long tn = p0 >> 32, da = p0 & 0xffffffffll, de = p1 >> 32, jn = p1 & 0xffffffffll;
    const long long ja = p2 >> 32, je = p2 & 0xffffffffll;
    // observation-only moments: warp-tree sums; spreads and the machines' gap_ave use the stored reciprocals
    const double inv_kt = fj_get_d(c.scal, FJ_S_INV_KT), inv_m = fj_get_d(c.scal, FJ_S_INV_M);
    fj_sum_d2(s_fr, s_gr);
    // (the means feed differences, so they stay true divisions; the spreads use the reciprocals)
    // the three means that need no further reduction: one division on lanes 0..2, then broadcast
    const long long tsum_m = fj_get_ll(c.scal, FJ_S_MENDSUM);
    double cro_ave, gap_ave, ct_ave;
    if (FJ_NL == 1) {
        cro_ave = fj_div(s_fr, (double)KT); gap_ave = fj_div(s_gr, (double)KT); ct_ave = fj_div((double)tsum_m, (double)M);
    } else {
        const double mean = fj_div(lane == 0 ? s_fr : lane == 1 ? s_gr : (double)tsum_m, lane < 2 ? (double)KT : (double)M);
        cro_ave = fj_bcast_d(mean, 0); gap_ave = fj_bcast_d(mean, 1 % FJ_NL); ct_ave = fj_bcast_d(mean, 2 % FJ_NL);
    }
    // second pass: variances
    double v_fr = 0.0, v_gr = 0.0;
    if (rounds == 1) {
        if (lane < KT) { v_fr = fj_mul(fr_keep - cro_ave, fr_keep - cro_ave); v_gr = fj_mul(gr_keep - gap_ave, gr_keep - gap_ave); }
    } else
    FJ_NOUNROLL
    for (int q = lane; q < KT; q += FJ_NL) {
        double fr, gr;
        if (VARIANT != FJSP_SO_FJSSP) { fr = c.urg[q]; gr = c.maxe[q]; }   // written by this lane above
        else {
            int residue = 0;
            FJ_NOUNROLL
            for (int s = 0; s < S; ++s) residue += c.cntunp[q * Sx + s];
            const int pr = c.proc[q];
            fr = fj_div((double)pr, (double)(residue + pr));
            const double fluid_unp = fj_sub((double)c.fstart[q], fj_mul(c.rsum[q], gt));
            gr = fj_div(fj_sub((double)residue, fluid_unp), (double)c.fstart[q]);
        }
        v_fr = fj_add(v_fr, fj_mul(fr - cro_ave, fr - cro_ave));
        v_gr = fj_add(v_gr, fj_mul(gr - gap_ave, gr - gap_ave));
    }
    // machines: completion-time spread; MO also the mean / spread of the machines' gap_ave.  Summed
    // over a machine's operation types, unprocessed - fluid_unprocessed telescopes to
    // (gap_time * sum of its fluid rates) - (dispatches on it since the last arrival); the exact
    // CPython-ordered value is only needed as the key of machine rule 4 and is computed there.
    double v_ct = 0.0, s_gm = 0.0, ga_l = 0.0;
    FJ_NOUNROLL
    for (int m = lane; m < M; m += FJ_NL) {
        const double dv = fj_sub((double)c.mend[m], ct_ave);
        v_ct = fj_add(v_ct, fj_mul(dv, dv));
        if (MO) {
            ga_l = fj_mul(fj_sub(fj_mul(gt, c.mF[m]), (double)c.mD[m]), c.invnkt[m]);
            c.gapave[m] = ga_l;      // same lane reads it back below
            s_gm = fj_add(s_gm, ga_l);
        }
    }
    fj_sum_d4(v_fr, v_gr, v_ct, s_gm);
    double gm_ave = 0.0, v_gm = 0.0;
    if (MO) {
        gm_ave = fj_div(s_gm, (double)M);
        FJ_NOUNROLL
        for (int m = lane; m < M; m += FJ_NL) { const double dv = fj_sub(c.gapave[m], gm_ave); v_gm = fj_add(v_gm, fj_mul(dv, dv)); }
        v_gm = fj_sum_d(v_gm);
    }
    // the four spreads: one square root per lane (lanes 0..3), written straight to the observation
    double cro_std = 0.0, gap_std = 0.0, ct_std = 0.0, gm_std = 0.0;
    if (FJ_NL == 1) {
        cro_std = sqrt(fj_mul(v_fr, inv_kt)); gap_std = sqrt(fj_mul(v_gr, inv_kt)); ct_std = sqrt(fj_mul(v_ct, inv_m));
        if (MO) gm_std = sqrt(fj_mul(v_gm, inv_m));
    } else if (lane < 4) {
        const double sd = sqrt(fj_mul(lane == 0 ? v_fr : lane == 1 ? v_gr : lane == 2 ? v_ct : v_gm, lane < 2 ? inv_kt : inv_m));
        if (MO) c.obs2[lane == 0 ? 6 : lane == 1 ? 8 : lane == 2 ? 3 : 10] = sd;
        else if (lane < 3) c.obs2[lane == 0 ? 3 : lane == 1 ? 5 : 1] = sd;
    }
    if (lane < 5 || FJ_NL == 1) {   // the four delay rates and MO's fluid-available ratio: one division per lane
        FJ_NOUNROLL
        for (int k = (FJ_NL == 1 ? 0 : lane); k < (MO ? 5 : 4); k += (FJ_NL == 1 ? 1 : 5)) {
            const long long num = k == 0 ? da : k == 1 ? de : k == 2 ? ja : k == 3 ? je : (long long)nfav;
            const double den = k < 2 ? (double)tn : k < 4 ? (double)jn : fj_add((double)nav, 1e-08);
            const double r = fj_div((double)num, den);
            if (k < 4) c.obs2[(MO ? 11 : 6) + k] = rates_zero ? 0.0 : r; else c.obs2[4] = r;
        }
    }
    if (lane == 0) {
        fj_set_ll(c.scal, FJ_S_DELAY_UNPROC, dunp);
        c.scal[FJ_S_NAV] = nav; c.scal[FJ_S_NFAV] = nfav;
        double *o = c.obs2;
        if (MO) {
            uint64_t bits = (uint32_t)FJ_I(c, hdr)[5] | ((uint64_t)(uint32_t)FJ_I(c, hdr)[6] << 32);
            double ddt; memcpy(&ddt, &bits, 8);
            o[0] = ddt; o[1] = (double)M; o[2] = (double)S;
            o[5] = cro_ave; o[7] = gap_ave; o[9] = gm_ave;
            if (FJ_NL == 1) { o[3] = ct_std; o[6] = cro_std; o[8] = gap_std; o[10] = gm_std; }
        } else {
            o[0] = (double)M; o[2] = cro_ave; o[4] = gap_ave;
            if (FJ_NL == 1) { o[1] = ct_std; o[3] = cro_std; o[5] = gap_std; }
        }
    }
    fj_sync();
}

// ---------------------------------------------------------------- task_select
FJ_OUTLINE int fj_nth_set(const uint32_t *mask, int words, int nth)
{
    FJ_NOUNROLL
    for (int w = 0; w < words; ++w) {
        unsigned v = mask[w];
        int pc = fj_popc(v);
        if (nth < pc) { while (nth--) v &= v - 1; return w * 32 + fj_ffs0(v); }
        nth -= pc;
    }
    return -1;
}
FJ_FN int fj_mask_any(const uint32_t *mask, int words)
{
    unsigned v = 0;
    FJ_NOUNROLL
    for (int w = 0; w < words; ++w) v |= mask[w];
    return v != 0;
}

// kind_task_delivery_urgency of one operation type (SO_DFJSP.py:139-156): CPython's compensated
// sum over every unprocessed operation's estimated delay, divided by their number.
template <int SUM_MODE>
FJ_FN double fj_urgency(const FjCtx &c, int q, double td)
{
    const int S = c.S, Sx = c.Sx;
    const FjDueRO due = fj_due(c);
    const double f = c.tsum[q];
    double kd = 0.0, sf = 0.0, sc = 0.0;
    int residue = 0;
    FJ_NOUNROLL
    for (int s = 0; s < S; ++s) {
        const int cnt = c.cntunp[q * Sx + s];
        if (cnt == 0) continue;
        residue += cnt;
        const double dd = (double)due[s];
        FJ_UNROLL4
        for (int i = 0; i < cnt; ++i) {
            kd += 1.0;
            fj_neumaier<SUM_MODE>(sf, sc, fj_sub(fj_add(td, fj_mul(f, kd)), dd));
        }
    }
    if (SUM_MODE != 0 && sc != 0.0 && isfinite(sc)) sf = fj_add(sf, sc);
    return fj_div(sf, (double)residue);
}

// kind_task_delay_time_e of one operation type (SO_DFJSP.py:152-154): the largest estimated
// delay over its unprocessed operations = the largest last-of-order term (monotone inside an order)
FJ_FN double fj_max_est_delay(const FjCtx &c, int q, double td)
{
    const int S = c.S, Sx = c.Sx;
    const FjDueRO due = fj_due(c);
    const double f = c.tsum[q];
    int kpos = 0; bool first = true; double mx = 0.0;
    FJ_NOUNROLL
    for (int s = 0; s < S; ++s) {
        const int cnt = c.cntunp[q * Sx + s];
        if (cnt == 0) continue;
        kpos += cnt;
        const double ve = fj_sub(fj_add(td, fj_mul(f, (double)kpos)), (double)due[s]);
        if (first || ve > mx) mx = ve;
        first = false;
    }
    return mx;
}

// SO_DFJSP.py:270-301 / MO_DFJSP.py:300-352.  Warp-cooperative: lanes own operation types,
// one argmax/argmin reduction (lowest index on ties = Python's first extremal element).
template <int VARIANT, int SUM_MODE>
FJ_FN int fj_task_select(FjCtx &c, int rule, uint32_t rnd)
{
    const int lane = fj_lane();
    const bool MO = (VARIANT == FJSP_MO_DFJSP || VARIANT == FJSP_MO_BREAKDOWN);
    const int KT = c.KT, S = c.S, Sx = c.Sx, Mx = c.Mx;
    const int KTW = (KT + 31) / 32;
    const int nav = c.scal[FJ_S_NAV], nfav = c.scal[FJ_S_NFAV];
    if (nav <= 0) return -1;
    // key: 0 urgency 1 est. delay 2 actual delay 3 gap (all max) 4 min due 5 min energy 6 min time (all min)
    const uint32_t *set = c.avmask;
    int key = -1;
    bool fluid_m = false;   // energy / time over the fluid machines
    const uint32_t *flav = nfav > 0 ? c.favmask : c.avmask;
    switch (rule) {
    case 1: if (fj_mask_any(c.demask, KTW)) { set = c.demask; key = 1; } else key = 0; break;
    case 2: if (fj_mask_any(c.damask, KTW)) { set = c.damask; key = 2; } else key = 0; break;
    case 3: set = flav; key = 3; break;
    case 4: set = flav; key = 0; break;
    case 5: set = flav; key = 4; break;
    default:
        if (!MO) { if (rule == 6) return fj_nth_set(c.avmask, KTW, (int)(rnd % (uint32_t)nav)); return -1; }
        switch (rule) {
        case 6: key = 4; break;
        case 7: set = flav; key = 5; fluid_m = nfav > 0; break;
        case 8: key = 5; break;
        case 9: set = flav; key = 6; fluid_m = nfav > 0; break;
        case 10: key = 6; break;
        case 11: return nfav > 0 ? fj_nth_set(c.favmask, KTW, (int)(rnd % (uint32_t)nfav))
                                 : fj_nth_set(c.avmask, KTW, (int)(rnd % (uint32_t)nav));
        case 12: return fj_nth_set(c.avmask, KTW, (int)(rnd % (uint32_t)nav));
        default: return -1;
        }
    }
    const FjDueRO due = fj_due(c);
    const FjEligRO elig = fj_elig(c);
    const int t = c.scal[FJ_S_TIME];
    const unsigned idle = ~(unsigned)c.scal[FJ_S_BUSY] & c.mmask;
    const double gt = fj_get_d(c.scal, FJ_S_GAPTIME);
    const bool want_max = key <= 3;
    FjBest b; fj_best_init(b);
    FJ_NOUNROLL
    for (int q = lane; q < KT; q += FJ_NL) {
        if (!(set[q >> 5] >> (q & 31) & 1u)) continue;
        double k;
        if (key == 0) k = (VARIANT == FJSP_SO_FJSSP) ? c.urg[q] : fj_urgency<SUM_MODE>(c, q, (double)t);
        else if (key == 1) k = (VARIANT == FJSP_SO_FJSSP) ? c.maxe[q] : fj_max_est_delay(c, q, (double)t);
        else if (key == 2 && VARIANT == FJSP_SO_FJSSP) k = (double)((long long)t - c.mindue[q]);
        else if (key == 4 && VARIANT == FJSP_SO_FJSSP) {   // min due date over the waiting jobs
            const int r = fj_rjkind(c)[q], jb = fj_jobbase(c)[r];
            int mind = 0x7fffffff, n;
            if (fj_rjstage(c)[q] == 0) {
                const int arrived = fj_cum(c)[c.scal[FJ_S_NEXTORDER] * c.Kx + r];
                FJ_NOUNROLL
                for (n = arrived - c.qlen[q]; n < arrived; ++n) { const int d = c.duejob[jb + n]; mind = d < mind ? d : mind; }
            } else {
                n = c.qhead[q];
                FJ_NOUNROLL
                for (int i = 0; i < c.qlen[q]; ++i) { const int d = c.duejob[jb + n]; mind = d < mind ? d : mind; n = c.next[jb + n]; }
            }
            k = (double)mind;
        }
        else if (key == 2) {   // max over unprocessed operations of (t - due)
            int mind = 0x7fffffff;
            FJ_NOUNROLL
            for (int s = 0; s < S; ++s) if (c.cntunp[q * Sx + s] > 0 && due[s] < mind) mind = due[s];
            k = (double)((long long)t - mind);
        } else if (key == 3) {
            int residue = 0;
            FJ_NOUNROLL
            for (int s = 0; s < S; ++s) residue += c.cntunp[q * Sx + s];
            k = fj_sub((double)residue, fj_sub((double)c.fstart[q], fj_mul(c.rsum[q], gt)));
        } else if (key == 4) {
            int mind = 0x7fffffff;
            FJ_NOUNROLL
            for (int s = 0; s < S; ++s) if (c.cntnow[q * Sx + s] > 0 && due[s] < mind) mind = due[s];
            k = (double)mind;
        } else {   // MO_DFJSP.py:429-451: min energy / time over the idle (fluid) machines
            unsigned sm = (fluid_m ? c.flmask[q] : (unsigned)elig[q]) & idle;
            const FjRO tab = (key == 5 ? FJ_I(c, energy) : FJ_I(c, ptime)) + q * Mx;
            int mn = 0x7fffffff;
            FJ_NOUNROLL
            while (sm) { const int m = fj_ffs0(sm); sm &= sm - 1; const int v = tab[m]; mn = v < mn ? v : mn; }
            k = (double)mn;
        }
        if (want_max) fj_best_max(b, k, q); else fj_best_min(b, k, q);
    }
    fj_best_reduce(b, want_max ? 1 : 0);
    return b.idx == 0x7fffffff ? -1 : b.idx;
}

// ---------------------------------------------------------------- machine_select
// Candidate lists are list(set(idle) & set(machine_tuple)) in CPython's set iteration
// order (see oracle/pyemu.h).  Five or more members iterate in ascending order; up to four
// live in an 8-slot table whose layout depends on insertion order, emulated here on a
// 64-bit register (one byte per slot).
FJ_OUTLINE unsigned fj_small_set_order(unsigned seq, int n)   // seq: members, one byte each, insertion order
{
    if (n <= 1) return seq;
    // slot of every member in the 8-slot table (probe i -> 5i+1; the perturbation is 0 for v < 32)
    unsigned used = 0, slots = 0;   // slots: 4 bits per member
    FJ_NOUNROLL
    for (int e = 0; e < n; ++e) {
        unsigned i = (seq >> (8 * e)) & 7u;
        while (used >> i & 1u) i = (i * 5u + 1u) & 7u;
        used |= 1u << i;
        slots |= i << (4 * e);
    }
    // iteration order = ascending slot: a member's rank is the number of members in lower slots
    unsigned out = 0;
    FJ_NOUNROLL
    for (int e = 0; e < n; ++e) {
        const unsigned se = (slots >> (4 * e)) & 7u;
        const int rank = fj_popc(used & ((1u << se) - 1u));
        out |= ((seq >> (8 * e)) & 0xffu) << (8 * rank);
    }
    return out;
}

struct FjCand { unsigned mask; unsigned packed; int n; };   // n >= 5: ascending over mask; else packed order

// list(set(idle) & set(other)); other given by its mask, its size and, when it has at most four
// members, its iteration order packed one byte per member (five or more iterate ascending)
FJ_OUTLINE FjCand fj_selectable(unsigned idle, unsigned omask, int nother, unsigned ord_packed)
{
    FjCand r;
    r.mask = idle & omask; r.n = fj_popc(r.mask); r.packed = 0;
    if (r.n >= 5 || r.n == 0) return r;
    const int ni = fj_popc(idle);
    unsigned seq = 0; int k = 0;
    if (nother > ni) {   // iterate set(idle)
        if (ni >= 5) { unsigned mk = r.mask; while (mk) { seq |= (unsigned)fj_ffs0(mk) << (8 * k++); mk &= mk - 1; } }
        else {
            unsigned a = 0; int na = 0; unsigned mk = idle;
            FJ_NOUNROLL
            while (mk) { a |= (unsigned)fj_ffs0(mk) << (8 * na++); mk &= mk - 1; }
            const unsigned ao = fj_small_set_order(a, na);
            FJ_NOUNROLL
            for (int i = 0; i < na; ++i) { const unsigned v = (ao >> (8 * i)) & 0xffu; if (omask >> v & 1u) seq |= v << (8 * k++); }
        }
    } else {             // iterate set(other)
        if (nother >= 5) { unsigned mk = r.mask; while (mk) { seq |= (unsigned)fj_ffs0(mk) << (8 * k++); mk &= mk - 1; } }
        else for (int i = 0; i < nother; ++i) { const unsigned v = (ord_packed >> (8 * i)) & 0xffu; if (idle >> v & 1u) seq |= v << (8 * k++); }
    }
    r.packed = fj_small_set_order(seq, r.n);
    return r;
}

// SO_DFJSP.py:303-325 / MO_DFJSP.py:354-398.  Warp-cooperative only for the exact gap_ave
// keys of rule 4; the pick itself is a short scalar loop every lane runs redundantly.
template <int VARIANT, int SUM_MODE>
FJ_FN int fj_machine_select(FjCtx &c, int rule, int q, uint32_t rnd)
{
    const bool MO = (VARIANT == FJSP_MO_DFJSP || VARIANT == FJSP_MO_BREAKDOWN);
    const int Mx = c.Mx, M = c.M;
    const unsigned idle = ~(unsigned)c.scal[FJ_S_BUSY] & c.mmask;
    const double gt = fj_get_d(c.scal, FJ_S_GAPTIME);
    int key_kind;       // 0 gap_rj(max) 1 time(min) 2 gap_ave(max) 3 energy(min) 4 idle power(min) 5 random
    bool use_f;
    if (!MO) {
        switch (rule) {
        case 1: key_kind = 0; use_f = true; break;
        case 2: key_kind = 0; use_f = false; break;
        case 3: key_kind = 1; use_f = false; break;
        case 4: key_kind = 2; use_f = true; break;
        case 5: key_kind = 5; use_f = false; break;
        default: return -1;
        }
    } else {
        switch (rule) {
        case 1: key_kind = 0; use_f = true; break;
        case 2: key_kind = 1; use_f = true; break;
        case 3: key_kind = 1; use_f = false; break;
        case 4: key_kind = 2; use_f = true; break;
        case 5: key_kind = 3; use_f = true; break;
        case 6: key_kind = 3; use_f = false; break;
        case 7: key_kind = 4; use_f = true; break;
        case 8: key_kind = 4; use_f = false; break;
        case 9: key_kind = 5; use_f = true; break;
        case 10: key_kind = 5; use_f = false; break;
        default: return -1;
        }
    }
    const unsigned em = (unsigned)fj_elig(c)[q], fm = c.flmask[q];
    if ((idle & em) == 0) return -1;
    FjCand cand;
    // set(fluid_machine_list) (members in x.items() order) or set(machine_tuple): both iteration
    // orders are kept packed in the record's hot prefix
    if (use_f && (idle & fm) != 0) cand = fj_selectable(idle, fm, fj_popc(fm), c.h_fo[q]);
    else cand = fj_selectable(idle, em, fj_popc(em), c.h_mtpack[q]);
    if (key_kind == 2) {   // exact gap_ave of the candidates, one lane per machine
        FJ_NOUNROLL
        for (int m = fj_lane(); m < M; m += FJ_NL)
            if (cand.mask >> m & 1u) c.gapave[m] = fj_machine_gap_ave<SUM_MODE, 1>(m, gt);
        fj_sync();
    }
    const int n = cand.n;
    if (key_kind == 5) {
        const int pick = (int)(rnd % (uint32_t)n);
        if (n >= 5) { unsigned mk = cand.mask; for (int i = 0; i < pick; ++i) mk &= mk - 1; return fj_ffs0(mk); }
        return (int)((cand.packed >> (8 * pick)) & 0xffu);
    }
    int best = -1; double bk = 0.0;
    unsigned mk = cand.mask;
    FJ_NOUNROLL
    for (int i = 0; i < n; ++i) {
        int m;
        if (n >= 5) { m = fj_ffs0(mk); mk &= mk - 1; } else m = (int)((cand.packed >> (8 * i)) & 0xffu);
        double k;
        switch (key_kind) {
        case 0: k = fj_gap_mrj(q, m, gt); break;
        case 1: k = (double)FJ_I(c, ptime)[q * Mx + m]; break;
        case 2: k = c.gapave[m]; break;
        case 3: k = (double)FJ_I(c, energy)[q * Mx + m]; break;
        default: k = (double)FJ_I(c, idlep)[m]; break;
        }
        const bool want_max = (key_kind == 0 || key_kind == 2);
        if (best < 0 || (want_max ? k > bk : k < bk)) { best = m; bk = k; }
    }
    return best;
}

// ---------------------------------------------------------------- suspension
// An environment whose clock loop (or auto-reset) reaches an order arrival needs the fluid
// LP.  In the main step kernel (SUSPEND = 1) it parks itself on the pending list; the LP
// kernel solves all parked LPs one CTA each; the resume kernel (SUSPEND = 0) picks the
// solution up and finishes the step and the rest of the launch, solving any FURTHER
// arrival of the same launch in line.
enum { FJ_PH_RUN = 0, FJ_PH_LP_STEP = 1, FJ_PH_LP_RESET = 2 };

FJ_FN void fj_suspend(FjCtx &c, const FjParams &P, const FjStepArgs &A, int env, int phase, int tt)
{
    if (fj_lane() == 0) {
        c.scal[FJ_S_PHASE] = phase; c.scal[FJ_S_TT] = tt;
#ifdef FJ_DEVICE_CODE
        const int idx = atomicAdd(A.park_count, 1);
#else
        const int idx = (*A.park_count)++;
#endif
        A.park_env[idx] = env;
        c.scal[FJ_S_LPSLOT] = idx < P.lp_slots ? idx : -1;
    }
    fj_sync();
}

// picks up the parked arrival's LP solution (or solves in line when no slot was free)
template <int SUM_MODE>
FJ_FN void fj_arrival_resume(FjCtx &c, const FjParams &P)
{
    const int slot = c.scal[FJ_S_LPSLOT];
    if (slot >= 0) fj_arrival_finish<SUM_MODE>(c, P.lp_x + (size_t)slot * P.d.NPx, P.lp_meta[2 * slot], P.lp_meta[2 * slot + 1]);
    else fj_order_arrives_inline<SUM_MODE>(0, 0);
    if (fj_lane() == 0) c.scal[FJ_S_PHASE] = FJ_PH_RUN;
    fj_sync();
}

// ---------------------------------------------------------------- reset
// reset() of the reference incl. its re-reset quirks (oracle/fjsp_oracle.c explains them):
// busy flags, order_arrive_time and the `done` seen by the first observation survive.
FJ_FN void fj_reset_begin(FjCtx &c, int fresh)
{
    const int lane = fj_lane();
    const int KT = c.KT, Sx = c.Sx, M = c.M;
    int was_done = 0;
    if (fresh) {
        const FjRO ielig = FJ_I(c, elig), ikind = FJ_I(c, rjkind), istage = FJ_I(c, rjstage), ilast = FJ_I(c, rjlast);
        const FjRO idue = FJ_I(c, due);
        FJ_NOUNROLL
        for (int q = lane; q < KT; q += FJ_NL) {
            c.h_elig[q] = (uint32_t)ielig[q];
            c.h_mtpack[q] = (uint32_t)FJ_I(c, mtpack)[q];
            c.h_fo[q] = 0;
            c.h_rjinfo[q] = (uint16_t)((ikind[q] << 8) | (istage[q] << 1) | (ilast[q] & 1));
        }
        FJ_NOUNROLL
        for (int k = lane; k < c.S; k += FJ_NL) c.h_due[k] = idue[k];
        {
            const FjRO icum = FJ_I(c, cum), ijb = FJ_I(c, jobbase);
            FJ_NOUNROLL
            for (int k = lane; k < (c.Sx + 1) * c.Kx; k += FJ_NL) c.h_cum[k] = icum[k];
            FJ_NOUNROLL
            for (int k = lane; k < c.K; k += FJ_NL) c.h_jobbase[k] = ijb[k];
        }
        FJ_NOUNROLL
        for (int i = lane; i < FJ_S_COUNT; i += FJ_NL) c.scal[i] = 0;
        FJ_NOUNROLL
        for (int i = lane; i < 16; i += FJ_NL) c.obs[i] = 0.0;
        fj_sync();
        {   // reciprocals of the static divisors of the observation-only means (x * (1/n) instead of x / n)
            const FjRO inkt = FJ_I(c, mnkt);
            FJ_NOUNROLL
            for (int m = lane; m < M; m += FJ_NL) c.invnkt[m] = fj_div(1.0, (double)inkt[m]);
            if (lane == 0) { fj_set_d(c.scal, FJ_S_INV_KT, fj_div(1.0, (double)KT)); fj_set_d(c.scal, FJ_S_INV_M, fj_div(1.0, (double)M)); }
        }
        fj_sync();
    } else {
        was_done = c.scal[FJ_S_DONE];
    }
    FJ_NOUNROLL
    for (int q = lane; q < KT; q += FJ_NL) {
        c.qhead[q] = 0xFFFF; c.qtail[q] = 0xFFFF; c.qlen[q] = 0; c.proc[q] = 0;
        FJ_NOUNROLL
        for (int s = 0; s < Sx; ++s) { c.cntunp[q * Sx + s] = 0; c.cntnow[q * Sx + s] = 0; }
    }
    FJ_NOUNROLL
    for (int m = lane; m < M; m += FJ_NL) { c.mend[m] = 0; c.mlast[m] = 0; c.mjob[m] = -1; }
    if (c.P->variant == FJSP_SO_FJSSP) {
        FJ_NOUNROLL
        for (int i = lane; i < KT * c.P->d.NWx; i += FJ_NL) c.unpmask[i] = 0;
    }
    fj_sync();
    if (lane == 0) {
        c.scal[FJ_S_NEXTORDER] = 1; c.scal[FJ_S_TIME] = 0; c.scal[FJ_S_STEPS] = 0; c.scal[FJ_S_HASTASK] = 0;
        c.scal[FJ_S_LEFT] = 0; fj_set_ll(c.scal, FJ_S_MENDSUM, 0);
        c.scal[FJ_S_COMPLETION] = 0; c.scal[FJ_S_COMPLETION_LAST] = 0; c.scal[FJ_S_WASDONE] = was_done;
        fj_set_ll(c.scal, FJ_S_ENERGY, 0); fj_set_ll(c.scal, FJ_S_ENERGY_LAST, 0);
        fj_set_ll(c.scal, FJ_S_DELAY_PROC, 0); fj_set_ll(c.scal, FJ_S_DELAY_LAST, 0);
        fj_set_ll(c.scal, FJ_S_DELAY_UNPROC, 0);
        fj_set_d(c.scal, FJ_S_GAPTIME, 0.0);
    }
    fj_sync();
    fj_arrival_begin(c, 0);
}

template <int VARIANT, int SUM_MODE>
FJ_FN void fj_reset_finish(FjCtx &c)
{
    const int lane = fj_lane();
    fj_observe<VARIANT, SUM_MODE>(c.scal[FJ_S_WASDONE]);
    FJ_NOUNROLL
    for (int i = lane; i < 16; i += FJ_NL) c.obs[i] = c.obs2[i];
    if (lane == 0) c.scal[FJ_S_DONE] = 0;
    fj_sync();
}

// The order-0 fluid LP of an instance is the same at every reset (same counts, all later
// queues empty), so after the first reset() its solution is cached per instance and an
// auto-reset needs no LP at all.
template <int VARIANT, int SUM_MODE>
FJ_FN int fj_reset_from_plan(FjCtx &c, const FjParams &P)
{
    if (!P.plan_ok || !P.plan_ok[c.inst]) return 0;
    fj_reset_begin(c, 0);
    fj_arrival_finish<SUM_MODE>(c, P.plan_x + (size_t)c.inst * P.d.NPx, P.plan_meta[2 * c.inst], P.plan_meta[2 * c.inst + 1]);
    fj_reset_finish<VARIANT, SUM_MODE>(c);
    if (fj_lane() == 0) c.scal[FJ_S_EPISODES] += 1;
    fj_sync();
    return 1;
}

// ---------------------------------------------------------------- step
// task_select + machine_select + dispatch; returns 0 when nothing could be dispatched.
// `rec`: this step's row of the dispatch record output (8 int32), or null
template <int VARIANT, int SUM_MODE>
FJ_FN_NOINLINE int fj_step_front(int task_rule0, int mach_rule0, uint32_t rnd_task, uint32_t rnd_mach, int32_t *rec)
{
    FjCtx &c = FJ_CTX;
    const int lane = fj_lane();
    const bool MO = (VARIANT == FJSP_MO_DFJSP || VARIANT == FJSP_MO_BREAKDOWN);
    const int Mx = c.Mx, Sx = c.Sx, Kx = c.Kx;
    const FjKindRO rjkind = fj_rjkind(c);
    const FjStageRO rjstage = fj_rjstage(c);
    const FjLastRO rjlast = fj_rjlast(c);
    const FjDueRO cum = fj_cum(c), jobbase = fj_jobbase(c);
    const FjDueRO due = fj_due(c);
    const int t = c.scal[FJ_S_TIME];
    // ---- task_select / machine_select on the keys the previous observation left
    const int trule = task_rule0 + 1, mrule = mach_rule0 + 1;
    int q = fj_task_select<VARIANT, SUM_MODE>(c, trule, rnd_task), m = -1;
    if (q >= 0) m = fj_machine_select<VARIANT, SUM_MODE>(c, mrule, q, rnd_mach);
    q = fj_bcast_i(q, 0); m = fj_bcast_i(m, 0);
    if (q < 0 || m < 0) {
        if (lane == 0) c.scal[FJ_S_ERROR] |= (q < 0 ? FJ_E_NO_TASK : FJ_E_NO_MACHINE);
        if (rec) for (int k = lane; k < 8; k += FJ_NL) rec[k] = -1;
        fj_sync();
        return 0;
    }
    // ---- dispatch (one lane; a handful of scalar updates)
    if (lane == 0) {
        const int r = rjkind[q], stage = rjstage[q];
        const int norder = c.scal[FJ_S_NEXTORDER];
        int n;
        if (stage == 0) n = cum[norder * Kx + r] - c.qlen[q];
        else { n = c.qhead[q]; c.qhead[q] = c.next[jobbase[r] + n]; }
        c.qlen[q] = (uint16_t)(c.qlen[q] - 1);
        if (stage > 0 && c.qlen[q] == 0) { c.qhead[q] = 0xFFFF; c.qtail[q] = 0xFFFF; }
        const int s = fj_order_of(c.h_cum, c.S, c.Kx, r, n);
        c.cntunp[q * Sx + s] = (uint16_t)(c.cntunp[q * Sx + s] - 1);
        c.cntnow[q * Sx + s] = (uint16_t)(c.cntnow[q * Sx + s] - 1);
        c.proc[q] += 1;
        const int dur = FJ_I(c, ptime)[q * Mx + m];
        int t_begin = t, t_end = t + dur, m_end = t_end;
        if (VARIANT == FJSP_MO_BREAKDOWN) {   // MO_DFJSP_breakdown.py:204-231
            const FjRO bp = FJ_I(c, bdptr), bs_ = FJ_I(c, bds), be_ = FJ_I(c, bde);
            FJ_NOUNROLL
            for (int i = bp[m]; i < bp[m + 1]; ++i) {
                const int bs = bs_[i], be = be_[i];
                if (bs <= t && t < be) { int d = be - t; t_begin += d; t_end += d; m_end = t_end; }
                else if (t < bs && bs < t_end) { t_end += be - bs; m_end = t_end; }
                else if (bs == t_end) { m_end += be - bs; }
                else if (bs > t_end) break;
            }
        }
        const int prev_last = c.mlast[m];
        const unsigned bit = 1u << m;
        const int had = ((unsigned)c.scal[FJ_S_HASTASK] & bit) != 0;
        fj_set_ll(c.scal, FJ_S_MENDSUM, fj_get_ll(c.scal, FJ_S_MENDSUM) + m_end - c.mend[m]);
        c.mend[m] = m_end; c.mlast[m] = t_end; c.mjob[m] = (q << 16) | n;
        c.scal[FJ_S_BUSY] = (int)((unsigned)c.scal[FJ_S_BUSY] | bit);
        c.scal[FJ_S_HASTASK] = (int)((unsigned)c.scal[FJ_S_HASTASK] | bit);
        const int sl = c.slot[q * Mx + m];
        if (sl == 0xFFFF) c.pk[q * Mx + m] = (uint16_t)(c.pk[q * Mx + m] + 1);
        else c.fu[sl] = fj_sub(c.fu[sl], 1.0);
        c.mD[m] += 1;
        if (MO) {
            if (t_end > c.scal[FJ_S_COMPLETION]) c.scal[FJ_S_COMPLETION] = t_end;
            long long en = fj_get_ll(c.scal, FJ_S_ENERGY) + FJ_I(c, energy)[q * Mx + m];
            if (had) en += (long long)(t - prev_last) * FJ_I(c, idlep)[m];
            fj_set_ll(c.scal, FJ_S_ENERGY, en);
        }
        if (VARIANT == FJSP_SO_FJSSP) c.unpmask[q * c.P->d.NWx + (n >> 5)] &= ~(1u << (n & 31));
        if (rjlast[q]) {
            c.scal[FJ_S_LEFT] -= 1;
            long long late = (long long)t_end - (VARIANT == FJSP_SO_FJSSP ? c.duejob[jobbase[r] + n] : due[s]);
            if (late > 0) fj_set_ll(c.scal, FJ_S_DELAY_PROC, fj_get_ll(c.scal, FJ_S_DELAY_PROC) + late);
        }
        if (rec) {
            rec[0] = q; rec[1] = r; rec[2] = stage; rec[3] = n; rec[4] = m;
            rec[5] = t_begin; rec[6] = t_end; rec[7] = m_end;
        }
    }
    fj_sync();
    return 1;
}

// advance the clock while nothing can be dispatched (SO_DFJSP.py:206-253).
// returns FJ_CLK_PARKED when parked on an order arrival (SUSPEND only; `resume` re-enters
// right after that arrival), FJ_CLK_DONE when the episode finished, else 0.
enum { FJ_CLK_PARKED = 1, FJ_CLK_DONE = 2 };
template <int SUM_MODE, int SUSPEND>
FJ_FN_NOINLINE int fj_clock(int resume)
{
    FjCtx &c = FJ_CTX;
    int done = 0;
    const int lane = fj_lane();
    const int M = c.M, KT = c.KT, S = c.S, Sx = c.Sx;
    const FjKindRO rjkind = fj_rjkind(c);
    const FjLastRO rjlast = fj_rjlast(c);
    const FjEligRO elig = fj_elig(c);
    const FjRO arrive = FJ_I(c, arrive);
    const FjDueRO jobbase = fj_jobbase(c);
    int t = c.scal[FJ_S_TIME];
    FJ_NOUNROLL
    for (;;) {
        long long left = 1;
        int norder, arr_time;
        if (!resume) {
            const unsigned idle = ~(unsigned)c.scal[FJ_S_BUSY] & c.mmask;
            int any = 0;
            FJ_NOUNROLL
            for (int x = lane; x < KT; x += FJ_NL) any |= (c.qlen[x] > 0 && ((unsigned)elig[x] & idle) != 0);
            if (fj_any(any)) break;
            int tmin = 0x7fffffff;
            FJ_NOUNROLL
            for (int i = lane; i < M; i += FJ_NL) { int e = c.mend[i]; if (e > t && e < tmin) tmin = e; }
            tmin = fj_min_i(tmin);
            if (tmin == 0x7fffffff) { if (lane == 0) c.scal[FJ_S_ERROR] |= FJ_E_NO_EVENT; break; }
            t = tmin;
            unsigned rel = 0;   // machines that complete now and hold a job
            FJ_NOUNROLL
            for (int i = lane; i < M; i += FJ_NL) if (c.mend[i] == t && c.mjob[i] >= 0) rel |= 1u << i;
            rel = fj_or_u(rel);
            if (lane == 0) {   // they release their jobs in ascending machine order
                while (rel) {
                    const int i = fj_ffs0(rel); rel &= rel - 1;
                    const int jq = c.mjob[i] >> 16, n = c.mjob[i] & 0xffff;
                    if (rjlast[jq]) continue;
                    const int q2 = jq + 1, r = rjkind[jq];
                    if (c.qlen[q2] == 0) c.qhead[q2] = (uint16_t)n;
                    else c.next[jobbase[r] + c.qtail[q2]] = (uint16_t)n;
                    c.qtail[q2] = (uint16_t)n;
                    c.qlen[q2] = (uint16_t)(c.qlen[q2] + 1);
                    const int s = fj_order_of(c.h_cum, c.S, c.Kx, r, n);
                    c.cntnow[q2 * Sx + s] = (uint16_t)(c.cntnow[q2 * Sx + s] + 1);
                }
            }
            fj_sync();
            left = c.scal[FJ_S_LEFT];   // maintained at dispatch / arrival
            norder = c.scal[FJ_S_NEXTORDER];
            arr_time = c.scal[FJ_S_ARRTIME];
            fj_sync();
            const int by_time = norder < S && arrive[norder] <= t;
            if (by_time || (norder < S && left == 0)) {
                arr_time = arrive[norder];
                if (!by_time) t = arr_time;
                if (lane == 0) { c.scal[FJ_S_NEXTORDER] = norder + 1; c.scal[FJ_S_ARRTIME] = arr_time; }
                fj_sync();
                fj_arrival_begin(c, norder);
                if (SUSPEND) {
                    if (lane == 0) c.scal[FJ_S_TIME] = t;
                    fj_sync();
                    return FJ_CLK_PARKED;
                } else {
                    fj_order_arrives_inline<SUM_MODE>(norder, 0);
                    ++norder; left = 1;
                }
            }
        } else {
            resume = 0;
            norder = c.scal[FJ_S_NEXTORDER];
            arr_time = c.scal[FJ_S_ARRTIME];
        }
        unsigned freed = 0;
        FJ_NOUNROLL
        for (int i = lane; i < M; i += FJ_NL) if (c.mend[i] <= t) freed |= 1u << i;
        freed = fj_or_u(freed);
        if (lane == 0) {
            c.scal[FJ_S_BUSY] = (int)((unsigned)c.scal[FJ_S_BUSY] & ~freed);
            fj_set_d(c.scal, FJ_S_GAPTIME, (double)(t - arr_time));
        }
        fj_sync();
        if (norder >= S && left == 0) { done = 1; break; }
    }
    if (lane == 0) c.scal[FJ_S_TIME] = t;
    fj_sync();
    return done ? FJ_CLK_DONE : 0;
}

// bookkeeping after the clock loop: new observation, rule cache; returns the reward (lane 0)
template <int VARIANT, int SUM_MODE>
FJ_FN_NOINLINE double fj_step_back(int done, int reward_policy, double completion_n, double tardiness_n, double energy_n)
{
    FjCtx &c = FJ_CTX;
    double rew = 0.0;
    const int lane = fj_lane();
    const bool MO = (VARIANT == FJSP_MO_DFJSP || VARIANT == FJSP_MO_BREAKDOWN);
    if (lane == 0) {
        c.scal[FJ_S_STEPS] += 1;
        if (done) c.scal[FJ_S_DONE] = 1;
    }
    fj_sync();
    fj_observe<VARIANT, SUM_MODE>(done);
    if (lane == 0) {
        const long long dsum = fj_get_ll(c.scal, FJ_S_DELAY_PROC) + fj_get_ll(c.scal, FJ_S_DELAY_UNPROC);
        const long long dlast = fj_get_ll(c.scal, FJ_S_DELAY_LAST);
        if (!MO) rew = -(double)(dsum - dlast);
        else {
            const long long comp = c.scal[FJ_S_COMPLETION], compl_ = c.scal[FJ_S_COMPLETION_LAST];
            const long long en = fj_get_ll(c.scal, FJ_S_ENERGY), enl = fj_get_ll(c.scal, FJ_S_ENERGY_LAST);
            if (reward_policy == 0) rew = (double)(compl_ - comp);
            else if (reward_policy == 1) rew = (double)(dlast - dsum);
            else if (reward_policy == 2) rew = (double)(enl - en);
            else if (reward_policy == 3) {
                const double a = fj_div((double)(compl_ - comp), completion_n);
                const double e = fj_div((double)(enl - en), energy_n);
                if (tardiness_n > 0) rew = fj_add(fj_add(a, fj_div((double)(dlast - dsum), tardiness_n)), e);
                else rew = fj_add(a, e);
            }
            c.scal[FJ_S_COMPLETION_LAST] = (int)comp;
            fj_set_ll(c.scal, FJ_S_ENERGY_LAST, en);
        }
        fj_set_ll(c.scal, FJ_S_DELAY_LAST, dsum);
    }
    fj_sync();
    return rew;
}

// ---------------------------------------------------------------- per-env driver
// Steps tt0..T-1 of one environment: emit [v(t+1), v(t+1)-v(t)], reward, done and the
// dispatch record; auto-reset a finished environment before its next action.
// SUSPEND = 1: main kernel, parks on the first LP it needs.  SUSPEND = 0: resume kernel
// (starts from the parked phase / step) and the one-lane host build's second pass.
FJ_FN void fj_stage_copy(unsigned char *dst, const unsigned char *src, int bytes)
{
    // records are 16-byte aligned multiples of 16 bytes: 128-bit coalesced copies
    const int n = bytes >> 4;
#ifdef FJ_DEVICE_CODE
    const uint4 *s4 = (const uint4 *)src; uint4 *d4 = (uint4 *)dst;
    FJ_NOUNROLL
    for (int i = fj_lane(); i < n; i += FJ_NL) d4[i] = s4[i];
#else
    memcpy(dst, src, (size_t)n * 16);
#endif
    fj_sync();
}

template <int VARIANT, int SUM_MODE, int SUSPEND>
FJ_FN void fj_env_rollout_body(const FjParams &P, const FjStepArgs &A, int env, unsigned char *lp, unsigned char *hot);

// `stage`: this warp's shared-memory slab (hot part of the env record lives there for the
// whole launch) or null (work on the record in HBM/L2 directly).
template <int VARIANT, int SUM_MODE, int SUSPEND>
FJ_FN void fj_env_rollout(const FjParams &Pin, const FjStepArgs &A, int env, unsigned char *lp, unsigned char *stage = nullptr)
{
    const FjParams &P = fj_params_bind(Pin);
    unsigned char *G = P.env + (size_t)env * P.eo.stride;
    if (stage) {
        fj_stage_copy(stage, G, P.eo.hot);
    }
    fj_env_rollout_body<VARIANT, SUM_MODE, SUSPEND>(P, A, env, lp, stage);
    if (stage) fj_stage_copy(G, stage, P.eo.hot);
}

template <int VARIANT, int SUM_MODE, int SUSPEND>
FJ_FN void fj_env_rollout_body(const FjParams &P, const FjStepArgs &A, int env, unsigned char *lp, unsigned char *hot)
{
    const int lane = fj_lane();
    FjCtx &c = fj_ctx_init(P, env, lp, hot);
    const int nobs = P.nobs, ns = 2 * nobs;
    int tt = c.scal[FJ_S_PHASE] == FJ_PH_RUN ? 0 : c.scal[FJ_S_TT];   // a parked env continues at the step it parked in
    FJ_NOUNROLL
    for (; tt < A.T; ++tt) {
        const size_t i = (size_t)tt * P.B + env;
        int phase = c.scal[FJ_S_PHASE];
        if (phase == FJ_PH_RUN && c.scal[FJ_S_DONE]) {
            if (!A.autoreset) {   // a finished env without auto-reset repeats its terminal output
                if (A.done && lane == 0) A.done[i] = 1;
                if (A.reward && lane == 0) A.reward[i] = 0.0;
                FJ_NOUNROLL
                for (int k = lane; k < ns; k += FJ_NL) {
                    double v = k < nobs ? c.obs[k] : 0.0;
                    if (A.state) A.state[i * ns + k] = v;
                    if (A.state32) A.state32[i * ns + k] = (float)v;
                }
                if (A.rec) for (int k = lane; k < 8; k += FJ_NL) A.rec[i * 8 + k] = -1;
                continue;
            }
            if (!fj_reset_from_plan<VARIANT, SUM_MODE>(c, P)) {
                fj_reset_begin(c, 0);
                if (lane == 0) c.scal[FJ_S_EPISODES] += 1;
                fj_sync();
                if (SUSPEND) { fj_suspend(c, P, A, env, FJ_PH_LP_RESET, tt); return; }
                else {
                    fj_order_arrives_inline<SUM_MODE>(0, 0);
                    fj_reset_finish<VARIANT, SUM_MODE>(c);
                }
            }
        } else if (phase == FJ_PH_LP_RESET) {
            fj_arrival_resume<SUM_MODE>(c, P);
            fj_reset_finish<VARIANT, SUM_MODE>(c);
            phase = FJ_PH_RUN;
        }
        double out_reward = 0.0;
        int out_done = 0;
        int resume = 0, ok = 1;
        if (phase == FJ_PH_LP_STEP) {
            fj_arrival_resume<SUM_MODE>(c, P);
            resume = 1;
        } else {
            ok = fj_step_front<VARIANT, SUM_MODE>(A.actions[2 * i], A.actions[2 * i + 1], A.rnd ? A.rnd[2 * i] : 0u,
                                                  A.rnd ? A.rnd[2 * i + 1] : 0u, A.rec ? A.rec + i * 8 : nullptr);
        }
        if (ok) {
            const int ck = fj_clock<SUM_MODE, SUSPEND>(resume);
            if (ck == FJ_CLK_PARKED) { fj_suspend(c, P, A, env, FJ_PH_LP_STEP, tt); return; }
            out_done = ck == FJ_CLK_DONE;
            out_reward = fj_step_back<VARIANT, SUM_MODE>(out_done, A.reward_policy, A.completion, A.tardiness, A.energy);
        } else {
            out_done = c.scal[FJ_S_DONE];
            FJ_NOUNROLL
            for (int k = lane; k < nobs; k += FJ_NL) c.obs2[k] = c.obs[k];
            fj_sync();
        }
        // outputs: lanes stream the state vector, lane 0 the scalars
        FJ_NOUNROLL
        for (int k = lane; k < ns; k += FJ_NL) {
            const int j = k < nobs ? k : k - nobs;
            const double v = k < nobs ? c.obs2[j] : fj_sub(c.obs2[j], c.obs[j]);
            if (A.state) A.state[i * ns + k] = v;
            if (A.state32) A.state32[i * ns + k] = (float)v;
        }
        fj_sync();
        FJ_NOUNROLL
        for (int k = lane; k < nobs; k += FJ_NL) c.obs[k] = c.obs2[k];
        if (lane == 0) {
            if (A.reward) A.reward[i] = out_reward;
            if (A.done) A.done[i] = out_done;
        }
        fj_sync();
    }
}

// ---------------------------------------------------------------- main kernel driver
// The step kernel's CTAs have two roles.
//   * ENV CTAs: every warp plays one environment copy.  The warps run in LOCKSTEP slots, one named
//     barrier per slot: all of them execute the same few KB of straight-line code at the same time and
//     share its instruction-cache lines (the per-step code is ~70 KB against a 32 KB L1.5 instruction
//     cache; free-running warps re-fetched it from L2: "no instruction" was the top stall,
//     profiles/README.md).  A slot is one env step.
//   * LP-SERVER CTAs (the grid's first P.srv_ctas CTAs) play no environment: their warps form groups
//     that serve fluid LPs.  An env warp whose clock reaches an order arrival writes what the LP is
//     built from (unprocessed operations per type, which waiting queues are empty) to its request
//     record in HBM/L2, takes a ticket on the queue and leaves the lockstep group: it free-runs,
//     polling for the solution, while its CTA-mates keep stepping.  A server group claims the ticket,
//     solves the LP with B^-1 in its shared memory and flags the solution ready.
//     (Round 1 stopped the whole CTA for every LP: 45 % of all warp stalls were that barrier.  The
//     first round-2 version kept a 4-warp LP team inside every env CTA: its warps ran code no other
//     warp of the SM was running and starved on instruction fetch -- "no instruction" was their top
//     stall, 23 cycles per instruction, 0.5 M cycles per LP.  On a server SM only LP code runs.)
#define FJ_BAR_ENV 1
#define FJ_BAR_ROUND 3
#define FJ_BAR_SRV0 4                  // server groups: barrier FJ_BAR_SRV0 + group
#define FJ_SLOT_DETACHED 0x40000000   // P.order flag: the env free-runs (it is expected to meet a fluid LP or a reset in this launch)

FJ_FN void fj_emit_state(const FjCtx &c, const FjStepArgs &A, size_t i, int nobs, int terminal)
{
    const int ns = 2 * nobs;
    FJ_NOUNROLL
    for (int k = fj_lane(); k < ns; k += FJ_NL) {
        const int j = k < nobs ? k : k - nobs;
        double v;
        if (terminal) v = k < nobs ? c.obs[j] : 0.0;
        else v = k < nobs ? c.obs2[j] : fj_sub(c.obs2[j], c.obs[j]);
        if (A.state) A.state[i * ns + k] = v;
        if (A.state32) A.state32[i * ns + k] = (float)v;
    }
}

// per-CTA shared state of the env role
struct FjLpBoard {
    int meta[64];     // host build: iterations, return code of the in-line LP
    unsigned long long mbar[32];   // one mbarrier per env warp: TMA stage-in of its record's hot prefix
};

// per-warp view of the main kernel's env role
struct FjCtaCtx {
    int warp, nwarps, cta_lp;    // env warp index, env warps of the CTA, 1: LP servers present
    int gslot;                   // env-warp slot of the launch: env CTA index * env warps + warp (request / response records)
    unsigned char *slab;         // host build: LP scratch of the in-line solve
    double *xbuf;                // this warp's LP solution buffer [NPx] (host build: the one buffer)
    FjLpBoard *board;
    FjCtaGroup group;            // host build: one thread
};

#ifdef __CUDACC__
FJ_FN int fj_env_vote(int pred, int nthreads, int bar)   // barrier `bar` of the lockstep group + number of its threads with pred
{
    unsigned r;
    asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %1, 0;\n bar.red.popc.u32 %0, %2, %3, q;\n}"
                 : "=r"(r) : "r"(pred), "r"(bar), "r"(nthreads) : "memory");
    return (int)r;
}
FJ_FN int fj_env_count(int pred, int nthreads)   // barrier of ALL env warps + number of threads with pred
{
    unsigned r;
    asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %1, 0;\n bar.red.popc.u32 %0, %2, %3, q;\n}"
                 : "=r"(r) : "r"(pred), "n"(FJ_BAR_ROUND), "r"(nthreads) : "memory");
    return (int)r;
}
FJ_FN unsigned fj_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// TMA bulk copies of a record's hot prefix (16-byte aligned, multiple of 16 bytes): global ->
// shared completes on the warp's mbarrier, shared -> global is a bulk group of the issuing lane
FJ_FN void fj_tma_load(unsigned char *dst, const unsigned char *src, unsigned bytes, unsigned long long *mbar, unsigned parity)
{
    // the slab was last touched through the generic proxy (previous round) and read by the bulk store
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const unsigned bar = fj_smem_addr(mbar);
    if (fj_lane() == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(fj_smem_addr(dst)), "l"(src), "r"(bytes), "r"(bar) : "memory");
    }
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
FJ_FN void fj_tma_store(unsigned char *dst, const unsigned char *src, unsigned bytes)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // every lane's writes, then the issuing lane
    __syncwarp();
    if (fj_lane() == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(fj_smem_addr(src)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
}
FJ_FN void fj_tma_prefetch_l2(const unsigned char *src, unsigned bytes)   // one lane
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
#endif

// FJ_TRACE builds (tools/cta_trace.py): every warp accumulates the cycles it spends in each
// phase of a step (barrier waits excluded) so that CTA / SM imbalance can be read off.
#if defined(FJ_TRACE) && defined(FJ_DEVICE_CODE)
#define FJ_TR_DECL long long tr_[6] = {0, 0, 0, 0, 0, 0}; long long tr_t_ = clock64(); const long long tr_begin_ = tr_t_
#define FJ_TR_MARK() (tr_t_ = clock64())
#define FJ_TR_ACC(k) do { const long long n_ = clock64(); tr_[k] += n_ - tr_t_; tr_t_ = n_; } while (0)
#define FJ_TR_COUNT(k) (tr_[k] += 1)
#define FJ_TR_FLUSH(P, K) do { if ((P).trace && fj_lane() == 0) { long long *o_ = (P).trace + ((size_t)blockIdx.x * FJ_TRACE_ROWS + (K).warp) * 8; \
    unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); \
    o_[0] += clock64() - tr_begin_; for (int k_ = 0; k_ < 6; ++k_) o_[1 + k_] += tr_[k_]; o_[7] = sm_; } } while (0)
#else
#define FJ_TR_DECL
#define FJ_TR_MARK()
#define FJ_TR_ACC(k)
#define FJ_TR_COUNT(k)
#define FJ_TR_FLUSH(P, K)
#endif

// one LP for the env whose context is c2, solved in line by group g on `slab` (the one-lane host build)
FJ_FN void fj_lp_for_ctx(const FjParams &P, const FjCtaGroup &g, FjCtx &c2, unsigned char *slab, double *x, int *meta)
{
    FjLp L;
    fj_lp_carve(L, slab, slab + (size_t)P.d.Rx * P.d.Rx * 8, P.d);
    int iters = 0;
    const int rc = fj_lp_solve(g, c2, L, x, &iters);
    if (g.rank() == 0) { meta[0] = iters; meta[1] = rc; }
    g.sync();
}

#ifdef __CUDACC__
// ---- the LP queue (HBM/L2).  Tickets are 32-bit and never reset: an entry carries ticket + 1, so a
// stale entry of an earlier lap (or launch) never matches.
// env warp (all lanes): write the request, take a ticket
FJ_FN void fj_lp_post(const FjParams &P, const FjCtx &c, int gslot, int env)
{
    const int lane = fj_lane();
    int *rq = P.lp_req + (size_t)gslot * P.lp_req_stride;
    const int KT = c.KT, KTx = P.d.KTx;
    for (int q0 = 0; q0 < KT + 1; q0 += 32) {   // (the mask's last word covers bit KT, which nothing reads)
        const int q = q0 + lane;
        if (q < KT) rq[2 + q] = c.fstart[q];
        const unsigned e = __ballot_sync(0xffffffffu, q < KT && c.qlen[q] == 0);
        if (lane == 0) rq[2 + KTx + (q0 >> 5)] = (int)e;
    }
    if (lane == 0) { rq[0] = env; P.lp_resp[4 * gslot] = 0; }
    __threadfence();      // the request is in L2 before its ticket
    __syncwarp();
    if (lane == 0) {
        const unsigned t = atomicAdd(P.lpq + 1, 1u);
        *(volatile unsigned long long *)(P.lpq_ring + (t & (FJ_LPQ_RING - 1))) = ((unsigned long long)(t + 1u) << 32) | (unsigned)gslot;
    }
}

// A server group's loop: the leader claims a ticket, the group solves that LP into the requester's
// solution buffer and flags it ready.  Leaves when every env CTA has finished its launch.
// (out of line: its register allocation is its own, not the env role's)
FJ_FN_NOINLINE void fj_lp_server_loop(int gid, int gw, unsigned char *gsmem, int env_ctas, int slab_index)
{
    const FjParams &P = fj_sP;
    FjCtaGroup g;
    g.red = nullptr; g.flip = 0; g.base = gid * gw * 32; g.nthr = gw * 32; g.bar = FJ_BAR_SRV0 + gid; g.dbar = FJ_BAR_SRV0 + 6 + gid;
    const FjLpFastSmem S = fj_lpf_state(gsmem, P.d);
    unsigned char *scratch = gsmem + fj_lpf_state_bytes(P.d);
    const int scratch_bytes = P.srv_group_smem - (int)fj_lpf_state_bytes(P.d);
    unsigned char *slab = P.lp + (size_t)slab_index * P.lp_stride;
    volatile unsigned *q = P.lpq;
    volatile unsigned *finished = (volatile unsigned *)P.pend_count + FJ_ROUNDS + 1;   // env CTAs done in this launch (zeroed per launch)
    for (;;) {
        if (g.rank() == 0) {
            // the leader alone polls (the group's other threads wait at the barrier below: an idle group executes
            // no instructions but its leader's; with the whole group going round the loop per poll the idle servers
            // were 20 % of the launch's warp instructions)
            int got = -1;
            unsigned h = 0;
            for (;;) {
                h = q[0];
                const unsigned t = q[1];
                if ((int)(t - h) > 0) {
                    if (atomicCAS(P.lpq, h, h + 1u) == h) { got = (int)(h & 0x3fffffffu); break; }
                } else if (*finished >= (unsigned)env_ctas) {
                    // every env CTA has finished (all their requests were answered before): leave unless a ticket slipped in between
                    const unsigned h2 = q[0], t2 = q[1];
                    if (h2 == t2) { got = -2; break; }
                } else __nanosleep(200);
            }
            S.ctl()[2] = got;
            if (got >= 0) {
                // wait for the entry of ticket h (its poster writes it right after taking the ticket)
                const volatile unsigned long long *e = P.lpq_ring + (h & (FJ_LPQ_RING - 1));
                unsigned long long v;
                do { v = *e; } while ((unsigned)(v >> 32) != h + 1u);
                S.ctl()[3] = (int)(unsigned)v;     // the requester's env-warp slot
                __threadfence();
            }
        }
        g.sync();
        const int got = S.ctl()[2];
        const int gslot = S.ctl()[3];
        g.sync();               // everyone has read the two words before the leader's next round rewrites them
        if (got == -2) break;
        if (got < 0) continue;
        const int *rq = P.lp_req + (size_t)gslot * P.lp_req_stride;
        FjLpIn in;
        const int env = __ldcg(rq);
        in.I = P.inst + (size_t)P.env_inst[env] * P.io.stride;
        in.M = __ldg(in.I + P.io.hdr); in.KT = __ldg(in.I + P.io.hdr + 2); in.NP = __ldg(in.I + P.io.hdr + 7);
        in.fstart = rq + 2; in.qlen = nullptr; in.empty = (const uint32_t *)(rq + 2 + P.d.KTx);
        double *x = P.cta_x + (size_t)gslot * P.d.NPx;
        int iters = 0;
        int rc = fj_lp_solve_fast(g, P, in, S, slab, x, &iters, scratch, scratch_bytes);
        if (g.rank() == 0) {
            if (rc < 0) { rc = 4; iters = 0; }   // beyond the fast path's row limit: reported as an LP error of the env
            P.lp_resp[4 * gslot + 1] = iters; P.lp_resp[4 * gslot + 2] = rc;
        }
        __threadfence();        // every thread's part of x is in L2 ...
        g.sync();
        if (g.rank() == 0) *(volatile int *)(P.lp_resp + 4 * gslot) = 1;   // ... before the flag
    }
}
#endif

enum { FJ_ST_IDLE = 0, FJ_ST_FRONT = 1, FJ_ST_WAIT = 2 };

// One round of a CTA's env warps: this warp plays `env` for A.T steps (or just keeps the slot
// barriers when it has no env).  `hotbuf`: the warp's shared-memory slab for the record's hot
// prefix (staged by TMA for the whole launch) or null (work on the record in HBM/L2 directly).
// `next_env`: the env this warp plays in the CTA's next round (-1 none): its record is
// prefetched into L2 while this round runs.
template <int VARIANT, int SUM_MODE>
FJ_FN void fj_cta_rollout(const FjParams &Pin, const FjStepArgs &A, const FjCtaCtx &K, int env, int active, unsigned char *hotbuf = nullptr,
                          int next_env = -1, unsigned parity = 0, int lock_threads = 0, int lock_bar = FJ_BAR_ENV)
{
    const FjParams &P = fj_params_bind(Pin);
    const int lane = fj_lane();
    FjCtx &c = FJ_CTX;
    unsigned char *G = P.env + (size_t)env * P.eo.stride;
    if (active) {
        if (hotbuf) {
#ifdef FJ_DEVICE_CODE
            fj_tma_load(hotbuf, G, (unsigned)P.eo.hot, &K.board->mbar[K.warp], parity);
            if (next_env >= 0 && lane == 0) fj_tma_prefetch_l2(P.env + (size_t)next_env * P.eo.stride, (unsigned)P.eo.stride);
#else
            fj_stage_copy(hotbuf, G, P.eo.hot);
#endif
        }
        fj_ctx_init(P, env, K.slab, hotbuf);
    }
    (void)next_env; (void)parity;
    const int nobs = P.nobs;
    FJ_TR_DECL;
    const size_t batch = (size_t)P.B;
    size_t i = (size_t)env;             // row of this env in the [T][B] inputs / outputs of its step tt
    int tt = 0;
    int st = active && A.T > 0 ? FJ_ST_FRONT : FJ_ST_IDLE;
    // lock_threads: threads of this round's lockstep group (this warp is one of them), 0: this warp
    // free-runs (no slot barrier; it polls for its LP solutions)
    for (;;) {
        if (st != FJ_ST_IDLE) {
            int kind = 0;       // 0 nothing to emit, 1 dispatched (clock, observation), 2 nothing dispatchable (emit unchanged)
            int resume = 0;
            FJ_TR_MARK();
#ifdef FJ_DEVICE_CODE
            if (st == FJ_ST_WAIT) {
                int ready = lane == 0 ? *(volatile int *)(P.lp_resp + 4 * K.gslot) : 0;
                ready = fj_bcast_i(ready, 0);
                if (!ready) { __nanosleep(lock_threads ? 200 : 100); FJ_TR_ACC(2); goto vote; }
                __threadfence();   // the solution was written by another SM: fj_arrival_finish reads it through L2
                fj_arrival_finish<SUM_MODE>(c, K.xbuf, __ldcg(P.lp_resp + 4 * K.gslot + 1), __ldcg(P.lp_resp + 4 * K.gslot + 2));
                resume = 1; kind = 1; st = FJ_ST_FRONT;
                FJ_TR_ACC(2);
            } else
#endif
            {
                // ---- auto-reset / task_select / machine_select / dispatch
                if (c.scal[FJ_S_DONE]) {
                    if (!A.autoreset) {   // a finished env without auto-reset repeats its terminal output
                        if (A.done && lane == 0) A.done[i] = 1;
                        if (A.reward && lane == 0) A.reward[i] = 0.0;
                        fj_emit_state(c, A, i, nobs, 1);
                        if (A.rec) for (int k = lane; k < 8; k += FJ_NL) A.rec[i * 8 + k] = -1;
                    } else if (!fj_reset_from_plan<VARIANT, SUM_MODE>(c, P)) {
                        // no cached order-0 solution: the LP / resume kernels finish this env's launch
                        fj_reset_begin(c, 0);
                        if (lane == 0) c.scal[FJ_S_EPISODES] += 1;
                        fj_sync();
                        fj_suspend(c, P, A, env, FJ_PH_LP_RESET, tt);
                        st = FJ_ST_IDLE;
                        FJ_TR_ACC(0);
                        goto vote;
                    }
                }
                if (!c.scal[FJ_S_DONE]) {
                    const int ok = fj_step_front<VARIANT, SUM_MODE>(A.actions[2 * i], A.actions[2 * i + 1], A.rnd ? A.rnd[2 * i] : 0u,
                                                                   A.rnd ? A.rnd[2 * i + 1] : 0u, A.rec ? A.rec + i * 8 : nullptr);
                    kind = ok ? 1 : 2;
                }
                FJ_TR_ACC(0);
            }
            // ---- discrete-event clock.  At an order arrival the env needs the fluid LP of its
            // new state: an LP-server group solves it (device) / it is solved in line (host build)
            int done = 0;
            if (kind == 1) {
                for (;;) {
                    const int ck = fj_clock<SUM_MODE, 1>(resume);
                    if (ck != FJ_CLK_PARKED) { done = ck == FJ_CLK_DONE; break; }
                    if (!K.cta_lp) {             // no LP servers in this launch: park for the LP / resume kernels
                        fj_suspend(c, P, A, env, FJ_PH_LP_STEP, tt);
                        st = FJ_ST_IDLE; kind = 0;
                        break;
                    }
#ifdef FJ_DEVICE_CODE
                    FJ_TR_COUNT(4);
                    {   // a long queue (every copy meeting its arrivals at once): solve in line, every warp its own LP
                        int over = 0;
                        if (lane == 0 && K.slab && P.lp_overflow > 0) {
                            const volatile unsigned *q = P.lpq;
                            over = (int)(q[1] - q[0]) >= P.lp_overflow;
                        }
                        if (fj_bcast_i(over, 0)) {
                            fj_order_arrives_inline<SUM_MODE>(0, 0);
                            resume = 1;
                            continue;
                        }
                    }
                    fj_lp_post(P, c, K.gslot, env);
                    st = FJ_ST_WAIT; kind = 0;
                    break;
#else
                    fj_lp_for_ctx(P, K.group, c, K.slab, K.xbuf, K.board->meta);
                    fj_arrival_finish<SUM_MODE>(c, K.xbuf, K.board->meta[0], K.board->meta[1]);
                    resume = 1;
#endif
                }
                FJ_TR_ACC(1);
                if (st != FJ_ST_FRONT) goto vote;
            }
            // ---- observation, reward, outputs
            if (kind) {
                double out_reward = 0.0;
                int out_done;
                if (kind == 1) {
                    out_reward = fj_step_back<VARIANT, SUM_MODE>(done, A.reward_policy, A.completion, A.tardiness, A.energy);
                    out_done = done;
                } else {
                    out_done = c.scal[FJ_S_DONE];
                    FJ_NOUNROLL
                    for (int k = lane; k < nobs; k += FJ_NL) c.obs2[k] = c.obs[k];
                    fj_sync();
                }
                fj_emit_state(c, A, i, nobs, 0);
                fj_sync();
                FJ_NOUNROLL
                for (int k = lane; k < nobs; k += FJ_NL) c.obs[k] = c.obs2[k];
                if (lane == 0) {
                    if (A.reward) A.reward[i] = out_reward;
                    if (A.done) A.done[i] = out_done;
                }
                fj_sync();
            }
            FJ_TR_ACC(3);
            ++tt; i += batch;
            if (tt >= A.T) st = FJ_ST_IDLE;
#ifdef FJ_DEVICE_CODE
            if (A.prog_count && ((tt & ((1 << A.prog_shift) - 1)) == 0 || tt >= A.T)) {
                // this env's outputs of a whole chunk of steps are written: count it; the last env of the batch
                // tells the host (which copies the chunk out while the launch goes on)
                __syncwarp();
                if (lane == 0) {
                    const int ch = (tt - 1) >> A.prog_shift;
                    __threadfence();
                    if (atomicAdd(A.prog_count + ch, 1u) + 1u == (unsigned)P.B) {
                        __threadfence_system();
                        *(volatile unsigned *)(A.prog_flag + ch) = A.prog_seq;
                    }
                }
            }
#endif
        }
    vote:
#ifdef FJ_DEVICE_CODE
        // slot barrier of the lockstep group.  A warp stays in the group only while it steps: one that
        // has done its T steps leaves, and so does one that waits for an LP (it free-runs from here
        // on: its mates must not slip a slot per poll); the vote tells the others the new group size.
        // (P.lock_mask = K - 1: the group meets every K steps only; a warp that leaves announces it at the
        // group's next meeting)
        if (lock_threads && ((tt & P.lock_mask) == 0 || st != FJ_ST_FRONT)) {
            const int stay = st == FJ_ST_FRONT;
            const int n = fj_env_vote(stay, lock_threads, lock_bar);
            lock_threads = stay ? n : 0;
        }
        if (st == FJ_ST_IDLE) break;
#else
        if (st == FJ_ST_IDLE) break;
#endif
    }
    if (active && hotbuf) {
#ifdef FJ_DEVICE_CODE
        fj_tma_store(G, hotbuf, (unsigned)P.eo.hot);
#else
        fj_stage_copy(G, hotbuf, P.eo.hot);
#endif
    }
    FJ_TR_FLUSH(P, K);
}

// reset() entry: phase 1 parks every env on its order-0 LP, phase 2 finishes
// `fresh`: the first reset() after creation (a new environment object); later calls are reset() of a USED
// object, with the reference's quirks (busy flags, order_arrive_time and the `done` seen by the first
// observation survive), exactly like the fused auto-reset
FJ_FN void fj_env_reset_begin(const FjParams &Pin, int env, int fresh = 1)
{
    const FjParams &P = fj_params_bind(Pin);
    FjCtx &c = fj_ctx_init(P, env, nullptr);
    fj_reset_begin(c, fresh);
    if (fj_lane() == 0) {
        c.scal[FJ_S_PHASE] = FJ_PH_LP_RESET;
        c.scal[FJ_S_LPSLOT] = env < P.lp_slots ? env : -1;
        P.pend_env[env] = env;   // list 0
    }
    fj_sync();
}

template <int VARIANT, int SUM_MODE>
FJ_FN void fj_env_reset_finish(const FjParams &Pin, int env, unsigned char *lp, double *state_out, float *state32_out)
{
    const FjParams &P = fj_params_bind(Pin);
    const int lane = fj_lane();
    FjCtx &c = fj_ctx_init(P, env, lp);
    fj_arrival_resume<SUM_MODE>(c, P);
    fj_reset_finish<VARIANT, SUM_MODE>(c);
    const int nobs = P.nobs, ns = 2 * nobs;
    FJ_NOUNROLL
    for (int k = lane; k < ns; k += FJ_NL) {
        const double v = k < nobs ? c.obs[k] : 0.0;
        if (state_out) state_out[(size_t)env * ns + k] = v;
        if (state32_out) state32_out[(size_t)env * ns + k] = (float)v;
    }
}

// one parked LP, solved by a whole CTA (device) / one thread (host build)
FJ_FN void fj_lp_service(const FjParams &Pin, const FjCtaGroup &g, const int *list, int idx, unsigned char *binv, unsigned char *small_, unsigned char *fast_slab = nullptr,
                         unsigned char *fast_smem = nullptr, int fast_smem_bytes = 0)
{
    const FjParams &P = fj_params_bind(Pin);
    const int env = list[idx];
    FjCtx &c = fj_ctx_init(P, env, nullptr);
    FjLp L;
    fj_lp_carve(L, binv, small_, P.d);
    int iters = 0;
    int rc = -1;
#ifdef FJ_DEVICE_CODE
    if (fast_slab && fast_smem && (size_t)fast_smem_bytes >= fj_lpf_state_bytes(P.d)) {
        const FjLpFastSmem S = fj_lpf_state(fast_smem, P.d);
        FjLpIn in;
        fj_lpin_from_ctx(in, c);
        const int sb = (int)fj_lpf_state_bytes(P.d);
        rc = fj_lp_solve_fast(g, P, in, S, fast_slab, P.lp_x + (size_t)idx * P.d.NPx, &iters, fast_smem + sb, fast_smem_bytes - sb);
    }
#endif
    (void)fast_slab; (void)fast_smem; (void)fast_smem_bytes;
    if (rc < 0) rc = fj_lp_solve(g, c, L, P.lp_x + (size_t)idx * P.d.NPx, &iters);
    if (g.rank() == 0) { P.lp_meta[2 * idx] = iters; P.lp_meta[2 * idx + 1] = rc; }
    g.sync();
}
