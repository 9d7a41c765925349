// TEST INFRASTRUCTURE ONLY.  Compiles the kernel source (csrc/fjsp_core.cuh) and the
// table builder (csrc/fjsp_host.h) with g++ as a ONE-LANE program, behind the same entry
// points as the CUDA library, so `pytest -m "not gpu"` can check the compressed state
// machine, the rule cache, the set-order emulation and the in-kernel simplex against the
// oracle on a machine without a GPU.  It is never loaded by the product package; the
// 32-lane cooperation itself is covered by the `-m gpu` parity tests.
// Build: g++ -O2 -ffp-contract=off -fPIC -shared (see tests/hostsim/build.py)
#include <stdlib.h>
#include <string>
#include <vector>
#include "../../deep_reinforcement_learning_for_fjsp_b200/csrc/fjsp_host.h"
#include "../../deep_reinforcement_learning_for_fjsp_b200/csrc/fjsp_core.cuh"

struct HostVec {
    FjTables tb;
    FjParams P;
    std::vector<int32_t> env_inst, pend_env, lp_meta;
    int pend_counts[FJ_ROUNDS + 1];
    std::vector<unsigned char> env, lp;
    std::vector<double> lp_x, plan_x;
    std::vector<int32_t> plan_meta, plan_ok;
    int pend_count;
    int variant, sum_mode;
    int n_resets = 0;
};

static void run_lp_service(HostVec *h, int round = 0)
{
    const int *list = h->pend_env.data() + (size_t)(round & 1) * h->P.B;
    h->pend_count = h->pend_counts[round];
    // the CTA-group LP kernel, one thread: Binv and the small arrays on the host slab
    const FjDims &d = h->tb.d;
    unsigned char *binv = h->lp.data();
    unsigned char *small_ = h->lp.data() + (size_t)d.Rx * d.Rx * 8;
    FjCtaGroup g; g.red = nullptr; g.flip = 0; g.whole_cta();
    int n = h->pend_count < h->P.lp_slots ? h->pend_count : h->P.lp_slots;
    for (int i = 0; i < n; ++i) fj_lp_service(h->P, g, list, i, binv, small_);
}

static std::string g_err;

template <int V, int SM>
static void run_reset(HostVec *h, double *state)
{
    for (int e = 0; e < h->P.B; ++e) fj_env_reset_begin(h->P, e, h->n_resets == 0 ? 1 : 0);
    h->n_resets += 1;
    h->pend_counts[0] = h->P.B;
    run_lp_service(h, 0);
    if (!h->P.plan_ok && !getenv("FJSP_HOSTSIM_NO_PLAN")) {   // cache each instance's order-0 LP solution
        const int np = h->tb.d.NPx, ni = h->tb.n_instances;
        h->plan_x.assign((size_t)ni * np, 0.0); h->plan_meta.assign((size_t)ni * 2, 0); h->plan_ok.assign(ni, 0);
        for (int e = h->P.B - 1; e >= 0; --e) {
            if (e >= h->P.lp_slots) continue;
            const int ii = h->env_inst[e];
            for (int j = 0; j < np; ++j) h->plan_x[(size_t)ii * np + j] = h->lp_x[(size_t)e * np + j];
            h->plan_meta[2 * ii] = h->lp_meta[2 * e]; h->plan_meta[2 * ii + 1] = h->lp_meta[2 * e + 1];
            h->plan_ok[ii] = 1;
        }
        h->P.plan_x = h->plan_x.data(); h->P.plan_meta = h->plan_meta.data(); h->P.plan_ok = h->plan_ok.data();
    }
    for (int e = 0; e < h->P.B; ++e) fj_env_reset_finish<V, SM>(h->P, e, h->lp.data(), state, nullptr);
}
template <int V, int SM>
static void run_step(HostVec *h, const FjStepArgs &A)
{
    // FJSP_HOSTSIM_STAGE=1: run on a staged copy of the record's hot part (the shared-memory path)
    static std::vector<unsigned char> slab;
    slab.assign(h->tb.eo.hot + 16, 0);
    unsigned char *stage = getenv("FJSP_HOSTSIM_STAGE") ? slab.data() : nullptr;
    for (int r = 0; r <= FJ_ROUNDS; ++r) h->pend_counts[r] = 0;
    FjStepArgs B_ = A;
    B_.park_count = &h->pend_counts[0]; B_.park_env = h->pend_env.data();
    FjCtaCtx K;
    static FjLpBoard board;
    K.warp = 0; K.nwarps = 1; K.cta_lp = h->P.cta_lp;
    K.slab = h->lp.data(); K.gslot = 0;
    K.xbuf = (double *)(K.slab + (size_t)h->tb.d.Rx * h->tb.d.Rx * 8 + (fj_lp_small_bytes(h->tb.d) + 7) / 8 * 8);
    K.board = &board; K.group.red = nullptr; K.group.flip = 0; K.group.whole_cta();
    for (int e = 0; e < h->P.B; ++e) fj_cta_rollout<V, SM>(h->P, B_, K, e, 1, stage);               // main kernel
    for (int r = 0; r < FJ_ROUNDS; ++r) {
        run_lp_service(h, r);                                                                       // LP kernel
        const int *list = h->pend_env.data() + (size_t)(r & 1) * h->P.B;
        const int n = h->pend_counts[r];
        B_.park_count = &h->pend_counts[r + 1]; B_.park_env = h->pend_env.data() + (size_t)((r + 1) & 1) * h->P.B;
        for (int i = 0; i < n; ++i) {                                                               // resume kernel
            if (r + 1 < FJ_ROUNDS) fj_env_rollout<V, SM, 1>(h->P, B_, list[i], h->lp.data(), nullptr);
            else fj_env_rollout<V, SM, 0>(h->P, B_, list[i], h->lp.data(), nullptr);
        }
    }
}

#define DISPATCH(fn, ...)                                                                      \
    do {                                                                                       \
        int key = h->variant * 2 + (h->sum_mode ? 1 : 0);                                      \
        switch (key) {                                                                         \
        case 0: fn<FJSP_SO_DFJSP, 0>(__VA_ARGS__); break;                                      \
        case 1: fn<FJSP_SO_DFJSP, 1>(__VA_ARGS__); break;                                      \
        case 2: fn<FJSP_MO_DFJSP, 0>(__VA_ARGS__); break;                                      \
        case 3: fn<FJSP_MO_DFJSP, 1>(__VA_ARGS__); break;                                      \
        case 4: fn<FJSP_MO_BREAKDOWN, 0>(__VA_ARGS__); break;                                  \
        case 5: fn<FJSP_MO_BREAKDOWN, 1>(__VA_ARGS__); break;                                  \
        case 6: fn<FJSP_SO_FJSSP, 0>(__VA_ARGS__); break;                                      \
        case 7: fn<FJSP_SO_FJSSP, 1>(__VA_ARGS__); break;                                      \
        }                                                                                      \
    } while (0)

extern "C" {

const char *fjsp_hostsim_last_error(void) { return g_err.c_str(); }

int fjsp_hostsim_create(const int32_t *blobs, const int64_t *offsets, int n_inst, const int32_t *env_instance,
                        int n_envs, int variant, int sum_mode, void **out)
{
    if (variant < 0 || variant > 3) { g_err = "unknown variant"; return -2; }
    HostVec *h = new HostVec();
    if (!fj_build_tables(blobs, offsets, n_inst, h->tb, g_err, variant)) { delete h; return -1; }
    h->variant = variant; h->sum_mode = sum_mode;
    h->env_inst.assign(env_instance, env_instance + n_envs);
    h->env.assign((size_t)n_envs * h->tb.eo.stride, 0);
    h->lp.assign(fj_lp_scratch_bytes(h->tb.d), 0);
    FjParams &P = h->P;
    P.d = h->tb.d; P.io = h->tb.io; P.eo = h->tb.eo;
    P.inst = h->tb.inst.data(); P.env_inst = h->env_inst.data(); P.env = h->env.data();
    P.lp = h->lp.data(); P.lp_stride = 0;
    h->pend_env.assign((size_t)2 * n_envs, 0);
    const char *ov = getenv("FJSP_HOSTSIM_LP_SLOTS");   // tests: force the "no free slot" in-line path
    P.lp_slots = ov ? atoi(ov) : n_envs;
    h->lp_x.assign((size_t)(P.lp_slots > 0 ? P.lp_slots : 1) * h->tb.d.NPx, 0.0);
    h->lp_meta.assign((size_t)(P.lp_slots > 0 ? P.lp_slots : 1) * 2, 0);
    P.stage = 0; P.env_warps = 1; P.cta_x = nullptr; P.stage_stride = h->tb.eo.hot;
    P.cta_lp = getenv("FJSP_HOSTSIM_NO_CTA_LP") ? 0 : 1;
    P.plan_x = nullptr; P.plan_meta = nullptr; P.plan_ok = nullptr;
    h->pend_count = 0;
    P.pend_count = &h->pend_count; P.pend_env = h->pend_env.data(); P.lp_x = h->lp_x.data(); P.lp_meta = h->lp_meta.data();
    P.B = n_envs; P.variant = variant; P.sum_mode = sum_mode;
    P.nobs = (variant == FJSP_SO_DFJSP || variant == FJSP_SO_FJSSP) ? 10 : 15;
    *out = h;
    return 0;
}

int fjsp_hostsim_destroy(void *v) { delete (HostVec *)v; return 0; }

int fjsp_hostsim_reset(void *v, double *state)
{
    HostVec *h = (HostVec *)v;
    DISPATCH(run_reset, h, state);
    return 0;
}

int fjsp_hostsim_step(void *v, int T, const int32_t *actions, const uint32_t *rnd, int reward_policy,
                      double completion, double tardiness, double energy, int autoreset,
                      double *state, float *state32, double *reward, int32_t *done, int32_t *rec)
{
    HostVec *h = (HostVec *)v;
    FjStepArgs A;
    A.T = T; A.actions = actions; A.rnd = rnd; A.reward_policy = reward_policy; A.autoreset = autoreset;
    A.completion = completion; A.tardiness = tardiness; A.energy = energy;
    A.state = state; A.state32 = state32; A.reward = reward; A.done = done; A.rec = rec;
    A.park_count = nullptr; A.park_env = nullptr;
    DISPATCH(run_step, h, A);
    return 0;
}

int fjsp_hostsim_info(void *v, int64_t *out)
{
    HostVec *h = (HostVec *)v;
    for (int e = 0; e < h->P.B; ++e) {
        const int32_t *s = (const int32_t *)(h->env.data() + (size_t)e * h->tb.eo.stride + h->tb.eo.scal);
        int64_t *o = out + (size_t)e * 12;
        long long dp = *(const long long *)(s + FJ_S_DELAY_PROC), du = *(const long long *)(s + FJ_S_DELAY_UNPROC);
        o[0] = s[FJ_S_TIME]; o[1] = s[FJ_S_STEPS]; o[2] = s[FJ_S_COMPLETION]; o[3] = dp + du;
        o[4] = *(const long long *)(s + FJ_S_ENERGY); o[5] = s[FJ_S_LPSOLVES]; o[6] = s[FJ_S_LPITERS];
        o[7] = s[FJ_S_ERROR]; o[8] = s[FJ_S_DONE]; o[9] = s[FJ_S_NEXTORDER]; o[10] = s[FJ_S_EPISODES]; o[11] = du;
    }
    return 0;
}

int fjsp_hostsim_query(void *v, int64_t *out8)
{
    HostVec *h = (HostVec *)v;
    out8[0] = h->P.B; out8[1] = 2 * h->P.nobs; out8[2] = h->tb.eo.stride; out8[3] = (int64_t)h->tb.io.stride * 4;
    out8[4] = 0; out8[5] = 0; out8[6] = (int64_t)fj_lp_scratch_bytes(h->tb.d); out8[7] = 0;
    return 0;
}

/* CPython set-order emulation of the table builder, exported for tests/test_pyemu.py */
int fjsp_hostsim_pyset_order(const int *seq, int n, int *out) { return fj_pyset_order(seq, n, out); }
int fjsp_hostsim_selectable(unsigned idle, const int *other_ord, int nother, int *out)
{
    unsigned omask = 0;
    for (int i = 0; i < nother; ++i) omask |= 1u << other_ord[i];
    unsigned packed = 0;   // other_ord: iteration order of set(other); packed when it has at most 4 members
    if (nother <= 4) for (int i = 0; i < nother; ++i) packed |= (unsigned)other_ord[i] << (8 * i);
    FjCand cand = fj_selectable(idle, omask, nother, packed);
    unsigned mk = cand.mask;
    for (int i = 0; i < cand.n; ++i) {
        if (cand.n >= 5) { out[i] = __builtin_ctz(mk); mk &= mk - 1; }
        else out[i] = (int)((cand.packed >> (8 * i)) & 0xffu);
    }
    return cand.n;
}
}
