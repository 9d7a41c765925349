// CUDA kernels and the C ABI (include/fjsp_b200.h) of the B200 vector FJSP environment.
// One warp owns one environment copy; a persistent grid (a multiple of the SM count)
// walks the batch.  Build: nvcc -gencode arch=compute_100a,code=sm_100a (build.py).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#include <string>
#include <vector>
#include "../../include/fjsp_b200.h"
#include "fjsp_host.h"
#include "fjsp_core.cuh"

#define FJ_BLOCK 128
#define FJ_PROG_CHUNKS 64    // progress chunks of one launch (fjsp_vec_step_host)
#define FJ_WARPS_PER_BLOCK (FJ_BLOCK / 32)

static thread_local std::string g_err;
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            return -10;                                                                            \
        }                                                                                          \
    } while (0)

// main step kernel: every env; parks an env on the first fluid LP it needs
#ifndef FJ_STEP_THREADS
#define FJ_STEP_THREADS 1024     // up to 32 warps in lockstep phases, one CTA per SM at 64 registers: one instruction
                                 // stream per SM and every thread of the SM on a CTA-served LP (71.0 M vs 69.1 M at 2 x 512)
#endif
#ifndef FJ_STEP_MIN_BLOCKS
#define FJ_STEP_MIN_BLOCKS (1024 / FJ_STEP_THREADS)
#endif
template <int VARIANT, int SUM_MODE>
__global__ void __launch_bounds__(FJ_STEP_THREADS, FJ_STEP_MIN_BLOCKS) fjsp_step_kernel(const __grid_constant__ FjParams P, const __grid_constant__ FjStepArgs A)
{
    fj_params_to_shared(P);
    extern __shared__ __align__(16) unsigned char stage_smem[];
    __shared__ FjLpBoard board;
    const int warp = threadIdx.x >> 5;
    const int nenv = P.env_warps;                 // env warps of an env CTA (= warp slots of a virtual CTA)
    const int nsrv = P.cta_lp == 1 ? P.srv_ctas : 0;
    if ((int)blockIdx.x < nsrv) {
        // LP-server CTA (the first CTAs of the grid: scheduled first, so an env CTA never waits for a
        // server that is not resident): groups of warps serve the LP queue until every env CTA is done
        const int gw = P.srv_group_warps, gid = warp / gw;
        if (gid >= P.srv_groups) return;
        fj_lp_server_loop(gid, gw, stage_smem + (size_t)gid * P.srv_group_smem, (int)gridDim.x - nsrv, (int)blockIdx.x * P.srv_groups + gid);
        return;
    }
    const int ecta = (int)blockIdx.x - nsrv, nectas = (int)gridDim.x - nsrv;
    for (int k = threadIdx.x; k < 32; k += blockDim.x)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fj_smem_addr(&board.mbar[k])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned char *hotbuf = P.stage && warp < nenv ? stage_smem + (size_t)warp * P.stage_stride : nullptr;
    FjCtaCtx K;
    K.warp = warp; K.nwarps = nenv; K.cta_lp = nsrv > 0;
    K.gslot = ecta * nenv + warp;
    K.slab = (P.lp_own && K.gslot < P.lp_own_slots) ? P.lp_own + (size_t)K.gslot * P.lp_stride : nullptr;   // overflow path: solve in line
    K.xbuf = P.cta_x + (size_t)K.gslot * P.d.NPx;
    K.board = &board;
    K.group.red = nullptr; K.group.flip = 0; K.group.base = 0; K.group.nthr = 32; K.group.bar = 0;
    const int total = nectas * nenv;
    // P.order maps warp slots to envs: slot = virtual CTA * nenv + warp, -1 = empty.  Env CTA b plays
    // the virtual CTAs b, b + env CTAs, ...
    if (P.cta_lp == 2) {
        // free-running warps: no CTA coupling at all; a warp that reaches an order arrival solves
        // the fluid LP itself on its own scratch slab (diagnostic mode, FJSP_FREE_RUN)
        unsigned char *lp = P.lp + (size_t)(ecta * nenv + warp) * P.lp_stride;
        for (int slot = ecta * nenv + warp; slot < P.n_slots; slot += total) {
            const int env = P.order[slot];
            if (env >= 0) fj_env_rollout<VARIANT, SUM_MODE, 0>(P, A, env, lp, nullptr);
        }
        return;
    }
    unsigned uses = 0;
    for (int base = ecta * nenv; base < P.n_slots; base += total) {
        const int raw = P.order[base + warp];
        const int env = raw < 0 ? -1 : raw & (FJ_SLOT_DETACHED - 1);
        const int detached = raw >= 0 && (raw & FJ_SLOT_DETACHED);
        int next_env = base + total < P.n_slots ? P.order[base + total + warp] : -1;
        if (next_env >= 0) next_env &= FJ_SLOT_DETACHED - 1;
        // the round's lockstep group: env warps that hold an env which is not expected to meet an LP or
        // a reset (those free-run: their long steps would stall every mate at the slot barrier)
        // (P.lock_groups = 2 / 4: the warps of each pair of SM sub-partitions / of each sub-partition form a group
        // of their own -- a shorter wait for the group's slowest warp, two / four instruction streams per SM)
        const int lg = P.lock_groups, mygrp = lg == 4 ? (warp & 3) : lg == 2 ? ((warp >> 1) & 1) : 0;
        int lock_threads = 0;
        for (int gi = 0; gi < lg; ++gi) {
            const int n = fj_env_count(env >= 0 && !detached && mygrp == gi, nenv * 32);
            if (mygrp == gi) lock_threads = n;
        }
        if (env < 0) continue;
        const int lock_bar = mygrp == 0 ? FJ_BAR_ENV : mygrp == 1 ? 2 : 10 + mygrp;
        fj_cta_rollout<VARIANT, SUM_MODE>(P, A, K, env, 1, hotbuf, next_env, uses & 1u, detached ? 0 : lock_threads, lock_bar);
        if (hotbuf) ++uses;
    }
    // every env warp (the free-running ones included) has finished its last round: one more CTA the servers need not wait for
    fj_env_count(0, nenv * 32);
    if (threadIdx.x == 0 && nsrv > 0) { __threadfence(); atomicAdd((unsigned *)P.pend_count + FJ_ROUNDS + 1, 1u); }
    // ... and from here on this CTA serves too: the envs that are still playing are the ones that wait for LPs, and
    // the launch ends with the last of them.  (Every warp's staging slab has been stored and waited for; the group
    // barriers and the groups' shared memory are laid out as in a server CTA, slabs after the servers'.)
    if (nsrv > 0 && P.srv_join) {
        const int gw = P.srv_group_warps, gid = warp / gw;
        if (gid < P.srv_groups)
            fj_lp_server_loop(gid, gw, stage_smem + (size_t)gid * P.srv_group_smem, nectas, (nsrv + ecta) * P.srv_groups + gid);
    }
}

// LP-aware packing, run before every step launch.  The warps of a CTA serve each other's fluid
// LPs in lockstep, so a CTA's launch time is  (lockstep steps of its envs) + (its LPs), and the
// launch ends with the slowest CTA.  Measured (profiles/README.md r01_v5): an episode meets all
// its order-arrival LPs within its first few dozen steps, an LP costs 0.15 M (small instances) to
// 0.9 M cycles (R ~ 110 rows) against 2.2 M cycles for 32 lockstep steps of a full CTA, and a
// warp alone (no CTA-mates to wait for) needs only 1.45 M for its 32 steps.  So the envs that
// are about to meet an LP get a virtual CTA of their own whose remaining warp slots are filled
// the less the heavier the instance's LP is; every other env fills the remaining virtual CTAs
// in the static order (similar walk lengths share a CTA).
//   slot = virtual CTA * wpb + warp;  the step kernel's CTA b plays virtual CTAs b, b + grid, ...
#define FJ_PACK_THREADS 1024
__device__ __forceinline__ int fj_lp_class(const FjParams &P, int env, int T, int r_heavy, int r_medium)
{
    // 0: no LP expected in the next T steps; 1 / 2 / 3: expected, heavy / medium / light LP
    const int32_t *s = (const int32_t *)(P.env + (size_t)env * P.eo.stride + P.eo.scal);
    const int32_t *I = P.inst + (size_t)P.env_inst[env] * P.io.stride;
    const int32_t *h = I + P.io.hdr;
    // likely = an order is still to come (they all arrive early in an episode), or the episode ends
    // and restarts within the launch (the restart needs ~10 steps to reach the first arrival)
    const int likely = s[FJ_S_DONE] || s[FJ_S_NEXTORDER] < h[3] || h[8] - s[FJ_S_STEPS] <= T - 8;
    if (!likely) return 0;
    const int rub = h[0] + 2 * h[2] - h[1];          // upper bound of the LP's rows: M + 2 KT - K
    return rub >= r_heavy ? 1 : rub >= r_medium ? 2 : 3;
}
__global__ void fjsp_flag_kernel(FjParams P, const int32_t *static_order, unsigned char *flags, int T, int r_heavy, int r_medium)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.B) flags[i] = (unsigned char)fj_lp_class(P, static_order[i], T, r_heavy, r_medium);
}
struct FjPackPlan { int e1, e2, e3, f1, f2, f3; };   // virtual CTAs per class, free slots per CTA of the class
__device__ __forceinline__ int fj_pack_free(const FjPackPlan &p, int wpb, int m)   // free slots of virtual CTAs < m
{
    int f = 0, x = m;
    int n = x < p.e1 ? x : p.e1; f += n * p.f1; x -= n;
    n = x < p.e2 ? x : p.e2; f += n * p.f2; x -= n;
    n = x < p.e3 ? x : p.e3; f += n * p.f3; x -= n;
    return f + x * wpb;
}
__global__ void __launch_bounds__(FJ_PACK_THREADS) fjsp_pack_kernel(FjParams P, const int32_t *static_order, const unsigned char *flags,
                                                                    int32_t *order_out, int wpb, int cap1, int cap2, int cap3, int detach)
{
    __shared__ int wsum[3][32];
    __shared__ FjPackPlan s_plan;
    const int B = P.B, tid = threadIdx.x, nt = blockDim.x;
    const int V = P.n_slots / wpb;
    for (int i = tid; i < P.n_slots; i += nt) order_out[i] = -1;
    const int per = (B + nt - 1) / nt, lo = tid * per < B ? tid * per : B, hi = lo + per < B ? lo + per : B;
    // per-class counts of this thread's chunk, scanned over the block (three 32-bit scans: any batch size)
    int cnt[3] = {0, 0, 0};
    for (int i = lo; i < hi; ++i) { const int c = flags[i]; if (c) ++cnt[c - 1]; }
    int inc[3];
    for (int k = 0; k < 3; ++k) {
        int v = cnt[k];
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, v, d); if ((tid & 31) >= d) v += o; }
        inc[k] = v;
        if ((tid & 31) == 31) wsum[k][tid >> 5] = v;
    }
    __syncthreads();
    if (tid < 32) {
        int tot[3];
        for (int k = 0; k < 3; ++k) {
            const int v = tid < (nt >> 5) ? wsum[k][tid] : 0;
            int w = v;
            for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, w, d); if (tid >= d) w += o; }
            wsum[k][tid] = w - v;
            tot[k] = w;
        }
        if (tid == 31) {
            const int n1 = tot[0], n2 = tot[1], n3 = tot[2];
            FjPackPlan p;
            p.e1 = n1 < V ? n1 : V; p.e2 = n2 < V - p.e1 ? n2 : V - p.e1; p.e3 = n3 < V - p.e1 - p.e2 ? n3 : V - p.e1 - p.e2;
            int c1 = cap1 < wpb ? cap1 : wpb, c2 = cap2 < wpb ? cap2 : wpb, c3 = cap3 < wpb ? cap3 : wpb;
            if (c1 < 1) c1 = 1; if (c2 < c1) c2 = c1; if (c3 < c2) c3 = c2;
            // every env needs a slot: widen the light, then the medium, then the heavy CTAs until they do
            for (;;) {
                const long long capacity = (long long)p.e1 * c1 + (long long)p.e2 * c2 + (long long)p.e3 * c3 + (long long)(V - p.e1 - p.e2 - p.e3) * wpb;
                if (capacity >= B) break;
                if (c3 < wpb) ++c3; else if (c2 < wpb) ++c2; else ++c1;
            }
            p.f1 = c1 - 1; p.f2 = c2 - 1; p.f3 = c3 - 1;
            s_plan = p;
        }
    }
    __syncthreads();
    const FjPackPlan p = s_plan;
    // per-class likely envs before this thread's chunk
    int r1 = wsum[0][tid >> 5] + inc[0] - cnt[0], r2 = wsum[1][tid >> 5] + inc[1] - cnt[1], r3 = wsum[2][tid >> 5] + inc[2] - cnt[2];
    for (int i = lo; i < hi; ++i) {
        const int env = static_order[i], c = flags[i];
        int slot = -1;
        if (c == 1 && r1 < p.e1) slot = r1 * wpb;
        else if (c == 2 && r2 < p.e2) slot = (p.e1 + r2) * wpb;
        else if (c == 3 && r3 < p.e3) slot = (p.e1 + p.e2 + r3) * wpb;
        if (slot < 0) {
            // rank among the envs that do not own a virtual CTA, then the CTA whose free slots hold it
            const int rB = i - ((r1 < p.e1 ? r1 : p.e1) + (r2 < p.e2 ? r2 : p.e2) + (r3 < p.e3 ? r3 : p.e3));
            int a = 0, b = V - 1;                    // largest v with free(v) <= rB
            while (a < b) {
                const int m = (a + b + 1) >> 1;
                if (fj_pack_free(p, wpb, m) <= rB) a = m; else b = m - 1;
            }
            slot = a * wpb + (a < p.e1 + p.e2 + p.e3 ? 1 : 0) + (rB - fj_pack_free(p, wpb, a));
        }
        order_out[slot] = c && detach ? (env | FJ_SLOT_DETACHED) : env;
        r1 += c == 1; r2 += c == 2; r3 += c == 3;
    }
}

// resume kernel: parked envs only; picks up the LP solution, finishes the launch
template <int VARIANT, int SUM_MODE, int SUSPEND>
__global__ void __launch_bounds__(FJ_BLOCK) fjsp_resume_kernel(const __grid_constant__ FjParams P, FjStepArgs A, const int *count_in, const int *list_in)
{
    fj_params_to_shared(P);
    const int gw = blockIdx.x * FJ_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int total = gridDim.x * FJ_WARPS_PER_BLOCK;
    unsigned char *lp = P.lp + (size_t)gw * P.lp_stride;
    const int n = *count_in;
    for (int i = gw; i < n; i += total) fj_env_rollout<VARIANT, SUM_MODE, SUSPEND>(P, A, list_in[i], lp, nullptr);
}

// LP kernel: one CTA per parked LP, basis inverse in shared memory when it fits
#ifndef FJ_LP_THREADS
#ifndef FJ_LP_THREADS
#define FJ_LP_THREADS 256
#endif
#endif
static_assert(FJ_STEP_THREADS / 32 <= FJ_CTX_WARPS && FJ_LP_THREADS / 32 <= FJ_CTX_WARPS && FJ_BLOCK / 32 <= FJ_CTX_WARPS,
              "fj_sC holds one context per warp of a CTA");
template <int SMEM_BINV>
__global__ void __launch_bounds__(FJ_LP_THREADS) fjsp_lp_kernel(const __grid_constant__ FjParams P, const int *count_in, const int *list_in)
{
    fj_params_to_shared(P);
    extern __shared__ __align__(16) unsigned char smem[];
    int n = *count_in;
    if (n > P.lp_slots) n = P.lp_slots;
    const size_t binv_bytes = (size_t)P.d.Rx * P.d.Rx * 8;
    unsigned char *binv = SMEM_BINV ? smem : P.lp + (size_t)blockIdx.x * P.lp_stride;
    unsigned char *small_ = SMEM_BINV ? smem + binv_bytes : smem;
    unsigned char *red = smem + ((SMEM_BINV ? binv_bytes : 0) + fj_lp_small_bytes(P.d) + 15) / 16 * 16;   // 16-byte aligned
    FjCtaGroup g;
    g.red = (int4 *)red; g.flip = 0; g.whole_cta();
    // the fast path carves the same dynamic shared memory its own way (it is one or the other per LP)
    const int fast_bytes = (int)(red - smem);
    for (int i = blockIdx.x; i < n; i += gridDim.x)
        fj_lp_service(P, g, list_in, i, binv, small_, P.lp + (size_t)blockIdx.x * P.lp_stride, smem, fast_bytes);
}

// after the first reset(): copy each instance's order-0 LP solution from its representative env
__global__ void fjsp_plan_kernel(FjParams P, const int *rep_env, int n_inst, double *plan_x, int *plan_meta, int *plan_ok)
{
    const int ii = blockIdx.x;
    if (ii >= n_inst) return;
    const int env = rep_env[ii];
    if (env < 0 || env >= P.lp_slots) { if (threadIdx.x == 0) plan_ok[ii] = 0; return; }
    const int np = P.d.NPx;
    for (int j = threadIdx.x; j < np; j += blockDim.x) plan_x[(size_t)ii * np + j] = P.lp_x[(size_t)env * np + j];
    if (threadIdx.x == 0) { plan_meta[2 * ii] = P.lp_meta[2 * env]; plan_meta[2 * ii + 1] = P.lp_meta[2 * env + 1]; plan_ok[ii] = 1; }
}

__global__ void __launch_bounds__(FJ_BLOCK) fjsp_reset_begin_kernel(const __grid_constant__ FjParams P, int fresh)
{
    fj_params_to_shared(P);
    const int gw = blockIdx.x * FJ_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int total = gridDim.x * FJ_WARPS_PER_BLOCK;
    for (int env = gw; env < P.B; env += total) fj_env_reset_begin(P, env, fresh);
    if (blockIdx.x == 0 && threadIdx.x == 0) *P.pend_count = P.B;
}

template <int VARIANT, int SUM_MODE>
__global__ void __launch_bounds__(FJ_BLOCK) fjsp_reset_finish_kernel(const __grid_constant__ FjParams P, double *state64, float *state32)
{
    fj_params_to_shared(P);
    const int gw = blockIdx.x * FJ_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int total = gridDim.x * FJ_WARPS_PER_BLOCK;
    unsigned char *lp = P.lp + (size_t)gw * P.lp_stride;
    for (int env = gw; env < P.B; env += total) fj_env_reset_finish<VARIANT, SUM_MODE>(P, env, lp, state64, state32);
}

struct fjsp_vec {
    FjTables tb;
    FjParams P;
    int variant, sum_mode, B, device, grid, step_grid, step_threads, resume_grid, lp_grid, lp_smem_binv, nstate;
    size_t lp_smem_bytes, stage_bytes, step_smem_bytes;
    int *d_pend_count, *d_pend_env, *d_lp_meta, *d_rep_env, *d_plan_meta, *d_plan_ok;
    double *d_lp_x, *d_plan_x;
    int n_resets;
    int n_inst, plan_ready, plan_all;   // plan_all: every instance's order-0 LP solution gets cached by the first reset
    int32_t *d_inst, *d_env_inst, *d_order, *d_order_dyn, *d_order_static, *last_order;
    int n_slots, pack_cap[3], pack_rows[2], multi_round, env_warps, srv_ctas, env_ctas, detach;
    double *d_cta_x;
    unsigned int *d_lpq; unsigned long long *d_lpq_ring; int *d_lp_req, *d_lp_resp; unsigned char *d_lp_own;
    unsigned char *d_flags;
    int pack;
    unsigned char *d_env, *d_lp;
    long long launches;
    // staging for the host-buffer entry points
    cudaStream_t stream, copy_stream;
    cudaEvent_t chunk_done, dev_done;   // dev_done: last work queued through the device entry points (caller's stream)
    int dev_pending;
    int stage_T, stage_out_T;
    // progress reporting of a launch (fjsp_vec_step_host: chunks of steps copied out while the kernel runs)
    unsigned *d_prog, *h_prog, *h_prog_dev; unsigned prog_seq; int prog_shift;   // prog_shift < 0: off for the next launch
    // pipelined host calls (fjsp_vec_step_host_begin / _wait): two input staging slots
    int32_t *d_actions2[2]; uint32_t *d_rnd2[2]; unsigned char *d_out2[2]; int pipe_T; cudaEvent_t pipe_in[2], pipe_k[2], pipe_done[2];
    long long pipe_begun, pipe_waited; cudaStream_t copy_out_stream;
    int32_t *d_actions, *d_done, *d_rec;
    uint32_t *d_rnd;
    double *d_state64, *d_reward;
    float *d_state32;
    long long *d_trace;          // FJ_TRACE builds only
};

static void launch_lp(fjsp_vec *v, cudaStream_t st, int round)
{
    const int *cnt = v->d_pend_count + round, *lst = v->d_pend_env + (size_t)(round & 1) * v->B;
    if (v->lp_smem_binv) fjsp_lp_kernel<1><<<v->lp_grid, FJ_LP_THREADS, v->lp_smem_bytes, st>>>(v->P, cnt, lst);
    else fjsp_lp_kernel<0><<<v->lp_grid, FJ_LP_THREADS, v->lp_smem_bytes, st>>>(v->P, cnt, lst);
}

template <typename F> static int dispatch(fjsp_vec *v, F f)
{
    switch (v->variant * 2 + (v->sum_mode ? 1 : 0)) {
    case 0: return f(std::integral_constant<int, FJSP_SO_DFJSP>(), std::integral_constant<int, 0>());
    case 1: return f(std::integral_constant<int, FJSP_SO_DFJSP>(), std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, FJSP_MO_DFJSP>(), std::integral_constant<int, 0>());
    case 3: return f(std::integral_constant<int, FJSP_MO_DFJSP>(), std::integral_constant<int, 1>());
    case 4: return f(std::integral_constant<int, FJSP_MO_BREAKDOWN>(), std::integral_constant<int, 0>());
    case 5: return f(std::integral_constant<int, FJSP_MO_BREAKDOWN>(), std::integral_constant<int, 1>());
    case 6: return f(std::integral_constant<int, FJSP_SO_FJSSP>(), std::integral_constant<int, 0>());
    case 7: return f(std::integral_constant<int, FJSP_SO_FJSSP>(), std::integral_constant<int, 1>());
    }
    g_err = "unsupported variant";
    return -2;
}

// One handle can be driven through the device entry points (caller's stream) and the host-buffer entry
// points (the handle's own stream): work queued on the caller's stream is recorded here and waited
// for by the host calls; the host calls end synchronised, so the reverse order needs nothing.
static void note_device_work(fjsp_vec *v, cudaStream_t st)
{
    if (st == v->stream) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return; }
    if (cudaEventRecord(v->dev_done, st) == cudaSuccess) v->dev_pending = 1; else cudaGetLastError();
}
static void wait_device_work(fjsp_vec *v)
{
    if (v->dev_pending) { cudaStreamWaitEvent(v->stream, v->dev_done, 0); v->dev_pending = 0; }
}

extern "C" {

const char *fjsp_last_error(void) { return g_err.c_str(); }
int fjsp_abi_version(void) { return 2; }

int fjsp_vec_destroy(fjsp_vec *v);

int fjsp_vec_create(const int32_t *blobs, const int64_t *blob_offsets, int n_instances,
                    const int32_t *env_instance, int n_envs, int variant, int sum_mode, int device,
                    fjsp_vec **out)
{
    if (!blobs || !blob_offsets || !env_instance || !out || n_instances < 1 || n_envs < 1) {
        g_err = "fjsp_vec_create: null or empty argument"; return -1;
    }
    if (variant < 0 || variant > 3) {
        g_err = "fjsp_vec_create: variant must be 0 (SO_DFJSP), 1 (MO_DFJSP), 2 (MO_DFJSP_breakdown) or 3 (SO_FJSSP)"; return -2;
    }
    for (int e = 0; e < n_envs; ++e)
        if (env_instance[e] < 0 || env_instance[e] >= n_instances) { g_err = "fjsp_vec_create: env_instance out of range"; return -1; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_err = "fjsp_vec_create: no CUDA device (this library has no CPU path)"; return -3;
    }
    if (n_envs >= FJ_SLOT_DETACHED) { g_err = "fjsp_vec_create: more than 2^30 - 1 environment copies"; return -1; }
    CK(cudaSetDevice(device));
    fjsp_vec *v = new fjsp_vec();
    if (!fj_build_tables(blobs, blob_offsets, n_instances, v->tb, g_err, variant)) { delete v; return -1; }
    // from here on a failure frees what has been allocated so far
#undef CK
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            cudaGetLastError();                                                                    \
            fjsp_vec_destroy(v);                                                                   \
            return -10;                                                                            \
        }                                                                                          \
    } while (0)
    v->variant = variant; v->sum_mode = sum_mode ? 1 : 0; v->B = n_envs; v->device = device; v->launches = 0;
    v->nstate = (variant == FJSP_SO_DFJSP || variant == FJSP_SO_FJSSP) ? 20 : 30;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    // persistent grid: a multiple of the SM count, 8 CTAs (32 warps) per SM at most
    int want = (n_envs + FJ_WARPS_PER_BLOCK - 1) / FJ_WARPS_PER_BLOCK;
    int cap = prop.multiProcessorCount * 8;
    if (getenv("FJSP_GRID_PER_SM")) cap = prop.multiProcessorCount * atoi(getenv("FJSP_GRID_PER_SM"));
    v->grid = want < cap ? want : cap;
    int lp_mode = getenv("FJSP_FREE_RUN") ? 2 : getenv("FJSP_NO_CTA_LP") ? 0 : 1;
    int srv_groups = 0, srv_group_warps = 0;
    {
        // Shape of the main kernel at run time.  One CTA per SM: `srv_ctas` LP-server CTAs and env CTAs of
        // `env_warps` warps (one environment copy each, lockstep slots); a batch larger than
        // env CTAs x env_warps is played in rounds.
        const int nsm = prop.multiProcessorCount, wmax = FJ_STEP_THREADS / 32;
        // LP load per env step (group-cycles): an episode meets S - 1 order-arrival LPs in `ops` steps; model of an
        // LP: ~0.7 R iterations of ~(2500 + R^2 / 16) cycles
        double load = 0.0; int any_arrival = 0;
        for (int i = 0; i < n_instances; ++i) {
            const int32_t *b = blobs + blob_offsets[i];
            const int M = b[2], K = b[3], KT = b[4], S = b[5];
            const int32_t *ntask = b + 16, *count = b + 16 + K + 3 * KT + 3 * KT * M + M + 2 * S;
            double ops = 0.0;
            for (int r = 0; r < K; ++r) { double c = 0.0; for (int s2 = 0; s2 < S; ++s2) c += count[s2 * K + r]; ops += c * ntask[r]; }
            const double R = M + KT + 0.5 * (KT - K);
            if (S > 1) any_arrival = 1;
            load += (S - 1) / (ops > 1.0 ? ops : 1.0) * 0.7 * R * (2500.0 + R * R / 16.0);
        }
        load /= n_instances;
        if (lp_mode == 1 && !any_arrival) lp_mode = 0;   // no order ever arrives after reset(): nothing to serve
        int srv = 0;
        if (lp_mode == 1) {
            // Measured (profiles/README.md r02): an LP costs ~2.5 x that model on a server group; an env CTA retires
            // one env step per ~2000 cycles; the queue stays short while the groups are busy < 55 % of the time;
            // two groups per server CTA.  10 machines / 3 orders (Instance_generate.py profile): 12-13 of 148 SMs.
            const double groups_per_env_cta = load * 2.5 / 2000.0 / 0.55;
            srv = (int)(nsm * groups_per_env_cta / (2.0 + groups_per_env_cta) + 0.5);
            if (srv < 2) srv = 2;
            if (srv > nsm / 3) srv = nsm / 3;
            if (n_envs <= 32) srv = 1;
            if (getenv("FJSP_LP_SERVERS")) srv = atoi(getenv("FJSP_LP_SERVERS"));
            if (srv < 1) srv = 1;
            if (srv > nsm - 1) srv = nsm - 1;
        }
        const int gmax = nsm - srv;
        const long long want_slots = (long long)n_envs + (getenv("FJSP_SPARE") ? n_envs * (long long)atoi(getenv("FJSP_SPARE")) / 100 : 0);
        v->multi_round = (long long)n_envs > (long long)gmax * wmax;
        int wpb = (int)((want_slots + gmax - 1) / gmax);
        if (wpb > wmax) wpb = wmax;
        if (wpb < 4) wpb = 4;
        if (getenv("FJSP_STEP_WARPS")) wpb = atoi(getenv("FJSP_STEP_WARPS"));
        if (wpb > wmax) wpb = wmax;
        if (wpb < 1) wpb = 1;
        long long ctas = (want_slots + wpb - 1) / wpb;              // virtual CTAs
        v->env_ctas = (int)(ctas < gmax ? ctas : gmax);
        // SMs the env CTAs leave over serve LPs too (4096 copies: 133 env CTAs of 31 warps whether 13 or 15 SMs serve)
        if (lp_mode == 1 && !getenv("FJSP_LP_SERVERS") && n_envs > 32 && srv < nsm - v->env_ctas) {
            const int model = srv;
            srv = nsm - v->env_ctas;
            if (srv > 2 * model) srv = 2 * model;                   // (a small batch leaves most SMs over: no use for them)
            if (srv > nsm / 3) srv = nsm / 3;
        }
        v->env_warps = wpb; v->srv_ctas = srv;
        v->step_threads = wpb * 32;
        v->step_grid = v->env_ctas + srv;
        const long long rounds = (ctas + v->env_ctas - 1) / v->env_ctas;
        v->n_slots = (int)(rounds * v->env_ctas * wpb);
        if (srv) {
            // server groups: 4 per CTA when the CTA has the warps for it (a driver warp + helpers each)
            // server groups: two per CTA (up to 16 warps and ~100 KB of shared memory each: B^-1 of a 110-row LP
            // fits; four of the group's warps drive the iteration, the others apply the rank-1 updates)
            srv_groups = wpb >= 12 ? 2 : 1;
            if (getenv("FJSP_LP_GROUPS")) srv_groups = atoi(getenv("FJSP_LP_GROUPS"));
            if (srv_groups < 1) srv_groups = 1;
            if (srv_groups > 6) srv_groups = 6;
            if (srv_groups > wpb) srv_groups = wpb;
            srv_group_warps = wpb / srv_groups;
        }
    }
    const unsigned long long lp_stride = fj_lp_scratch_bytes(v->tb.d);
    // the resume kernel (in-line LP fallback) and the LP kernel (global Binv fallback) share the slabs
    int rcap = prop.multiProcessorCount * 4;
    v->resume_grid = want < rcap ? want : rcap;
    v->lp_grid = prop.multiProcessorCount * 2;
    const size_t small_b = fj_lp_small_bytes_host(v->tb.d) + 16 + 64 * 16;   // + alignment slack + reduction scratch
    const size_t binv_b = (size_t)v->tb.d.Rx * v->tb.d.Rx * 8;
    v->lp_smem_binv = (binv_b + small_b <= 200 * 1024) ? 1 : 0;
    v->lp_smem_bytes = v->lp_smem_binv ? binv_b + small_b : small_b;
    if (v->lp_smem_bytes > 200 * 1024) { g_err = "fjsp_vec_create: instance too large for the LP kernel's shared memory"; delete v; return -5; }
    if (v->lp_smem_binv && binv_b + small_b > 100 * 1024) v->lp_grid = prop.multiProcessorCount;
    int slabs = v->resume_grid * FJ_WARPS_PER_BLOCK;
    if (v->lp_grid > slabs) slabs = v->lp_grid;
    if ((v->srv_ctas + v->env_ctas) * srv_groups > slabs) slabs = (v->srv_ctas + v->env_ctas) * srv_groups;   // one slab per server group of the main kernel (env CTAs serve once their envs are done)
    if (lp_mode == 2 && v->env_ctas * v->env_warps > slabs) slabs = v->env_ctas * v->env_warps;
    const size_t lp_bytes = (size_t)lp_stride * slabs;
    const size_t env_bytes = (size_t)n_envs * v->tb.eo.stride;
    CK(cudaMalloc(&v->d_inst, v->tb.inst.size() * 4));
    CK(cudaMalloc(&v->d_env_inst, (size_t)n_envs * 4));
    CK(cudaMalloc(&v->d_order, (size_t)n_envs * 4));
    CK(cudaMalloc(&v->d_order_dyn, (size_t)v->n_slots * 4));
    CK(cudaMalloc(&v->d_flags, (size_t)n_envs));
    {
        // static walk length of an env: jobs of its largest kind x warp rounds over its operation types
        std::vector<long long> key(n_instances);
        for (int i = 0; i < n_instances; ++i) {
            const int32_t *b = blobs + blob_offsets[i];
            key[i] = (long long)b[10] * ((b[4] + 31) / 32);
        }
        std::vector<int32_t> order(n_envs);
        for (int e = 0; e < n_envs; ++e) order[e] = e;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return key[env_instance[x]] > key[env_instance[y]]; });
        CK(cudaMemcpy(v->d_order, order.data(), (size_t)n_envs * 4, cudaMemcpyHostToDevice));
        order.resize(v->n_slots, -1);   // without packing: the static order, spare slots empty
        CK(cudaMemcpy(v->d_order_dyn, order.data(), (size_t)v->n_slots * 4, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&v->d_order_static, (size_t)v->n_slots * 4));
        CK(cudaMemcpy(v->d_order_static, order.data(), (size_t)v->n_slots * 4, cudaMemcpyHostToDevice));
        v->last_order = v->d_order_dyn;
    }
    CK(cudaMalloc(&v->d_env, env_bytes));
    CK(cudaMalloc(&v->d_lp, lp_bytes));
    {   // LP service of the main kernel: queue, one request / response record and one solution buffer per env warp
        const size_t gslots = (size_t)v->env_ctas * v->env_warps;
        const int req_stride = 2 + v->tb.d.KTx + v->tb.d.KTW + 1;
        CK(cudaMalloc(&v->d_cta_x, gslots * v->tb.d.NPx * 8 + 16));
        CK(cudaMalloc(&v->d_lpq, 16));
        CK(cudaMalloc(&v->d_lpq_ring, (size_t)FJ_LPQ_RING * 8));
        CK(cudaMalloc(&v->d_lp_req, gslots * req_stride * 4));
        CK(cudaMalloc(&v->d_lp_resp, gslots * 16));
        CK(cudaMemset(v->d_lpq, 0, 16));
        CK(cudaMemset(v->d_lpq_ring, 0, (size_t)FJ_LPQ_RING * 8));
        CK(cudaMemset(v->d_lp_req, 0, gslots * req_stride * 4));
        CK(cudaMemset(v->d_lp_resp, 0, gslots * 16));
        v->P.lpq = v->d_lpq; v->P.lpq_ring = v->d_lpq_ring; v->P.lp_req = v->d_lp_req; v->P.lp_resp = v->d_lp_resp;
        v->P.lp_req_stride = req_stride;
        v->P.srv_ctas = v->srv_ctas; v->P.srv_groups = srv_groups; v->P.srv_group_warps = srv_group_warps;
        // Overflow: when every copy meets its order arrivals at once (right after reset(): 2 x 4096 LPs in the first
        // launch against ~26 server groups = 70 ms) an env warp that finds the queue long solves its LP itself, on
        // a scratch slab of its own in HBM/L2 (the warp-level generic solver: ~1-3 ms, but every warp at once).
        v->P.lp_own = nullptr; v->P.lp_own_slots = 0; v->P.lp_overflow = 0;
        // FJSP_SRV_JOIN=1: env CTAs whose envs are done serve LPs until the launch ends.  Off by default: measured
        // neutral at the default server count (102.4 vs 101.8 M env-steps/s: the launch ends with the env CTAs'
        // lockstep steps, not with a queue of LPs), a gain only when the servers are under-provisioned (3 servers: 79 M)
        v->P.srv_join = getenv("FJSP_SRV_JOIN") ? atoi(getenv("FJSP_SRV_JOIN")) : 0;
        if (v->srv_ctas > 0) {
            size_t own = gslots;
            const size_t budget = (size_t)(getenv("FJSP_LP_OWN_MB") ? atoi(getenv("FJSP_LP_OWN_MB")) : 4096) << 20;
            if (own * lp_stride > budget) own = budget / lp_stride;
            if (own > 0) {
                CK(cudaMalloc(&v->d_lp_own, own * lp_stride));
                v->P.lp_own = v->d_lp_own; v->P.lp_own_slots = (int)own;
                v->P.lp_overflow = 16 * v->srv_ctas * srv_groups;
                if (getenv("FJSP_LP_OVERFLOW")) v->P.lp_overflow = atoi(getenv("FJSP_LP_OVERFLOW"));
            }
        }
    }
    CK(cudaMemcpy(v->d_inst, v->tb.inst.data(), v->tb.inst.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(v->d_env_inst, env_instance, (size_t)n_envs * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(v->d_env, 0, env_bytes));
    CK(cudaMemset(v->d_lp, 0, lp_bytes));
    // parked-LP list and solution slots (one per env unless that would exceed 2 GiB)
    size_t slots = (size_t)n_envs;
    const size_t per_slot = (size_t)v->tb.d.NPx * 8;
    if (slots * per_slot > ((size_t)2 << 30)) slots = ((size_t)2 << 30) / per_slot;
    if (getenv("FJSP_LP_SLOTS") && (size_t)atoi(getenv("FJSP_LP_SLOTS")) < slots) slots = (size_t)(atoi(getenv("FJSP_LP_SLOTS")) > 1 ? atoi(getenv("FJSP_LP_SLOTS")) : 1);   // test knob: force the no-slot fallback paths
    CK(cudaMalloc(&v->d_pend_count, 4 * (FJ_ROUNDS + 2)));
    CK(cudaMalloc(&v->d_pend_env, (size_t)n_envs * 4 * 2));
    CK(cudaMalloc(&v->d_lp_x, slots * per_slot));
    CK(cudaMalloc(&v->d_lp_meta, slots * 8));
    CK(cudaMemset(v->d_pend_count, 0, 4 * (FJ_ROUNDS + 2)));
    {   // per-instance cache of the order-0 LP solution (filled by the first reset)
        std::vector<int> rep(n_instances, -1);
        for (int e = n_envs - 1; e >= 0; --e) rep[env_instance[e]] = e;
        v->n_inst = n_instances; v->plan_ready = 0;
        v->plan_all = 1;
        for (int i2 = 0; i2 < n_instances; ++i2) if (rep[i2] < 0 || (size_t)rep[i2] >= slots) v->plan_all = 0;
        CK(cudaMalloc(&v->d_rep_env, (size_t)n_instances * 4));
        CK(cudaMalloc(&v->d_plan_x, (size_t)n_instances * per_slot));
        CK(cudaMalloc(&v->d_plan_meta, (size_t)n_instances * 8));
        CK(cudaMalloc(&v->d_plan_ok, (size_t)n_instances * 4));
        CK(cudaMemcpy(v->d_rep_env, rep.data(), (size_t)n_instances * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(v->d_plan_ok, 0, (size_t)n_instances * 4));
    }
    CK(cudaFuncSetAttribute(fjsp_lp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v->lp_smem_bytes));
    CK(cudaFuncSetAttribute(fjsp_lp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v->lp_smem_bytes));
    FjParams &P = v->P;
    P.d = v->tb.d; P.io = v->tb.io; P.eo = v->tb.eo;
    P.order = v->d_order_dyn;   // rewritten before every step launch by fjsp_pack_kernel
    P.n_slots = v->n_slots;
    P.inst = v->d_inst; P.env_inst = v->d_env_inst; P.env = v->d_env; P.lp = v->d_lp; P.lp_stride = lp_stride;
    P.B = n_envs; P.variant = variant; P.sum_mode = v->sum_mode; P.nobs = v->nstate / 2;
    P.cta_x = v->d_cta_x; P.env_warps = v->env_warps;
    P.pend_count = v->d_pend_count; P.pend_env = v->d_pend_env; P.lp_x = v->d_lp_x; P.lp_meta = v->d_lp_meta;
    P.lp_slots = (int)slots;
    P.plan_x = v->d_plan_x; P.plan_meta = v->d_plan_meta; P.plan_ok = nullptr;
    P.trace = nullptr; v->d_trace = nullptr;
#ifdef FJ_TRACE
    CK(cudaMalloc(&v->d_trace, (size_t)v->step_grid * FJ_TRACE_ROWS * 8 * 8));
    CK(cudaMemset(v->d_trace, 0, (size_t)v->step_grid * FJ_TRACE_ROWS * 8 * 8));
    P.trace = v->d_trace;
#endif
    // the main kernel stages the hot prefix of its env warps' records in shared memory for the whole
    // launch (TMA bulk copies) when the slabs fit next to the static shared memory
    P.stage_stride = v->tb.eo.hot;
    v->stage_bytes = (size_t)v->env_warps * P.stage_stride;
    P.stage = v->stage_bytes <= 200 * 1024 ? 1 : 0;
    if (getenv("FJSP_NO_STAGE")) P.stage = 0;
    if (!P.stage) v->stage_bytes = 0;
    P.cta_lp = lp_mode;
    P.lock_mask = 0;
    P.lock_groups = 1;
    if (getenv("FJSP_LOCK_GROUPS")) { const int k = atoi(getenv("FJSP_LOCK_GROUPS")); P.lock_groups = k >= 4 ? 4 : k >= 2 ? 2 : 1; }
    if (getenv("FJSP_LOCKSTEP_K")) { int k = atoi(getenv("FJSP_LOCKSTEP_K")); int m = 1; while (m * 2 <= k) m *= 2; P.lock_mask = m - 1; }
    v->pack = (P.cta_lp == 1 && !getenv("FJSP_NO_PACK")) ? 1 : 0;
    {   // The packing kernel deals the envs that will meet an LP in this launch one per virtual CTA first
        // (they free-run and wait for the LP servers; their CTA-mates keep stepping) and fills the rest in the static order.
        // FJSP_PACK="cap1,cap2,cap3,rows_heavy,rows_medium" limits the env warps of a virtual CTA that
        // holds an env with a heavy / medium / light LP ahead (needs spare slots, FJSP_SPARE).
        const int w = v->env_warps;
        v->pack_cap[0] = v->pack_cap[1] = v->pack_cap[2] = w;
        v->pack_rows[0] = 70; v->pack_rows[1] = 45;
        if (getenv("FJSP_PACK")) sscanf(getenv("FJSP_PACK"), "%d,%d,%d,%d,%d", &v->pack_cap[0], &v->pack_cap[1], &v->pack_cap[2], &v->pack_rows[0], &v->pack_rows[1]);
    }
    {   // shared memory of a server group: the solver's vectors, the column descriptors and positions, and B^-1
        // when it still fits (else B^-1 stays on the group's slab in HBM/L2).  Env and server CTAs are one
        // launch: the dynamic shared memory is the larger of the two needs (FJSP_SRV_SMEM_KB overrides the
        // servers' share).
        P.srv_group_smem = 0;
        v->step_smem_bytes = v->stage_bytes;
        if (P.cta_lp == 1 && v->srv_ctas > 0) {
            size_t rub = 1;
            for (int i = 0; i < n_instances; ++i) { const int32_t *b = blobs + blob_offsets[i]; const size_t r = (size_t)b[2] + 2 * b[4] - b[3]; if (r > rub) rub = r; }
            const size_t C = (size_t)v->tb.d.NPx + 1;
            const size_t state_b = ((4 * ((size_t)v->tb.d.Rx + 2)) * 8 + ((size_t)v->tb.d.Rx + 2) * 4 + 16 + (size_t)v->tb.d.KTx * 2 + 15) / 16 * 16;
            const size_t small_b = (C * 24 + ((C + rub + 1) & ~(size_t)1) * 4 + 15) / 16 * 16;
            const size_t want = state_b + small_b + ((rub + 1) * (rub | 1) * 8 + 15) / 16 * 16;
            // by default no more than the env CTAs stage (or 100 KB): a larger request moves the whole launch to a
            // bigger shared-memory carve-out, i.e. less L1 for the env CTAs (164 KB -> 228 KB: -4 %)
            const size_t dflt = v->stage_bytes > (size_t)100 * 1024 ? v->stage_bytes : (size_t)100 * 1024;
            size_t cap = (getenv("FJSP_SRV_SMEM_KB") ? (size_t)atoi(getenv("FJSP_SRV_SMEM_KB")) * 1024 : dflt) / P.srv_groups;
            size_t give = want <= cap ? want : cap;
            if (give < state_b + small_b) give = state_b + small_b;      // the vectors and descriptors must be on chip
            give = (give + 15) / 16 * 16;
            if (give * P.srv_groups > (size_t)200 * 1024) { g_err = "fjsp_vec_create: instance too large for the LP servers' shared memory"; fjsp_vec_destroy(v); return -5; }
            P.srv_group_smem = (int)give;
            if (give * P.srv_groups > v->step_smem_bytes) v->step_smem_bytes = give * P.srv_groups;
        }
    }
    v->detach = getenv("FJSP_NO_DETACH") ? 0 : 1;
    if (dispatch(v, [&](auto V, auto SM) {
            auto kern = fjsp_step_kernel<decltype(V)::value, decltype(SM)::value>;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v->step_smem_bytes) != cudaSuccess) {
                cudaGetLastError();
                g_err = "fjsp_vec_create: the device refuses the step kernel's dynamic shared memory";
                return -5;
            }
            if (getenv("FJSP_CARVEOUT"))   // percent of the unified L1 / shared memory given to shared memory
                cudaFuncSetAttribute(fjsp_step_kernel<decltype(V)::value, decltype(SM)::value>,
                                     cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("FJSP_CARVEOUT")));
            return 0;
        })) { fjsp_vec_destroy(v); return -5; }
    CK(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&v->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&v->chunk_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&v->dev_done, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&v->copy_out_stream, cudaStreamNonBlocking));
    v->prog_shift = -1;
    CK(cudaMalloc(&v->d_prog, 4 * FJ_PROG_CHUNKS));
    CK(cudaHostAlloc(&v->h_prog, 4 * FJ_PROG_CHUNKS, cudaHostAllocMapped));
    memset(v->h_prog, 0, 4 * FJ_PROG_CHUNKS);
    CK(cudaHostGetDevicePointer(&v->h_prog_dev, v->h_prog, 0));
    for (int k = 0; k < 2; ++k) { CK(cudaEventCreateWithFlags(&v->pipe_in[k], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&v->pipe_k[k], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&v->pipe_done[k], cudaEventDisableTiming)); }
    v->dev_pending = 0;
    v->stage_T = 0; v->stage_out_T = 0;
    v->d_actions = v->d_done = v->d_rec = nullptr; v->d_rnd = nullptr;
    v->d_state64 = v->d_reward = nullptr; v->d_state32 = nullptr;
    *out = v;
    return 0;
#undef CK
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            return -10;                                                                            \
        }                                                                                          \
    } while (0)
}

static void free_stage(fjsp_vec *v)
{
    cudaFree(v->d_actions); cudaFree(v->d_rnd); cudaFree(v->d_done); cudaFree(v->d_rec);
    cudaFree(v->d_state64); cudaFree(v->d_state32); cudaFree(v->d_reward);
    v->d_actions = v->d_done = v->d_rec = nullptr; v->d_rnd = nullptr;
    v->d_state64 = v->d_reward = nullptr; v->d_state32 = nullptr;
    v->stage_T = 0; v->stage_out_T = 0;
}

int fjsp_vec_destroy(fjsp_vec *v)
{
    if (!v) return 0;
    cudaSetDevice(v->device);
    free_stage(v);
    cudaFree(v->d_order); cudaFree(v->d_order_dyn); cudaFree(v->d_order_static); cudaFree(v->d_flags);
    cudaFree(v->d_inst); cudaFree(v->d_env_inst); cudaFree(v->d_env); cudaFree(v->d_lp);
    cudaFree(v->d_pend_count); cudaFree(v->d_pend_env); cudaFree(v->d_lp_x); cudaFree(v->d_lp_meta);
    cudaFree(v->d_trace); cudaFree(v->d_cta_x);
    cudaFree(v->d_lpq); cudaFree(v->d_lpq_ring); cudaFree(v->d_lp_req); cudaFree(v->d_lp_resp); cudaFree(v->d_lp_own);
    cudaFree(v->d_prog); if (v->h_prog) cudaFreeHost(v->h_prog);
    cudaFree(v->d_rep_env); cudaFree(v->d_plan_x); cudaFree(v->d_plan_meta); cudaFree(v->d_plan_ok);
    if (v->stream) cudaStreamDestroy(v->stream);
    if (v->copy_stream) cudaStreamDestroy(v->copy_stream);
    if (v->chunk_done) cudaEventDestroy(v->chunk_done);
    if (v->dev_done) cudaEventDestroy(v->dev_done);
    if (v->copy_out_stream) cudaStreamDestroy(v->copy_out_stream);
    for (int k = 0; k < 2; ++k) { if (v->pipe_in[k]) cudaEventDestroy(v->pipe_in[k]); if (v->pipe_k[k]) cudaEventDestroy(v->pipe_k[k]); if (v->pipe_done[k]) cudaEventDestroy(v->pipe_done[k]); cudaFree(v->d_actions2[k]); cudaFree(v->d_rnd2[k]); cudaFree(v->d_out2[k]); }
    delete v;
    return 0;
}

int fjsp_vec_query(fjsp_vec *v, int64_t *o)
{
    if (!v || !o) { g_err = "fjsp_vec_query: null argument"; return -1; }
    o[0] = v->B; o[1] = v->nstate; o[2] = v->tb.eo.stride; o[3] = (int64_t)v->tb.io.stride * 4;
    o[4] = v->step_grid; o[5] = v->step_threads; o[6] = (int64_t)v->P.lp_stride; o[7] = v->launches;
    o[8] = v->env_warps; o[9] = v->srv_ctas; o[10] = v->n_slots; o[11] = (int64_t)v->step_smem_bytes;
    return 0;
}

int fjsp_vec_reset(fjsp_vec *v, void *stream, double *d_state64, float *d_state32)
{
    if (!v) { g_err = "fjsp_vec_reset: null handle"; return -1; }
    CK(cudaSetDevice(v->device));
    cudaStream_t st = (cudaStream_t)stream;
    // the first reset() is that of a new environment object; later ones reset a used object (reference quirks included)
    fjsp_reset_begin_kernel<<<v->grid, FJ_BLOCK, 0, st>>>(v->P, v->n_resets == 0 ? 1 : 0);
    v->n_resets += 1;
    launch_lp(v, st, 0);
    if (!v->plan_ready) {
        fjsp_plan_kernel<<<v->n_inst, 128, 0, st>>>(v->P, v->d_rep_env, v->n_inst, v->d_plan_x, v->d_plan_meta, v->d_plan_ok);
        v->plan_ready = 1;
        v->P.plan_ok = v->d_plan_ok;   // later launches may reset from the cache
    }
    int rc = dispatch(v, [&](auto V, auto SM) {
        // one LP slab per warp (the in-line fallback of an env without a solution slot writes it): the resume kernel's grid
        fjsp_reset_finish_kernel<decltype(V)::value, decltype(SM)::value><<<v->resume_grid, FJ_BLOCK, 0, st>>>(v->P, d_state64, d_state32);
        return 0;
    });
    if (rc) return rc;
    v->launches += 3;
    CK(cudaGetLastError());
    note_device_work(v, st);
    return 0;
}

int fjsp_vec_step(fjsp_vec *v, void *stream, int T, const int32_t *d_actions, const uint32_t *d_rnd,
                  int reward_policy, double completion, double tardiness, double energy, int autoreset,
                  double *d_state64, float *d_state32, double *d_reward, int32_t *d_done, int32_t *d_rec)
{
    if (!v || !d_actions || T < 1) { g_err = "fjsp_vec_step: null handle/actions or T < 1"; return -1; }
    if (v->variant != FJSP_SO_DFJSP && v->variant != FJSP_SO_FJSSP && (reward_policy < 0 || reward_policy > 3)) {
        g_err = "fjsp_vec_step: reward_policy must be 0..3 (the reference raises MyError otherwise)"; return -4;
    }
    CK(cudaSetDevice(v->device));
    FjStepArgs A;
    A.T = T; A.actions = d_actions; A.rnd = d_rnd; A.reward_policy = reward_policy; A.autoreset = autoreset;
    A.completion = completion; A.tardiness = tardiness; A.energy = energy;
    A.state = d_state64; A.state32 = d_state32; A.reward = d_reward; A.done = d_done; A.rec = d_rec;
    A.park_count = nullptr; A.park_env = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    if (v->prog_shift >= 0 && v->d_prog) {
        CK(cudaMemsetAsync(v->d_prog, 0, 4 * FJ_PROG_CHUNKS, st));
        A.prog_count = v->d_prog; A.prog_flag = v->h_prog_dev; A.prog_seq = v->prog_seq; A.prog_shift = v->prog_shift;
    }
    // parked-LP counts of the fallback rounds + the env CTAs finished in this launch (the LP queue's tickets are never reset)
    CK(cudaMemsetAsync(v->d_pend_count, 0, 4 * (FJ_ROUNDS + 2), st));
    int rc = dispatch(v, [&](auto V, auto SM) {
        constexpr int VV = decltype(V)::value, MM = decltype(SM)::value;
        A.park_count = v->d_pend_count; A.park_env = v->d_pend_env;
        // LP-aware packing pays for itself on rollouts (T >= 8: 16 us against >= 0.3 ms); a launch of a few steps
        // (T = 1 is the agents' step()) uses the static map: an env that meets an LP leaves its lockstep group anyway
        const bool pack = v->pack && T >= 8;
        FjParams Pl = v->P;
        if (pack) {
            fjsp_flag_kernel<<<(v->B + 255) / 256, 256, 0, st>>>(v->P, v->d_order, v->d_flags, T, v->pack_rows[0], v->pack_rows[1]);
            fjsp_pack_kernel<<<1, FJ_PACK_THREADS, 0, st>>>(v->P, v->d_order, v->d_flags, v->d_order_dyn, v->env_warps,
                                                            v->pack_cap[0], v->pack_cap[1], v->pack_cap[2], v->detach);
        } else Pl.order = v->d_order_static;
        v->last_order = (int32_t *)Pl.order;
        v->launches += 1 + 2 * (pack ? 1 : 0);
        fjsp_step_kernel<VV, MM><<<v->step_grid, v->step_threads, v->step_smem_bytes, st>>>(Pl, A);
        // resume rounds: an env can meet a reset and further order arrivals inside one launch; the
        // last round solves whatever is left in line
        // With the CTA-served LP and every instance's order-0 solution cached nothing can park
        // (an env parks only for a reset that finds no cached solution): skip the fallback rounds.
        const bool can_park = !(v->P.cta_lp == 1 && v->plan_ready && v->plan_all);
        for (int r = 0; can_park && r < FJ_ROUNDS; ++r) {
            launch_lp(v, st, r);
            const int *cnt = v->d_pend_count + r, *lst = v->d_pend_env + (size_t)(r & 1) * v->B;
            A.park_count = v->d_pend_count + r + 1; A.park_env = v->d_pend_env + (size_t)((r + 1) & 1) * v->B;
            if (r + 1 < FJ_ROUNDS) fjsp_resume_kernel<VV, MM, 1><<<v->resume_grid, FJ_BLOCK, 0, st>>>(v->P, A, cnt, lst);
            else fjsp_resume_kernel<VV, MM, 0><<<v->resume_grid, FJ_BLOCK, 0, st>>>(v->P, A, cnt, lst);
        }
        return 0;
    });
    if (rc) return rc;
    v->launches += (v->P.cta_lp == 1 && v->plan_ready && v->plan_all) ? 0 : 2 * FJ_ROUNDS;
    CK(cudaGetLastError());
    note_device_work(v, st);
    return 0;
}

static double fj_now_ms()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// staging for the host-buffer entry points: inputs always (H2D copies), outputs only for host buffers
// the device cannot write directly
static int ensure_stage(fjsp_vec *v, int T, bool outputs)
{
    const size_t n = (size_t)T * v->B;
    if (T > v->stage_T) {
        cudaFree(v->d_actions); cudaFree(v->d_rnd); v->d_actions = nullptr; v->d_rnd = nullptr;
        v->stage_T = 0;
        CK(cudaMalloc(&v->d_actions, n * 2 * 4));
        CK(cudaMalloc(&v->d_rnd, n * 2 * 4));
        v->stage_T = T;
    }
    if (outputs && T > v->stage_out_T) {
        cudaFree(v->d_done); cudaFree(v->d_rec); cudaFree(v->d_state64); cudaFree(v->d_state32); cudaFree(v->d_reward);
        v->d_done = v->d_rec = nullptr; v->d_state64 = v->d_reward = nullptr; v->d_state32 = nullptr;
        v->stage_out_T = 0;
        CK(cudaMalloc(&v->d_done, n * 4));
        CK(cudaMalloc(&v->d_rec, n * 8 * 4));
        CK(cudaMalloc(&v->d_state64, n * v->nstate * 8));
        CK(cudaMalloc(&v->d_state32, n * v->nstate * 4));
        CK(cudaMalloc(&v->d_reward, n * 8));
        v->stage_out_T = T;
    }
    return 0;
}

// device address of a host buffer the kernel can write directly (page-locked and mapped: cudaHostAlloc /
// cudaHostRegister memory under unified addressing), or null
static void *mapped_host(void *h)
{
    if (!h) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, h) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return (a.type == cudaMemoryTypeHost && a.devicePointer) ? a.devicePointer : nullptr;
}

int fjsp_vec_step_host(fjsp_vec *v, int T, const int32_t *h_actions, const uint32_t *h_rnd,
                       int reward_policy, double completion, double tardiness, double energy, int autoreset,
                       double *h_state64, float *h_state32, double *h_reward, int32_t *h_done, int32_t *h_rec)
{
    if (!v || !h_actions || T < 1) { g_err = "fjsp_vec_step_host: null handle/actions or T < 1"; return -1; }
    const double t_enter = getenv("FJSP_HOST_DEBUG") ? fj_now_ms() : 0.0;
    CK(cudaSetDevice(v->device));
    const size_t n = (size_t)T * v->B;
    cudaStream_t st = v->stream, cp = v->copy_stream;
    // How the outputs reach the host buffers:
    //  * page-locked (mapped) buffers are written by the step kernel itself while it runs: stores over the link, no
    //    staging and no copy after the launch (measured on the same env steps: 0.97 x the device-timed rate, against
    //    0.88 x for the chunked copies below).  FJSP_ZEROCOPY=0 turns this off, =state maps only the observation.
    //  * other buffers, T >= 16: the kernel writes device staging and reports, per chunk of 8 steps, when every env
    //    has written it; this thread polls the report (a word per chunk in page-locked memory) and queues the
    //    chunk's device-to-host copies on the copy stream while the kernel plays the next steps, so that only the
    //    last chunk's copy is left when the kernel ends (FJSP_PROGRESSIVE=0: off).
    //  * otherwise: device staging, copies after the launch (rollouts of T >= 64 in chunks of 32 steps).
    const char *zc = getenv("FJSP_ZEROCOPY"), *pg = getenv("FJSP_PROGRESSIVE");
    const bool can_park = !(v->P.cta_lp == 1 && v->plan_ready && v->plan_all);
    const int zmode = !zc ? 2 : (zc[0] == '0' ? 0 : (zc[0] == 's' ? 1 : 2));
    double *m_state64 = zmode ? (double *)mapped_host(h_state64) : nullptr, *m_reward = zmode == 2 ? (double *)mapped_host(h_reward) : nullptr;
    float *m_state32 = zmode ? (float *)mapped_host(h_state32) : nullptr;
    int32_t *m_done = zmode == 2 ? (int32_t *)mapped_host(h_done) : nullptr, *m_rec = zmode == 2 ? (int32_t *)mapped_host(h_rec) : nullptr;
    const bool staged = (h_state64 && !m_state64) || (h_state32 && !m_state32) || (h_reward && !m_reward) || (h_done && !m_done) || (h_rec && !m_rec);
    const bool progressive = staged && T >= 16 && !can_park && v->d_prog && !(pg && pg[0] == '0');
    int rc = ensure_stage(v, T, staged);
    if (rc) return rc;
    wait_device_work(v);
    CK(cudaMemcpyAsync(v->d_actions, h_actions, n * 2 * 4, cudaMemcpyHostToDevice, st));
    if (h_rnd) CK(cudaMemcpyAsync(v->d_rnd, h_rnd, n * 2 * 4, cudaMemcpyHostToDevice, st));
    // long rollouts are cut into chunks of 32 steps when outputs are staged: the device-to-host copy of one
    // chunk's outputs (copy stream) overlaps the kernels of the next chunk (compute stream).  Shorter
    // chunks measured slower (per-launch staging and launch overheads), so T < 64 is one chunk.
    const int nchunk = staged && T >= 64 ? T / 32 : 1;
    const size_t B = (size_t)v->B, ns = (size_t)v->nstate;
    if (progressive) {
        int shift = 3;
        while (((T - 1) >> shift) + 1 > FJ_PROG_CHUNKS) ++shift;
        const int nprog = ((T - 1) >> shift) + 1;
        v->prog_seq += 1;
        v->prog_shift = shift;
        rc = fjsp_vec_step(v, st, T, v->d_actions, h_rnd ? v->d_rnd : nullptr, reward_policy, completion, tardiness, energy, autoreset,
                           h_state64 ? (m_state64 ? m_state64 : v->d_state64) : nullptr, h_state32 ? (m_state32 ? m_state32 : v->d_state32) : nullptr,
                           h_reward ? (m_reward ? m_reward : v->d_reward) : nullptr, h_done ? (m_done ? m_done : v->d_done) : nullptr,
                           h_rec ? (m_rec ? m_rec : v->d_rec) : nullptr);
        v->prog_shift = -1;
        if (rc) return rc;
        auto copy_chunk = [&](int c) -> int {
            const int t0 = c << shift, t1 = std::min(T, (c + 1) << shift);
            const size_t o = (size_t)t0 * B, m = (size_t)(t1 - t0) * B;
            if (h_state64 && !m_state64) CK(cudaMemcpyAsync(h_state64 + o * ns, v->d_state64 + o * ns, m * ns * 8, cudaMemcpyDeviceToHost, cp));
            if (h_state32 && !m_state32) CK(cudaMemcpyAsync(h_state32 + o * ns, v->d_state32 + o * ns, m * ns * 4, cudaMemcpyDeviceToHost, cp));
            if (h_reward && !m_reward) CK(cudaMemcpyAsync(h_reward + o, v->d_reward + o, m * 8, cudaMemcpyDeviceToHost, cp));
            if (h_done && !m_done) CK(cudaMemcpyAsync(h_done + o, v->d_done + o, m * 4, cudaMemcpyDeviceToHost, cp));
            if (h_rec && !m_rec) CK(cudaMemcpyAsync(h_rec + o * 8, v->d_rec + o * 8, m * 8 * 4, cudaMemcpyDeviceToHost, cp));
            return 0;
        };
        const volatile unsigned *flag = v->h_prog;
        int next = 0;
        unsigned spins = 0;
        const bool dbg = getenv("FJSP_HOST_DEBUG") != nullptr;      // host-side time line of the call on stderr
        double t_flag[FJ_PROG_CHUNKS];
        const double t_launched = dbg ? fj_now_ms() : 0.0;
        while (next < nprog) {
            if (flag[next] == v->prog_seq) { if (dbg) t_flag[next] = fj_now_ms(); if (copy_chunk(next)) return -10; ++next; continue; }
            if ((++spins & 1023u) == 0 && cudaStreamQuery(st) != cudaErrorNotReady) break;   // finished (or failed): no more reports to wait for
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        const int seen = next;
        CK(cudaStreamSynchronize(st));
        const double t_kernel = dbg ? fj_now_ms() : 0.0;
        for (; next < nprog; ++next) if (copy_chunk(next)) return -10;
        CK(cudaStreamSynchronize(cp));
        if (dbg) {
            const double t_end = fj_now_ms();
            fprintf(stderr, "step_host T=%d: enter->launched %.3f ms, chunks reported %d/%d at", T, t_launched - t_enter, seen, nprog);
            for (int c = 0; c < seen; ++c) fprintf(stderr, " %.3f", t_flag[c] - t_launched);
            fprintf(stderr, ", kernel done %.3f, copies done %.3f (ms after launch)\n", t_kernel - t_launched, t_end - t_launched);
        }
        return 0;
    }
    for (int c = 0; c < nchunk; ++c) {
        const int t0 = (int)((long long)T * c / nchunk), t1 = (int)((long long)T * (c + 1) / nchunk);
        if (t1 == t0) continue;
        const size_t o = (size_t)t0 * B, m = (size_t)(t1 - t0) * B;
        rc = fjsp_vec_step(v, st, t1 - t0, v->d_actions + o * 2, h_rnd ? v->d_rnd + o * 2 : nullptr, reward_policy,
                           completion, tardiness, energy, autoreset,
                           h_state64 ? (m_state64 ? m_state64 : v->d_state64) + o * ns : nullptr,
                           h_state32 ? (m_state32 ? m_state32 : v->d_state32) + o * ns : nullptr,
                           h_reward ? (m_reward ? m_reward : v->d_reward) + o : nullptr,
                           h_done ? (m_done ? m_done : v->d_done) + o : nullptr,
                           h_rec ? (m_rec ? m_rec : v->d_rec) + o * 8 : nullptr);
        if (rc) return rc;
        if (!staged) continue;
        CK(cudaEventRecord(v->chunk_done, st));
        CK(cudaStreamWaitEvent(cp, v->chunk_done, 0));
        if (h_state64 && !m_state64) CK(cudaMemcpyAsync(h_state64 + o * ns, v->d_state64 + o * ns, m * ns * 8, cudaMemcpyDeviceToHost, cp));
        if (h_state32 && !m_state32) CK(cudaMemcpyAsync(h_state32 + o * ns, v->d_state32 + o * ns, m * ns * 4, cudaMemcpyDeviceToHost, cp));
        if (h_reward && !m_reward) CK(cudaMemcpyAsync(h_reward + o, v->d_reward + o, m * 8, cudaMemcpyDeviceToHost, cp));
        if (h_done && !m_done) CK(cudaMemcpyAsync(h_done + o, v->d_done + o, m * 4, cudaMemcpyDeviceToHost, cp));
        if (h_rec && !m_rec) CK(cudaMemcpyAsync(h_rec + o * 8, v->d_rec + o * 8, m * 8 * 4, cudaMemcpyDeviceToHost, cp));
    }
    CK(cudaStreamSynchronize(st));
    if (staged) CK(cudaStreamSynchronize(cp));
    return 0;
}

// Pipelined form of fjsp_vec_step_host for callers whose actions do not depend on the previous call's
// outputs (rule-based rollouts, replays): _begin queues the host-to-device copy of the inputs (input copy
// stream), the launch (compute stream, outputs into one of two device staging slots) and the device-to-host
// copy of the outputs (output copy stream) and returns; _wait blocks until the OLDEST call begun has
// delivered its outputs.  Up to two calls in flight: the copies of one run beside the kernels of the other,
// each direction on its own copy engine.  (Outputs go through staging + DMA here, not through kernel stores
// into host memory: SM stores of 120-byte rows reach ~10 GB/s on the link, the copy engine 2-3 x that, and
// with two calls in flight the copy is hidden anyway.)  Host buffers should be page-locked; pageable ones
// make the copies synchronous (correct, not overlapped).
int fjsp_vec_step_host_begin(fjsp_vec *v, int T, const int32_t *h_actions, const uint32_t *h_rnd,
                             int reward_policy, double completion, double tardiness, double energy, int autoreset,
                             double *h_state64, float *h_state32, double *h_reward, int32_t *h_done, int32_t *h_rec)
{
    if (!v || !h_actions || T < 1) { g_err = "fjsp_vec_step_host_begin: null handle/actions or T < 1"; return -1; }
    if (v->pipe_begun - v->pipe_waited >= 2) { g_err = "fjsp_vec_step_host_begin: two calls already in flight (call fjsp_vec_step_host_wait first)"; return -7; }
    CK(cudaSetDevice(v->device));
    const int slot = (int)(v->pipe_begun & 1);
    const size_t n = (size_t)T * v->B, ns = (size_t)v->nstate;
    if (T > v->pipe_T) {
        CK(cudaDeviceSynchronize());
        for (int k = 0; k < 2; ++k) {
            cudaFree(v->d_actions2[k]); cudaFree(v->d_rnd2[k]); cudaFree(v->d_out2[k]);
            v->d_actions2[k] = nullptr; v->d_rnd2[k] = nullptr; v->d_out2[k] = nullptr;
            CK(cudaMalloc(&v->d_actions2[k], n * 2 * 4));
            CK(cudaMalloc(&v->d_rnd2[k], n * 2 * 4));
            CK(cudaMalloc(&v->d_out2[k], n * (ns * 12 + 8 + 4 + 32)));   // state64 | state32 | reward | done | rec
        }
        v->pipe_T = T;
    }
    unsigned char *o = v->d_out2[slot];
    double *d_state64 = (double *)o; o += n * ns * 8;
    double *d_reward = (double *)o; o += n * 8;
    float *d_state32 = (float *)o; o += n * ns * 4;
    int32_t *d_done = (int32_t *)o; o += n * 4;
    int32_t *d_rec = (int32_t *)o;
    cudaStream_t st = v->stream, cin = v->copy_stream, cout = v->copy_out_stream;
    wait_device_work(v);
    // the slot's previous user (call begun - 2) was waited for by the host: its copies are done
    // The (small) input copy goes on the compute stream, behind the previous call's kernels (on a copy stream of
    // its own it can end up queued behind the previous call's output copy, which waits for that call's kernels);
    // only the output copy (17 MB per call in the bench) needs to overlap the kernels.
    (void)cin;
    CK(cudaMemcpyAsync(v->d_actions2[slot], h_actions, n * 2 * 4, cudaMemcpyHostToDevice, st));
    if (h_rnd) CK(cudaMemcpyAsync(v->d_rnd2[slot], h_rnd, n * 2 * 4, cudaMemcpyHostToDevice, st));
    int rc = fjsp_vec_step(v, st, T, v->d_actions2[slot], h_rnd ? v->d_rnd2[slot] : nullptr, reward_policy, completion, tardiness, energy,
                           autoreset, h_state64 ? d_state64 : nullptr, h_state32 ? d_state32 : nullptr, h_reward ? d_reward : nullptr,
                           h_done ? d_done : nullptr, h_rec ? d_rec : nullptr);
    if (rc) return rc;
    CK(cudaEventRecord(v->pipe_k[slot], st));
    CK(cudaStreamWaitEvent(cout, v->pipe_k[slot], 0));
    if (h_state64) CK(cudaMemcpyAsync(h_state64, d_state64, n * ns * 8, cudaMemcpyDeviceToHost, cout));
    if (h_state32) CK(cudaMemcpyAsync(h_state32, d_state32, n * ns * 4, cudaMemcpyDeviceToHost, cout));
    if (h_reward) CK(cudaMemcpyAsync(h_reward, d_reward, n * 8, cudaMemcpyDeviceToHost, cout));
    if (h_done) CK(cudaMemcpyAsync(h_done, d_done, n * 4, cudaMemcpyDeviceToHost, cout));
    if (h_rec) CK(cudaMemcpyAsync(h_rec, d_rec, n * 8 * 4, cudaMemcpyDeviceToHost, cout));
    CK(cudaEventRecord(v->pipe_done[slot], cout));
    v->pipe_begun += 1;
    return 0;
}

int fjsp_vec_step_host_wait(fjsp_vec *v)
{
    if (!v) { g_err = "fjsp_vec_step_host_wait: null handle"; return -1; }
    if (v->pipe_waited >= v->pipe_begun) return 0;
    CK(cudaSetDevice(v->device));
    CK(cudaEventSynchronize(v->pipe_done[v->pipe_waited & 1]));
    v->pipe_waited += 1;
    return 0;
}

int fjsp_vec_reset_host(fjsp_vec *v, double *h_state64, float *h_state32)
{
    if (!v) { g_err = "fjsp_vec_reset_host: null handle"; return -1; }
    CK(cudaSetDevice(v->device));
    int rc = ensure_stage(v, 1, true);
    if (rc) return rc;
    cudaStream_t st = v->stream;
    wait_device_work(v);
    rc = fjsp_vec_reset(v, st, h_state64 ? v->d_state64 : nullptr, h_state32 ? v->d_state32 : nullptr);
    if (rc) return rc;
    const size_t n = (size_t)v->B * v->nstate;
    if (h_state64) CK(cudaMemcpyAsync(h_state64, v->d_state64, n * 8, cudaMemcpyDeviceToHost, st));
    if (h_state32) CK(cudaMemcpyAsync(h_state32, v->d_state32, n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

int fjsp_vec_slots(fjsp_vec *v, int32_t *h_out, int capacity)
{
    if (!v) { g_err = "fjsp_vec_slots: null handle"; return -1; }
    if (!h_out) return v->n_slots;
    if (capacity < v->n_slots) { g_err = "fjsp_vec_slots: buffer too small"; return -1; }
    CK(cudaSetDevice(v->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_out, v->last_order, (size_t)v->n_slots * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < v->n_slots; ++i) if (h_out[i] >= 0) h_out[i] &= FJ_SLOT_DETACHED - 1;   // bit 30: the env free-runs in that launch
    return v->n_slots;
}

int fjsp_vec_trace(fjsp_vec *v, int64_t *h_out, int clear)
{
    if (!v) { g_err = "fjsp_vec_trace: null handle"; return -1; }
    if (!v->d_trace) { g_err = "fjsp_vec_trace: not a trace build (compile with -DFJ_TRACE)"; return -6; }
    CK(cudaSetDevice(v->device));
    CK(cudaDeviceSynchronize());
    const size_t bytes = (size_t)v->step_grid * FJ_TRACE_ROWS * 8 * 8;
    if (h_out) CK(cudaMemcpy(h_out, v->d_trace, bytes, cudaMemcpyDeviceToHost));
    if (clear) CK(cudaMemset(v->d_trace, 0, bytes));
    return 0;
}

__global__ void fjsp_info_kernel(FjParams P, long long *out)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= P.B) return;
    const int32_t *s = (const int32_t *)(P.env + (size_t)e * P.eo.stride + P.eo.scal);
    long long *o = out + (size_t)e * 12;
    const long long dp = *(const long long *)(s + FJ_S_DELAY_PROC), du = *(const long long *)(s + FJ_S_DELAY_UNPROC);
    o[0] = s[FJ_S_TIME]; o[1] = s[FJ_S_STEPS]; o[2] = s[FJ_S_COMPLETION]; o[3] = dp + du;
    o[4] = *(const long long *)(s + FJ_S_ENERGY); o[5] = s[FJ_S_LPSOLVES]; o[6] = s[FJ_S_LPITERS];
    o[7] = s[FJ_S_ERROR]; o[8] = s[FJ_S_DONE]; o[9] = s[FJ_S_NEXTORDER]; o[10] = s[FJ_S_EPISODES]; o[11] = du;
}

int fjsp_vec_info(fjsp_vec *v, int64_t *h_out)
{
    if (!v || !h_out) { g_err = "fjsp_vec_info: null argument"; return -1; }
    CK(cudaSetDevice(v->device));
    CK(cudaDeviceSynchronize());   // steps may have been queued on a caller's stream
    long long *d = nullptr;
    CK(cudaMalloc(&d, (size_t)v->B * 12 * 8));
    fjsp_info_kernel<<<(v->B + 255) / 256, 256, 0, v->stream>>>(v->P, d);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_out, d, (size_t)v->B * 12 * 8, cudaMemcpyDeviceToHost, v->stream));
    CK(cudaStreamSynchronize(v->stream));
    CK(cudaFree(d));
    return 0;
}
}
